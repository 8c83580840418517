// Fused squeeze/excite bottleneck of flex_conv_dilate (core/backbones.py:45-55 se_res_bottleneck with
// add_se='max_pool', called at :84-88):
//     pooled = flex_pool(x, nbr)                      user_ops/kernels/flex_pool_kernel_gpu.cu.cc:30-63 (max only)
//     gate   = sigmoid(W2^T relu(W1^T pooled + b1) + b2)      two feature_conv1d_1 layers without BN
//     out    = relu(x + x * gate)
// Unfused this is four launches (flex_pool, two skinny GEMMs with K or N = C/4, se_excite) that write and
// re-read the pooled features, the hidden layer and the gate: 201 us per 32 x 8192 x 64 step (r1n op table)
// for ~210 MB of compulsory traffic.  Here C/4 lanes own one point (4 channels each, 16-byte accesses): they
// gather-max the K neighbour rows, multiply their 4 pooled channels into all H = C/4 hidden partial sums,
// reduce-scatter those across the lane group with shuffles (lane j ends up with hidden unit j), all-gather the
// H activations back and finish the 4 gate channels they own.  Nothing but x, nbr and out touches HBM.
// Exact fp32 FFMA arithmetic (no tensor cores: 2 x C x H MACs per point is ~1 GFLOP per step).
#include <float.h>

#include "common.cuh"

namespace dh3d {

template <int C>
struct SeCfg {
  static constexpr int H = C / 4;     // hidden units == lanes per point
  static constexpr int LP = C / 4;
  static constexpr int PP = 32 / LP;  // points per warp iteration
  static constexpr bool kW1Reg = (C == 64);
};

template <int C>
__global__ void __launch_bounds__(256, 2)
se_pool_excite_kernel(const float* __restrict__ x, const int32_t* __restrict__ nbr,
                      const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out,
                      int rows, int n, int K) {
  using Cfg = SeCfg<C>;
  constexpr int H = Cfg::H, LP = Cfg::LP, PP = Cfg::PP;
  __shared__ __align__(16) float s_w2[H * C];
  __shared__ __align__(16) float s_w1[Cfg::kW1Reg ? 4 : C * H];

  for (int e = threadIdx.x; e < H * C / 4; e += blockDim.x)
    reinterpret_cast<float4*>(s_w2)[e] = ldg4(w2 + 4 * e);
  if constexpr (!Cfg::kW1Reg)   // [i][j/4][q][4]: the lanes of a group read consecutive 16-byte chunks
    for (int e = threadIdx.x; e < C * H / 4; e += blockDim.x) {
      const int row = e / (H / 4), jq = e - row * (H / 4);
      reinterpret_cast<float4*>(s_w1)[((row & 3) * (H / 4) + jq) * LP + (row >> 2)] = ldg4(w1 + 4 * e);
    }

  const int lane = threadIdx.x & 31;
  const int q = lane & (LP - 1);          // channel chunk 4q..4q+3 and, after the reduction, hidden unit q
  const int grp0 = lane & ~(LP - 1);      // first lane of this point's lane group
  const int sub = lane / LP;              // point of the warp iteration

  unsigned long long w1r[Cfg::kW1Reg ? 4 : 1][Cfg::kW1Reg ? H / 2 : 1];   // W1[4q+i][2j], [2j+1] packed
  if constexpr (Cfg::kW1Reg) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < H; j += 4) {
        const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(w1 + (4 * q + i) * H + j));
        w1r[i][j / 2] = t.x; w1r[i][j / 2 + 1] = t.y;
      }
  }
  const float bias1 = __ldg(b1 + q);
  const float4 bias2 = ldg4(b2 + 4 * q);
  __syncthreads();

  const int groups = (rows + PP - 1) / PP;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool k8 = (K == 8);
  const int nshift = (n & (n - 1)) == 0 ? __ffs(n) - 1 : -1;   // power-of-two clouds: no integer division per point

  int4 ia = make_int4(0, 0, 0, 0), ib = ia;   // neighbour ids of the NEXT iteration (K == 8 only)
  auto row_of = [&](int g) { const int r = g * PP + sub; return r < rows ? r : rows - 1; };
  if (k8 && warp0 < groups) {
    const int4* p = reinterpret_cast<const int4*>(nbr + (long long)row_of(warp0) * 8);
    ia = __ldg(p); ib = __ldg(p + 1);
  }
  for (int g = warp0; g < groups; g += nwarps) {
    const int r = row_of(g);
    const bool valid = g * PP + sub < rows;
    const int cloud0 = nshift >= 0 ? ((r >> nshift) << nshift) : (r / n) * n;
    const float* fb = x + (long long)cloud0 * C + 4 * q;
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    const float4 xv = ldg4(x + (long long)r * C + 4 * q);
    if (k8) {
      const int id[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
      float4 f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = ldg4(fb + (long long)id[k] * C);
      if (g + nwarps < groups) {
        const int4* p = reinterpret_cast<const int4*>(nbr + (long long)row_of(g + nwarps) * 8);
        ia = __ldg(p); ib = __ldg(p + 1);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // first maximum wins (flex_pool_kernel.cc:41-57); only the value is used
        if (best.x < f[k].x) best.x = f[k].x;
        if (best.y < f[k].y) best.y = f[k].y;
        if (best.z < f[k].z) best.z = f[k].z;
        if (best.w < f[k].w) best.w = f[k].w;
      }
    } else {
      const int32_t* nb = nbr + (long long)r * K;
      for (int k = 0; k < K; ++k) {
        const float4 t = ldg4(fb + (long long)__ldg(nb + k) * C);
        if (best.x < t.x) best.x = t.x;
        if (best.y < t.y) best.y = t.y;
        if (best.z < t.z) best.z = t.z;
        if (best.w < t.w) best.w = t.w;
      }
    }
    // hidden partial sums over this lane's 4 pooled channels (packed fp32x2 FMAs: two hidden units per instruction)
    unsigned long long part2[H / 2];
    const float pv[4] = {best.x, best.y, best.z, best.w};
#pragma unroll
    for (int j = 0; j < H / 2; ++j) part2[j] = 0ull;
    if constexpr (Cfg::kW1Reg) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < H / 2; ++j) ffma2(part2[j], w1r[i][j], pv[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < H; j += 4) {
          const ulonglong2 t = reinterpret_cast<const ulonglong2*>(s_w1)[(i * (H / 4) + j / 4) * LP + q];
          ffma2(part2[j / 2], t.x, pv[i]);
          ffma2(part2[j / 2 + 1], t.y, pv[i]);
        }
    }
    float part[H];
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
      const float2 t = unpack2(part2[j]);
      part[2 * j] = t.x; part[2 * j + 1] = t.y;
    }
    // reduce-scatter across the lane group: after the last step lane q holds the full sum of hidden unit q
#pragma unroll
    for (int s = H / 2; s >= 1; s >>= 1) {
      const bool up = (q & s) != 0;
#pragma unroll
      for (int t = 0; t < s; ++t) {
        const float send = up ? part[t] : part[t + s];
        const float keep = up ? part[t + s] : part[t];
        part[t] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    const float h = fmaxf(part[0] + bias1, 0.f);
    unsigned long long acc01, acc23;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc01) : "f"(bias2.x), "f"(bias2.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc23) : "f"(bias2.z), "f"(bias2.w));
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const float hj = __shfl_sync(0xffffffffu, h, grp0 | j);
      const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(s_w2 + j * C + 4 * q);
      ffma2(acc01, t.x, hj);
      ffma2(acc23, t.y, hj);
    }
    const float2 a01 = unpack2(acc01), a23 = unpack2(acc23);
    float4 o;   // sigmoid with the fast reciprocal (2 ulp): the gate feeds a 1e-4 feature tolerance
    o.x = fmaxf(xv.x + xv.x * __fdividef(1.f, 1.f + __expf(-a01.x)), 0.f);
    o.y = fmaxf(xv.y + xv.y * __fdividef(1.f, 1.f + __expf(-a01.y)), 0.f);
    o.z = fmaxf(xv.z + xv.z * __fdividef(1.f, 1.f + __expf(-a23.x)), 0.f);
    o.w = fmaxf(xv.w + xv.w * __fdividef(1.f, 1.f + __expf(-a23.y)), 0.f);
    if (valid) *reinterpret_cast<float4*>(out + (long long)r * C + 4 * q) = o;
  }
}

// x [B*N, C] point-major, nbr [B*N, K] (ids within the cloud), w1 [C, C/4], b1 [C/4], w2 [C/4, C], b2 [C]
int se_pool_excite_launch(const float* x, const int32_t* nbr, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* out, int B, int N, int K, int C, int H, cudaStream_t st) {
  if (!x || !nbr || !w1 || !b1 || !w2 || !b2 || !out) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || C <= 0 || H <= 0) return DH3D_ERR_DIM;
  if ((long long)B * N > 0x7fffffffLL) return DH3D_ERR_DIM;
  if (H * 4 != C || (C != 64 && C != 128)) return DH3D_ERR_UNSUPPORTED;
  if ((((uintptr_t)x | (uintptr_t)out | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)b2 | (uintptr_t)nbr) & 15) != 0)
    return DH3D_ERR_ALIGN;
  const int rows = B * N;
  if (C == 64) {
    const int groups = (rows + 1) / 2;
    int blocks = (groups + 7) / 8;
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    se_pool_excite_kernel<64><<<blocks, 256, 0, st>>>(x, nbr, w1, b1, w2, b2, out, rows, N, K);
  } else {
    int blocks = (rows + 7) / 8;
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    se_pool_excite_kernel<128><<<blocks, 256, 0, st>>>(x, nbr, w1, b1, w2, b2, out, rows, N, K);
  }
  return launch_status();
}

}  // namespace dh3d
