// fp32 SIMT GEMM with a fused per-column epilogue:  Y = act((X @ W) * scale + shift)
//   X [M,K] row-major (ldx), W [K,N] row-major, Y [M,N] row-major (ldy).
// Exact-fp32 (FFMA) path used for (1) the FlexConv contraction  A[N, 4*Din] @ Theta_ext[4*Din, Dout]
// and (2) the dense 1x1 stacks, whenever bit-faithful fp32 accumulation is wanted; the
// tensor-core (tcgen05, fp16-pair split) path lives in gemm_tc16.cu.
// 128 x BN x 16 CTA tile, 256 threads, 8 x (BN/16) register tile per thread, register-staged
// double buffering of the next K slab.
#include "common.cuh"

namespace dh3d {

__device__ __forceinline__ float gemm_act(float v, int act) {
  if (act == DH3D_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DH3D_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

template <int BN>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
             const float* __restrict__ scale, const float* __restrict__ shift, int act,
             float* __restrict__ Y, int ldy, int M, int K, int N) {
  constexpr int BM = 128, BK = 16;
  constexpr int TN = BN / 16;  // columns per thread (8 for BN=128, 4 for BN=64)
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid: rows ty*8.., cols tx*4.. (+64)

  // global->smem staging assignments
  // X tile: 128 rows x 16 cols = 512 float4; thread loads rows (tid>>2) and (tid>>2)+64, col4 = tid&3
  const int xr = tid >> 2, xc = (tid & 3) * 4;
  // W tile: 16 rows x BN cols = 4*BN float4
  constexpr int W4 = BK * BN / 4;           // float4 count
  constexpr int WPT = W4 / 256;             // per thread (2 for BN=128, 1 for BN=64)

  float4 xa[2], wb[WPT];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = m0 + xr + h * 64;
      const int c = k0 + xc;
      xa[h] = (r < M && c < K) ? ldg4(X + (long long)r * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < WPT; ++h) {
      const int e = tid + h * 256;
      const int r = e / (BN / 4), c = (e % (BN / 4)) * 4;
      wb[h] = (k0 + r < K && n0 + c < N) ? ldg4(W + (long long)(k0 + r) * N + n0 + c)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = xr + h * 64;
      As[buf][xc + 0][r] = xa[h].x; As[buf][xc + 1][r] = xa[h].y;
      As[buf][xc + 2][r] = xa[h].z; As[buf][xc + 3][r] = xa[h].w;
    }
#pragma unroll
    for (int h = 0; h < WPT; ++h) {
      const int e = tid + h * 256;
      const int r = e / (BN / 4), c = (e % (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][r][c]) = wb[h];
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][(j / 4) * 64 + tx * 4]);
        b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int j = 0; j < TN; j += 4) {
    const int c = n0 + (j / 4) * 64 + tx * 4;  // column groups 64 apart: conflict-free LDS.128
    if (c >= N) continue;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) sc = ldg4(scale + c);
    if (shift) sh = ldg4(shift + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + ty * 8 + i;
      if (r >= M) continue;
      float4 o;
      o.x = gemm_act(fmaf(acc[i][j + 0], sc.x, sh.x), act);
      o.y = gemm_act(fmaf(acc[i][j + 1], sc.y, sh.y), act);
      o.z = gemm_act(fmaf(acc[i][j + 2], sc.z, sh.z), act);
      o.w = gemm_act(fmaf(acc[i][j + 3], sc.w, sh.w), act);
      *reinterpret_cast<float4*>(Y + (long long)r * ldy + c) = o;
    }
  }
}

int linear_simt_launch(const float* x, int ldx, const float* w, const float* scale,
                       const float* shift, int act, float* y, int ldy, int M, int K, int N,
                       cudaStream_t st) {
  if (!x || !w || !y) return DH3D_ERR_NULL;
  if (M <= 0 || K <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (K % 4 || N % 4 || ldx % 4 || ldy % 4 || ldx < K || ldy < N) return DH3D_ERR_DIM;
  if ((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift) & 15) != 0)
    return DH3D_ERR_ALIGN;
  if (N > 64) {
    dim3 grid(ceil_div(M, 128), ceil_div(N, 128));
    sgemm_kernel<128><<<grid, 256, 0, st>>>(x, ldx, w, scale, shift, act, y, ldy, M, K, N);
  } else {
    dim3 grid(ceil_div(M, 128), ceil_div(N, 64));
    sgemm_kernel<64><<<grid, 256, 0, st>>>(x, ldx, w, scale, shift, act, y, ldy, M, K, N);
  }
  return launch_status();
}

}  // namespace dh3d
