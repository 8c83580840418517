// Backward passes of the custom ops and the transposed FlexConv.
//
// Reference:  FlexConvGrad        user_ops/kernels/flex_conv_kernel.cc:75-163 (CPU loops),
//                                 flex_conv_kernel_gpu.cu.cc:160-400,443-520 (BackwardThetaKernel: one CTA per
//                                 (dout,din) walking all B*N*K gathers; BackwardFeatureKernel: atomics)
//             FlexPoolGrad        flex_pool_kernel.cc:63-95, flex_pool_kernel_gpu.cu.cc:65-98
//             ConvPointsetGrad    conv_pointset_kernel.cc:72-147
//             FlexDeconv forward  flex_deconv_kernel.cc:25-70
//             GroupPointGrad      tf_ops/grouping/tf_grouping_g.cu:114-133; GatherPointGrad
//                                 tf_ops/sampling/tf_sampling_g.cu:183-192; ThreeInterpolateGrad
//                                 tf_ops/interpolation/tf_interpolate.cpp:131-153
//
// Same factoring as the forward pass (flexconv.cu): the FlexConv weight is affine in the offset, so with
//     q(n,k) = (1, dx, dy, dz),  d = p[nbr(n,k)] - p[nbr(n,0)]        (the backward centres on nbr(n,0), :134,:196)
//     A[n, p*Din + c] = sum_k q_p(n,k) * f[nbr(n,k), c]
//     Theta_ext = [bias ; theta_x ; theta_y ; theta_z]                 (4*Din x Dout)
// the three gradients are
//     [grad_bias ; grad_theta] = A^T @ G                               (a TN GEMM reduced over all B*N rows)
//     H = G @ Theta_ext^T                                              (rows x 4*Din)
//     grad_f[nbr(n,k), c] += sum_p q_p(n,k) * H[n, p*Din + c]          (vector atomics)
// i.e. 2 dense GEMMs in which K no longer appears plus one gather and one scatter, instead of the reference's
// B*N*K*Din*Dout scalar gathers per parameter gradient.  The parameter gradients are reduced in a fixed order
// (split-K partials + ordered sum: run-to-run deterministic); the feature gradient uses fp32 atomics like the
// reference (tensorflow::CudaAtomicAdd, :372), so its summation order is not fixed.
// FlexDeconv forward is the same scatter fed by  Z = F @ [bias | theta_x | theta_y | theta_z]  read at row nbr(n,0).
// All kernels are point-major ([rows, C]); the reference-layout entry points transpose through the workspace.
#include "common.cuh"

namespace dh3d {

int linear_simt_launch(const float* x, int ldx, const float* w, const float* scale, const float* shift, int act,
                       float* y, int ldy, int M, int K, int N, cudaStream_t st);
int transpose_launch(const void* src, void* dst, int B, int R, int C, cudaStream_t st);
int transpose_strided_launch(const void* src, long long sbs, int lds, void* dst, long long sbd, int ldd, int B,
                             int R, int C, cudaStream_t st);

static inline int pad4(int v) { return (v + 3) & ~3; }
static inline int ew_grid(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// moments centred on nbr(n,0):  A[r, p*din + c]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flex_moments_c0_kernel(const float* __restrict__ feat, const float* __restrict__ xyz,
                       const int32_t* __restrict__ nbr, float* __restrict__ A, long long rows, int n, int k,
                       int din) {
  const int cv = din >> 2;
  const long long total = rows * cv;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cv;
    const int c = (int)(e - r * cv) << 2;
    const long long b = r / n;
    const float* fbase = feat + b * n * (long long)din + c;
    const float* pbase = xyz + b * n * 3LL;
    const int32_t* nb = nbr + r * k;
    const int g0 = __ldg(nb);
    const float px = __ldg(pbase + g0 * 3LL), py = __ldg(pbase + g0 * 3LL + 1), pz = __ldg(pbase + g0 * 3LL + 2);
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), mx = m0, my = m0, mz = m0;
#pragma unroll 4
    for (int kk = 0; kk < k; ++kk) {
      const int g = __ldg(nb + kk);
      const float4 f = ldg4(fbase + (long long)g * din);
      const float dx = __ldg(pbase + g * 3LL) - px;
      const float dy = __ldg(pbase + g * 3LL + 1) - py;
      const float dz = __ldg(pbase + g * 3LL + 2) - pz;
      m0.x += f.x; m0.y += f.y; m0.z += f.z; m0.w += f.w;
      mx.x = fmaf(dx, f.x, mx.x); mx.y = fmaf(dx, f.y, mx.y); mx.z = fmaf(dx, f.z, mx.z); mx.w = fmaf(dx, f.w, mx.w);
      my.x = fmaf(dy, f.x, my.x); my.y = fmaf(dy, f.y, my.y); my.z = fmaf(dy, f.z, my.z); my.w = fmaf(dy, f.w, my.w);
      mz.x = fmaf(dz, f.x, mz.x); mz.y = fmaf(dz, f.y, mz.y); mz.z = fmaf(dz, f.z, mz.z); mz.w = fmaf(dz, f.w, mz.w);
    }
    float* a = A + r * 4LL * din + c;
    *reinterpret_cast<float4*>(a) = m0;
    *reinterpret_cast<float4*>(a + din) = mx;
    *reinterpret_cast<float4*>(a + 2 * din) = my;
    *reinterpret_cast<float4*>(a + 3 * din) = mz;
  }
}

// ---------------------------------------------------------------------------------------------
// TN GEMM with a split reduction:  P[s] = A[rows_s, Ka]^T @ G[rows_s, Nb]  (64x64 tile, 4x4 per thread),
// then an ordered sum over s.  Ka, Nb multiples of 4.
// ---------------------------------------------------------------------------------------------
constexpr int kTnTile = 64, kTnRows = 16;

__global__ void __launch_bounds__(256)
gemm_tn_partial_kernel(const float* __restrict__ A, int lda, const float* __restrict__ G, int ldg,
                       float* __restrict__ P, long long rows, int Ka, int Nb, int tiles_n, long long rows_per_split) {
  __shared__ __align__(16) float As[2][kTnRows][kTnTile];
  __shared__ __align__(16) float Gs[2][kTnRows][kTnTile];
  const int tid = threadIdx.x;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const int a0 = tm * kTnTile, g0 = tn * kTnTile;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  long long r_end = r_begin + rows_per_split;
  if (r_end > rows) r_end = rows;
  const int lr = tid >> 4, lc = (tid & 15) << 2;  // staging: row lr, 4 columns at lc
  const int ty = tid >> 4, tx = tid & 15;         // compute: A-columns ty*4.., G-columns tx*4..

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](long long r0, float4& av, float4& gv) {
    const long long r = r0 + lr;
    av = (r < r_end && a0 + lc < Ka) ? ldg4(A + r * lda + a0 + lc) : z4;
    gv = (r < r_end && g0 + lc < Nb) ? ldg4(G + r * ldg + g0 + lc) : z4;
  };
  float4 av, gv;
  fetch(r_begin, av, gv);
  int buf = 0;
  for (long long r0 = r_begin; r0 < r_end; r0 += kTnRows) {
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = av;
    *reinterpret_cast<float4*>(&Gs[buf][lr][lc]) = gv;
    __syncthreads();
    if (r0 + kTnRows < r_end) fetch(r0 + kTnRows, av, gv);
#pragma unroll
    for (int rr = 0; rr < kTnRows; ++rr) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][rr][ty * 4]);
      const float4 g = *reinterpret_cast<const float4*>(&Gs[buf][rr][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], gg[j], acc[i][j]);
    }
    buf ^= 1;  // the next store goes to the other buffer: one barrier per slab is enough
  }
  float* Ps = P + (long long)blockIdx.y * Ka * Nb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ar = a0 + ty * 4 + i, gc = g0 + tx * 4;
    if (ar < Ka && gc < Nb)
      *reinterpret_cast<float4*>(Ps + (long long)ar * Nb + gc) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// out element (row r of the padded [rows_p x Nb] product, column o) -> two destination blocks:
//   r <  split_row : dst0[(r - 0) ...]          r >= split_row : dst1[...]
// with padded -> logical index mapping (row blocks of `rowsp` padded / `rowsl` logical rows).
__global__ void gemm_tn_reduce_kernel(const float* __restrict__ P, int splits, int Ka, int Nb, float* __restrict__ dst0,
                                      float* __restrict__ dst1, int blocks0, int rowsp, int rowsl, int row_off,
                                      int colsl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ka * Nb) return;
  const int r = i / Nb, o = i % Nb;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += P[(long long)k * Ka * Nb + i];
  if (o >= colsl) return;
  const int rr = r - row_off;
  if (rr < 0) {                      // leading rows (ConvPointset: the all-ones column -> grad_bias)
    if (r == 0 && dst0) dst0[o] = s;
    return;
  }
  const int blk = rr / rowsp, c = rr % rowsp;
  if (c >= rowsl) return;
  if (blk < blocks0) {
    if (dst0) dst0[((long long)blk * rowsl + c) * colsl + o] = s;
  } else if (dst1) {
    dst1[((long long)(blk - blocks0) * rowsl + c) * colsl + o] = s;
  }
}

static int tn_splits(long long rows, int Ka, int Nb) {
  const int tiles = ceil_div(Ka, kTnTile) * ceil_div(Nb, kTnTile);
  long long s = ceil_div(kNumSMs * 4, tiles);
  const long long max_s = (rows + 255) / 256;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 65535) s = 65535;
  return (int)s;
}
static size_t tn_partial_bytes(long long rows, int Ka, int Nb) {
  return align_up((size_t)tn_splits(rows, Ka, Nb) * Ka * Nb * sizeof(float), 256);
}
static int gemm_tn_launch(const float* A, int lda, const float* G, int ldg, float* P, long long rows, int Ka, int Nb,
                          cudaStream_t st) {
  const int splits = tn_splits(rows, Ka, Nb);
  long long rps = (rows + splits - 1) / splits;
  rps = (rps + kTnRows - 1) / kTnRows * kTnRows;
  const int tiles_n = ceil_div(Nb, kTnTile);
  dim3 grid(ceil_div(Ka, kTnTile) * tiles_n, splits);
  gemm_tn_partial_kernel<<<grid, 256, 0, st>>>(A, lda, G, ldg, P, rows, Ka, Nb, tiles_n, rps);
  return launch_status();
}

// W[o, p*dinp + c] = Theta_ext[p*din + c, o]  (transposed; zero padding) -- the weight of  H = G @ Theta_ext^T
__global__ void theta_ext_t_kernel(const float* __restrict__ theta, const float* __restrict__ bias, float* __restrict__ W,
                                   int din, int dout, int dinp, int doutp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Kd = 4 * dinp;
  if (i >= Kd * doutp) return;
  const int o = i / Kd, row = i % Kd;
  const int c = row % dinp, p = row / dinp;
  float v = 0.f;
  if (c < din && o < dout) v = (p == 0) ? bias[(size_t)c * dout + o] : theta[((size_t)(p - 1) * din + c) * dout + o];
  W[i] = v;
}
// W[c, p*doutp + o] = Theta_ext[p*din + c, o]  -- the weight of  Z = F @ [bias | theta_x | theta_y | theta_z]
__global__ void theta_cat_kernel(const float* __restrict__ theta, const float* __restrict__ bias, float* __restrict__ W,
                                 int din, int dout, int dinp, int doutp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Nd = 4 * doutp;
  if (i >= dinp * Nd) return;
  const int c = i / Nd, col = i % Nd;
  const int o = col % doutp, p = col / doutp;
  float v = 0.f;
  if (c < din && o < dout) v = (p == 0) ? bias[(size_t)c * dout + o] : theta[((size_t)(p - 1) * din + c) * dout + o];
  W[i] = v;
}

// out[nbr(n,k), c] += sum_p q_p(n,k) * H[src, p*C + c],  src = n (FlexConv grad) or nbr(n,0) (FlexDeconv)
template <bool SRC_NBR0>
__global__ void __launch_bounds__(256)
flex_scatter_kernel(const float* __restrict__ H, const float* __restrict__ xyz, const int32_t* __restrict__ nbr,
                    float* __restrict__ out, long long rows, int n, int k, int C) {
  const int cv = C >> 2;
  const long long total = rows * cv;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cv;
    const int c = (int)(e - r * cv) << 2;
    const long long b = r / n;
    const float* pbase = xyz + b * n * 3LL;
    const int32_t* nb = nbr + r * k;
    const int g0 = __ldg(nb);
    const float px = __ldg(pbase + g0 * 3LL), py = __ldg(pbase + g0 * 3LL + 1), pz = __ldg(pbase + g0 * 3LL + 2);
    const long long src = SRC_NBR0 ? (b * n + g0) : r;
    const float* h = H + src * 4LL * C + c;
    const float4 h0 = ldg4(h), hx = ldg4(h + C), hy = ldg4(h + 2 * C), hz = ldg4(h + 3 * C);
    float* obase = out + b * n * (long long)C + c;
    for (int kk = 0; kk < k; ++kk) {
      const int g = __ldg(nb + kk);
      const float dx = __ldg(pbase + g * 3LL) - px;
      const float dy = __ldg(pbase + g * 3LL + 1) - py;
      const float dz = __ldg(pbase + g * 3LL + 2) - pz;
      float4 v;
      v.x = fmaf(dz, hz.x, fmaf(dy, hy.x, fmaf(dx, hx.x, h0.x)));
      v.y = fmaf(dz, hz.y, fmaf(dy, hy.y, fmaf(dx, hx.y, h0.y)));
      v.z = fmaf(dz, hz.z, fmaf(dy, hy.z, fmaf(dx, hx.z, h0.z)));
      v.w = fmaf(dz, hz.w, fmaf(dy, hy.w, fmaf(dx, hx.w, h0.w)));
      red_add4(obase + (long long)g * C, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FlexConv backward, point-major
// ---------------------------------------------------------------------------------------------
size_t flex_conv_grad_pm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  (void)K;
  if (B <= 0 || N <= 0 || Din <= 0 || Dout <= 0) return 0;
  const long long rows = (long long)B * N;
  return align_up((size_t)rows * 4 * Din * sizeof(float), 256) + tn_partial_bytes(rows, 4 * Din, Dout) +
         align_up((size_t)4 * Din * Dout * sizeof(float), 256);
}

// Din/Dout are the 4-aligned dims of feat/topdiff/grad_feat; theta/bias/grad_theta/grad_bias have the logical dims.
static int flex_conv_grad_pm_padded(const float* feat, const float* theta, const float* bias, const int32_t* nbr,
                                    const float* xyz, const float* topdiff, float* grad_feat, float* grad_theta,
                                    float* grad_bias, int B, int N, int K, int Din, int Dout, int din_l, int dout_l,
                                    void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!feat || !theta || !bias || !nbr || !xyz || !topdiff || !grad_feat || !grad_theta || !grad_bias)
    return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (Din % 4 || Dout % 4) return DH3D_ERR_UNSUPPORTED;
  const long long rows = (long long)B * N;
  if (rows > 0x7fffffffLL) return DH3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < flex_conv_grad_pm_workspace_bytes(B, N, K, Din, Dout)) return DH3D_ERR_WORKSPACE;
  if ((((uintptr_t)feat | (uintptr_t)topdiff | (uintptr_t)grad_feat | (uintptr_t)ws) & 15) != 0) return DH3D_ERR_ALIGN;
  char* p = reinterpret_cast<char*>(ws);
  float* A = reinterpret_cast<float*>(p);  // moments, later reused for H
  p += align_up((size_t)rows * 4 * Din * sizeof(float), 256);
  float* P = reinterpret_cast<float*>(p);
  p += tn_partial_bytes(rows, 4 * Din, Dout);
  float* Wt = reinterpret_cast<float*>(p);

  flex_moments_c0_kernel<<<ew_grid(rows * (Din / 4), 256), 256, 0, st>>>(feat, xyz, nbr, A, rows, N, K, Din);
  int rc = launch_status();
  if (rc != DH3D_OK) return rc;
  if ((rc = gemm_tn_launch(A, 4 * Din, topdiff, Dout, P, rows, 4 * Din, Dout, st)) != DH3D_OK) return rc;
  gemm_tn_reduce_kernel<<<ceil_div(4 * Din * Dout, 256), 256, 0, st>>>(
      P, tn_splits(rows, 4 * Din, Dout), 4 * Din, Dout, grad_bias, grad_theta, 1, Din, din_l, 0, dout_l);
  if ((rc = launch_status()) != DH3D_OK) return rc;

  theta_ext_t_kernel<<<ceil_div(4 * Din * Dout, 256), 256, 0, st>>>(theta, bias, Wt, din_l, dout_l, Din, Dout);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  if ((rc = linear_simt_launch(topdiff, Dout, Wt, nullptr, nullptr, DH3D_ACT_NONE, A, 4 * Din, (int)rows, Dout,
                               4 * Din, st)) != DH3D_OK)
    return rc;
  cudaError_t e = cudaMemsetAsync(grad_feat, 0, (size_t)rows * Din * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  flex_scatter_kernel<false><<<ew_grid(rows * (Din / 4), 256), 256, 0, st>>>(A, xyz, nbr, grad_feat, rows, N, K, Din);
  return launch_status();
}

int flex_conv_grad_pm(const float* feat, const float* theta, const float* bias, const int32_t* nbr, const float* xyz,
                      const float* topdiff, float* grad_feat, float* grad_theta, float* grad_bias, int B, int N, int K,
                      int Din, int Dout, void* ws, size_t ws_bytes, cudaStream_t st) {
  return flex_conv_grad_pm_padded(feat, theta, bias, nbr, xyz, topdiff, grad_feat, grad_theta, grad_bias, B, N, K, Din,
                                  Dout, Din, Dout, ws, ws_bytes, st);
}

// ---- reference layout: features [B,Din,N], neighborhood [B,K,N], positions [B,3,N], topdiff [B,Dout,N] ----------
struct CmBuffers {
  float* feat_pm; int32_t* nbr_pm; float* xyz_pm; float* top_pm; float* out_pm; char* rest; size_t rest_bytes;
};
static size_t cm_bytes(int B, int N, int K, int dinp, int doutp) {
  return align_up((size_t)B * N * dinp * 4, 256) * 2 + align_up((size_t)B * N * K * 4, 256) +
         align_up((size_t)B * N * 3 * 4, 256) + align_up((size_t)B * N * doutp * 4, 256);
}
static CmBuffers cm_carve(void* ws, size_t ws_bytes, int B, int N, int K, int dinp, int doutp) {
  CmBuffers c;
  char* p = reinterpret_cast<char*>(ws);
  c.feat_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * dinp * 4, 256);
  c.out_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * dinp * 4, 256);
  c.nbr_pm = reinterpret_cast<int32_t*>(p); p += align_up((size_t)B * N * K * 4, 256);
  c.xyz_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * 3 * 4, 256);
  c.top_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * doutp * 4, 256);
  c.rest = p;
  c.rest_bytes = ws_bytes - (size_t)(p - reinterpret_cast<char*>(ws));
  return c;
}
// [B,C,N] -> zero-padded [B,N,Cp]
static int cm_to_pm(const float* src_cm, float* dst_pm, int B, int C, int Cp, int N, cudaStream_t st) {
  if (Cp != C) {
    cudaError_t e = cudaMemsetAsync(dst_pm, 0, (size_t)B * N * Cp * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  return transpose_strided_launch(src_cm, (long long)C * N, N, dst_pm, (long long)N * Cp, Cp, B, C, N, st);
}
static int pm_to_cm(const float* src_pm, float* dst_cm, int B, int C, int Cp, int N, cudaStream_t st) {
  return transpose_strided_launch(src_pm, (long long)N * Cp, Cp, dst_cm, (long long)C * N, N, B, N, C, st);
}

size_t flex_conv_grad_cm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return 0;
  const int dinp = pad4(Din), doutp = pad4(Dout);
  return cm_bytes(B, N, K, dinp, doutp) + flex_conv_grad_pm_workspace_bytes(B, N, K, dinp, doutp);
}
int flex_conv_grad_cm(const float* feat_cm, const float* theta, const float* bias, const int32_t* nbr_cm,
                      const float* pos_cm, const float* topdiff_cm, float* grad_feat_cm, float* grad_theta,
                      float* grad_bias, int B, int N, int K, int Din, int Dout, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  if (!feat_cm || !theta || !bias || !nbr_cm || !pos_cm || !topdiff_cm || !grad_feat_cm || !grad_theta || !grad_bias)
    return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (!ws || ws_bytes < flex_conv_grad_cm_workspace_bytes(B, N, K, Din, Dout)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0) return DH3D_ERR_ALIGN;
  const int dinp = pad4(Din), doutp = pad4(Dout);
  CmBuffers c = cm_carve(ws, ws_bytes, B, N, K, dinp, doutp);
  int rc;
  if ((rc = cm_to_pm(feat_cm, c.feat_pm, B, Din, dinp, N, st)) != DH3D_OK) return rc;
  if ((rc = cm_to_pm(topdiff_cm, c.top_pm, B, Dout, doutp, N, st)) != DH3D_OK) return rc;
  if ((rc = transpose_launch(nbr_cm, c.nbr_pm, B, K, N, st)) != DH3D_OK) return rc;
  if ((rc = transpose_launch(pos_cm, c.xyz_pm, B, 3, N, st)) != DH3D_OK) return rc;
  rc = flex_conv_grad_pm_padded(c.feat_pm, theta, bias, c.nbr_pm, c.xyz_pm, c.top_pm, c.out_pm, grad_theta, grad_bias, B,
                                N, K, dinp, doutp, Din, Dout, c.rest, c.rest_bytes, st);
  if (rc != DH3D_OK) return rc;
  return pm_to_cm(c.out_pm, grad_feat_cm, B, Din, dinp, N, st);
}

// ---------------------------------------------------------------------------------------------
// FlexDeconv forward (flex_convolution_transpose), reference layout
// ---------------------------------------------------------------------------------------------
size_t flex_deconv_cm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return 0;
  const int dinp = pad4(Din), doutp = pad4(Dout);
  // feat_pm [rows,dinp], out_pm [rows,doutp] (held in the top_pm slot), nbr, xyz, Z [rows,4*doutp], Wcat
  return cm_bytes(B, N, K, dinp, doutp) + align_up((size_t)B * N * 4 * doutp * 4, 256) +
         align_up((size_t)dinp * 4 * doutp * 4, 256);
}
int flex_deconv_cm(const float* feat_cm, const float* theta, const float* bias, const int32_t* nbr_cm,
                   const float* pos_cm, float* out_cm, int B, int N, int K, int Din, int Dout, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  if (!feat_cm || !theta || !bias || !nbr_cm || !pos_cm || !out_cm) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (!ws || ws_bytes < flex_deconv_cm_workspace_bytes(B, N, K, Din, Dout)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0) return DH3D_ERR_ALIGN;
  const int dinp = pad4(Din), doutp = pad4(Dout);
  const long long rows = (long long)B * N;
  if (rows > 0x7fffffffLL) return DH3D_ERR_UNSUPPORTED;
  CmBuffers c = cm_carve(ws, ws_bytes, B, N, K, dinp, doutp);
  float* Z = reinterpret_cast<float*>(c.rest);
  float* Wcat = reinterpret_cast<float*>(c.rest + align_up((size_t)rows * 4 * doutp * 4, 256));
  float* out_pm = c.top_pm;  // [rows, doutp]
  int rc;
  if ((rc = cm_to_pm(feat_cm, c.feat_pm, B, Din, dinp, N, st)) != DH3D_OK) return rc;
  if ((rc = transpose_launch(nbr_cm, c.nbr_pm, B, K, N, st)) != DH3D_OK) return rc;
  if ((rc = transpose_launch(pos_cm, c.xyz_pm, B, 3, N, st)) != DH3D_OK) return rc;
  theta_cat_kernel<<<ceil_div(dinp * 4 * doutp, 256), 256, 0, st>>>(theta, bias, Wcat, Din, Dout, dinp, doutp);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  if ((rc = linear_simt_launch(c.feat_pm, dinp, Wcat, nullptr, nullptr, DH3D_ACT_NONE, Z, 4 * doutp, (int)rows, dinp,
                               4 * doutp, st)) != DH3D_OK)
    return rc;
  cudaError_t e = cudaMemsetAsync(out_pm, 0, (size_t)rows * doutp * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  flex_scatter_kernel<true><<<ew_grid(rows * (doutp / 4), 256), 256, 0, st>>>(Z, c.xyz_pm, c.nbr_pm, out_pm, rows, N, K,
                                                                              doutp);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  return pm_to_cm(out_pm, out_cm, B, Dout, doutp, N, st);
}

// ---------------------------------------------------------------------------------------------
// FlexPool backward, reference layout: grad_f[b,d,argmax[b,d,n]] += topdiff[b,d,n]
// ---------------------------------------------------------------------------------------------
__global__ void flex_pool_grad_cm_kernel(const float* __restrict__ topdiff, const int32_t* __restrict__ argmax,
                                         float* __restrict__ grad, long long total, int n) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / n;  // (b, d)
    atomicAdd(grad + row * n + __ldg(argmax + e), __ldg(topdiff + e));
  }
}
int flex_pool_grad_cm(const float* topdiff, const int32_t* argmax, float* grad_feat, int B, int N, int D,
                      cudaStream_t st) {
  if (!topdiff || !argmax || !grad_feat) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || D <= 0) return DH3D_ERR_DIM;
  const long long total = (long long)B * D * N;
  cudaError_t e = cudaMemsetAsync(grad_feat, 0, (size_t)total * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  flex_pool_grad_cm_kernel<<<ew_grid(total, 256), 256, 0, st>>>(topdiff, argmax, grad_feat, total, N);
  return launch_status();
}

// ---------------------------------------------------------------------------------------------
// ConvPointset backward, reference layout (Din is tiny -- 3 in DH3D -- so the kernels are per point)
//   A'[n] = (1, sum_k (f[:,nbr_k] - f[:,nbr_0]))      [grad_bias ; grad_theta] = A'^T @ G
//   t[n,j] = sum_l theta[j,l] g[n,l];  grad_f[j,nbr_k] += t;  grad_f[j,nbr_0] -= t   (per k, like the reference)
// ---------------------------------------------------------------------------------------------
__global__ void cp_prep_kernel(const float* __restrict__ feat_cm, const int32_t* __restrict__ nbr_cm,
                               float* __restrict__ Ap, long long rows, int n, int k, int din, int ka) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const long long b = r / n;
    const int i = (int)(r - b * n);
    const int32_t* nb = nbr_cm + b * (long long)k * n + i;
    const int g0 = __ldg(nb);
    float* a = Ap + r * ka;
    a[0] = 1.f;
    for (int j = 0; j < din; ++j) {
      const float* f = feat_cm + (b * din + j) * (long long)n;
      const float f0 = __ldg(f + g0);
      float s = 0.f;
      for (int kk = 0; kk < k; ++kk) s += __ldg(f + __ldg(nb + (long long)kk * n)) - f0;
      a[1 + j] = s;
    }
    for (int j = 1 + din; j < ka; ++j) a[j] = 0.f;
  }
}
__global__ void cp_scatter_kernel(const float* __restrict__ top_pm, const float* __restrict__ theta,
                                  const int32_t* __restrict__ nbr_cm, float* __restrict__ grad_cm, long long rows, int n,
                                  int k, int din, int dout, int doutp) {
  extern __shared__ float s_theta[];  // din*dout
  for (int i = threadIdx.x; i < din * dout; i += blockDim.x) s_theta[i] = theta[i];
  __syncthreads();
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const long long b = r / n;
    const int i = (int)(r - b * n);
    const int32_t* nb = nbr_cm + b * (long long)k * n + i;
    const int g0 = __ldg(nb);
    const float* g = top_pm + r * doutp;
    for (int j = 0; j < din; ++j) {
      float t = 0.f;
      for (int l = 0; l < dout; ++l) t = fmaf(s_theta[j * dout + l], __ldg(g + l), t);
      float* gf = grad_cm + (b * din + j) * (long long)n;
      for (int kk = 0; kk < k; ++kk) {
        atomicAdd(gf + __ldg(nb + (long long)kk * n), t);
        atomicAdd(gf + g0, -t);
      }
    }
  }
}
size_t conv_pointset_grad_cm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  (void)K;
  if (B <= 0 || N <= 0 || Din <= 0 || Dout <= 0) return 0;
  const long long rows = (long long)B * N;
  const int ka = pad4(1 + Din), doutp = pad4(Dout);
  return align_up((size_t)rows * ka * 4, 256) + align_up((size_t)rows * doutp * 4, 256) + tn_partial_bytes(rows, ka, doutp);
}
int conv_pointset_grad_cm(const float* feat_cm, const float* theta, const int32_t* nbr_cm, const float* topdiff_cm,
                          float* grad_feat_cm, float* grad_theta, float* grad_bias, int B, int N, int K, int Din,
                          int Dout, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!feat_cm || !theta || !nbr_cm || !topdiff_cm || !grad_feat_cm || !grad_theta || !grad_bias) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if ((size_t)Din * Dout * sizeof(float) > 48 * 1024) return DH3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < conv_pointset_grad_cm_workspace_bytes(B, N, K, Din, Dout)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0) return DH3D_ERR_ALIGN;
  const long long rows = (long long)B * N;
  if (rows > 0x7fffffffLL) return DH3D_ERR_UNSUPPORTED;
  const int ka = pad4(1 + Din), doutp = pad4(Dout);
  char* p = reinterpret_cast<char*>(ws);
  float* Ap = reinterpret_cast<float*>(p); p += align_up((size_t)rows * ka * 4, 256);
  float* top_pm = reinterpret_cast<float*>(p); p += align_up((size_t)rows * doutp * 4, 256);
  float* P = reinterpret_cast<float*>(p);
  int rc;
  if ((rc = cm_to_pm(topdiff_cm, top_pm, B, Dout, doutp, N, st)) != DH3D_OK) return rc;
  cp_prep_kernel<<<ew_grid(rows, 256), 256, 0, st>>>(feat_cm, nbr_cm, Ap, rows, N, K, Din, ka);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  if ((rc = gemm_tn_launch(Ap, ka, top_pm, doutp, P, rows, ka, doutp, st)) != DH3D_OK) return rc;
  // row 0 -> grad_bias[Dout]; rows 1..Din -> grad_theta[Din,Dout]
  gemm_tn_reduce_kernel<<<ceil_div(ka * doutp, 256), 256, 0, st>>>(P, tn_splits(rows, ka, doutp), ka, doutp, grad_bias,
                                                                   grad_theta, 0, ka, Din, 1, Dout);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  cudaError_t e = cudaMemsetAsync(grad_feat_cm, 0, (size_t)rows * Din * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  cp_scatter_kernel<<<ew_grid(rows, 256), 256, (size_t)Din * Dout * sizeof(float), st>>>(
      top_pm, theta, nbr_cm, grad_feat_cm, rows, N, K, Din, Dout, doutp);
  return launch_status();
}

// ---------------------------------------------------------------------------------------------
// Row scatter-adds of the PointNet++ ops (point-major, like the reference's tf_ops)
//   group_point_grad:        grad_points[b, idx[b,j,s], :] += grad_out[b,j,s,:]
//   three_interpolate_grad:  grad_points[b, idx[b,i,t], :] += grad_out[b,i,:] * weight[b,i,t]
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void row_scatter_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                   const float* __restrict__ weight, float* __restrict__ dst, long long src_rows,
                                   long long rows_per_batch, int per_src, int n_dst, int c) {
  // one work item = (source row, target slot t < per_src, column group)
  const int cv = c / VEC;
  const long long total = src_rows * per_src * cv;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long rt = e / cv;
    const int col = (int)(e - rt * cv) * VEC;
    const long long r = rt / per_src;
    const long long b = r / rows_per_batch;
    const int target = __ldg(idx + rt);
    const float w = weight ? __ldg(weight + rt) : 1.f;
    float* d = dst + (b * n_dst + target) * (long long)c + col;
    const float* s = src + r * c + col;
    if (VEC == 4) {
      const float4 v = ldg4(s);
      red_add4(d, make_float4(v.x * w, v.y * w, v.z * w, v.w * w));
    } else {
      atomicAdd(d, __ldg(s) * w);
    }
  }
}
static int row_scatter_launch(const float* src, const int32_t* idx, const float* weight, float* dst, int b,
                              long long rows_per_batch, int per_src, int n_dst, int c, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)b * n_dst * c * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const long long src_rows = (long long)b * rows_per_batch;
  const bool v4 = (c % 4 == 0) && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0);
  if (v4)
    row_scatter_kernel<4><<<ew_grid(src_rows * per_src * (c / 4), 256), 256, 0, st>>>(src, idx, weight, dst, src_rows,
                                                                                      rows_per_batch, per_src, n_dst, c);
  else
    row_scatter_kernel<1><<<ew_grid(src_rows * per_src * c, 256), 256, 0, st>>>(src, idx, weight, dst, src_rows,
                                                                                rows_per_batch, per_src, n_dst, c);
  return launch_status();
}
// grad_out [b,m,nsample,c], idx [b,m,nsample] -> grad_points [b,n,c]
int group_point_grad_launch(int b, int n, int c, int m, int nsample, const float* grad_out, const int32_t* idx,
                            float* grad_points, cudaStream_t st) {
  if (!grad_out || !idx || !grad_points) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || c <= 0 || m <= 0 || nsample <= 0) return DH3D_ERR_DIM;
  // every (j,s) pair is its own source row of c floats
  return row_scatter_launch(grad_out, idx, nullptr, grad_points, b, (long long)m * nsample, 1, n, c, st);
}
// grad_out [b,n,c], idx [b,n,3], weight [b,n,3] -> grad_points [b,m,c]
int three_interpolate_grad_launch(int b, int n, int c, int m, const float* grad_out, const int32_t* idx,
                                  const float* weight, float* grad_points, cudaStream_t st) {
  if (!grad_out || !idx || !weight || !grad_points) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || c <= 0 || m <= 0) return DH3D_ERR_DIM;
  return row_scatter_launch(grad_out, idx, weight, grad_points, b, n, 3, m, c, st);
}

}  // namespace dh3d
