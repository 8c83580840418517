// FlexConv forward (reference: user_ops/kernels/flex_conv_kernel_gpu.cu.cc:44-158).
//
// The reference evaluates  out[n,o] = sum_k sum_c (bias[c,o] + theta[:,c,o].(p[nbr_k]-p[n])) * f[c,nbr_k]
// with 9 flops per (n,k,c,o) and 4-byte scattered loads from a channel-major feature map.
// The weight is affine in the offset, so the sum factors exactly into
//     A[n, p'*Din + c] = sum_k dext[n,k,p'] * f[nbr(n,k), c],   dext = (1, dx, dy, dz)
//     out[n, :]        = A[n, :] @ Theta_ext,   Theta_ext = [bias ; theta_x ; theta_y ; theta_z]  (4*Din x Dout)
// i.e. a neighbour gather-reduce into 4*Din moments (HBM/L2-bandwidth bound: each neighbour is
// one contiguous Din-float row in point-major layout, read with 16-byte loads) followed by ONE
// dense GEMM in which K no longer appears (flops 9*N*K*Din*Dout -> 8*N*K*Din + 8*N*Din*Dout).
// The GEMM + feature-bias/BatchNorm/ReLU epilogue of the two-kernel form is gemm_simt.cu.
#include <stdlib.h>

#include "common.cuh"

namespace dh3d {

int linear_launch(const float* x, int ldx, const float* w, const float* scale, const float* shift,
                  int act, float* y, int ldy, int M, int K, int N, cudaStream_t st);
bool exact_fp32();   // capi.cu: DH3D_EXACT_FP32=1 -> the two-kernel form with the FFMA GEMM
// the fused kernel (flexconv_ca.cu) needs whole 32-channel groups and 16-byte output rows
static bool flexconv_fused_supported(int Din, int Dout) { return Din % 32 == 0 && Dout % 4 == 0; }
static bool flexconv_use_fused(int Din, int Dout) { return !exact_fp32() && flexconv_fused_supported(Din, Dout); }
int flexconv_ca_launch(const float* feat, const float* xyz, const int32_t* nbr, const void* theta_packed,
                       const float* scale, const float* shift, int act, float* out, int rows, int n_per_cloud,
                       int K, int Din, int Dout, cudaStream_t st);
int transpose_launch(const void* src, void* dst, int B, int R, int C, cudaStream_t st);
int transpose_strided_launch(const void* src, long long sbs, int lds, void* dst, long long sbd,
                             int ldd, int B, int R, int C, cudaStream_t st);

// thread = (point row r, 4-channel group); lanes of a warp cover consecutive channel groups of
// the same / adjacent rows, so neighbour-row reads are fully coalesced 16-byte accesses.
__global__ void __launch_bounds__(256)
flexconv_moments_kernel(const float* __restrict__ feat, const float* __restrict__ xyz,
                        const int32_t* __restrict__ nbr, float* __restrict__ A, long long rows,
                        int n, int k, int din) {
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
    const float px = __ldg(xyz + r * 3), py = __ldg(xyz + r * 3 + 1), pz = __ldg(xyz + r * 3 + 2);
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

static size_t moments_bytes(int B, int N, int Din) {
  return align_up((size_t)B * N * 4 * Din * sizeof(float), 256);
}
// Theta_ext in BOTH operand forms, back to back: plain [4*Din, Dout] fp32 (FFMA GEMM of the two-kernel form), then
// {hi^T, lo^T} [Dout, 4*Din] tf32 pairs (fused tcgen05 kernel).  A prepacked buffer therefore never depends on
// which path the process runs.
static size_t theta_plane_bytes(int Din, int Dout) { return align_up((size_t)4 * Din * Dout * sizeof(float), 256); }
static size_t theta_ext_bytes(int Din, int Dout) { return 3 * theta_plane_bytes(Din, Dout); }

size_t flex_conv_pm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  (void)K;
  if (B <= 0 || N <= 0 || Din <= 0 || Dout <= 0) return 0;
  return moments_bytes(B, N, Din) + theta_ext_bytes(Din, Dout);
}

// Theta_ext[p*DinP + c][o] = (p == 0 ? bias[c][o] : theta[p-1][c][o]), zero in the padding
__global__ void theta_ext_kernel(const float* __restrict__ theta, const float* __restrict__ bias,
                                 float* __restrict__ ext, int din, int dout, int dinp, int doutp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * dinp * doutp) return;
  const int o = i % doutp;
  const int row = i / doutp;
  const int c = row % dinp, p = row / dinp;
  float v = 0.f;
  if (c < din && o < dout)
    v = (p == 0) ? bias[(size_t)c * dout + o] : theta[((size_t)(p - 1) * din + c) * dout + o];
  ext[i] = v;
}

// Same matrix in the tensor-core GEMM's packed form: {hi^T, lo^T} with Theta_ext^T [Dout, 4*Din]
__global__ void theta_ext_packed_kernel(const float* __restrict__ theta, const float* __restrict__ bias,
                                        float* __restrict__ hi, float* __restrict__ lo, int din, int dout,
                                        int dinp, int doutp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Kd = 4 * dinp;
  if (i >= Kd * doutp) return;
  const int o = i / Kd, row = i % Kd;
  const int c = row % dinp, p = row / dinp;
  float v = 0.f;
  if (c < din && o < dout)
    v = (p == 0) ? bias[(size_t)c * dout + o] : theta[((size_t)(p - 1) * din + c) * dout + o];
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  const float h = __uint_as_float(b & 0xFFFFE000u);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v - h));
  hi[i] = h;
  lo[i] = __uint_as_float(b & 0xFFFFE000u);
}

// y = act((x + fb) * scale + shift) == act(x*scale + (fb*scale + shift)): fold fb into the shift.
__global__ void fold_bias_kernel(const float* __restrict__ fb, const float* __restrict__ scale,
                                 const float* __restrict__ shift, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = scale ? scale[i] : 1.f;
    out[i] = fmaf(fb ? fb[i] : 0.f, s, shift ? shift[i] : 0.f);
  }
}

size_t flex_conv_pm_total_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  size_t base = flex_conv_pm_workspace_bytes(B, N, K, Din, Dout);
  return base ? base + align_up((size_t)Dout * sizeof(float), 256) : 0;
}

// ---- weight-only preparation (Theta_ext in the contraction's operand form + the folded shift) ----------
// packed = [Theta_ext plain | Theta_ext^T hi | Theta_ext^T lo | fshift[Dout]].  Depends on the
// layer's weights only, so an inference caller prepares it once (dh3d_flex_conv_prepack) instead of
// re-deriving it in every forward (two extra launches per layer, ~6 us each with their gaps).
size_t flex_conv_prepack_bytes(int Din, int Dout) {
  if (Din <= 0 || Dout <= 0) return 0;
  return theta_ext_bytes(Din, Dout) + align_up((size_t)Dout * sizeof(float), 256);
}

static int flex_conv_prepack_padded(const float* theta, const float* bias, const float* feature_bias,
                                    const float* scale, const float* shift, int Din, int Dout,
                                    int din_logical, int dout_logical, void* packed, cudaStream_t st) {
  float* theta_ext = reinterpret_cast<float*>(packed);
  float* fshift = reinterpret_cast<float*>(reinterpret_cast<char*>(packed) + theta_ext_bytes(Din, Dout));
  // Theta_ext = [bias ; theta_x ; theta_y ; theta_z]  (logical dims may be smaller than the padded ones)
  float* hi = reinterpret_cast<float*>(reinterpret_cast<char*>(theta_ext) + theta_plane_bytes(Din, Dout));
  float* lo = reinterpret_cast<float*>(reinterpret_cast<char*>(theta_ext) + 2 * theta_plane_bytes(Din, Dout));
  theta_ext_kernel<<<ceil_div(4 * Din * Dout, 256), 256, 0, st>>>(theta, bias, theta_ext, din_logical,
                                                                 dout_logical, Din, Dout);
  theta_ext_packed_kernel<<<ceil_div(4 * Din * Dout, 256), 256, 0, st>>>(
      theta, bias, hi, lo, din_logical, dout_logical, Din, Dout);
  fold_bias_kernel<<<ceil_div(Dout, 128), 128, 0, st>>>(feature_bias, scale, shift, fshift, Dout);
  return launch_status();
}

int flex_conv_prepack(const float* theta, const float* bias, const float* feature_bias, const float* scale,
                      const float* shift, int Din, int Dout, void* packed, cudaStream_t st) {
  if (!theta || !bias || !packed) return DH3D_ERR_NULL;
  if (Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (Din % 4 || Dout % 4) return DH3D_ERR_UNSUPPORTED;
  if (((uintptr_t)packed & 255) != 0) return DH3D_ERR_ALIGN;
  return flex_conv_prepack_padded(theta, bias, feature_bias, scale, shift, Din, Dout, Din, Dout, packed, st);
}

static int flex_conv_run(const float* feat, const float* theta_ext, const float* eff_shift, const int32_t* nbr,
                         const float* xyz, float* out, int B, int N, int K, int Din, int Dout,
                         const float* scale, int act, float* A, cudaStream_t st);

size_t flex_conv_pm_packed_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  (void)K;
  if (B <= 0 || N <= 0 || Din <= 0 || Dout <= 0) return 0;
  // only the two-kernel form (DH3D_EXACT_FP32 / dims the fused kernel does not cover) needs the moment matrix
  if (flexconv_use_fused(Din, Dout)) return 256;
  return moments_bytes(B, N, Din);
}

// scale must be the one given to flex_conv_prepack (the shift is already inside `packed`)
int flex_conv_pm_packed(const float* feat, const void* packed, const int32_t* nbr, const float* xyz, float* out,
                        int B, int N, int K, int Din, int Dout, const float* scale, int act, void* ws,
                        size_t ws_bytes, cudaStream_t st) {
  if (!feat || !packed || !nbr || !xyz || !out) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (Din % 4 || Dout % 4 || K > 64) return DH3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < flex_conv_pm_packed_workspace_bytes(B, N, K, Din, Dout)) return DH3D_ERR_WORKSPACE;
  if ((((uintptr_t)feat | (uintptr_t)out | (uintptr_t)ws) & 15) != 0 || ((uintptr_t)packed & 255) != 0)
    return DH3D_ERR_ALIGN;
  const float* theta_ext = reinterpret_cast<const float*>(packed);
  const float* fshift =
      reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) + theta_ext_bytes(Din, Dout));
  return flex_conv_run(feat, theta_ext, fshift, nbr, xyz, out, B, N, K, Din, Dout, scale, act,
                       reinterpret_cast<float*>(ws), st);
}

// Din/Dout are the (4-aligned) dims of feat/out; theta/bias have the logical dims
// din_logical x dout_logical (equal to Din/Dout except for the padded channel-major entry).
int flex_conv_pm_padded(const float* feat, const float* theta, const float* bias,
                        const int32_t* nbr, const float* xyz, float* out, int B, int N, int K, int Din,
                        int Dout, int din_logical, int dout_logical, const float* feature_bias,
                        const float* scale, const float* shift, int act, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  if (!feat || !theta || !bias || !nbr || !xyz || !out) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (Din % 4 || Dout % 4 || K > 64) return DH3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < flex_conv_pm_total_workspace_bytes(B, N, K, Din, Dout))
    return DH3D_ERR_WORKSPACE;
  if ((((uintptr_t)feat | (uintptr_t)out | (uintptr_t)ws) & 15) != 0) return DH3D_ERR_ALIGN;

  char* p = reinterpret_cast<char*>(ws);
  float* A = reinterpret_cast<float*>(p);
  p += moments_bytes(B, N, Din);
  void* packed = p;   // [Theta_ext | fshift], the layout flex_conv_prepack writes
  int rc = flex_conv_prepack_padded(theta, bias, feature_bias, scale, shift, Din, Dout, din_logical,
                                    dout_logical, packed, st);
  if (rc != DH3D_OK) return rc;
  const float* theta_ext = reinterpret_cast<const float*>(packed);
  const float* fshift = reinterpret_cast<const float*>(p + theta_ext_bytes(Din, Dout));
  return flex_conv_run(feat, theta_ext, fshift, nbr, xyz, out, B, N, K, Din, Dout, scale, act, A, st);
}

static int flex_conv_run(const float* feat, const float* theta_ext_c, const float* eff_shift, const int32_t* nbr,
                         const float* xyz, float* out, int B, int N, int K, int Din, int Dout,
                         const float* scale, int act, float* A, cudaStream_t st) {
  float* theta_ext = const_cast<float*>(theta_ext_c);   // the launchers take non-const operand pointers
  const long long rows = (long long)B * N;
  if (rows > 0x7fffffffLL) return DH3D_ERR_UNSUPPORTED;
  if (flexconv_use_fused(Din, Dout)) {
    // cp.async-staged gather + tcgen05 contraction in one kernel; staging A/B (TMA tile::gather4 1.46x slower,
    // per-thread LDG 1.31x slower at 64->64 x 262144 points): profiles/flexconv_staging_ab_r2f.json
    const float* hi = reinterpret_cast<const float*>(reinterpret_cast<const char*>(theta_ext) +
                                                     theta_plane_bytes(Din, Dout));
    return flexconv_ca_launch(feat, xyz, nbr, hi, scale, eff_shift, act, out, (int)rows, N, K, Din, Dout, st);
  }
  long long blocks = (rows * (Din / 4) + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  flexconv_moments_kernel<<<(int)blocks, 256, 0, st>>>(feat, xyz, nbr, A, rows, N, K, Din);
  int rc = launch_status();
  if (rc != DH3D_OK) return rc;
  return linear_launch(A, 4 * Din, theta_ext, scale, eff_shift, act, out, Dout, (int)rows, 4 * Din, Dout, st);
}

int flex_conv_pm(const float* feat, const float* theta, const float* bias, const int32_t* nbr,
                 const float* xyz, float* out, int B, int N, int K, int Din, int Dout,
                 const float* feature_bias, const float* scale, const float* shift, int act,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  return flex_conv_pm_padded(feat, theta, bias, nbr, xyz, out, B, N, K, Din, Dout, Din, Dout,
                             feature_bias, scale, shift, act, ws, ws_bytes, st);
}

// ---- reference-layout (channel-major) entry: transposes in, native kernel, transpose out -----
static inline int pad4(int v) { return (v + 3) & ~3; }

size_t flex_conv_cm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return 0;
  const int dinp = pad4(Din), doutp = pad4(Dout);
  return align_up((size_t)B * N * dinp * 4, 256) + align_up((size_t)B * N * K * 4, 256) +
         align_up((size_t)B * N * 3 * 4, 256) + align_up((size_t)B * N * doutp * 4, 256) +
         flex_conv_pm_total_workspace_bytes(B, N, K, dinp, doutp);
}

// Any Din/Dout (the reference's own fixture uses Din=2, Dout=6): channels are zero-padded to a
// multiple of 4 inside the workspace.
int flex_conv_cm(const float* feat_cm, const float* theta, const float* bias, const int32_t* nbr_cm,
                 const float* pos_cm, float* out_cm, int B, int N, int K, int Din, int Dout, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  if (!feat_cm || !theta || !bias || !nbr_cm || !pos_cm || !out_cm) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0 || Din <= 0 || Dout <= 0) return DH3D_ERR_DIM;
  if (!ws || ws_bytes < flex_conv_cm_workspace_bytes(B, N, K, Din, Dout)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0) return DH3D_ERR_ALIGN;
  const int dinp = pad4(Din), doutp = pad4(Dout);
  char* p = reinterpret_cast<char*>(ws);
  float* feat_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * dinp * 4, 256);
  int32_t* nbr_pm = reinterpret_cast<int32_t*>(p); p += align_up((size_t)B * N * K * 4, 256);
  float* xyz_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * 3 * 4, 256);
  float* out_pm = reinterpret_cast<float*>(p); p += align_up((size_t)B * N * doutp * 4, 256);
  int rc;
  if (dinp != Din) {
    cudaError_t e = cudaMemsetAsync(feat_pm, 0, (size_t)B * N * dinp * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  if ((rc = transpose_strided_launch(feat_cm, (long long)Din * N, N, feat_pm, (long long)N * dinp,
                                     dinp, B, Din, N, st)) != DH3D_OK) return rc;
  if ((rc = transpose_launch(nbr_cm, nbr_pm, B, K, N, st)) != DH3D_OK) return rc;
  if ((rc = transpose_launch(pos_cm, xyz_pm, B, 3, N, st)) != DH3D_OK) return rc;
  rc = flex_conv_pm_padded(feat_pm, theta, bias, nbr_pm, xyz_pm, out_pm, B, N, K, dinp, doutp, Din,
                           Dout, nullptr, nullptr, nullptr, DH3D_ACT_NONE, p,
                           ws_bytes - (size_t)(p - reinterpret_cast<char*>(ws)), st);
  if (rc != DH3D_OK) return rc;
  return transpose_strided_launch(out_pm, (long long)N * doutp, doutp, out_cm, (long long)Dout * N, N,
                                  B, N, Dout, st);
}

}  // namespace dh3d
