// Farthest point sampling (reference: tf_ops/sampling/tf_sampling_g.cu:105-170, <<<32,512>>>).
//
// The reference runs at most 32 CTAs, keeps the running min-distance array `temp` in GLOBAL
// memory and pays 2 x 9 __syncthreads per round for a shared-memory tree argmax.  Here one
// 1024-thread CTA owns one cloud (all B clouds run concurrently), the points and their running
// min-distance live in REGISTERS (8 points/thread at n = 8192), the argmax is one REDUX.MAX +
// ballot per warp and one barrier per round (double-buffered per-warp slots).
//
// Bit-exact parity with the reference's selection order.  The reference picks, among the points
// with maximal d2, the one with the smallest (k mod 512, k): thread tid scans k=tid,tid+512,..
// with a strict `>` (:130-149) and the tree keeps the lower slot on ties (:158).  We give thread
// t the points whose tie rank  tk(k) = (k mod 512)*V + (k div 512),  V = ceil(n/512),  lies in
// [t*PPT, (t+1)*PPT), scan them in increasing tk with a strict `>`, and break cross-thread ties
// towards the lowest thread -- the same total order.  d = fma(dz,dz,fma(dx,dx,dy*dy)) is the
// contraction nvcc applies to :142; d2 = fminf(d, temp).
#include "common.cuh"

namespace dh3d {

constexpr int kFpsThreads = 1024;
constexpr int kFpsWarps = kFpsThreads / 32;

// argmax over the block of (value, index); ties -> lowest thread.  One barrier.
__device__ __forceinline__ int fps_block_argmax(float best, int besti, int* s_val, int* s_idx,
                                                int lane, int warp) {
  // d2 >= 0 for every valid candidate and -1.0f marks "no candidate": non-negative floats order
  // like their bit patterns as signed ints and -1.0f's pattern is a negative int.
  const int bits = __float_as_int(best);
  const int wmax = __reduce_max_sync(0xffffffffu, bits);
  const unsigned m = __ballot_sync(0xffffffffu, bits == wmax);
  const int wi = __shfl_sync(0xffffffffu, besti, __ffs(m) - 1);
  if (lane == 0) { s_val[warp] = wmax; s_idx[warp] = wi; }
  __syncthreads();
  const int v = s_val[lane];
  const int i = s_idx[lane];
  const int bmax = __reduce_max_sync(0xffffffffu, v);
  const unsigned m2 = __ballot_sync(0xffffffffu, v == bmax);
  return __shfl_sync(0xffffffffu, i, __ffs(m2) - 1);
}

// Register-resident variant: n <= 1024 * PPT, xyz cached in shared memory for the centre lookup.
template <int PPT>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_reg_kernel(int n, int m, const float* __restrict__ dataset, int32_t* __restrict__ idxs) {
  extern __shared__ __align__(16) float s_xyz[];  // n*3 floats
  __shared__ int s_val[2][kFpsWarps];
  __shared__ int s_idx[2][kFpsWarps];

  const int b = blockIdx.x;
  const float* ds = dataset + (long long)b * n * 3;
  int32_t* out = idxs + (long long)b * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int j = tid; j < n * 3; j += kFpsThreads) s_xyz[j] = ds[j];
  __syncthreads();

  const int V = (n + 511) >> 9;
  const unsigned magic = 0xffffffffu / (unsigned)V + 1u;  // tk / V == umulhi(tk, magic), tk < 2^16
  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int tk = tid * PPT + i;
    const int k = (tk % V) * 512 + tk / V;
    const bool valid = (tk < 512 * V) && (k < n);
    px[i] = valid ? s_xyz[k * 3 + 0] : 0.f;
    py[i] = valid ? s_xyz[k * 3 + 1] : 0.f;
    pz[i] = valid ? s_xyz[k * 3 + 2] : 0.f;
    td[i] = valid ? 1e38f : -1.f;  // -1 pins d2 = min(d,-1) = -1, which never beats best = -1
  }

  int old = 0;
  if (tid == 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
    float best = -1.f;
    int bslot = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float dx = px[i] - x1, dy = py[i] - y1, dz = pz[i] - z1;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      const float d2 = fminf(d, td[i]);
      td[i] = d2;
      if (d2 > best) { best = d2; bslot = i; }
    }
    const unsigned tk = (unsigned)fps_block_argmax(best, tid * PPT + bslot, s_val[j & 1],
                                                   s_idx[j & 1], lane, warp);
    const unsigned q = (V == 1) ? tk : __umulhi(tk, magic);
    old = (int)((tk - q * (unsigned)V) * 512u + q);
    if (tid == 0) out[j] = old;
  }
}

// Large-n variant (n > 8192): running min-distance in shared memory, xyz through L1/L2.
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_smem_kernel(int n, int m, const float* __restrict__ dataset, int32_t* __restrict__ idxs) {
  extern __shared__ __align__(16) float s_td[];  // 512*V floats, indexed by tie rank
  __shared__ int s_val[2][kFpsWarps];
  __shared__ int s_idx[2][kFpsWarps];

  const int b = blockIdx.x;
  const float* ds = dataset + (long long)b * n * 3;
  int32_t* out = idxs + (long long)b * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = (n + 511) >> 9;
  const int total = 512 * V;
  const int ppt = (total + kFpsThreads - 1) / kFpsThreads;

  for (int j = tid; j < total; j += kFpsThreads) s_td[j] = 1e38f;
  __syncthreads();

  int old = 0;
  if (tid == 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = __ldg(ds + old * 3 + 0), y1 = __ldg(ds + old * 3 + 1),
                z1 = __ldg(ds + old * 3 + 2);
    float best = -1.f;
    int besti = 0;
    for (int i = 0; i < ppt; ++i) {
      const int tk = tid * ppt + i;
      if (tk >= total) break;
      const int k = (tk % V) * 512 + tk / V;
      if (k >= n) continue;
      const float dx = __ldg(ds + k * 3 + 0) - x1, dy = __ldg(ds + k * 3 + 1) - y1,
                  dz = __ldg(ds + k * 3 + 2) - z1;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      const float d2 = fminf(d, s_td[tk]);
      s_td[tk] = d2;
      if (d2 > best) { best = d2; besti = k; }
    }
    old = fps_block_argmax(best, besti, s_val[j & 1], s_idx[j & 1], lane, warp);
    if (tid == 0) out[j] = old;
  }
}

template <int PPT>
static int fps_launch_reg(int b, int n, int m, const float* inp, int32_t* out, cudaStream_t st) {
  size_t smem = (size_t)n * 3 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fps_reg_kernel<PPT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  fps_reg_kernel<PPT><<<b, kFpsThreads, smem, st>>>(n, m, inp, out);
  return launch_status();
}

int fps_launch(int b, int n, int m, const float* inp, int32_t* out, cudaStream_t st) {
  if (!inp || !out) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || m < 0) return DH3D_ERR_DIM;
  if (m == 0) return DH3D_OK;  // reference kernel returns immediately (tf_sampling_g.cu:106-107)
  if (n > 65536) return DH3D_ERR_UNSUPPORTED;
  const int total = 512 * ((n + 511) / 512);
  const int ppt = ceil_div(total, kFpsThreads);
  if (ppt <= 1) return fps_launch_reg<1>(b, n, m, inp, out, st);
  if (ppt <= 2) return fps_launch_reg<2>(b, n, m, inp, out, st);
  if (ppt <= 4) return fps_launch_reg<4>(b, n, m, inp, out, st);
  if (ppt <= 8) return fps_launch_reg<8>(b, n, m, inp, out, st);
  size_t smem = (size_t)total * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  fps_smem_kernel<<<b, kFpsThreads, smem, st>>>(n, m, inp, out);
  return launch_status();
}

}  // namespace dh3d
