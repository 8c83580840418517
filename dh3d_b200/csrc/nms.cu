// Keypoint non-maximum suppression on the detector's attention.
//
// Reference: core/utils.py:15-43 `single_nms` (host numpy + sklearn ball tree), called per cloud by
// evaluate/local_eval/localdesc_extract.py:92-98 with nms_radius 0.5, min_response_ratio 0.01, max_keypoints 512:
//   distances, indices = 50-NN of every point (self first)                                   :17-18
//   attention[distances[:,7] > 2.0] = 0                       (remove_noise)                 :19-22
//   knn_attention = attention[indices]; knn_attention[distances > nms_radius] = 0            :24-26
//   is_max = argmax(knn_attention, axis=1) == 0                                              :27
//   keep m in is_max with attention[m] > max(attention) * min_response_ratio                 :30-32
//   sort (attention, index) descending, first max_keypoints                                  :33-40
//
// Here: the 50-NN lists come from this library's k-NN engine (knn.cu, K = 50, fp32 keys); every THRESHOLD test
// (> 2.0, > nms_radius) recomputes the pair distance in fp64 from the coordinates, which is what sklearn
// compares (it works in float64), so a pair cannot flip sides of a radius because of fp32 rounding.  Selection =
// per-cloud candidate compaction + an O(C^2) rank (C = local maxima, a few thousand): out[rank] = index, which
// is the reference's descending (attention, index) order without a sort.
#include "common.cuh"

namespace dh3d {

size_t knn_workspace_bytes(int B, int N);
int knn_launch(const float* pos, int B, int N, int K, long long sb, int sp, int sd, int32_t* ids, float* dists,
               void* workspace, size_t workspace_bytes, cudaStream_t st);

constexpr int kNmsK = 50;        // n_neighbors of the reference (core/utils.py:17)
constexpr int kNmsNoiseRank = 7; // distances[:, 7]
constexpr double kNmsNoiseDist = 2.0;

__device__ __forceinline__ unsigned ordered_key(float v) {  // monotone float -> unsigned
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_val(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ double pair_dist(const float* __restrict__ p, int i, int j) {
  const double dx = (double)__ldg(p + 3LL * i) - (double)__ldg(p + 3LL * j);
  const double dy = (double)__ldg(p + 3LL * i + 1) - (double)__ldg(p + 3LL * j + 1);
  const double dz = (double)__ldg(p + 3LL * i + 2) - (double)__ldg(p + 3LL * j + 2);
  return sqrt(dx * dx + dy * dy + dz * dz);
}

// att2 = attention with noise points zeroed; per-cloud maximum of att2
__global__ void nms_prepare_kernel(const float* __restrict__ xyz, const float* __restrict__ att,
                                   const int32_t* __restrict__ ids, float* __restrict__ att2,
                                   unsigned* __restrict__ maxkey, int n, int remove_noise) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned key = 0u;
  if (i < n) {
    const float* p = xyz + (long long)b * n * 3;
    float a = __ldg(att + (long long)b * n + i);
    if (remove_noise) {
      const int j = __ldg(ids + ((long long)b * n + i) * kNmsK + kNmsNoiseRank);
      if (j >= 0 && pair_dist(p, i, j) > kNmsNoiseDist) a = 0.f;
    }
    att2[(long long)b * n + i] = a;
    key = ordered_key(a);
  }
  key = __reduce_max_sync(0xffffffffu, key);
  if ((threadIdx.x & 31) == 0 && key) atomicMax(maxkey + b, key);
}

__global__ void nms_select_kernel(const float* __restrict__ xyz, const float* __restrict__ att2,
                                  const int32_t* __restrict__ ids, const unsigned* __restrict__ maxkey,
                                  int32_t* __restrict__ cand, int32_t* __restrict__ ncand, int n, double radius,
                                  double ratio) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + (long long)b * n * 3;
  const float* a2 = att2 + (long long)b * n;
  const int32_t* nb = ids + ((long long)b * n + i) * kNmsK;
  const int first = __ldg(nb);
  const float v0 = (first >= 0 && pair_dist(p, i, first) > radius) ? 0.f : __ldg(a2 + (first >= 0 ? first : i));
  bool is_max = true;
  for (int j = 1; j < kNmsK; ++j) {
    const int id = __ldg(nb + j);
    if (id < 0) break;
    const float a = (pair_dist(p, i, id) > radius) ? 0.f : __ldg(a2 + id);
    if (a > v0) { is_max = false; break; }  // argmax returns the FIRST maximum: slot 0 wins ties
  }
  const double thresh = (double)ordered_val(__ldg(maxkey + b)) * ratio;
  if (is_max && (double)__ldg(a2 + i) > thresh) cand[(long long)b * n + atomicAdd(ncand + b, 1)] = i;
}

// rank of each candidate in descending (attention, index) order; out[rank] = index for rank < max_kp
__global__ void nms_rank_kernel(const float* __restrict__ att2, const int32_t* __restrict__ cand,
                                const int32_t* __restrict__ ncand, int32_t* __restrict__ out,
                                int32_t* __restrict__ out_cnt, int n, int max_kp) {
  __shared__ float s_a[256];
  __shared__ int s_i[256];
  const int b = blockIdx.y;
  const int c = __ldg(ncand + b);
  if (blockIdx.x * blockDim.x >= c && blockIdx.x > 0) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t* cl = cand + (long long)b * n;
  const float* a2 = att2 + (long long)b * n;
  int mine = -1;
  float ma = 0.f;
  if (t < c) { mine = __ldg(cl + t); ma = __ldg(a2 + mine); }
  int rank = 0;
  for (int base = 0; base < c; base += 256) {
    const int u = base + threadIdx.x;
    __syncthreads();
    if (u < c) { s_i[threadIdx.x] = __ldg(cl + u); s_a[threadIdx.x] = __ldg(a2 + s_i[threadIdx.x]); }
    __syncthreads();
    const int lim = min(256, c - base);
    if (mine >= 0)
      for (int e = 0; e < lim; ++e) rank += (s_a[e] > ma) || (s_a[e] == ma && s_i[e] > mine);
  }
  if (mine >= 0 && rank < max_kp) out[(long long)b * max_kp + rank] = mine;
  if (t == 0) out_cnt[b] = min(c, max_kp);
}

size_t keypoint_nms_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  const size_t rows = (size_t)B * N;
  return align_up(knn_workspace_bytes(B, N), 256) + 2 * align_up(rows * kNmsK * 4, 256) + 2 * align_up(rows * 4, 256) +
         2 * align_up((size_t)B * 4, 256);
}

int keypoint_nms_launch(const float* xyz, const float* attention, int B, int N, float nms_radius,
                        float min_response_ratio, int max_keypoints, int remove_noise, int32_t* out_idx,
                        int32_t* out_cnt, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!xyz || !attention || !out_idx || !out_cnt) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || max_keypoints <= 0) return DH3D_ERR_DIM;
  if (N < kNmsK || B > 65535) return DH3D_ERR_UNSUPPORTED;  // sklearn raises for n_neighbors > n_samples
  if (!ws || ws_bytes < keypoint_nms_workspace_bytes(B, N)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)ws & 255) != 0) return DH3D_ERR_ALIGN;
  const size_t rows = (size_t)B * N;
  char* p = reinterpret_cast<char*>(ws);
  void* knn_ws = p; p += align_up(knn_workspace_bytes(B, N), 256);
  int32_t* ids = reinterpret_cast<int32_t*>(p); p += align_up(rows * kNmsK * 4, 256);
  float* dists = reinterpret_cast<float*>(p); p += align_up(rows * kNmsK * 4, 256);
  float* att2 = reinterpret_cast<float*>(p); p += align_up(rows * 4, 256);
  int32_t* cand = reinterpret_cast<int32_t*>(p); p += align_up(rows * 4, 256);
  unsigned* maxkey = reinterpret_cast<unsigned*>(p); p += align_up((size_t)B * 4, 256);
  int32_t* ncand = reinterpret_cast<int32_t*>(p);

  int rc = knn_launch(xyz, B, N, kNmsK, 3LL * N, 3, 1, ids, dists, knn_ws, knn_workspace_bytes(B, N), st);
  if (rc != DH3D_OK) return rc;
  cudaError_t e = cudaMemsetAsync(maxkey, 0, align_up((size_t)B * 4, 256) * 2, st);  // maxkey + ncand
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(out_idx, 0xff, (size_t)B * max_keypoints * sizeof(int32_t), st);  // -1 padding
  if (e != cudaSuccess) return (int)e;
  dim3 grid(ceil_div(N, 256), B);
  nms_prepare_kernel<<<grid, 256, 0, st>>>(xyz, attention, ids, att2, maxkey, N, remove_noise);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  nms_select_kernel<<<grid, 256, 0, st>>>(xyz, att2, ids, maxkey, cand, ncand, N, (double)nms_radius,
                                          (double)min_response_ratio);
  if ((rc = launch_status()) != DH3D_OK) return rc;
  nms_rank_kernel<<<grid, 256, 0, st>>>(att2, cand, ncand, out_idx, out_cnt, N, max_keypoints);
  return launch_status();
}

// ---------------------------------------------------------------------------------------------
// The rest of the reference's --perform_nms output mode (evaluate/local_eval/localdesc_extract.py:92-104):
//   response = 1 - attention                       (:95)       -> affine_kernel
//   xyzfeatatt_nms = res[max_indices, :]           (:99)       -> gather_rows_kernel (idx < 0 = padding -> zero row)
// ---------------------------------------------------------------------------------------------
__global__ void affine_kernel(const float* __restrict__ x, float a, float b, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], b);
}

int affine_launch(const float* x, float a, float b, float* y, size_t n, cudaStream_t st) {
  if (!x || !y) return DH3D_ERR_NULL;
  if (n == 0) return DH3D_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
  affine_kernel<<<(int)blocks, 256, 0, st>>>(x, a, b, y, n);
  return launch_status();
}

// out[b, j, 0:c] = idx[b, j] >= 0 ? src[b, idx[b, j], 0:c] : 0     (out rows ldo floats apart)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                   float* __restrict__ out, long long rows, int m, int n, int c, int ldo) {
  const long long total = rows * c;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / c;
    const int col = (int)(e - r * c);
    const long long b = r / m;
    const int ii = __ldg(idx + r);
    out[r * ldo + col] = (ii >= 0 && ii < n) ? __ldg(src + ((long long)b * n + ii) * c + col) : 0.f;
  }
}

int gather_rows_launch(int b, int n, int c, int m, const float* src, const int32_t* idx, float* out, int ldo,
                       cudaStream_t st) {
  if (!src || !idx || !out) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || c <= 0 || m <= 0 || ldo < c) return DH3D_ERR_DIM;
  const long long rows = (long long)b * m;
  long long blocks = (rows * c + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  gather_rows_kernel<<<(int)blocks, 256, 0, st>>>(src, idx, out, rows, m, n, c, ldo);
  return launch_status();
}

}  // namespace dh3d
