// k-NN inside each cloud, results identical to the exhaustive search
// (reference: user_ops/kernels/knn_bruteforce_kernel_gpu.cu.cc).
//
// The reference launches one CTA per QUERY point and radix-sorts all N keys to extract K of them
// (N <= 8192, followed by a device sync).  Here:
//   1. knn_sort_kernel (one CTA per cloud): counting sort of the points by the Morton code of a
//      16^3 grid cell -> spatially coherent float4 array (x,y,z,original index) + one bounding box
//      per 64-point chunk.
//   2. knn_query_kernel: one thread per query (queries taken in sorted order, so a warp's 32 queries
//      are neighbours) with a sorted top-K list in registers.  Each warp walks the chunks outward from
//      its own; a chunk is skipped when, for every lane, a lower bound of the distance to the chunk's
//      box already exceeds that lane's current K-th distance.  Surviving chunks are read with
//      warp-uniform 16-byte loads (the cloud is L1-resident) at 3 FADD + FMUL + 2 FFMA + compare per pair.
//
// Output order is the reference's:
//   key   = sqrt.rn(fma(dz,dz,fma(dy,dy,dx*dx)))   (kernel_gpu.cu.cc:102-107, nvcc-contracted)
//   ties  = cub::BlockRadixSort blocked-order stability: point x sits at rank
//           s(x) = (x mod T)*V + (x div T)   (T,V picked by N, :181-216); plain index above 8192.
// Candidates are not visited in rank order, so the list is ordered by the explicit pair (key, rank).
// The pair compare runs on d^2 against a conservative bound (every d^2 whose sqrt.rn could be <= the
// current K-th key passes); the exact compare happens only on that rare path.  The box lower bound
// uses the same mul/fma sequence as the distance, so by monotonicity of each rounded operation it never
// exceeds the computed d^2 of any point inside the box: skipping is exact.
// DH3D_KNN=tiled selects the exhaustive shared-memory-tiled scan of knn_tiled.cu instead.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace dh3d {

size_t knn_tiled_workspace_bytes(int B, int N);
int knn_tiled_launch(const float* pos, int B, int N, int K, long long sb, int sp, int sd, int32_t* ids,
                     float* dists, void* workspace, size_t workspace_bytes, cudaStream_t st);

constexpr int kKnnChunk = 64;       // candidates per bounding box
constexpr int kKnnThreads = 128;    // queries per CTA
constexpr int kSortThreads = 1024;
constexpr int kCells = 4096;        // 16^3 Morton-ordered grid cells

struct KnnOrder { int T; int logT; int logV; };

static KnnOrder knn_order(int N) {
  int T, V;
  if (N <= 32) { T = 32; V = 1; }
  else if (N <= 64) { T = 64; V = 1; }
  else if (N <= 128) { T = 128; V = 1; }
  else if (N <= 256) { T = 128; V = 2; }
  else if (N <= 512) { T = 128; V = 4; }
  else if (N <= 1024) { T = 256; V = 4; }
  else if (N <= 2048) { T = 256; V = 8; }
  else if (N <= 4096) { T = 512; V = 8; }
  else if (N <= 8192) { T = 1024; V = 8; }
  else { T = 1 << 30; V = 1; }  // no reference order above 8192: rank = index
  KnnOrder o;
  o.T = T;
  o.logT = 0;
  while ((1 << o.logT) < T) ++o.logT;
  o.logV = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : 3;
  return o;
}

__device__ __forceinline__ int knn_rank_of(int x, int T, int logT, int logV) {
  return ((x & (T - 1)) << logV) + (x >> logT);
}
__device__ __forceinline__ int knn_point_of(int r, int T, int logV) {
  return (r & ((1 << logV) - 1)) * T + (r >> logV);
}

static inline int knn_padded(int N) { return ceil_div(N, kKnnChunk) * kKnnChunk; }

static bool knn_use_tiled() {
  static const bool t = [] {
    const char* e = getenv("DH3D_KNN");
    return e && (e[0] == 't' || e[0] == 'T');
  }();
  return t;
}

size_t knn_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  const size_t np = knn_padded(N);
  const size_t pruned = (size_t)B * (np * sizeof(float4) + (np / kKnnChunk) * 2 * sizeof(float4));
  const size_t tiled = knn_tiled_workspace_bytes(B, N);
  return pruned > tiled ? pruned : tiled;
}

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 4 bits -> every third bit
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

// ---------------------------------------------------------------------------------------------
// sort: positions (any layout via strides) -> Morton-cell-sorted float4 (x,y,z,idx) + chunk boxes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads, 1)
knn_sort_kernel(const float* __restrict__ pos, int N, int Np, long long sb, int sp, int sd,
                float4* __restrict__ sorted, float4* __restrict__ boxes) {
  __shared__ int s_hist[kCells];
  __shared__ float s_red[6][kSortThreads / 32];
  __shared__ float s_box[6];
  __shared__ int s_wsum[kSortThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = pos + (long long)b * sb;
  float4* out = sorted + (long long)b * Np;
  float4* bx = boxes + (long long)b * (Np / kKnnChunk) * 2;

  // (1) cloud bounding box
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
  float mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int i = tid; i < N; i += kSortThreads) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = P[(long long)i * sp + d * sd];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if (lane == 0) { s_red[d][warp] = mn[d]; s_red[3 + d][warp] = mx[d]; }
  }
  for (int i = tid; i < kCells; i += kSortThreads) s_hist[i] = 0;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float a = s_red[d][lane], c = s_red[3 + d][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
        c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));
      }
      if (lane == 0) { s_box[d] = a; s_box[3 + d] = c; }
    }
  }
  __syncthreads();
  float lo[3], inv[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = s_box[d];
    const float ext = s_box[3 + d] - s_box[d];
    inv[d] = (ext > 0.f && ext < CUDART_INF_F) ? 16.f / ext : 0.f;
  }
  auto cell_of = [&](int i) -> int {
    unsigned c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = (P[(long long)i * sp + d * sd] - lo[d]) * inv[d];
      int q = (v >= 0.f) ? ((v < 15.f) ? (int)v : 15) : 0;  // NaN -> 0; only the scan order depends on it
      c[d] = (unsigned)q;
    }
    return (int)(spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2));
  };

  // (2) histogram, (3) exclusive scan, (4) scatter
  for (int i = tid; i < N; i += kSortThreads) atomicAdd(&s_hist[cell_of(i)], 1);
  __syncthreads();
  {
    constexpr int PER = kCells / kSortThreads;
    int v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { v[j] = s_hist[tid * PER + j]; sum += v[j]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_wsum[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      s_wsum[lane] = wi - w;
    }
    __syncthreads();
    int run = s_wsum[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) { s_hist[tid * PER + j] = run; run += v[j]; }
  }
  __syncthreads();
  for (int i = tid; i < N; i += kSortThreads) {
    const int p = atomicAdd(&s_hist[cell_of(i)], 1);
    out[p] = make_float4(P[(long long)i * sp], P[(long long)i * sp + sd], P[(long long)i * sp + 2 * sd],
                         __int_as_float(i));
  }
  for (int i = N + tid; i < Np; i += kSortThreads)
    out[i] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(-1));
  __syncthreads();

  // (5) one bounding box per chunk of 64 sorted points (2 per lane)
  for (int c = warp; c < Np / kKnnChunk; c += kSortThreads / 32) {
    float bmn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
    float bmx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 q = out[c * kKnnChunk + h * 32 + lane];
      if (__float_as_int(q.w) >= 0) {
        bmn[0] = fminf(bmn[0], q.x); bmx[0] = fmaxf(bmx[0], q.x);
        bmn[1] = fminf(bmn[1], q.y); bmx[1] = fmaxf(bmx[1], q.y);
        bmn[2] = fminf(bmn[2], q.z); bmx[2] = fmaxf(bmx[2], q.z);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        bmn[d] = fminf(bmn[d], __shfl_xor_sync(0xffffffffu, bmn[d], o));
        bmx[d] = fmaxf(bmx[d], __shfl_xor_sync(0xffffffffu, bmx[d], o));
      }
    if (lane == 0) {
      bx[2 * c] = make_float4(bmn[0], bmn[1], bmn[2], 0.f);
      bx[2 * c + 1] = make_float4(bmx[0], bmx[1], bmx[2], 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// query
// ---------------------------------------------------------------------------------------------
// KC = compiled list length (>= K).  EXACT: K == KC and outputs are 16-byte aligned.
template <int KC, bool EXACT>
__global__ void __launch_bounds__(kKnnThreads)
knn_query_kernel(const float4* __restrict__ sorted, const float4* __restrict__ boxes, int N, int Np,
                 int T, int logT, int logV, int K, int32_t* __restrict__ ids,
                 float* __restrict__ dists) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * kKnnThreads + threadIdx.x;  // sorted position of this thread's query
  const float4* cloud = sorted + (long long)b * Np;
  const float4* bx = boxes + (long long)b * (Np / kKnnChunk) * 2;
  const int nchunks = Np / kKnnChunk;

  float qx = 0.f, qy = 0.f, qz = 0.f;
  int y = -1;
  if (p < Np) {
    const float4 q = __ldg(cloud + p);
    y = __float_as_int(q.w);
    if (y >= 0) { qx = q.x; qy = q.y; qz = q.z; }
  }
  const bool active = y >= 0;

  float sq[KC];
  int rk[KC];
#pragma unroll
  for (int j = 0; j < KC; ++j) { sq[j] = 3.402823466e+38f; rk[j] = INT_MAX; }  // reference padding lanes
  float kth_s = 3.402823466e+38f;  // current K-th key ...
  int kth_r = INT_MAX;             // ... and its rank
  float thr2 = active ? CUDART_INF_F : -1.f;  // d^2 bound implied by kth_s (inactive lanes never pass)

  const int c0 = min((int)((blockIdx.x * kKnnThreads + (threadIdx.x & ~31)) / kKnnChunk), nchunks - 1);
  // walk outward: c0, c0+1, c0-1, c0+2, ...
  for (int step = 0; step < 2 * nchunks; ++step) {
    const int c = (step & 1) ? c0 + ((step + 1) >> 1) : c0 - (step >> 1);
    if (c < 0 || c >= nchunks) continue;
    const float4 lo4 = __ldg(bx + 2 * c), hi4 = __ldg(bx + 2 * c + 1);
    // lower bound of the computed d^2 to any point in the box (same op sequence as below)
    const float ex = fmaxf(fmaxf(lo4.x - qx, qx - hi4.x), 0.f);
    const float ey = fmaxf(fmaxf(lo4.y - qy, qy - hi4.y), 0.f);
    const float ez = fmaxf(fmaxf(lo4.z - qz, qz - hi4.z), 0.f);
    const float lb = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
    if (!__any_sync(0xffffffffu, lb <= thr2)) continue;
    const float4* cand = cloud + c * kKnnChunk;
#pragma unroll 8
    for (int j = 0; j < kKnnChunk; ++j) {
      const float4 v = __ldg(cand + j);
      const float dx = v.x - qx, dy = v.y - qy, dz = v.z - qz;
      const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      if (d2 <= thr2) {
        const float s = __fsqrt_rn(d2);
        const int x = __float_as_int(v.w);
        const int r = knn_rank_of(x, T, logT, logV);
        if (x >= 0 && (s < kth_s || (s == kth_s && r < kth_r))) {
#pragma unroll
          for (int i = KC - 1; i > 0; --i) {
            const bool before_prev = s < sq[i - 1] || (s == sq[i - 1] && r < rk[i - 1]);
            const bool before_here = s < sq[i] || (s == sq[i] && r < rk[i]);
            if (before_prev) { sq[i] = sq[i - 1]; rk[i] = rk[i - 1]; }
            else if (before_here) { sq[i] = s; rk[i] = r; }
          }
          if (s < sq[0] || (s == sq[0] && r < rk[0])) { sq[0] = s; rk[0] = r; }
          if constexpr (EXACT) {
            kth_s = sq[KC - 1]; kth_r = rk[KC - 1];
          } else {
            kth_s = sq[0]; kth_r = rk[0];
#pragma unroll
            for (int i = 1; i < KC; ++i) { kth_s = (i < K) ? sq[i] : kth_s; kth_r = (i < K) ? rk[i] : kth_r; }
          }
          // every d2 with sqrt.rn(d2) <= kth_s satisfies d2 <= kth_s^2*(1+2^-23) < the bound below
          thr2 = __fmul_rn(__fmul_rn(kth_s, kth_s), 1.000001f);
        }
      }
    }
  }

  if (active) {
    int32_t* oi = ids + ((long long)b * N + y) * K;
    float* od = dists + ((long long)b * N + y) * K;
    int id[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) id[j] = (rk[j] == INT_MAX) ? -1 : knn_point_of(rk[j], T, logV);
    if constexpr (EXACT && KC % 4 == 0) {
#pragma unroll
      for (int j = 0; j < KC; j += 4) {
        *reinterpret_cast<int4*>(oi + j) = make_int4(id[j], id[j + 1], id[j + 2], id[j + 3]);
        *reinterpret_cast<float4*>(od + j) = make_float4(sq[j], sq[j + 1], sq[j + 2], sq[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < KC; ++j)
        if (j < K) { oi[j] = id[j]; od[j] = sq[j]; }
    }
  }
}

int knn_launch(const float* pos, int B, int N, int K, long long sb, int sp, int sd, int32_t* ids,
               float* dists, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (knn_use_tiled())
    return knn_tiled_launch(pos, B, N, K, sb, sp, sd, ids, dists, workspace, workspace_bytes, st);
  if (!pos || !ids || !dists) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0) return DH3D_ERR_DIM;
  if (K > 32 || N > 65536 || B > 65535) return DH3D_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < knn_workspace_bytes(B, N)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 127) != 0) return DH3D_ERR_ALIGN;
  const KnnOrder o = knn_order(N);
  const int Np = knn_padded(N);
  float4* sorted = reinterpret_cast<float4*>(workspace);
  float4* boxes = sorted + (size_t)B * Np;
  knn_sort_kernel<<<B, kSortThreads, 0, st>>>(pos, N, Np, sb, sp, sd, sorted, boxes);
  int rc = launch_status();
  if (rc != DH3D_OK) return rc;
  dim3 grid(ceil_div(Np, kKnnThreads), B);
  const bool vec_ok = (((uintptr_t)ids | (uintptr_t)dists) & 15) == 0;
#define DH3D_KNN(KC, EX)                                                                        \
  knn_query_kernel<KC, EX><<<grid, kKnnThreads, 0, st>>>(sorted, boxes, N, Np, o.T, o.logT, o.logV, K, \
                                                         ids, dists)
  if (K == 8 && vec_ok) DH3D_KNN(8, true);
  else if (K == 16 && vec_ok) DH3D_KNN(16, true);
  else if (K == 32 && vec_ok) DH3D_KNN(32, true);
  else if (K <= 4) DH3D_KNN(4, false);
  else if (K <= 8) DH3D_KNN(8, false);
  else if (K <= 16) DH3D_KNN(16, false);
  else DH3D_KNN(32, false);
#undef DH3D_KNN
  return launch_status();
}

}  // namespace dh3d
