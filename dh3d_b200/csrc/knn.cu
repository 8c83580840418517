// k-NN inside each cloud, results identical to the exhaustive search
// (reference: user_ops/kernels/knn_bruteforce_kernel_gpu.cu.cc).
//
// The reference launches one CTA per QUERY point and radix-sorts all N keys to extract K of them
// (N <= 8192, followed by a device sync).  Here:
//   1. knn_sort_kernel (one CTA per cloud): counting sort of the points by the Hilbert index of a
//      16^3 grid cell -> spatially coherent float4 array (x,y,z,original index) + one bounding box
//      per 32-point chunk + one per 16 chunks (two-level: a warp tests 16 coarse boxes, then only the fine
//      boxes inside the coarse ones that can still matter).
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
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace dh3d {

constexpr int kKnnChunk = 32;       // candidates per bounding box (one per lane of the box-building warp)
constexpr int kKnnSuper = 16;       // chunks per second-level box (512 points)
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
// float4s of box storage per cloud: (min,max) per chunk, then (min,max) per super-chunk
static inline size_t knn_box_f4(int Np) {
  return 2 * (size_t)(Np / kKnnChunk) + 2 * (size_t)ceil_div(Np / kKnnChunk, kKnnSuper);
}

size_t knn_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  const size_t np = knn_padded(N);
  return (size_t)B * (np + knn_box_f4((int)np)) * sizeof(float4);
}

__device__ __forceinline__ long long knn_box_f4_dev(int Np) {
  const int nc = Np / kKnnChunk;
  return 2LL * nc + 2LL * ((nc + kKnnSuper - 1) / kKnnSuper);
}

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 4 bits -> every third bit
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

// ---------------------------------------------------------------------------------------------
// sort: positions (any layout via strides) -> Morton-cell-sorted float4 (x,y,z,idx) + chunk boxes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads, 1)
knn_sort_kernel(const float* __restrict__ pos, int N, int Np, long long sb, int sp, int sd, int T, int logT, int logV,
                float4* __restrict__ sorted, float4* __restrict__ boxes) {
  __shared__ int s_hist[kCells];
  __shared__ float s_red[6][kSortThreads / 32];
  __shared__ float s_box[6];
  __shared__ int s_wsum[kSortThreads / 32];
  __shared__ int s_axis[3][64];   // per-axis occupancy histogram of the grid range (gap trimming)
  __shared__ int s_changed[6];    // one flag per trimming round (no reuse: no reset race)
  __shared__ int s_nout;          // points outside the trimmed grid range: their own bucket after the last cell
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = pos + (long long)b * sb;
  float4* out = sorted + (long long)b * Np;
  float4* bx = boxes + (long long)b * knn_box_f4_dev(Np);

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
  // Gap trimming of the grid range.  The 16^3 grid spans [s_box lo, hi]; a few far points (the reference pads
  // short clouds with points at 1e5 when randsample=False, core/utils.py:107-108; stray returns) would stretch it
  // until every real point shares one cell and no box prunes anything (measured: k-NN 0.27 -> 1.24 ms).  Up to 6
  // rounds: histogram each axis over 64 bins of its current range; if some run of EMPTY bins covers at least half
  // of the occupied range, keep only the side of the gap that holds more points, then tighten to the occupied
  // bins.  Well-spread clouds never have such a gap and keep their exact bounding box (round 0 changes nothing).  Points left outside the range are sorted into one extra bucket after the
  // last cell (their own chunks, with whatever boxes they get), so they cannot inflate the boxes of real chunks.
  // This only changes the SCAN ORDER and the pruning power, never the result.
  for (int round = 0; round < 6; ++round) {
    if (tid < 3 * 64) (&s_axis[0][0])[tid] = 0;
    if (tid == 0) s_changed[round] = 0;
    __syncthreads();
    {
      float rlo[3], rinv[3], rhi[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        rlo[d] = s_box[d];
        rhi[d] = s_box[3 + d];
        const float ext = rhi[d] - rlo[d];
        rinv[d] = (ext > 0.f && ext < CUDART_INF_F) ? 64.f / ext : 0.f;
      }
      for (int i = tid; i < N; i += kSortThreads) {
        float v[3];
        bool in = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          v[d] = P[(long long)i * sp + d * sd];
          in = in && (v[d] >= rlo[d]) && (v[d] <= rhi[d]);
        }
        if (in) {
#pragma unroll
          for (int d = 0; d < 3; ++d) atomicAdd(&s_axis[d][min(63, (int)((v[d] - rlo[d]) * rinv[d]))], 1);
        }
      }
    }
    __syncthreads();
    if (tid < 3) {
      const int d = tid;
      int first = -1, last = -1;
      for (int bin = 0; bin < 64; ++bin)
        if (s_axis[d][bin] != 0) { if (first < 0) first = bin; last = bin; }
      if (first >= 0) {
        int best_a = 0, best_len = 0, run_a = 0, run_len = 0;   // longest run of empty bins strictly inside
        for (int bin = first; bin <= last; ++bin) {
          if (s_axis[d][bin] == 0) {
            if (run_len == 0) run_a = bin;
            ++run_len;
            if (run_len > best_len) { best_len = run_len; best_a = run_a; }
          } else {
            run_len = 0;
          }
        }
        const int span = last - first + 1;
        const float l0 = s_box[d], w = (s_box[3 + d] - s_box[d]) * (1.f / 64.f);
        if (2 * best_len >= span && best_len >= 8) {        // a gap over half of the occupied range: keep the heavier side
          int left = 0, right = 0;
          for (int bin = first; bin < best_a; ++bin) left += s_axis[d][bin];
          for (int bin = best_a + best_len; bin <= last; ++bin) right += s_axis[d][bin];
          if (left >= right) { s_box[d] = l0 + w * (float)first; s_box[3 + d] = l0 + w * (float)best_a; }
          else { s_box[d] = l0 + w * (float)(best_a + best_len); s_box[3 + d] = l0 + w * (float)(last + 1); }
          s_changed[round] = 1;
        } else if (span <= 48 && round > 0) {              // after a cut: tighten to the occupied bins
          s_box[d] = l0 + w * (float)first;
          s_box[3 + d] = l0 + w * (float)(last + 1);
          s_changed[round] = 1;
        }
      }
    }
    __syncthreads();
    if (!s_changed[round]) break;
  }
  if (tid == 0) s_nout = 0;
  float lo[3], hi[3], inv[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = s_box[d];
    hi[d] = s_box[3 + d];
    const float ext = s_box[3 + d] - s_box[d];
    inv[d] = (ext > 0.f && ext < CUDART_INF_F) ? 16.f / ext : 0.f;
  }
  auto cell_of = [&](int i) -> int {
    unsigned c[3];
    bool outside = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float p = P[(long long)i * sp + d * sd];
      outside = outside || (p < lo[d]) || (p > hi[d]);
      const float v = (p - lo[d]) * inv[d];
      int q = (v >= 0.f) ? ((v < 15.f) ? (int)v : 15) : 0;  // NaN -> 0; only the scan order depends on it
      c[d] = (unsigned)q;
    }
    if (outside) return kCells;
    // Hilbert index of the cell (Skilling's axes-to-transpose, 4 bits per axis): consecutive cells are always
    // face neighbours, so a window of consecutive sorted points has a compact bounding box (a Morton window
    // that straddles a block boundary spans far more space: ~80 chunks per warp had to be scanned, ncu r1l)
    constexpr unsigned Mb = 8u;  // 1 << (bits - 1)
#pragma unroll
    for (unsigned Q = Mb; Q > 1u; Q >>= 1) {
      const unsigned Pm = Q - 1u;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (c[d] & Q) c[0] ^= Pm;
        else { const unsigned t = (c[0] ^ c[d]) & Pm; c[0] ^= t; c[d] ^= t; }
      }
    }
    c[1] ^= c[0];
    c[2] ^= c[1];
    unsigned t = 0u;
#pragma unroll
    for (unsigned Q = Mb; Q > 1u; Q >>= 1)
      if (c[2] & Q) t ^= Q - 1u;
    c[0] ^= t; c[1] ^= t; c[2] ^= t;
    return (int)(spread3(c[2]) | (spread3(c[1]) << 1) | (spread3(c[0]) << 2));
  };

  // (2) histogram, (3) exclusive scan, (4) scatter
  for (int i = tid; i < N; i += kSortThreads) {
    const int c = cell_of(i);
    atomicAdd(c < kCells ? &s_hist[c] : &s_nout, 1);
  }
  __syncthreads();
  const int n_inside = N - s_nout;
  __syncthreads();
  if (tid == 0) s_nout = 0;   // now the running position inside the outside bucket
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
    const int c = cell_of(i);
    const int p = c < kCells ? atomicAdd(&s_hist[c], 1) : n_inside + atomicAdd(&s_nout, 1);
    out[p] = make_float4(P[(long long)i * sp], P[(long long)i * sp + sd], P[(long long)i * sp + 2 * sd],
                         __int_as_float(i));
  }
  for (int i = N + tid; i < Np; i += kSortThreads)
    out[i] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(-1));
  __syncthreads();

  // (5) one bounding box per chunk of 32 sorted points (one per lane), (6) one per 16 chunks
  const int nchunks = Np / kKnnChunk;
  for (int c = warp; c < nchunks; c += kSortThreads / 32) {
    float bmn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
    float bmx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    const float4 q = out[c * kKnnChunk + lane];
    int mr = INT_MAX;   // smallest tie rank in the chunk: with the box bound, a lower bound of every packed (key, rank)
    int mi = INT_MAX;   // smallest point index (the 3-NN metric's tie rank): one sorted copy serves both metrics
    if (__float_as_int(q.w) >= 0) {
      bmn[0] = q.x; bmx[0] = q.x;
      bmn[1] = q.y; bmx[1] = q.y;
      bmn[2] = q.z; bmx[2] = q.z;
      mr = knn_rank_of(__float_as_int(q.w), T, logT, logV);
      mi = __float_as_int(q.w);
    }
    mr = __reduce_min_sync(0xffffffffu, mr);
    mi = __reduce_min_sync(0xffffffffu, mi);
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        bmn[d] = fminf(bmn[d], __shfl_xor_sync(0xffffffffu, bmn[d], o));
        bmx[d] = fmaxf(bmx[d], __shfl_xor_sync(0xffffffffu, bmx[d], o));
      }
    if (lane == 0) {
      bx[2 * c] = make_float4(bmn[0], bmn[1], bmn[2], __int_as_float(mr));
      bx[2 * c + 1] = make_float4(bmx[0], bmx[1], bmx[2], __int_as_float(mi));
    }
  }
  __syncthreads();
  const int nsuper = (nchunks + kKnnSuper - 1) / kKnnSuper;
  float4* sbx = bx + 2 * nchunks;
  for (int s = warp; s < nsuper; s += kSortThreads / 32) {
    const int c = s * kKnnSuper + (lane & (kKnnSuper - 1));
    float4 lo4 = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(INT_MAX));
    float4 hi4 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, __int_as_float(INT_MAX));
    if (c < nchunks) { lo4 = bx[2 * c]; hi4 = bx[2 * c + 1]; }
    float bmn[3] = {lo4.x, lo4.y, lo4.z}, bmx[3] = {hi4.x, hi4.y, hi4.z};
    int mr = __float_as_int(lo4.w), mi = __float_as_int(hi4.w);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      mr = min(mr, __shfl_xor_sync(0xffffffffu, mr, o));
      mi = min(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        bmn[d] = fminf(bmn[d], __shfl_xor_sync(0xffffffffu, bmn[d], o));
        bmx[d] = fmaxf(bmx[d], __shfl_xor_sync(0xffffffffu, bmx[d], o));
      }
    if (lane == 0) {
      sbx[2 * s] = make_float4(bmn[0], bmn[1], bmn[2], __int_as_float(mr));
      sbx[2 * s + 1] = make_float4(bmx[0], bmx[1], bmx[2], __int_as_float(mi));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// query
// ---------------------------------------------------------------------------------------------
// Distance policies.  d2() is the exact operation sequence of the reference, key() the sort key,
// bound() a value such that every candidate whose key could still enter the list has d2 <= bound.
struct KnnMetric {      // knn_bruteforce_kernel_gpu.cu.cc:102-107 (nvcc-contracted), key = sqrt.rn
  static __device__ __forceinline__ float d2(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  }
  // two candidates per instruction (FMUL2 / FFMA2): the same operations in the same order, lane by lane
  static constexpr bool kPacked = true;
  static __device__ __forceinline__ unsigned long long d2v(unsigned long long dx, unsigned long long dy,
                                                           unsigned long long dz) {
    return ffma2v(dz, dz, ffma2v(dy, dy, fmul2(dx, dx)));
  }
  static __device__ __forceinline__ float key(float d) { return __fsqrt_rn(d); }
  // every d2 with sqrt.rn(d2) <= k satisfies d2 <= k^2*(1+2^-23) < the bound below
  static __device__ __forceinline__ float bound(float k) { return __fmul_rn(__fmul_rn(k, k), 1.000001f); }
  static __device__ __forceinline__ int rank_of(int x, int T, int logT, int logV) {
    return knn_rank_of(x, T, logT, logV);
  }
  static __device__ __forceinline__ int point_of(int r, int T, int logV) { return knn_point_of(r, T, logV); }
  static constexpr uint32_t kInitKey = 0x7f7fffffu;  // FLT_MAX, the reference's padding key
  static constexpr int kInitRank = INT_MAX;
  static constexpr bool kRankIsIndex = false;   // boxes: smallest tie rank of the box in lo.w
};
struct ThreeNnMetric {  // tf_interpolate.cpp:60-103: un-fused (dx*dx + dy*dy) + dz*dz, key = d2, ties to low index
  static __device__ __forceinline__ float d2(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  }
  // no packed form: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (seen in SASS; 3-NN indices changed),
  // and this metric is defined by its UN-fused roundings -- the scalar __fmul_rn / __fadd_rn path is used
  static constexpr bool kPacked = false;
  static __device__ __forceinline__ float key(float d) { return d; }
  static __device__ __forceinline__ float bound(float k) { return k; }
  static __device__ __forceinline__ int rank_of(int x, int, int, int) { return x; }
  static __device__ __forceinline__ int point_of(int r, int, int) { return r; }
  static constexpr uint32_t kInitKey = 0x7f800000u;  // best = 1e40 -> +inf in float, index 0
  static constexpr int kInitRank = 0;
  static constexpr bool kRankIsIndex = true;    // boxes: smallest point index of the box in hi.w
};


// One thread per query, KC-entry sorted list of packed (key bits, rank) pairs in registers.  A non-negative
// float orders like its bit pattern, so one 64-bit integer compare is the reference's (key, tie-rank) order.
// The scan loop never branches into the insertion: it only sets one bit per candidate that is under the lane's
// bound; at the end of the chunk the warp inserts the marked candidates together (coordinates read back from the
// staged chunk, distance recomputed with the same IEEE operations), which is where the bound tightens.  A stale
// bound is only looser, so nothing is lost.  KC >= K; EXACT: K == KC and 16-byte aligned outputs.
// SELF: queries are the candidates themselves (k-NN; the walk starts at the warp's own chunk), otherwise
// the walk starts at the chunk whose box is nearest to the warp's first query (3-NN).
template <int KC, bool EXACT, bool SELF, class M>
__global__ void __launch_bounds__(kKnnThreads)
knn_query_kernel(const float4* __restrict__ qsorted, int Nq, int Nqp, const float4* __restrict__ sorted,
                 const float4* __restrict__ boxes, int Np, int T, int logT, int logV, int K,
                 int32_t* __restrict__ ids, float* __restrict__ dists) {
  __shared__ __align__(16) float4 s_chunk[kKnnThreads / 32][kKnnChunk];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int p = blockIdx.x * kKnnThreads + tid;  // sorted position of this thread's query
  const float4* cloud = sorted + (long long)b * Np;
  const float4* bx = boxes + (long long)b * knn_box_f4_dev(Np);
  const int nchunks = Np / kKnnChunk;
  const int nsuper = (nchunks + kKnnSuper - 1) / kKnnSuper;
  const float4* sbx = bx + 2 * nchunks;

  float qx = 0.f, qy = 0.f, qz = 0.f;
  int y = -1;
  if (p < Nqp) {
    const float4 q = __ldg(qsorted + (long long)b * Nqp + p);
    y = __float_as_int(q.w);
    if (y >= 0) { qx = q.x; qy = q.y; qz = q.z; }
  }
  const bool active = y >= 0;

  constexpr unsigned long long kInit = ((unsigned long long)M::kInitKey << 32) | (uint32_t)M::kInitRank;
  unsigned long long L[KC];
#pragma unroll
  for (int j = 0; j < KC; ++j) L[j] = kInit;
  float thr2 = active ? CUDART_INF_F : -1.f;  // inactive lanes never pass
  unsigned long long kth_pk = kInit;           // the lane's current K-th packed (key, rank)

  // Inserts the candidates of the staged chunk whose bit is set in `mask` (the scan's "under the lane's bound" bits)
  // into the lane's list.  The candidate's coordinates are read back from the warp's staged chunk (pair-interleaved
  // floats: candidate j = 2P + h has x, y at floats 8P + h, 8P + 2 + h, z and its index at 8P + 4 + h, 8P + 6 + h) and
  // its distance recomputed with the scalar form of the metric (the same IEEE operations as the packed scan).
  auto flush = [&](uint32_t mask, const float* cf) {
    int m = __popc(mask);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    for (int e = 0; e < m; ++e) {
      if (mask != 0u) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1u;
        const float* c = cf + 8 * (j >> 1) + (j & 1);
        const int x = __float_as_int(c[6]);
        const float d2 = M::d2(c[0] - qx, c[2] - qy, c[4] - qz);
        // re-test against the bound as tightened by this flush's own insertions: most marked candidates of a chunk
        // fall out here, before the sqrt / pack / 64-bit insertion network
        if (x >= 0 && d2 <= thr2) {
          const unsigned long long pk = ((unsigned long long)__float_as_uint(M::key(d2)) << 32) |
                                        (uint32_t)M::rank_of(x, T, logT, logV);
          if (pk < kth_pk) {
#pragma unroll
            for (int i = KC - 1; i > 0; --i) {
              const bool before_prev = pk < L[i - 1];
              const bool before_here = pk < L[i];
              L[i] = before_prev ? L[i - 1] : (before_here ? pk : L[i]);
            }
            if (pk < L[0]) L[0] = pk;
            unsigned long long kth = L[KC - 1];
            if constexpr (!EXACT) {
              kth = L[0];
#pragma unroll
              for (int i = 1; i < KC; ++i) kth = (i < K) ? L[i] : kth;
            }
            kth_pk = kth;
            thr2 = M::bound(__uint_as_float((uint32_t)(kth >> 32)));
          }
        }
      }
    }
  };

  // Can a box still hold a candidate that enters this lane's list?  lb = lower bound of every d2 in the box
  // (same op sequence as the distance, monotone roundings), minrank = smallest tie rank in the box: every
  // candidate's packed (key, rank) is >= (key(lb), minrank) because key() is monotone.  The rank half is what
  // keeps clouds with masses of EQUAL keys (duplicated padding points; the all-zero padding clouds of the
  // reference's extractors, where every box bound ties with every list) from degenerating into a full scan.
  auto box_can_enter = [&](float lb, int minrank) -> bool {
    if (!(lb <= thr2)) return false;
    const unsigned long long pk = ((unsigned long long)__float_as_uint(M::key(lb)) << 32) | (uint32_t)minrank;
    return pk < kth_pk;
  };

  auto box_lb = [&](const float4* boxp, int c) -> float {
    const float4 lo4 = __ldg(boxp + 2 * c), hi4 = __ldg(boxp + 2 * c + 1);
    // lower bound of the computed d2 to any point in the box (same op sequence, monotone roundings)
    const float ex = fmaxf(fmaxf(lo4.x - qx, qx - hi4.x), 0.f);
    const float ey = fmaxf(fmaxf(lo4.y - qy, qy - hi4.y), 0.f);
    const float ez = fmaxf(fmaxf(lo4.z - qz, qz - hi4.z), 0.f);
    return M::d2(ex, ey, ez);
  };

  int c0;
  if constexpr (SELF) {
    c0 = min((int)((blockIdx.x * kKnnThreads + (tid & ~31)) / kKnnChunk), nchunks - 1);
  } else {
    float best = CUDART_INF_F;
    int bc = 0;
    for (int c = 0; c < nchunks; ++c) {
      const float lb = box_lb(bx, c);
      if (lb < best) { best = lb; bc = c; }
    }
    // the warp's first active lane decides (queries of a warp are spatial neighbours)
    const unsigned act = __ballot_sync(0xffffffffu, active);
    c0 = __shfl_sync(0xffffffffu, bc, act ? __ffs(act) - 1 : 0);
  }

  // bounding box of the warp's 32 queries (for the lane-parallel coarse test of 16 fine boxes at once)
  float wlo[3] = {active ? qx : CUDART_INF_F, active ? qy : CUDART_INF_F, active ? qz : CUDART_INF_F};
  float whi[3] = {active ? qx : -CUDART_INF_F, active ? qy : -CUDART_INF_F, active ? qz : -CUDART_INF_F};
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wlo[d] = fminf(wlo[d], __shfl_xor_sync(0xffffffffu, wlo[d], o));
      whi[d] = fmaxf(whi[d], __shfl_xor_sync(0xffffffffu, whi[d], o));
    }
  const int lane = tid & 31;

  // two-level walk: super-chunks outward from the warp's own (s0, s0+1, s0-1, ...).  Inside a surviving
  // super-chunk lane j loads fine box j ONCE (one coalesced request instead of 16 dependent uniform loads) and
  // bounds the distance between it and the warp's query box; a chunk is then visited only if that coarse bound
  // is within the largest per-lane bound, and scanned only if the exact per-lane test (box broadcast by shuffles)
  // passes.  Chunks nearest (in sorted order) to the warp's own chunk come first, so the bounds tighten early.
  const int s0 = c0 / kKnnSuper;
  const int sreach = max(s0, nsuper - 1 - s0);
  for (int sstep = 0; sstep <= 2 * sreach; ++sstep) {
    const int sc = (sstep & 1) ? s0 + ((sstep + 1) >> 1) : s0 - (sstep >> 1);
    if (sc < 0 || sc >= nsuper) continue;
    if (!__any_sync(0xffffffffu, box_can_enter(box_lb(sbx, sc),
                                               __float_as_int(__ldg(sbx + 2 * sc + (M::kRankIsIndex ? 1 : 0)).w))))
      continue;
    const int cbeg = sc * kKnnSuper, cend = min(cbeg + kKnnSuper, nchunks);
    const int cn = cend - cbeg;
    float4 blo = make_float4(0.f, 0.f, 0.f, 0.f), bhi = blo;
    float lbw = CUDART_INF_F;
    {
      const int j = lane & (kKnnSuper - 1);
      if (j < cn) {
        blo = __ldg(bx + 2 * (cbeg + j));
        bhi = __ldg(bx + 2 * (cbeg + j) + 1);
        // p in [blo,bhi], q in [wlo,whi]: fl(p - q) >= fl(blo - whi) and fl(q - p) >= fl(wlo - bhi) (monotone rounding)
        const float ex = fmaxf(fmaxf(blo.x - whi[0], wlo[0] - bhi.x), 0.f);
        const float ey = fmaxf(fmaxf(blo.y - whi[1], wlo[1] - bhi.y), 0.f);
        const float ez = fmaxf(fmaxf(blo.z - whi[2], wlo[2] - bhi.z), 0.f);
        lbw = M::d2(ex, ey, ez);
      }
    }
    // start inside this super-chunk: own chunk, or the end facing the own super-chunk
    const int cs = (sc == s0) ? c0 : (sc > s0 ? cbeg : cend - 1);
    for (int step = 0; step < 2 * cn; ++step) {
      const int c = (step & 1) ? cs + ((step + 1) >> 1) : cs - (step >> 1);
      if (c < cbeg || c >= cend) continue;
      const int j = c - cbeg;
      // largest bound in the warp (non-negative floats order like their bit patterns; inactive lanes hold -1)
      const float tmax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(thr2)));
      if (!(__shfl_sync(0xffffffffu, lbw, j) <= tmax)) continue;
      {
        const float lx = __shfl_sync(0xffffffffu, blo.x, j), ly = __shfl_sync(0xffffffffu, blo.y, j),
                    lz = __shfl_sync(0xffffffffu, blo.z, j);
        const float hx = __shfl_sync(0xffffffffu, bhi.x, j), hy = __shfl_sync(0xffffffffu, bhi.y, j),
                    hz = __shfl_sync(0xffffffffu, bhi.z, j);
        const int mr = __shfl_sync(0xffffffffu, __float_as_int(M::kRankIsIndex ? bhi.w : blo.w), j);
        const float ex = fmaxf(fmaxf(lx - qx, qx - hx), 0.f);
        const float ey = fmaxf(fmaxf(ly - qy, qy - hy), 0.f);
        const float ez = fmaxf(fmaxf(lz - qz, qz - hz), 0.f);
        if (!__any_sync(0xffffffffu, box_can_enter(M::d2(ex, ey, ez), mr))) continue;
      }
      // one coalesced 512-byte load per chunk, staged in the warp's shared-memory slot and read back as LDS.128
      // broadcasts
      // The staged chunk is pair-interleaved -- slot 2P = (x0,x1,y0,y1), slot 2P+1 = (z0,z1,w0,w1) of candidates
      // 2P, 2P+1 -- so two broadcast LDS.128 feed packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): 6 instead of
      // 12 FP instructions per candidate pair, bit-identical distances.  Lane pairs swap halves with one shuffle
      // pair so that every lane still writes one 16-byte slot.
      const float4 mine = __ldg(cloud + c * kKnnChunk + lane);
      const bool odd = lane & 1;
      const float s0v = __shfl_xor_sync(0xffffffffu, odd ? mine.x : mine.z, 1);
      const float s1v = __shfl_xor_sync(0xffffffffu, odd ? mine.y : mine.w, 1);
      __syncwarp();
      s_chunk[tid >> 5][lane] = odd ? make_float4(s0v, mine.z, s1v, mine.w) : make_float4(mine.x, s0v, mine.y, s1v);
      __syncwarp();
      const ulonglong2* cand = reinterpret_cast<const ulonglong2*>(s_chunk[tid >> 5]);
      // The scan only records WHICH candidates are under the lane's bound, one bit each: 12 instructions per candidate
      // pair (2 LDS.128, 6 packed FP, 2 FSETP, 2 predicated LOP3) with no dependent address arithmetic; the buffered
      // (d2, index) append it replaces cost 21 (predicated STS.64 + counter + address chain per candidate).
      uint32_t mask = 0u;
#pragma unroll
      for (int j0 = 0; j0 < kKnnChunk; j0 += 2) {
        const ulonglong2 xy = cand[j0], zw = cand[j0 + 1];
        float2 d2;
        if constexpr (M::kPacked) {
          d2 = unpack2(M::d2v(fsub2s(xy.x, qx), fsub2s(xy.y, qy), fsub2s(zw.x, qz)));
        } else {
          const float2 x = unpack2(xy.x), yv = unpack2(xy.y), z = unpack2(zw.x);
          d2.x = M::d2(x.x - qx, yv.x - qy, z.x - qz);
          d2.y = M::d2(x.y - qx, yv.y - qy, z.y - qz);
        }
        if (d2.x <= thr2) mask |= 1u << j0;
        if (d2.y <= thr2) mask |= 2u << j0;
      }
      if (__any_sync(0xffffffffu, mask != 0u)) flush(mask, reinterpret_cast<const float*>(s_chunk[tid >> 5]));
    }
  }

  if (active) {
    int32_t* oi = ids + ((long long)b * Nq + y) * K;
    float* od = dists + ((long long)b * Nq + y) * K;
    int id[KC];
    float ky[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) {
      const int r = (int)(uint32_t)(L[j] & 0xffffffffull);
      ky[j] = __uint_as_float((uint32_t)(L[j] >> 32));
      id[j] = (SELF && r == INT_MAX) ? -1 : M::point_of(r, T, logV);
    }
    if constexpr (EXACT && KC % 4 == 0) {
#pragma unroll
      for (int j = 0; j < KC; j += 4) {
        *reinterpret_cast<int4*>(oi + j) = make_int4(id[j], id[j + 1], id[j + 2], id[j + 3]);
        *reinterpret_cast<float4*>(od + j) = make_float4(ky[j], ky[j + 1], ky[j + 2], ky[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < KC; ++j)
        if (j < K) { oi[j] = id[j]; od[j] = ky[j]; }
    }
  }
}

// sort only: positions (any layout via strides) -> the workspace's cell-sorted copy + chunk boxes
int knn_sort_launch(const float* pos, int B, int N, long long sb, int sp, int sd, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  if (!pos) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (N > 65536 || B > 65535) return DH3D_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < knn_workspace_bytes(B, N)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 127) != 0) return DH3D_ERR_ALIGN;
  const KnnOrder o = knn_order(N);
  const int Np = knn_padded(N);
  float4* sorted = reinterpret_cast<float4*>(workspace);
  float4* boxes = sorted + (size_t)B * Np;
  knn_sort_kernel<<<B, kSortThreads, 0, st>>>(pos, N, Np, sb, sp, sd, o.T, o.logT, o.logV, sorted, boxes);
  return launch_status();
}

// query only: the workspace as knn_sort_launch (of the same B, N) left it
int knn_query_sorted_launch(const void* workspace, int B, int N, int K, int32_t* ids, float* dists, cudaStream_t st) {
  if (!workspace || !ids || !dists) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0) return DH3D_ERR_DIM;
  if (K > 64 || N > 65536 || B > 65535) return DH3D_ERR_UNSUPPORTED;
  if (((uintptr_t)workspace & 127) != 0) return DH3D_ERR_ALIGN;
  const KnnOrder o = knn_order(N);
  const int Np = knn_padded(N);
  const float4* sorted = reinterpret_cast<const float4*>(workspace);
  const float4* boxes = sorted + (size_t)B * Np;
  dim3 grid(ceil_div(Np, kKnnThreads), B);
  const bool vec_ok = (((uintptr_t)ids | (uintptr_t)dists) & 15) == 0;
#define DH3D_KNN(KC, EX)                                                                              \
  knn_query_kernel<KC, EX, true, KnnMetric><<<grid, kKnnThreads, 0, st>>>(sorted, N, Np, sorted, boxes, Np, \
                                                                          o.T, o.logT, o.logV, K, ids,       \
                                                                          dists)
  if (K == 8 && vec_ok) DH3D_KNN(8, true);
  else if (K == 16 && vec_ok) DH3D_KNN(16, true);
  else if (K == 32 && vec_ok) DH3D_KNN(32, true);
  else if (K <= 4) DH3D_KNN(4, false);
  else if (K <= 8) DH3D_KNN(8, false);
  else if (K <= 16) DH3D_KNN(16, false);
  else if (K <= 32) DH3D_KNN(32, false);
  else DH3D_KNN(64, false);   // keypoint NMS uses K = 50 (core/utils.py:17)
#undef DH3D_KNN
  return launch_status();
}

int knn_launch(const float* pos, int B, int N, int K, long long sb, int sp, int sd, int32_t* ids,
               float* dists, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!pos || !ids || !dists) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0) return DH3D_ERR_DIM;
  if (K > 64 || N > 65536 || B > 65535) return DH3D_ERR_UNSUPPORTED;
  int rc = knn_sort_launch(pos, B, N, sb, sp, sd, workspace, workspace_bytes, st);
  if (rc != DH3D_OK) return rc;
  return knn_query_sorted_launch(workspace, B, N, K, ids, dists, st);
}

// ---------------------------------------------------------------------------------------------
// three_nn through the same engine: Morton-sort the known points (candidates + chunk boxes) and the
// query points (spatially coherent warps), scan with the reference's un-fused arithmetic.
// workspace: B * { float4[np2] candidates, float4[2*np2/64] their boxes, float4[np1] queries,
//                   float4[2*np1/64] query-chunk boxes (written by the sort, unused) }
// ---------------------------------------------------------------------------------------------
size_t three_nn_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  const size_t np1 = knn_padded(n), np2 = knn_padded(m);
  return (size_t)b * (np2 + knn_box_f4(np2) + np1 + knn_box_f4(np1)) * sizeof(float4);
}

// sorted1: optional -- the query cloud as the k-NN of the same points left it at the start of its workspace (float4
// (x,y,z,index) [b][knn_padded(n)], any cell order): the query sort is then skipped (DH3D runs the k-NN of the dense
// cloud anyway; its sort is 30 us of the side stream's chain)
static int three_nn_pruned_impl(int b, int n, int m, const float* xyz1, const float4* sorted1, const float* xyz2,
                                const float4* sorted2, float* dist, int32_t* idx, void* workspace,
                                size_t workspace_bytes, cudaStream_t st) {
  if ((!xyz1 && !sorted1) || (!xyz2 && !sorted2) || !dist || !idx) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || m <= 0) return DH3D_ERR_DIM;
  if (b > 65535) return DH3D_ERR_UNSUPPORTED;
  const bool need_ws = !sorted1 || !sorted2;
  if (need_ws && (!workspace || workspace_bytes < three_nn_workspace_bytes(b, n, m))) return DH3D_ERR_WORKSPACE;
  if ((((uintptr_t)workspace | (uintptr_t)sorted1 | (uintptr_t)sorted2) & 127) != 0) return DH3D_ERR_ALIGN;
  const int np1 = knn_padded(n), np2 = knn_padded(m);
  float4* cands = reinterpret_cast<float4*>(workspace);
  float4* boxes = cands + (size_t)b * np2;
  float4* queries = boxes + (size_t)b * knn_box_f4(np2);
  float4* qboxes = queries + (size_t)b * np1;
  int rc;
  // tie rank of the 3-NN metric = the index itself: the sort kernel leaves the smallest INDEX of every box in hi.w
  // whatever rank parameters it ran with, so a k-NN workspace of the same points serves as well as an own sort
  if (!sorted1) {
    knn_sort_kernel<<<b, kSortThreads, 0, st>>>(xyz1, n, np1, 3LL * n, 3, 1, 1 << 30, 30, 0, queries, qboxes);
    if ((rc = launch_status()) != DH3D_OK) return rc;
  }
  const float4* c_pts = cands;
  const float4* c_box = boxes;
  if (sorted2) {   // a k-NN workspace of xyz2: [b][np2] sorted points, then the boxes
    c_pts = sorted2;
    c_box = sorted2 + (size_t)b * np2;
  } else {
    knn_sort_kernel<<<b, kSortThreads, 0, st>>>(xyz2, m, np2, 3LL * m, 3, 1, 1 << 30, 30, 0, cands, boxes);
    if ((rc = launch_status()) != DH3D_OK) return rc;
  }
  dim3 grid(ceil_div(np1, kKnnThreads), b);
  knn_query_kernel<4, false, false, ThreeNnMetric><<<grid, kKnnThreads, 0, st>>>(
      sorted1 ? sorted1 : queries, n, np1, c_pts, c_box, np2, 0, 0, 0, 3, idx, dist);
  return launch_status();
}

int three_nn_pruned_launch(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                           int32_t* idx, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  return three_nn_pruned_impl(b, n, m, xyz1, nullptr, xyz2, nullptr, dist, idx, workspace, workspace_bytes, st);
}

int three_nn_presorted_launch(int b, int n, int m, const void* knn_workspace_of_xyz1, const float* xyz2, float* dist,
                              int32_t* idx, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!knn_workspace_of_xyz1) return DH3D_ERR_NULL;
  return three_nn_pruned_impl(b, n, m, nullptr, reinterpret_cast<const float4*>(knn_workspace_of_xyz1), xyz2, nullptr,
                              dist, idx, workspace, workspace_bytes, st);
}

// both clouds pre-sorted (k-NN workspaces of xyz1 and of xyz2): one launch, no workspace
int three_nn_presorted2_launch(int b, int n, int m, const void* knn_workspace_of_xyz1, const void* knn_workspace_of_xyz2,
                               float* dist, int32_t* idx, cudaStream_t st) {
  if (!knn_workspace_of_xyz1 || !knn_workspace_of_xyz2) return DH3D_ERR_NULL;
  return three_nn_pruned_impl(b, n, m, nullptr, reinterpret_cast<const float4*>(knn_workspace_of_xyz1), nullptr,
                              reinterpret_cast<const float4*>(knn_workspace_of_xyz2), dist, idx, nullptr, 0, st);
}

}  // namespace dh3d
