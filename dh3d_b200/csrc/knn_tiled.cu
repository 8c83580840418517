// Exhaustive tiled k-NN scan (the round-1 first kernel; kept as the DH3D_KNN=tiled A/B path and as the
// fallback form of knn.cu's pruned search).  Reference: user_ops/kernels/knn_bruteforce_kernel_gpu.cu.cc.
//
// The reference launches one CTA per QUERY point and radix-sorts all N keys to extract K of
// them.  Here one thread owns one query and keeps a sorted top-K list in registers; the cloud is
// streamed through shared memory in 16 KB tiles (1-D bulk TMA, double buffered) as packed
// float4 candidates, so every LDS.128 is a warp-wide broadcast and the inner loop is 3 FADD,
// 1 FMUL, 2 FFMA, 1 compare per pair.
//
// Bit-exact parity with the reference's output order:
//   key   = sqrt.rn(fma(dz,dz,fma(dy,dy,dx*dx)))   (kernel_gpu.cu.cc:102-107, nvcc-contracted)
//   ties  = cub::BlockRadixSort blocked-order stability: point x sits at rank
//           s(x) = (x mod T)*V + (x div T)   (T,V picked by N, :181-216).
// The pack kernel lays the candidates out in rank order, the scan visits them in that order and
// inserts only on strictly-smaller keys, so ties come out in rank order without carrying the
// rank.  The compare runs on d^2 against a conservative bound (every d^2 whose sqrt.rn could be
// below the current k-th key passes); the exact sqrt.rn compare happens only on that rare path.
#include "common.cuh"

namespace dh3d {

constexpr int kTKnnTile = 1024;     // candidates per smem tile (16 KB as float4)
constexpr int kTKnnThreads = 128;   // queries per CTA

struct TKnnOrder { int T; int logV; int P; };  // P = T*V rounded up to a tile multiple

static TKnnOrder tknn_order(int N) {
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
  else { T = N; V = 1; }
  TKnnOrder o;
  o.T = T;
  o.logV = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : 3;
  o.P = ceil_div(T * V, kTKnnTile) * kTKnnTile;
  return o;
}

// rank position p -> point index x
__device__ __forceinline__ int tknn_rank_to_point(int p, int T, int logV) {
  return (p & ((1 << logV) - 1)) * T + (p >> logV);
}

// positions (any layout via strides) -> rank-ordered float4 candidates; invalid lanes get +inf.
__global__ void tknn_pack_kernel(const float* __restrict__ pos, int N, long long sb, int sp, int sd,
                                int T, int logV, int P, float4* __restrict__ packed) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int b = blockIdx.y;
  if (p >= P) return;
  int x = tknn_rank_to_point(p, T, logV);
  float4 v;
  if (x < N && p < (T << logV)) {
    const float* q = pos + (long long)b * sb + (long long)x * sp;
    v = make_float4(q[0], q[sd], q[2 * sd], 0.f);
  } else {
    v = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
  }
  packed[(long long)b * P + p] = v;
}

// KC = compiled list length (>= K).  EXACT: K == KC and outputs are 16-byte aligned.
template <int KC, bool EXACT>
__global__ void __launch_bounds__(kTKnnThreads)
tknn_scan_kernel(const float4* __restrict__ packed, const float* __restrict__ pos, int N,
                long long sb, int sp, int sd, int T, int logV, int P, int K,
                int32_t* __restrict__ ids, float* __restrict__ dists) {
  __shared__ __align__(128) float4 s_tile[2][kTKnnTile];
  __shared__ __align__(8) uint64_t s_full[2];

  const int b = blockIdx.y;
  const int y = blockIdx.x * kTKnnThreads + threadIdx.x;
  const bool active = y < N;
  const float4* cloud = packed + (long long)b * P;
  const int ntiles = P / kTKnnTile;
  constexpr uint32_t kTileBytes = kTKnnTile * sizeof(float4);

  if (threadIdx.x == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&s_full[0], kTileBytes);
    tma_load_1d(&s_tile[0][0], cloud, kTileBytes, &s_full[0]);
    if (ntiles > 1) {
      mbar_arrive_expect_tx(&s_full[1], kTileBytes);
      tma_load_1d(&s_tile[1][0], cloud + kTKnnTile, kTileBytes, &s_full[1]);
    }
  }

  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float* q = pos + (long long)b * sb + (long long)y * sp;
    qx = q[0]; qy = q[sd]; qz = q[2 * sd];
  }

  float sq[KC];
  int id[KC];
#pragma unroll
  for (int j = 0; j < KC; ++j) { sq[j] = 3.402823466e+38f; id[j] = -1; }  // reference padding lanes
  float kth = 3.402823466e+38f;   // current K-th key (sq[K-1])
  float thr2 = CUDART_INF_F;      // d^2 bound implied by kth

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    mbar_wait(&s_full[buf], (t >> 1) & 1);
    const float4* tile = s_tile[buf];
    const int base = t * kTKnnTile;
#pragma unroll 8
    for (int c = 0; c < kTKnnTile; ++c) {
      const float4 p = tile[c];
      const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
      const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
      if (d2 <= thr2) {
        const float s = __fsqrt_rn(d2);
        if (s < kth) {
          const int x = tknn_rank_to_point(base + c, T, logV);
#pragma unroll
          for (int j = KC - 1; j > 0; --j) {
            if (s < sq[j - 1]) { sq[j] = sq[j - 1]; id[j] = id[j - 1]; }
            else if (s < sq[j]) { sq[j] = s; id[j] = x; }
          }
          if (s < sq[0]) { sq[0] = s; id[0] = x; }
          if constexpr (EXACT) {
            kth = sq[KC - 1];
          } else {
            kth = sq[0];
#pragma unroll
            for (int j = 1; j < KC; ++j) kth = (j < K) ? sq[j] : kth;
          }
          // every d2 with sqrt.rn(d2) < kth satisfies d2 < kth^2 (exact) <= the bound below
          thr2 = __fmul_rn(__fmul_rn(kth, kth), 1.000001f);
        }
      }
    }
    __syncthreads();  // everyone is done reading s_tile[buf]
    if (threadIdx.x == 0 && t + 2 < ntiles) {
      mbar_arrive_expect_tx(&s_full[buf], kTileBytes);
      tma_load_1d(&s_tile[buf][0], cloud + (long long)(t + 2) * kTKnnTile, kTileBytes,
                  &s_full[buf]);
    }
  }

  if (active) {
    int32_t* oi = ids + ((long long)b * N + y) * K;
    float* od = dists + ((long long)b * N + y) * K;
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

size_t knn_tiled_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * tknn_order(N).P * sizeof(float4);
}

int knn_tiled_launch(const float* pos, int B, int N, int K, long long sb, int sp, int sd, int32_t* ids,
               float* dists, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!pos || !ids || !dists) return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0 || K <= 0) return DH3D_ERR_DIM;
  if (K > 32 || N > 65536 || B > 65535) return DH3D_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < knn_tiled_workspace_bytes(B, N)) return DH3D_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 127) != 0) return DH3D_ERR_ALIGN;
  TKnnOrder o = tknn_order(N);
  float4* packed = reinterpret_cast<float4*>(workspace);
  tknn_pack_kernel<<<dim3(ceil_div(o.P, 256), B), 256, 0, st>>>(pos, N, sb, sp, sd, o.T, o.logV, o.P,
                                                              packed);
  dim3 grid(ceil_div(N, kTKnnThreads), B);
  const bool vec_ok = (((uintptr_t)ids | (uintptr_t)dists) & 15) == 0;
#define DH3D_KNN(KC, EX)                                                                       \
  tknn_scan_kernel<KC, EX><<<grid, kTKnnThreads, 0, st>>>(packed, pos, N, sb, sp, sd, o.T, o.logV, \
                                                        o.P, K, ids, dists)
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
