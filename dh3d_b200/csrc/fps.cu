// Farthest point sampling (reference: tf_ops/sampling/tf_sampling_g.cu:105-170, <<<32,512>>>).
//
// The reference runs at most 32 CTAs, keeps the running min-distance array `temp` in GLOBAL
// memory and pays 2 x 9 __syncthreads per round for a shared-memory tree argmax.  Here 1024 threads
// own one cloud (all B clouds run concurrently; a 4-CTA cluster of 256 threads each, see below), the
// points and their running min-distance live in REGISTERS (8 points/thread at n = 8192) and the argmax
// is one REDUX.MAX + ballot per warp.
//
// Bit-exact parity with the reference's selection order.  The reference picks, among the points
// with maximal d2, the one with the smallest (k mod 512, k): thread tid scans k=tid,tid+512,..
// with a strict `>` (:130-149) and the tree keeps the lower slot on ties (:158).  We give thread
// t the points whose tie rank  tk(k) = (k mod 512)*V + (k div 512),  V = ceil(n/512),  lies in
// [t*PPT, (t+1)*PPT), scan them in increasing tk with a strict `>`, and break cross-thread ties
// towards the lowest thread -- the same total order.  d = fma(dz,dz,fma(dx,dx,dy*dy)) is the
// contraction nvcc applies to :142; d2 = fminf(d, temp).
#include <stdlib.h>

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

// ---- 4-CTA cluster kernel (n <= 8192) ---------------------------------------------------------
// A single 1024-thread CTA per cloud (the first version of this file, 0.71 ms) is ISSUE-bound, not latency-bound: 32 warps x ~105 instructions per
// round on one SM's four schedulers = ~840 of the ~1375 cycles a round takes (ncu r1h), while 116
// of the 148 SMs idle at B = 32.  Here a cluster of 4 CTAs x 256 threads owns one cloud: the same
// 32 warps, two per scheduler over four SMs.  There is NO barrier inside the round loop: lane r of
// every warp stores the warp's (d2 bits | round tag | tie rank) as one 8-byte word into slot
// [round parity][global warp] of CTA r's shared memory (st.shared::cluster -- distributed shared
// memory), and every warp spins on its own CTA's 32 slots until all carry this round's tag, then
// reduces them with one REDUX.MAX + ballot.  A slot of parity p is rewritten two rounds later,
// which its writer can only reach after receiving every warp's next-round word, i.e. after every
// reader is done with it.  Same selection order as above (global thread g = rank*256 + tid owns
// tie ranks [g*PPT, (g+1)*PPT); lower slot = lower threads wins ties).
constexpr int kFpsClusterCtas = 4;
constexpr int kFpsClusterThreads = 256;
constexpr int kFpsClusterWarps = kFpsClusterThreads / 32;           // per CTA
constexpr int kFpsClusterSlots = kFpsClusterCtas * kFpsClusterWarps;  // 32 = one per lane

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_b64(uint32_t addr, unsigned long long v) {
  asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_poll_b64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.relaxed.cluster.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int PPT>
__global__ void __launch_bounds__(kFpsClusterThreads, 1)
fps_cluster_kernel(int n, int m, const float* __restrict__ dataset, int32_t* __restrict__ idxs, int xyz_in_smem) {
  // xyz_in_smem: each CTA keeps the whole cloud (n*3 floats, 96 KB at n = 8192) to read the coordinates of the
  // point chosen in the previous round; otherwise they are read through L1 (__ldg) and the CTA needs no dynamic
  // shared memory, so it can share an SM with the ~200 KB persistent kernels of the main stream (DH3D_FPS_XYZ)
  extern __shared__ __align__(16) float s_xyz[];
  __shared__ __align__(8) unsigned long long s_slot[2][kFpsClusterSlots];

  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / kFpsClusterCtas;
  const float* ds = dataset + (long long)b * n * 3;
  int32_t* out = idxs + (long long)b * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = (int)rank * kFpsClusterThreads + tid;  // thread id within the cloud

  if (xyz_in_smem)
    for (int j = tid; j < n * 3; j += kFpsClusterThreads) s_xyz[j] = ds[j];
  if (tid < 2 * kFpsClusterSlots) (&s_slot[0][0])[tid] = 0ull;  // tag 0 = no round
  __syncthreads();
  cluster_sync_all();  // every CTA's slots are initialised before any remote store lands

  const int V = (n + 511) >> 9;
  const unsigned magic = 0xffffffffu / (unsigned)V + 1u;
  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int tk = g * PPT + i;
    const int k = (tk % V) * 512 + tk / V;
    const bool valid = (tk < 512 * V) && (k < n);
    px[i] = valid ? __ldg(ds + k * 3 + 0) : 0.f;
    py[i] = valid ? __ldg(ds + k * 3 + 1) : 0.f;
    pz[i] = valid ? __ldg(ds + k * 3 + 2) : 0.f;
    td[i] = valid ? 1e38f : -1.f;
  }

  unsigned long long qx[PPT >= 2 ? PPT / 2 : 1], qy[PPT >= 2 ? PPT / 2 : 1], qz[PPT >= 2 ? PPT / 2 : 1];
  if constexpr (PPT >= 2) {
#pragma unroll
    for (int i = 0; i < PPT / 2; ++i) {
      qx[i] = pack2(px[2 * i], px[2 * i + 1]);
      qy[i] = pack2(py[2 * i], py[2 * i + 1]);
      qz[i] = pack2(pz[2 * i], pz[2 * i + 1]);
    }
  }

  // lane r < 4 sends to CTA r; slot index = global warp id
  const uint32_t slot0 = smem_u32(&s_slot[0][0]);
  const uint32_t my_slot_off = (uint32_t)((int)rank * kFpsClusterWarps + warp) * 8u;
  const uint32_t remote0 = mapa_shared(slot0 + my_slot_off, (uint32_t)(lane & (kFpsClusterCtas - 1)));
  const uint32_t poll0 = slot0 + (uint32_t)lane * 8u;

  int old = 0;
  if (g == 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    float x1, y1, z1;
    if (xyz_in_smem) {
      x1 = s_xyz[old * 3 + 0]; y1 = s_xyz[old * 3 + 1]; z1 = s_xyz[old * 3 + 2];
    } else {
      x1 = __ldg(ds + old * 3 + 0); y1 = __ldg(ds + old * 3 + 1); z1 = __ldg(ds + old * 3 + 2);
    }
    float best = -1.f;
    int bslot = 0;
    if constexpr (PPT >= 2) {
      // two points per instruction (sm_100 FADD2 / FMUL2 / FFMA2): same IEEE operations in the same order as
      // the scalar form below, 6 instead of 12 FP instructions per pair -- the rounds of this kernel share the
      // SM's issue slots with the k-NN scan of the main stream
#pragma unroll
      for (int i = 0; i < PPT / 2; ++i) {
        const unsigned long long dx = fsub2s(qx[i], x1), dy = fsub2s(qy[i], y1), dz = fsub2s(qz[i], z1);
        const float2 d = unpack2(ffma2v(dz, dz, ffma2v(dx, dx, fmul2(dy, dy))));
        const float da = fminf(d.x, td[2 * i]), db = fminf(d.y, td[2 * i + 1]);
        td[2 * i] = da;
        td[2 * i + 1] = db;
        if (da > best) { best = da; bslot = 2 * i; }
        if (db > best) { best = db; bslot = 2 * i + 1; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        const float dx = px[i] - x1, dy = py[i] - y1, dz = pz[i] - z1;
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        const float d2 = fminf(d, td[i]);
        td[i] = d2;
        if (d2 > best) { best = d2; bslot = i; }
      }
    }
    const int bits = __float_as_int(best);
    const int wmax = __reduce_max_sync(0xffffffffu, bits);
    const unsigned mk = __ballot_sync(0xffffffffu, bits == wmax);
    const int wi = __shfl_sync(0xffffffffu, g * PPT + bslot, __ffs(mk) - 1);
    const uint32_t tag = (uint32_t)j & 0xffffu;
    const uint32_t par = ((uint32_t)j & 1u) * (kFpsClusterSlots * 8u);
    if (lane < kFpsClusterCtas)
      st_cluster_b64(remote0 + par,
                     ((unsigned long long)(uint32_t)wmax << 32) | (tag << 16) | (uint32_t)wi);
    unsigned long long v;
    int spins = 0;
    do {
      v = ld_poll_b64(poll0 + par);
      if (++spins > (1 << 24)) __trap();  // a lost peer becomes an error, never a hang
    } while (!__all_sync(0xffffffffu, (((uint32_t)v >> 16) & 0xffffu) == tag));
    const int hv = (int)(uint32_t)(v >> 32);
    const int bmax = __reduce_max_sync(0xffffffffu, hv);
    const unsigned m2 = __ballot_sync(0xffffffffu, hv == bmax);
    const unsigned tk = (unsigned)__shfl_sync(0xffffffffu, (int)((uint32_t)v & 0xffffu), __ffs(m2) - 1);
    const unsigned q = (V == 1) ? tk : __umulhi(tk, magic);
    old = (int)((tk - q * (unsigned)V) * 512u + q);
    if (g == 0) out[j] = old;
  }
  cluster_sync_all();  // no CTA leaves while a peer could still store into its shared memory
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

// ---- box-pruned kernel on a pre-sorted cloud (n <= 8192) --------------------------------------------------------
// The exhaustive kernels update the running min-distance of ALL n points in every one of the m - 1 rounds, although a
// point's value only changes if the newly chosen point is closer to it than every earlier one -- after a few dozen
// rounds that is a small neighbourhood of the new point.  This kernel works on the cell-sorted copy of the cloud that
// the k-NN of the same points leaves in its workspace (knn.cu: float4 (x,y,z,index) along a Hilbert curve + one
// bounding box per 32-point chunk): each lane owns one chunk's box and the chunk's exact current maximum, and a round
//   1. bounds the distance from the new point to each box with the metric's own operation sequence (monotone rounding:
//      the bound never exceeds the computed distance of a point inside) -- a chunk whose bound is not below its maximum
//      cannot change and is skipped;
//   2. re-reduces only the remaining chunks (one point per lane, REDUX.MAX on the distance bits + REDUX.MIN on the tie
//      rank), chunks dealt round-robin to the 16 warps so that a spatial neighbourhood spreads over all of them;
//   3. takes the arg-max over the 256 chunk maxima (warp REDUX, 16 shared-memory slots, ONE barrier per round).
// Same result as the exhaustive scan, point for point: identical distance arithmetic on identical floats, and the
// reference's tie order (smallest (k mod 512, k) among maximal d2, see the header of this file) carried as the rank tk.
// One 512-thread CTA per cloud on 32 SMs instead of a 4-CTA cluster on 128.  Alone it is no faster than the exhaustive
// kernel (0.38 vs 0.42 ms: a round revisits ~17 of 256 chunks, ~4 of them on the busiest warp, and pays one barrier and
// six REDUX per round), but the k-NN that runs next to it keeps the other 116 SMs to itself.
constexpr int kFpsBThreads = 512;    // 16 warps: 0.381 ms for 32 x 8192 -> 1024 (8 warps 0.399, 32 warps 0.446: barrier + reductions)
constexpr int kFpsBChunk = 32;      // == kKnnChunk
constexpr int kFpsBSuper = 16;      // == kKnnSuper

__global__ void __launch_bounds__(kFpsBThreads, 1)
fps_bucket_kernel(int n, int np, int m, const float4* __restrict__ sorted, const float4* __restrict__ boxes,
                  long long box_stride, int32_t* __restrict__ idxs) {
  extern __shared__ __align__(16) unsigned char fb_smem[];
  float* sx = reinterpret_cast<float*>(fb_smem);
  float* sy = sx + np;
  float* sz = sy + np;
  float* std_ = sz + np;
  unsigned short* stk = reinterpret_cast<unsigned short*>(std_ + np);
  __shared__ unsigned long long s_slot[2][kFpsBThreads / 32];
  __shared__ int s_first;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* pts = sorted + (long long)b * np;
  const float4* bx = boxes + (long long)b * box_stride;
  int32_t* out = idxs + (long long)b * m;
  const int V = (n + 511) >> 9;
  const unsigned magic = 0xffffffffu / (unsigned)V + 1u;
  const int nchunks = np / kFpsBChunk;

  for (int i = tid; i < np; i += kFpsBThreads) {
    const float4 q = __ldg(pts + i);
    const int k = __float_as_int(q.w);
    const bool valid = k >= 0;
    sx[i] = q.x; sy[i] = q.y; sz[i] = q.z;
    std_[i] = valid ? 1e38f : -1.f;
    stk[i] = valid ? (unsigned short)((k & 511) * V + (k >> 9)) : (unsigned short)0xFFFF;
    if (k == 0) s_first = i;
  }
  // this lane's chunk (dealt round-robin: chunk c belongs to warp c % 16, lane c / 16 -- lanes 0..15 at n = 8192)
  constexpr int NW = kFpsBThreads / 32;
  const int c = lane * NW + warp;
  const bool own = c < nchunks;
  float4 blo = make_float4(0.f, 0.f, 0.f, 0.f), bhi = blo;
  if (own) { blo = __ldg(bx + 2 * c); bhi = __ldg(bx + 2 * c + 1); }
  // exact maximum of the chunk's running min-distances and (tie rank << 13 | sorted position) of its arg-max; before
  // the first update every valid chunk holds the initial 1e38 (any chunk with a finite box is updated in round 1)
  float cmax = (own && blo.x <= bhi.x) ? 1e38f : -1.f;   // an all-padding chunk has the box (inf, -inf)
  unsigned ckey = 0xFFFFFFFFu;
  __syncthreads();

  int pos = s_first;
  if (tid == 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = sx[pos], y1 = sy[pos], z1 = sz[pos];
    // lower bound of d to any point of the own chunk's box, in the metric's operation order
    const float ex = fmaxf(fmaxf(blo.x - x1, x1 - bhi.x), 0.f);
    const float ey = fmaxf(fmaxf(blo.y - y1, y1 - bhi.y), 0.f);
    const float ez = fmaxf(fmaxf(blo.z - z1, z1 - bhi.z), 0.f);
    const float lb = __fmaf_rn(ez, ez, __fmaf_rn(ex, ex, __fmul_rn(ey, ey)));
    // (round 1 visits every chunk that holds a point, whatever the bound: that is where the chunk maxima get their keys;
    //  !(lb >= cmax) rather than lb < cmax so that a NaN bound updates instead of skipping)
    unsigned mask = __ballot_sync(0xffffffffu, own && cmax >= 0.f && (j == 1 || !(lb >= cmax)));
    while (mask) {   // (measured slower: 4 chunks per iteration with interleaved REDUX chains -- most iterations then
                     //  repeat a chunk to fill the slots, 0.46 ms; two chunks per iteration on half-warps with
                     //  half-mask reductions, 0.72 ms)
      const int l = __ffs(mask) - 1;
      mask &= mask - 1u;
      const int i = (l * NW + warp) * kFpsBChunk + lane;
      const float dx = sx[i] - x1, dy = sy[i] - y1, dz = sz[i] - z1;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      const float nd = fminf(d, std_[i]);
      std_[i] = nd;
      const int bits = __float_as_int(nd);      // valid: >= 0 (orders like the float); padding: -1.0f, a negative int
      const int wmax = __reduce_max_sync(0xffffffffu, bits);
      const unsigned key = ((unsigned)stk[i] << 13) | (unsigned)i;
      const unsigned kmin = __reduce_min_sync(0xffffffffu, bits == wmax ? key : 0xFFFFFFFFu);
      if (lane == l) { cmax = __int_as_float(wmax); ckey = kmin; }
    }
    // arg-max over this warp's chunks, then over the warps
    const int cb = __float_as_int(cmax);
    const int wmax = __reduce_max_sync(0xffffffffu, cb);
    const unsigned wkey = __reduce_min_sync(0xffffffffu, cb == wmax ? ckey : 0xFFFFFFFFu);
    if (lane == 0) s_slot[j & 1][warp] = ((unsigned long long)(unsigned)wmax << 32) | wkey;
    __syncthreads();
    const unsigned long long v = s_slot[j & 1][lane & (NW - 1)];
    const int hv = (int)(unsigned)(v >> 32);
    const int gmax = __reduce_max_sync(0xffffffffu, hv);
    const unsigned gkey = __reduce_min_sync(0xffffffffu, hv == gmax ? (unsigned)v : 0xFFFFFFFFu);
    pos = (int)(gkey & 0x1FFFu);
    if (tid == 0) {
      const unsigned tk = gkey >> 13;
      const unsigned q = (V == 1) ? tk : __umulhi(tk, magic);
      out[j] = (int)((tk - q * (unsigned)V) * 512u + q);
    }
  }
}

// sorted / boxes: the workspace a k-NN call on the same [b,n,3] cloud left behind (knn.cu: float4[b][np] sorted points,
// then per cloud 2 * np/32 chunk-box float4s + the super-chunk boxes)
int fps_presorted_launch(int b, int n, int m, const void* knn_workspace, int32_t* out, cudaStream_t st) {
  if (!knn_workspace || !out) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || m < 0) return DH3D_ERR_DIM;
  if (m == 0) return DH3D_OK;
  if (n > 8192 || b > 65535) return DH3D_ERR_UNSUPPORTED;
  if (((uintptr_t)knn_workspace & 15) != 0) return DH3D_ERR_ALIGN;
  const int np = ceil_div(n, kFpsBChunk) * kFpsBChunk;
  const int nchunks = np / kFpsBChunk;
  const long long box_stride = 2LL * nchunks + 2LL * ceil_div(nchunks, kFpsBSuper);
  const float4* sorted = reinterpret_cast<const float4*>(knn_workspace);
  const float4* boxes = sorted + (size_t)b * np;
  const size_t smem = (size_t)np * (4 * sizeof(float) + sizeof(unsigned short));
  cudaError_t e = cudaFuncSetAttribute(fps_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  fps_bucket_kernel<<<b, kFpsBThreads, smem, st>>>(n, np, m, sorted, boxes, box_stride, out);
  return launch_status();
}

template <int PPT>
static int fps_launch_cluster(int b, int n, int m, const float* inp, int32_t* out, cudaStream_t st) {
  // the cloud's coordinates live in each CTA's shared memory (reading them through L1 instead, so that other
  // CTAs could share the SM, measured slower: 0.59 vs 0.45 ms in the step, r1v)
  const int xyz_in_smem = 1;
  size_t smem = xyz_in_smem ? (size_t)n * 3 * sizeof(float) : 0;
  cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<PPT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * kFpsClusterCtas));
  cfg.blockDim = dim3(kFpsClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kFpsClusterCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<PPT>, n, m, inp, out, xyz_in_smem);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  return launch_status();
}

int fps_launch(int b, int n, int m, const float* inp, int32_t* out, cudaStream_t st) {
  if (!inp || !out) return DH3D_ERR_NULL;
  if (b <= 0 || n <= 0 || m < 0) return DH3D_ERR_DIM;
  if (m == 0) return DH3D_OK;  // reference kernel returns immediately (tf_sampling_g.cu:106-107)
  if (n > 65536) return DH3D_ERR_UNSUPPORTED;
  const int total = 512 * ((n + 511) / 512);
  const int ppt = ceil_div(total, kFpsThreads);
  if (ppt <= 8) {  // n <= 8192: a 4-CTA cluster of 256 threads per cloud
    if (ppt <= 1) return fps_launch_cluster<1>(b, n, m, inp, out, st);
    if (ppt <= 2) return fps_launch_cluster<2>(b, n, m, inp, out, st);
    if (ppt <= 4) return fps_launch_cluster<4>(b, n, m, inp, out, st);
    return fps_launch_cluster<8>(b, n, m, inp, out, st);
  }
  // n > 8192 (outside DH3D's shapes): one CTA per cloud, min-distances in shared memory
  size_t smem = (size_t)total * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  fps_smem_kernel<<<b, kFpsThreads, smem, st>>>(n, m, inp, out);
  return launch_status();
}

}  // namespace dh3d
