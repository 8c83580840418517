// Attention-weighted NetVLAD (reference: core/backbones.py:202-279 global_netvald_block +
// :282-320 context_gating; ~15 TF library ops that write/read the [B*N,64] assignment ~6x and
// the [B*N,256] features 3x).
//
// Here the per-point work is ONE kernel that reads each feature row once and never writes a
// per-point tensor:  l2-normalise row -> 256x64 assignment (+folded cluster BN) -> softmax(64)
// -> x attention -> accumulate  V[d,k] += a[n,k]*x[n,d]  and  S[k] += a[n,k]  in registers.
// Each CTA reduces a slab of points; slabs are combined deterministically (no float atomics) by
// the small finalize kernel, which also does  V - S*W2, the per-cluster and global l2 norms and
// the feature-major flatten.  The 16384x256 projection is a split-K GEMV-like kernel (weights are
// read exactly once per batch), and the head kernel applies BN, context gating and the final
// l2-normalise.
//   D == 256 features, Kc == 64 clusters, out_dim == 256 (the shipped DH3D configuration).
#include <stdlib.h>

#include "common.cuh"

namespace dh3d {

constexpr int kVD = 256;    // feature dim == threads per CTA
constexpr int kVK = 64;     // clusters
constexpr int kVTP = 32;    // points per sub-tile
constexpr int kVXS = kVD + 4;  // padded row stride of the x tile (floats)
constexpr int kVMaxSlabs = 32; // upper bound of CTAs (slabs of points) per cloud
constexpr int kVSlice = 64;    // rows of hidden1_weights per projection CTA (256 CTAs: two per SM hide the weight-load latency)

struct __align__(16) VladSmem {
  float w[kVD][kVK];        // cluster_weights           64 KB
  float x[kVTP][kVXS];      // normalised feature tile    32.5 KB
  float a[kVTP][kVK];       // attention-weighted assign   8 KB
};

__global__ void __launch_bounds__(kVD)
netvlad_aggregate_kernel(const float* __restrict__ feat, const float* __restrict__ att, int N,
                         int slabs, const float* __restrict__ cw, const float* __restrict__ bn_scale,
                         const float* __restrict__ bn_shift, float* __restrict__ part_v,
                         float* __restrict__ part_s, unsigned int* zero_word) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VladSmem& sm = *reinterpret_cast<VladSmem*>(smem_raw);
  if (zero_word && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)   // the tail kernel's four counters
    *reinterpret_cast<uint4*>(zero_word) = make_uint4(0u, 0u, 0u, 0u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int per = (N + slabs - 1) / slabs;
  const int n_begin = slab * per;
  const int n_end = min(N, n_begin + per);

  for (int i = tid; i < kVD * kVK / 4; i += kVD)
    reinterpret_cast<float4*>(&sm.w[0][0])[i] = ldg4(cw + i * 4);

  float acc[kVK];
#pragma unroll
  for (int k = 0; k < kVK; ++k) acc[k] = 0.f;
  float ssum = 0.f;  // threads 0..63: S[k]

  // assignment GEMM thread mapping: rows 2*ty, 2*ty+1; cols 4*tx .. 4*tx+3
  const int tx = tid & 15, ty = tid >> 4;
  float4 bsc = ldg4(bn_scale + tx * 4), bsh = ldg4(bn_shift + tx * 4);

  for (int n0 = n_begin; n0 < n_end; n0 += kVTP) {
    const int cnt = min(kVTP, n_end - n0);
    __syncthreads();  // previous tile fully consumed (also covers the sm.w fill on the first pass)
    // (a) load + l2-normalise rows: warp w owns rows w*4 .. w*4+3
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = warp * 4 + rr;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (r < cnt) {
        const float* src = feat + ((long long)b * N + n0 + r) * kVD;
        v0 = ldg4(src + lane * 4);
        v1 = ldg4(src + 128 + lane * 4);
      }
      float ss = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y +
                 v1.z * v1.z + v1.w * v1.w;
      ss = warp_sum(ss);
      const float inv = rsqrtf(fmaxf(ss, 1e-12f));  // tf.nn.l2_normalize default epsilon
      v0.x *= inv; v0.y *= inv; v0.z *= inv; v0.w *= inv;
      v1.x *= inv; v1.y *= inv; v1.z *= inv; v1.w *= inv;
      *reinterpret_cast<float4*>(&sm.x[r][lane * 4]) = v0;
      *reinterpret_cast<float4*>(&sm.x[r][128 + lane * 4]) = v1;
    }
    __syncthreads();
    // (b) assignment logits for 2 rows x 4 clusters per thread
    float l0[4] = {0.f, 0.f, 0.f, 0.f}, l1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int d = 0; d < kVD; d += 4) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sm.x[2 * ty][d]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sm.x[2 * ty + 1][d]);
      const float a0[4] = {x0.x, x0.y, x0.z, x0.w}, a1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 w4 = *reinterpret_cast<const float4*>(&sm.w[d + e][tx * 4]);
        l0[0] = fmaf(a0[e], w4.x, l0[0]); l0[1] = fmaf(a0[e], w4.y, l0[1]);
        l0[2] = fmaf(a0[e], w4.z, l0[2]); l0[3] = fmaf(a0[e], w4.w, l0[3]);
        l1[0] = fmaf(a1[e], w4.x, l1[0]); l1[1] = fmaf(a1[e], w4.y, l1[1]);
        l1[2] = fmaf(a1[e], w4.z, l1[2]); l1[3] = fmaf(a1[e], w4.w, l1[3]);
      }
    }
    // (c) folded cluster BN, softmax over the 64 clusters (16 lanes x 4), x attention
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* l = h ? l1 : l0;
      l[0] = fmaf(l[0], bsc.x, bsh.x); l[1] = fmaf(l[1], bsc.y, bsh.y);
      l[2] = fmaf(l[2], bsc.z, bsh.z); l[3] = fmaf(l[3], bsc.w, bsh.w);
      float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float e0 = __expf(l[0] - mx), e1 = __expf(l[1] - mx), e2 = __expf(l[2] - mx),
            e3 = __expf(l[3] - mx);
      float sum = (e0 + e1) + (e2 + e3);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const int r = 2 * ty + h;
      float scale = 0.f;
      if (r < cnt) scale = __ldg(att + (long long)b * N + n0 + r) / sum;
      *reinterpret_cast<float4*>(&sm.a[r][tx * 4]) =
          make_float4(e0 * scale, e1 * scale, e2 * scale, e3 * scale);
    }
    __syncthreads();
    // (d) V[d=tid, :] += a[n, :] * x[n, d];  S[k] += a[n, k]
    for (int n = 0; n < cnt; ++n) {
      const float xv = sm.x[n][tid];
#pragma unroll
      for (int k = 0; k < kVK; k += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sm.a[n][k]);
        acc[k] = fmaf(a4.x, xv, acc[k]); acc[k + 1] = fmaf(a4.y, xv, acc[k + 1]);
        acc[k + 2] = fmaf(a4.z, xv, acc[k + 2]); acc[k + 3] = fmaf(a4.w, xv, acc[k + 3]);
      }
      if (tid < kVK) ssum += sm.a[n][tid];
    }
  }

  float* pv = part_v + (((long long)b * slabs + slab) * kVD + tid) * kVK;
#pragma unroll
  for (int k = 0; k < kVK; k += 4)
    *reinterpret_cast<float4*>(pv + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
  if (tid < kVK) part_s[((long long)b * slabs + slab) * kVK + tid] = ssum;
}

// ---- tail: slab combine + intra-norm -> 16384 x 256 projection -> BN / context gating / l2-norm, ONE launch ------------
// Three phases in ONE launch (was three launches: 17 + 31 + 33 us for 32 clouds, each mostly launch ramp and load
// latency).  One CTA per work item; a CTA takes a TICKET (atomic counter) when it starts and the ticket names its item:
// tickets [0, nA) are phase-A items, [nA, nA + nB) phase B, the rest phase C.  A later phase waits on a completion
// counter of the phase before it.  A CTA therefore only ever waits for CTAs with LOWER tickets, i.e. CTAs that have
// already started (resident or finished) and that wait, in turn, only for still lower tickets: progress never depends on
// a CTA that has not been scheduled yet.  That holds with any number of these launches (or anything else) running
// next to each other on other streams, unlike a spinning grid barrier, which deadlocks as soon as two half-resident
// grids hold each other's SM slots.  Data written in one phase is read in the next with plain (coherent) loads.
// Counters (workspace, zeroed by the aggregate kernel one launch ahead): [0] ticket, [1] phase-A items done,
// [2] phase-B items done, [3] phase-C items done.
constexpr int kVQ = 8;     // clusters per phase-A work item: B * 8 items

__device__ __forceinline__ void nv_signal(unsigned int* ctr) {   // whole CTA: its item's stores are done
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
  }
}

__device__ __forceinline__ void nv_wait(const unsigned int* ctr, unsigned int target) {   // whole CTA
  if (threadIdx.x == 0) {
    unsigned int v;
    do {   // relaxed polls (an acquire load costs a cache invalidation per poll), one fence once the count is in
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// phase A, work item (b, q): clusters 8q .. 8q+7 of cloud b; thread = feature d.  Combines the slabs, subtracts
// S*W2, intra-normalises each cluster column (over the 256 features, inside the CTA) and writes the feature-major
// flattening ([d*64 + k], backbones.py:258-260).  The global l2 norm of the flattened vector only needs the 64
// column norms: they go to coln[b][k] and the scalar is applied after the (linear) projection, in phase C.
__device__ __forceinline__ void nv_finalize_item(int b, int q, const float* __restrict__ part_v,
                                                 const float* __restrict__ part_s, int slabs,
                                                 const float* __restrict__ cw2, float* vlad, float* coln) {
  __shared__ float s_sum[kVQ];
  __shared__ float s_red[kVD / 32][kVQ];
  __shared__ float s_inv[kVQ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k0 = q * kVQ;
  __syncthreads();   // the previous item's readers of the shared arrays are done
  if (tid < kVQ) {
    float s = 0.f;
    for (int c = 0; c < slabs; ++c) s += part_s[((long long)b * slabs + c) * kVK + k0 + tid];
    s_sum[tid] = s;
  }
  float v[kVQ];
#pragma unroll
  for (int k = 0; k < kVQ; ++k) v[k] = 0.f;
#pragma unroll 4
  for (int c = 0; c < slabs; ++c) {
    const float* pv = part_v + (((long long)b * slabs + c) * kVD + tid) * kVK + k0;
#pragma unroll
    for (int k = 0; k < kVQ; k += 4) {
      const float4 t = ldg4(pv + k);
      v[k] += t.x; v[k + 1] += t.y; v[k + 2] += t.z; v[k + 3] += t.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kVQ; ++k) {
    v[k] -= s_sum[k] * __ldg(cw2 + tid * kVK + k0 + k);
    const float ss = warp_sum(v[k] * v[k]);
    if (lane == 0) s_red[warp][k] = ss;
  }
  __syncthreads();
  if (tid < kVQ) {
    float ss = 0.f;
#pragma unroll
    for (int w = 0; w < kVD / 32; ++w) ss += s_red[w][tid];
    const float inv = rsqrtf(fmaxf(ss, 1e-12f));
    s_inv[tid] = inv;
    coln[(long long)b * kVK + k0 + tid] = ss * inv * inv;  // squared norm of the normalised cluster column
  }
  __syncthreads();
  float* o = vlad + ((long long)b * kVD + tid) * kVK + k0;
#pragma unroll
  for (int k = 0; k < kVQ; k += 4)
    *reinterpret_cast<float4*>(o + k) =
        make_float4(v[k] * s_inv[k], v[k + 1] * s_inv[k + 1], v[k + 2] * s_inv[k + 2], v[k + 3] * s_inv[k + 3]);
}

// phase B, work item (slice, b0): split-K projection, rows [slice*64, slice*64+64) of hidden1_weights [16384, 256]
// for the clouds b0 .. b0+31; thread = output column.  part_h [slices][B][256].
__device__ __forceinline__ void nv_project_item(int slice, int b0, const float* vlad, const float* __restrict__ hw,
                                                int B, int KD, float* part_h, const unsigned int* done_a,
                                                unsigned int n_a) {
  __shared__ __align__(16) float s_x[32][kVSlice];
  const int tid = threadIdx.x;
  const int nb = min(32, B - b0);
  // the weight stream is what this phase waits for (64 KB per item, read once per batch): the first 16 coalesced row
  // loads are issued before the activations are staged
  const float* w = hw + ((long long)slice * kVSlice) * kVD + tid;
  float wv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) wv[j] = __ldg(w + (long long)j * kVD);
  nv_wait(done_a, n_a);   // every phase-A item (the vlad rows of all clouds) is written; the weight loads are in flight
  for (int i = tid; i < 32 * kVSlice; i += kVD) {
    const int bb = i / kVSlice, r = i % kVSlice;
    s_x[bb][r] = bb < nb ? vlad[(long long)(b0 + bb) * KD + (long long)slice * kVSlice + r] : 0.f;   // written in phase A
  }
  __syncthreads();
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int r0 = 0; r0 < kVSlice; r0 += 16) {
    float wn[16];
    if (r0 + 16 < kVSlice) {
#pragma unroll
      for (int j = 0; j < 16; ++j) wn[j] = __ldg(w + (long long)(r0 + 16 + j) * kVD);
    }
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(&s_x[i][r0 + j]);   // broadcast LDS.128 feeds four FMAs
        acc[i] = fmaf(x.x, wv[j], acc[i]);
        acc[i] = fmaf(x.y, wv[j + 1], acc[i]);
        acc[i] = fmaf(x.z, wv[j + 2], acc[i]);
        acc[i] = fmaf(x.w, wv[j + 3], acc[i]);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) wv[j] = wn[j];
  }
  for (int i = 0; i < nb; ++i)
    part_h[((long long)slice * B + b0 + i) * kVD + tid] = acc[i];
}

// phase C, work item b: sum the slices, apply the global l2 norm of the flattened VLAD (from the column norms) ->
// BN -> context gating (256x256 matvec, BN, sigmoid) -> optional final l2-normalise (core/model.py:205, eps 1e-8).
__device__ __forceinline__ void nv_head_item(int b, const float* part_h, int slices, int B, const float* coln,
                                             const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                             const float* __restrict__ gw, const float* __restrict__ g_scale,
                                             const float* __restrict__ g_shift, int final_l2norm,
                                             float* __restrict__ out) {
  __shared__ float s_h[kVD];
  __shared__ float s_red[kVD / 32];
  __shared__ float s_g[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (tid < kVK) {
    const float t = warp_sum(coln[(long long)b * kVK + tid]);
    if (lane == 0) s_g[warp] = t;
  }
  float h4[4] = {0.f, 0.f, 0.f, 0.f};   // (four chains: the 256 partial rows are what this phase waits for)
#pragma unroll 16
  for (int s = 0; s < slices; s += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (s + u < slices) h4[u] += part_h[((long long)(s + u) * B + b) * kVD + tid];
  }
  float h = (h4[0] + h4[1]) + (h4[2] + h4[3]);
  __syncthreads();
  h *= rsqrtf(fmaxf(s_g[0] + s_g[1], 1e-12f));
  h = fmaf(h, __ldg(bn_scale + tid), __ldg(bn_shift + tid));
  s_h[tid] = h;
  __syncthreads();
  float g4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 16
  for (int i = 0; i < kVD; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) g4[u] = fmaf(s_h[i + u], __ldg(gw + (i + u) * kVD + tid), g4[u]);
  }
  float g = (g4[0] + g4[1]) + (g4[2] + g4[3]);
  g = fmaf(g, __ldg(g_scale + tid), __ldg(g_shift + tid));
  float y = h * (1.f / (1.f + __expf(-g)));
  if (final_l2norm) {
    const float ss = warp_sum(y * y);
    if (lane == 0) s_red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kVD / 32; ++w) tot += s_red[w];
    y *= rsqrtf(fmaxf(tot, 1e-8f));
  }
  out[(long long)b * kVD + tid] = y;
}

struct NvTailArgs {
  const float* part_v; const float* part_s; int slabs; const float* cw2; float* vlad; float* coln;
  const float* hw; float* part_h;
  const float* bn_scale; const float* bn_shift; const float* gw; const float* g_scale; const float* g_shift;
  int final_l2norm; float* out; int B;
  unsigned int* barrier;   // four counters (ticket, A / B / C items done): zero at launch, left zero by the last item
};

__global__ void __launch_bounds__(kVD)
netvlad_tail_kernel(const NvTailArgs a) {
  __shared__ unsigned int s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicAdd(a.barrier, 1u);
  __syncthreads();
  const unsigned int t = s_ticket;
  const int slices = kVD * kVK / kVSlice;
  const unsigned int n_a = (unsigned)a.B * (kVK / kVQ);
  const unsigned int n_b = (unsigned)slices * (unsigned)((a.B + 31) / 32);
  const unsigned int n_c = (unsigned)a.B;
  if (t < n_a) {
    nv_finalize_item((int)t / (kVK / kVQ), (int)t % (kVK / kVQ), a.part_v, a.part_s, a.slabs, a.cw2, a.vlad, a.coln);
    nv_signal(a.barrier + 1);
  } else if (t < n_a + n_b) {
    const int w = (int)(t - n_a);
    nv_project_item(w % slices, (w / slices) * 32, a.vlad, a.hw, a.B, kVD * kVK, a.part_h, a.barrier + 1, n_a);
    nv_signal(a.barrier + 2);
  } else if (t < n_a + n_b + n_c) {
    nv_wait(a.barrier + 2, n_b);
    nv_head_item((int)(t - n_a - n_b), a.part_h, slices, a.B, a.coln, a.bn_scale, a.bn_shift, a.gw, a.g_scale,
                 a.g_shift, a.final_l2norm, a.out);
    __syncthreads();
    // the last item to finish leaves the counters zero again (nobody reads them any more: every other CTA is past its wait)
    if (threadIdx.x == 0 && atomicAdd(a.barrier + 3, 1u) == n_c - 1u)
      *reinterpret_cast<uint4*>(a.barrier) = make_uint4(0u, 0u, 0u, 0u);
  }
}

static int netvlad_tail_launch(const NvTailArgs& a, cudaStream_t st) {
  const int slices = kVD * kVK / kVSlice;
  const int grid = a.B * (kVK / kVQ) + slices * ((a.B + 31) / 32) + a.B;   // one CTA per work item of the three phases
  netvlad_tail_kernel<<<grid, kVD, 0, st>>>(a);
  return launch_status();
}

static size_t nv_part_v_bytes(int B) { return align_up((size_t)B * kVMaxSlabs * kVD * kVK * 4, 256); }
static size_t nv_part_s_bytes(int B) { return align_up((size_t)B * kVMaxSlabs * kVK * 4, 256); }
static size_t nv_vlad_bytes(int B) { return align_up((size_t)B * kVD * kVK * 4, 256); }
static size_t nv_coln_bytes(int B) { return align_up((size_t)B * kVK * 4, 256); }
static size_t nv_part_h_bytes(int B) {
  return align_up((size_t)(kVD * kVK / kVSlice) * B * kVD * 4, 256);
}
static size_t nv_barrier_bytes() { return 256; }

// netvlad_tc.cu: the aggregation on the tensor cores (default); DH3D_EXACT_FP32=1 selects the FFMA kernel above
size_t netvlad_tc_workspace_bytes();
int netvlad_tc_aggregate_launch(const float* features, const float* att, int B, int N, const float* cw,
                                const float* bn_scale, const float* bn_shift, float* part_v, float* part_s,
                                int* P_out, void* ws, unsigned int* zero_word, cudaStream_t st);
bool exact_fp32();   // capi.cu
static bool netvlad_use_tc() { return !exact_fp32(); }

size_t netvlad_workspace_bytes(int B, int N, int D, int Kc, int out_dim) {
  (void)N;
  if (B <= 0 || D != kVD || Kc != kVK || out_dim != kVD) return 0;
  return nv_part_v_bytes(B) + nv_part_s_bytes(B) + nv_vlad_bytes(B) + nv_coln_bytes(B) + nv_part_h_bytes(B) +
         nv_barrier_bytes() + align_up(netvlad_tc_workspace_bytes(), 256);
}

int netvlad_launch(const float* features, const float* att, int B, int N, int D, int Kc, int out_dim,
                   const float* cw, const float* cbn_scale, const float* cbn_shift, const float* cw2,
                   const float* hw, const float* bn_scale, const float* bn_shift, const float* gw,
                   const float* gbn_scale, const float* gbn_shift, int final_l2norm, float* out,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!features || !att || !cw || !cbn_scale || !cbn_shift || !cw2 || !hw || !bn_scale || !bn_shift ||
      !gw || !gbn_scale || !gbn_shift || !out)
    return DH3D_ERR_NULL;
  if (B <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (D != kVD || Kc != kVK || out_dim != kVD || B > 65535) return DH3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < netvlad_workspace_bytes(B, N, D, Kc, out_dim)) return DH3D_ERR_WORKSPACE;
  if ((((uintptr_t)features | (uintptr_t)ws | (uintptr_t)cw | (uintptr_t)cbn_scale |
        (uintptr_t)cbn_shift) & 15) != 0)
    return DH3D_ERR_ALIGN;
  char* p = reinterpret_cast<char*>(ws);
  float* part_v = reinterpret_cast<float*>(p); p += nv_part_v_bytes(B);
  float* part_s = reinterpret_cast<float*>(p); p += nv_part_s_bytes(B);
  float* vlad = reinterpret_cast<float*>(p); p += nv_vlad_bytes(B);
  float* coln = reinterpret_cast<float*>(p); p += nv_coln_bytes(B);
  float* part_h = reinterpret_cast<float*>(p); p += nv_part_h_bytes(B);
  unsigned int* barrier = reinterpret_cast<unsigned int*>(p); p += nv_barrier_bytes();
  void* tc_ws = p;
  // (the tail kernel's counters: the workspace is not assumed to be zeroed, the aggregate kernel zeroes them)
  NvTailArgs tail{part_v, part_s, 0, cw2, vlad, coln, hw, part_h, bn_scale, bn_shift, gw, gbn_scale, gbn_shift,
                  final_l2norm, out, B, barrier};

  int rc;
  if (netvlad_use_tc()) {
    int P = 0;
    rc = netvlad_tc_aggregate_launch(features, att, B, N, cw, cbn_scale, cbn_shift, part_v, part_s, &P, tc_ws, barrier, st);
    if (rc != DH3D_OK) return rc;
    tail.slabs = P;
    return netvlad_tail_launch(tail, st);
  }

  cudaError_t e = cudaFuncSetAttribute(netvlad_aggregate_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(VladSmem));
  if (e != cudaSuccess) return (int)e;
  // slabs per cloud: fill whole waves of (148 SMs x 2 resident CTAs), at least 64 points per slab
  int slabs = (2 * kNumSMs) / B;
  if (slabs < 1) slabs = 1;
  if (slabs > kVMaxSlabs) slabs = kVMaxSlabs;
  while (slabs > 1 && (N + slabs - 1) / slabs < 64) --slabs;
  netvlad_aggregate_kernel<<<dim3(slabs, B), kVD, sizeof(VladSmem), st>>>(
      features, att, N, slabs, cw, cbn_scale, cbn_shift, part_v, part_s, barrier);
  rc = launch_status();
  if (rc != DH3D_OK) return rc;
  tail.slabs = slabs;
  return netvlad_tail_launch(tail, st);
}

}  // namespace dh3d
