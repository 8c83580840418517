// Shared pieces of the fp16-pair split GEMM kernels (gemm_tc16.cu: streaming kernel + two-branch join;
// gemm_head16.cu: the resident-activation head kernel): split arithmetic, out-of-window row queue and fp32
// recompute, UMMA wrappers of the CTA-pair form, packed-weight layout helpers.  See gemm_tc16.cu for the scheme.
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace dh3d {

constexpr int kT16Threads = 320;
constexpr float kT16XScale = 16.f;         // 2^4
constexpr float kT16XScaleInv = 0.0625f;
// safe window of a row's largest |x| for the fixed-scale fp16 split, as (float bits << 1) (monotone in |x|;
// inf / NaN compare above every finite value)
constexpr uint32_t kT16HiBits2 = 0x456A6000u << 1;   // 3750.0f
constexpr uint32_t kT16LoBits2 = 0x3A000000u << 1;   // 2^-11
constexpr int kT16BadCap = 1024;                     // queued out-of-window rows per CTA (more: redo all its rows)
constexpr uint32_t kT16BadBytes = 16 + kT16BadCap * 4;

__device__ __forceinline__ uint32_t t16_absmax8(uint32_t m, const float4& a, const float4& b) {
  m = max(m, __float_as_uint(a.x) << 1); m = max(m, __float_as_uint(a.y) << 1);
  m = max(m, __float_as_uint(a.z) << 1); m = max(m, __float_as_uint(a.w) << 1);
  m = max(m, __float_as_uint(b.x) << 1); m = max(m, __float_as_uint(b.y) << 1);
  m = max(m, __float_as_uint(b.z) << 1); m = max(m, __float_as_uint(b.w) << 1);
  return m;
}
__device__ __forceinline__ bool t16_row_out_of_window(uint32_t m2) {
  return m2 > kT16HiBits2 || (m2 != 0u && m2 < kT16LoBits2);
}
// split-warp thread (chunk c = t & 3 of rows (t >> 2) + 32 i): reduce the 4 chunk owners of each row and queue
// the rows whose maximum left the window.  bad[0] = count, bad[4..] = global row indices.
__device__ __forceinline__ void t16_queue_bad_rows(uint32_t (&rmax)[4], int t, int row0, uint32_t* bad) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t m = rmax[i];
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
    rmax[i] = 0u;
    if ((t & 3) == 0 && t16_row_out_of_window(m)) {
      const uint32_t slot = atomicAdd(&bad[0], 1u);
      if (slot < (uint32_t)kT16BadCap) bad[4 + slot] = (uint32_t)(row0 + (t >> 2) + 32 * i);
    }
  }
}

struct T16Epilogue {
  const float* scale;     // [N] or null
  const float* shift;     // [N] or null
  const float* colscale;  // [N]: 2^-4 / sw[n] (from the prepack)
  int act;
  const float* w2;        // [N]: fused row-dot (ROWDOT mode)
  float b2;
  int act2;
  float* y2;              // [M]   (ROWDOT mode)
  // what the fp32 recompute of out-of-window rows reads / writes (raw pointers next to the tensor maps)
  const float* x; int ldx;
  const __half* wh; const __half* wl; int Kp;
  float* y; int ldy;
};

constexpr int kT16RB = 16;          // out-of-window rows recomputed together (one pass over W per batch)
constexpr int kT16FixMaxK = 1024;   // batched recompute stages RB x K floats in the (idle) stage ring

// 8 weights (wh + wl, exact in fp32) of column n at k0 .. k0+7
struct T16W8 { float w[8]; };
__device__ __forceinline__ T16W8 t16_load_w8(const __half* __restrict__ h, const __half* __restrict__ l, int k0) {
  const uint4 hv = __ldg(reinterpret_cast<const uint4*>(h + k0));
  const uint4 lv = __ldg(reinterpret_cast<const uint4*>(l + k0));
  const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
  T16W8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
    r.w[2 * i] = a.x + b.x;
    r.w[2 * i + 1] = a.y + b.y;
  }
  return r;
}

// acc[r] += x_r[k0 .. k0+7] . w8 for the staged rows (xs: [RB][K] floats in shared memory); K % 4 == 0 only
__device__ __forceinline__ void t16_dot_rows(float (&acc)[kT16RB], const float* xs, int K, int k0, const T16W8& w) {
  const bool tail = k0 + 4 < K;
#pragma unroll
  for (int r = 0; r < kT16RB; ++r) {
    const float4 x0 = *reinterpret_cast<const float4*>(xs + r * K + k0);
    const float4 x1 = tail ? *reinterpret_cast<const float4*>(xs + r * K + k0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float a = acc[r];
    a = fmaf(x0.x, w.w[0], a); a = fmaf(x0.y, w.w[1], a); a = fmaf(x0.z, w.w[2], a); a = fmaf(x0.w, w.w[3], a);
    a = fmaf(x1.x, w.w[4], a); a = fmaf(x1.y, w.w[5], a); a = fmaf(x1.z, w.w[6], a); a = fmaf(x1.w, w.w[7], a);
    acc[r] = a;
  }
}

// rows of the batch -> shared memory (zero rows beyond nrows), by the whole CTA
__device__ __forceinline__ void t16_stage_rows(float* xs, const float* x, int ldx, int K, const int* brow, int nrows) {
  const int kv = K >> 2;
  for (int i = threadIdx.x; i < kT16RB * kv; i += blockDim.x) {
    const int r = i / kv, c = i - r * kv;
    reinterpret_cast<float4*>(xs)[i] = r < nrows ? ldg4(x + (long long)brow[r] * ldx + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// fp32 recompute of up to RB output rows by the whole CTA (out-of-window rows only; see the header): the rows'
// activations are staged in the idle stage ring, one warp per output column reads the (wh + wl) column ONCE for
// the whole batch (lane l owns k = 8l .. 8l+7, +256 per round; 16-byte loads), reduces with shuffles and applies
// the epilogue.  y[row, n] = act((x[row, :] @ W[:, n]) * scale[n] + shift[n]) with W = (wh + wl) * colscale * 2^4.
template <bool ROWDOT>
__device__ __forceinline__ void t16_fixup_batch(const T16Epilogue& ep, const int* brow, int nrows, int K, int N, float* xs,
                                             float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  t16_stage_rows(xs, ep.x, ep.ldx, K, brow, nrows);
  __syncthreads();
  float part[kT16RB];
#pragma unroll
  for (int r = 0; r < kT16RB; ++r) part[r] = 0.f;
  for (int n = warp; n < N; n += nw) {
    const __half* h = ep.wh + (long long)n * ep.Kp;
    const __half* l = ep.wl + (long long)n * ep.Kp;
    float acc[kT16RB];
#pragma unroll
    for (int r = 0; r < kT16RB; ++r) acc[r] = 0.f;
    for (int k0 = lane * 8; k0 < K; k0 += 256) t16_dot_rows(acc, xs, K, k0, t16_load_w8(h, l, k0));
    const float cs = __ldg(ep.colscale + n) * kT16XScale;
    const float sc = ep.scale ? __ldg(ep.scale + n) : 1.f, sh = ep.shift ? __ldg(ep.shift + n) : 0.f;
    const float w2 = ROWDOT ? __ldg(ep.w2 + n) : 0.f;
#pragma unroll
    for (int r = 0; r < kT16RB; ++r) {
      const float v = tc_act(fmaf(warp_sum(acc[r]) * cs, sc, sh), ep.act);
      if constexpr (ROWDOT) part[r] = fmaf(v, w2, part[r]);
      else if (lane == 0 && r < nrows) ep.y[(long long)brow[r] * ep.ldy + n] = v;
    }
  }
  if constexpr (ROWDOT) {
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < kT16RB; ++r) red[warp * kT16RB + r] = part[r];
    }
    __syncthreads();
    if ((int)threadIdx.x < nrows) {
      float tot = 0.f;
      for (int w = 0; w < nw; ++w) tot += red[w * kT16RB + threadIdx.x];
      ep.y2[brow[threadIdx.x]] = tc_act(tot + ep.b2, ep.act2);
    }
  }
}

// x[row, 0:K] . (wh + wl)[n, 0:K] by one warp (rows too long for the batched form): every lane gets the sum
__device__ __forceinline__ float t16_row_dot(const float* __restrict__ xr, const __half* __restrict__ h,
                                             const __half* __restrict__ l, int K, int lane) {
  float acc = 0.f;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const T16W8 w = t16_load_w8(h, l, k0);
    const float4 x0 = ldg4(xr + k0);
    const float4 x1 = (k0 + 4 < K) ? ldg4(xr + k0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);   // K % 4 == 0 only
    acc = fmaf(x0.x, w.w[0], acc); acc = fmaf(x0.y, w.w[1], acc); acc = fmaf(x0.z, w.w[2], acc);
    acc = fmaf(x0.w, w.w[3], acc); acc = fmaf(x1.x, w.w[4], acc); acc = fmaf(x1.y, w.w[5], acc);
    acc = fmaf(x1.z, w.w[6], acc); acc = fmaf(x1.w, w.w[7], acc);
  }
  return warp_sum(acc);
}

// one row at a time (K > kT16FixMaxK): same arithmetic, activations read from global memory
template <bool ROWDOT>
__device__ __forceinline__ void t16_fixup_row(const T16Epilogue& ep, int row, int K, int N, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* xr = ep.x + (long long)row * ep.ldx;
  float part = 0.f;
  for (int n = warp; n < N; n += nw) {
    const float acc = t16_row_dot(xr, ep.wh + (long long)n * ep.Kp, ep.wl + (long long)n * ep.Kp, K, lane);
    float v = acc * (__ldg(ep.colscale + n) * kT16XScale);
    v = fmaf(v, ep.scale ? __ldg(ep.scale + n) : 1.f, ep.shift ? __ldg(ep.shift + n) : 0.f);
    v = tc_act(v, ep.act);
    if constexpr (ROWDOT) part = fmaf(v, __ldg(ep.w2 + n), part);
    else if (lane == 0) ep.y[(long long)row * ep.ldy + n] = v;
  }
  if constexpr (ROWDOT) {
    __syncthreads();
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < nw; ++w) tot += red[w];
      ep.y2[row] = tc_act(tot + ep.b2, ep.act2);
    }
  }
}

// The CTA's queued rows (or, if the queue overflowed, every row it owns) in batches of RB.
// ring: the stage ring (idle now): [RB*K floats | 16 ints | nw*RB floats]
template <bool ROWDOT, class NextTile>
__device__ __forceinline__ void t16_fixup_all(const T16Epilogue& ep, const uint32_t* bad, int M, int K, int N,
                                              uint8_t* ring, NextTile next_tile) {
  const uint32_t nbad = bad[0];
  float* xs = reinterpret_cast<float*>(ring);
  int* brow = reinterpret_cast<int*>(xs + kT16RB * (K <= kT16FixMaxK ? K : 0));
  float* red = reinterpret_cast<float*>(brow + kT16RB);
  auto run = [&](int nrows) {
    if (K <= kT16FixMaxK) {
      t16_fixup_batch<ROWDOT>(ep, brow, nrows, K, N, xs, red);
    } else {
      for (int r = 0; r < nrows; ++r) t16_fixup_row<ROWDOT>(ep, brow[r], K, N, red);
    }
  };
  if (nbad <= (uint32_t)kT16BadCap) {
    for (uint32_t base = 0; base < nbad; base += kT16RB) {
      const int nrows = min((int)(nbad - base), kT16RB);
      __syncthreads();
      if ((int)threadIdx.x < nrows) brow[threadIdx.x] = (int)bad[4 + base + threadIdx.x];
      __syncthreads();
      run(nrows);
    }
  } else {   // queue overflowed: redo every row this CTA owns
    for (int t = 0, mt; (mt = next_tile(t)) >= 0; ++t)
      for (int r0 = 0; r0 < kTcBM; r0 += kT16RB) {
        const int row0 = mt * kTcBM + r0;
        if (row0 >= M) break;
        const int nrows = min(M - row0, kT16RB);
        __syncthreads();
        if ((int)threadIdx.x < nrows) brow[threadIdx.x] = row0 + threadIdx.x;
        __syncthreads();
        run(nrows);
      }
  }
}

// K-major, 64B-swizzled operand tile: rows 64 B apart, 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;               // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                        // layout type SWIZZLE_64B
  return d;
}

// CG = 2: one instruction for the CTA pair (issued by the leader): each CTA's A tile [128 x 16] and N half of the
// B tile at the SAME shared-memory offsets in both CTAs, D rows 0-127 in the leader's TMEM, 128-255 in the peer's.
template <int CG = 1>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// barrier helpers of the pair form: arrives go to the LEADER CTA's barrier (rank 0 of the cluster), commits are multicast
// to both CTAs.  The arrive keeps the default (.release.cta) semantics: what it orders are the arriving thread's own
// shared-memory writes, already made visible to the async proxy by its fence.proxy.async, and its tcgen05.ld's (fenced
// by tcgen05.fence::before_thread_sync) -- an explicit .release.cluster / .acquire.cluster pair compiles to MEMBAR.ALL.GPU
// + ERRBAR per arrive and CCTL.IVALL per wait (ncu source view: 20 % of the kernel's stall samples, 0.33 -> 0.48 ms).
template <int MC>
__device__ __forceinline__ void mbar_arrive_x(uint64_t* bar) {
  if constexpr (MC == 1) {
    mbar_arrive(bar);
  } else {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, 0;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}" ::"r"(smem_u32(bar))
        : "memory");
  }
}
template <int MC>
__device__ __forceinline__ void umma_commit_x(uint64_t* bar) {
  if constexpr (MC == 1)
    umma_commit(bar);
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// 8 consecutive fp32 -> 8 fp16 high parts + 8 fp16 low parts (of x * 2^4), packed as two uint4
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x * kT16XScale, a.y * kT16XScale, a.z * kT16XScale, a.w * kT16XScale,
                      b.x * kT16XScale, b.y * kT16XScale, b.z * kT16XScale, b.w * kT16XScale};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- host: packed-weight layout, tensor maps, argument checks ------------------------------------------
static inline int t16_kp(int K) { return (K + 7) / 8 * 8; }
static inline size_t t16_plane_bytes(int K, int N) { return align_up((size_t)t16_kp(K) * N * sizeof(__half), 256); }

// 2-D fp16 tensor [rows, cols] with row pitch ld (elements); box = [box_rows x 32 cols], 64B swizzle.
static inline int make_map_f16(CUtensorMap* m, const __half* base, long long rows, long long cols, long long ld,
                        int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DH3D_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DH3D_OK : DH3D_ERR_UNSUPPORTED;
}

static inline int t16_check(const float* x, int ldx, const void* packed, int M, int K, int N) {
  if (!x || !packed) return DH3D_ERR_NULL;
  if (M <= 0 || K <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (K % 4 || N % 4 || ldx % 4 || ldx < K) return DH3D_ERR_DIM;
  if ((((uintptr_t)x | (uintptr_t)packed) & 15) != 0) return DH3D_ERR_ALIGN;
  return DH3D_OK;
}

struct T16Packed { const __half* wh; const __half* wl; const float* cs; };
static inline T16Packed t16_unpack(const void* packed, int K, int N) {
  const char* base = reinterpret_cast<const char*>(packed);
  T16Packed p;
  p.wh = reinterpret_cast<const __half*>(base);
  p.wl = reinterpret_cast<const __half*>(base + t16_plane_bytes(K, N));
  p.cs = reinterpret_cast<const float*>(base + 2 * t16_plane_bytes(K, N));
  return p;
}

}  // namespace dh3d
