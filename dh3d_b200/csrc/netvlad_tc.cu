// Attention NetVLAD aggregation on the tensor cores (reference: core/backbones.py:202-260).
//
// Per point n:  xn = x / |x|;  a[n,:] = softmax_k(BN(xn @ W)) * att[n];  V[d,k] += a[n,k] * xn[n,d];  S[k] += a[n,k].
// Both contractions are GEMMs -- the assignment (64 x 256 x 64 per 64-point tile) and the residual sum
// (256 x 64 x 64) -- and the FFMA kernel in netvlad.cu spends 0.66 ms on them at 6 % of the HBM roofline.
// Here they run as tcgen05.mma kind::f16 on 2-term fp16 splits (22 mantissa bits, 3 MMAs per product like
// gemm_tc16.cu), and one fp16 copy of the tile serves BOTH products:
//     xh/xl tile [64 points x 256 features], 4 slabs of [64 rows x 128 B], 128B swizzle
//       = K-major  A operand of  logits[64 x 64]  = Xn[64 x 256] . W[256 x 64]          (M = points)
//       = MN-major A operand of  V[256 x 64]     += Xn^T[256 x 64] . A[64 x 64]          (M = features)
// so each feature row is read from HBM exactly once and nothing per-point is ever written.
// Rows are l2-normalised BEFORE the split (|xn| = 1, scaled by 2^12) and the assignment lies in [0,1]
// (scaled by 2^14): no fp16 overflow for any input, low parts subnormal only below 2^-37 / 2^-39 absolute.
//
// One CTA works through consecutive 64-point tiles of one cloud (320 threads, 1 CTA/SM):
//   warp 0     TMA: 8 raw fp32 slabs [64 x 32] per tile into an 8-slab ring (the next tile's slabs stream in while
//              this one is being reduced); W hi/lo (64 KB fp16, pre-split by linear_prepack16) once per CTA
//   warps 2-5  row norms (pass 1 over the ring), then normalise + split into xh/xl (pass 2), freeing ring slabs
//   warp 1     48 MMAs 64x64x16 -> logits in TMEM; after the softmax 24 MMAs 128x64x16 -> V (2 x 64 TMEM columns,
//              accumulated over a sub-slab of 8 tiles = 512 points, then flushed: bounded accumulation depth)
//   warps 6-9  tcgen05.ld logits (M=64 layout: 16 rows per warp) -> folded BN -> softmax -> x attention ->
//              S[k] partial sums (16-lane shuffles) -> fp16 hi/lo assignment tile (B operand of the second product);
//              at the end of a sub-slab: V from TMEM -> part_v, S -> part_s (combined by netvlad_finalize_kernel)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace dh3d {

constexpr int kNT = 64;            // points per tile
constexpr int kND = 256;           // feature dim
constexpr int kNK = 64;            // clusters
constexpr int kNFlush = 8;         // tiles per sub-slab (TMEM accumulation depth 512 points)
constexpr int kNThreads = 320;     // 10 warps: TMA, MMA, 4 split, 4 softmax
constexpr uint32_t kRawSlab = kNT * 32 * 4;     // 8 KB
constexpr uint32_t kHSlab = kNT * 128;          // 8 KB: [64 rows x 64 fp16]
constexpr float kXScale = 4096.f;               // 2^12
constexpr float kAScale = 16384.f;              // 2^14

struct NvArgs {
  const float* att;        // [B, N]
  const float* colscale;   // [64] from linear_prepack16: 2^-4 / sw[k]
  const float* bn_scale;   // [64]
  const float* bn_shift;   // [64]
  float* part_v;           // [B][P][256][64]
  float* part_s;           // [B][P][64]
  int N, tiles_per_cta, subslabs, P;
  unsigned int* zero_word;   // the tail kernel's four counters (16-byte aligned): zeroed here, one launch ahead of their use
};

// MN-major, 128B-swizzled operand: 128-byte rows of 64 fp16 along M; 8-row K groups 1024 B apart;
// the next 64 M values (next slab) `lbo` bytes further.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;   // leading byte offset: stride between 128-byte groups along M
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset: stride between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void nv_umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void nv_split8(const float4& a, const float4& b, float s, uint4& hi, uint4& lo) {
  const float v[8] = {a.x * s, a.y * s, a.z * s, a.w * s, b.x * s, b.y * s, b.z * s, b.w * s};
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

// ===========================================================================================================
// The kernel.  A first version ran each 64-point tile as one serial chain (split -> logits MMAs -> softmax ->
// residual MMAs; ~14000 cycles per tile, ncu r1j: 239 us per 32 x 8192 points) because its single xh/xl operand tile
// was shared by both products and there was no room for a second one next to an 8-slab raw ring; it was deleted once
// this one measured 150 us (profiles/op_table_r1z.json vs r2a).  Here the raw fp32 slabs are converted IN PLACE: two raw slabs [64 rows x 32 fp32] of 8 KB (features 64j..64j+63)
// become the xh slab and the xl slab [64 rows x 64 fp16] of the same 16 KB, and every thread reads and writes
// only its own 128-byte rows, so the conversion needs no extra synchronisation.  That frees the 64 KB of xh/xl
// and turns the raw ring into TWO whole-tile buffers: tile i+1 is loaded and split while tile i is in the tensor
// core / softmax stages.  The logits accumulator is handed back as soon as the softmax warps hold it in registers
// (l_free), so the MMA thread issues logits(i+1) before the residual product of tile i; the softmax stage is what
// paces the loop.  Same arithmetic, same operand layouts and descriptors (only the slab addresses move).
// ===========================================================================================================
struct NvSmem2 {
  static constexpr uint32_t buf = 0;                          // 2 tiles x 8 slabs x 8 KB
  static constexpr uint32_t wh = buf + 2 * 8 * kRawSlab;      // 4 slabs x 8 KB  (W^T hi, K-major)
  static constexpr uint32_t wl = wh + 4 * kHSlab;
  static constexpr uint32_t ah = wl + 4 * kHSlab;             // [64 k rows x 64 points] fp16
  static constexpr uint32_t al = ah + kHSlab;
  static constexpr uint32_t ss = al + kHSlab;                 // [2 parity][2 halves][64] partial row sums of squares
  static constexpr uint32_t prm = ss + 2 * 2 * 64 * 4;        // [2][64]: logit scale, shift
  static constexpr uint32_t ssum = prm + 2 * 64 * 4;          // [4 warps][64] S partials at flush
  static constexpr uint32_t bars = ssum + 4 * 64 * 4;
  static constexpr uint32_t total = bars + 256;
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(kNThreads, 1)
netvlad_tc2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWh,
                   const __grid_constant__ CUtensorMap tmWl, const NvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (a.zero_word && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    *reinterpret_cast<uint4*>(a.zero_word) = make_uint4(0u, 0u, 0u, 0u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NvSmem2::bars);
  uint64_t* raw_full = bars;         // [2] count 1 + tx: a whole raw tile landed in buffer b
  uint64_t* x_full = bars + 2;       // [2] count 128: buffer b converted to xh/xl
  uint64_t* x_free = bars + 4;       // [2] count 1: residual product of the tile in buffer b done
  uint64_t* l_full = bars + 6;       // count 1: logits in TMEM
  uint64_t* l_free = bars + 7;       // count 128: logits copied to registers
  uint64_t* a_full = bars + 8;       // count 128: assignment tile written
  uint64_t* a_free = bars + 9;       // count 1: residual product done (assignment tile reusable)
  uint64_t* v_full = bars + 10;      // count 1: sub-slab accumulated
  uint64_t* v_empty = bars + 11;     // count 128: V read back
  uint64_t* w_full = bars + 12;      // count 1 + tx
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_ss = reinterpret_cast<float*>(smem + NvSmem2::ss);
  float* s_prm = reinterpret_cast<float*>(smem + NvSmem2::prm);
  float* s_sum = reinterpret_cast<float*>(smem + NvSmem2::ssum);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, c = blockIdx.x;
  const int tiles_cloud = (a.N + kNT - 1) / kNT;
  const int t_begin = c * a.tiles_per_cta;
  const int t_end = min(tiles_cloud, t_begin + a.tiles_per_cta);
  const int ntiles = max(0, t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&x_full[i], 128);
      mbar_init(&x_free[i], 1);
    }
    mbar_init(l_full, 1);
    mbar_init(l_free, 128);
    mbar_init(a_full, 128);
    mbar_init(a_free, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 128);
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 192 && threadIdx.x < 256) {  // epilogue constants: logit = raw * scale + shift
    const int k = threadIdx.x - 192;
    s_prm[k] = __ldg(a.colscale + k) * (16.f / kXScale) * __ldg(a.bn_scale + k);
    s_prm[64 + k] = __ldg(a.bn_shift + k);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_logits = tmem_base;        // 64 columns, M = 64 layout (16 lanes per warp quadrant)
  const uint32_t tm_v = tmem_base + 64;        // 2 x 64 columns, M = 128

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 8 * kHSlab);
      for (int s = 0; s < 4; ++s) {
        tma_load_2d(smem + NvSmem2::wh + s * kHSlab, &tmWh, s * 64, 0, w_full);
        tma_load_2d(smem + NvSmem2::wl + s * kHSlab, &tmWl, s * 64, 0, w_full);
      }
      for (int i = 0; i < ntiles; ++i) {
        const int bb = i & 1, n = i >> 1;
        const int row0 = b * a.N + (t_begin + i) * kNT;
        // the buffer of tile i+1 frees late (two tile buffers), so its HBM latency is taken now, into L2: the load
        // issued after the wait below then only pays the L2 -> shared-memory leg
        if (i + 1 < ntiles)
          for (int s = 0; s < 8; ++s) tma_prefetch_2d(&tmX, s * 32, row0 + kNT);
        mbar_wait(&x_free[bb], (uint32_t)(n & 1) ^ 1u);   // the tile that used this buffer two tiles ago is consumed
        mbar_arrive_expect_tx(&raw_full[bb], 8 * kRawSlab);
        for (int s = 0; s < 8; ++s)
          tma_load_2d(smem + NvSmem2::buf + (bb * 8 + s) * kRawSlab, &tmX, s * 32, row0, &raw_full[bb]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && ntiles > 0) {
      const uint32_t idesc1 = (1u << 4) | ((uint32_t)(kNK >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 15) | ((uint32_t)(kNK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t buf0 = smem_u32(smem + NvSmem2::buf);
      const uint32_t wh = smem_u32(smem + NvSmem2::wh), wl = smem_u32(smem + NvSmem2::wl);
      const uint32_t ah = smem_u32(smem + NvSmem2::ah), al = smem_u32(smem + NvSmem2::al);
      // logits(i) = Xn . W : xh slab j of buffer bb sits at slab 2j, xl slab j at slab 2j + 1
      auto issue_logits = [&](int i) {
        const uint32_t xb = buf0 + (uint32_t)(i & 1) * 8u * kRawSlab;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_h = umma_desc_sw128(xb + (2 * s) * kHSlab + k * 32);
            const uint64_t a_l = umma_desc_sw128(xb + (2 * s + 1) * kHSlab + k * 32);
            const uint64_t b_h = umma_desc_sw128(wh + s * kHSlab + k * 32), b_l = umma_desc_sw128(wl + s * kHSlab + k * 32);
            nv_umma_f16(tm_logits, a_l, b_h, idesc1, (s | k) != 0 ? 1u : 0u);
            nv_umma_f16(tm_logits, a_h, b_l, idesc1, 1u);
            nv_umma_f16(tm_logits, a_h, b_h, idesc1, 1u);
          }
        umma_commit(l_full);
      };
      // V += Xn^T . A
      auto issue_residual = [&](int i) {
        const int f = i / kNFlush, fi = i - f * kNFlush;
        const uint32_t xb = buf0 + (uint32_t)(i & 1) * 8u * kRawSlab;
        if (fi == 0 && f > 0) mbar_wait(v_empty, (f - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t aoff = k * 2048;   // 16 point rows per K step
            const uint64_t a_h = umma_desc_mn_sw128(xb + (4 * h) * kHSlab + aoff, 2 * kHSlab);
            const uint64_t a_l = umma_desc_mn_sw128(xb + (4 * h + 1) * kHSlab + aoff, 2 * kHSlab);
            const uint64_t b_h = umma_desc_sw128(ah + k * 32), b_l = umma_desc_sw128(al + k * 32);
            const uint32_t acc = (fi | k) != 0 ? 1u : 0u;
            nv_umma_f16(tm_v + h * 64, a_l, b_h, idesc2, acc);
            nv_umma_f16(tm_v + h * 64, a_h, b_l, idesc2, 1u);
            nv_umma_f16(tm_v + h * 64, a_h, b_h, idesc2, 1u);
          }
        umma_commit(&x_free[i & 1]);
        umma_commit(a_free);
        if (fi == kNFlush - 1 || i == ntiles - 1) umma_commit(v_full);
      };
      mbar_wait(w_full, 0);
      mbar_wait(&x_full[0], 0);
      issue_logits(0);
      for (int i = 0; i < ntiles; ++i) {
        bool need_logits = i + 1 < ntiles, need_residual = true;
        const int nb = (i + 1) & 1, nn = (i + 1) >> 1;
        while (need_logits || need_residual) {   // whichever is ready first; the tensor pipe runs them in issue order
          if (need_residual && mbar_test(a_full, i & 1)) {
            issue_residual(i);
            need_residual = false;
          } else if (need_logits && mbar_test(l_free, i & 1) && mbar_test(&x_full[nb], nn & 1)) {
            issue_logits(i + 1);
            need_logits = false;
          }
        }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ norms + in-place fp16 split (128 threads)
    const int t = threadIdx.x - 64;
    const int r = t & 63, half = t >> 6;   // row of the tile; raw slabs 4*half .. 4*half+3 (128 features)
    for (int i = 0; i < ntiles; ++i) {
      const int bb = i & 1, n = i >> 1;
      uint8_t* tile = smem + NvSmem2::buf + bb * 8 * kRawSlab;
      mbar_wait(&raw_full[bb], n & 1);
      // pass 1: sum of squares of this thread's half row
      float ssq = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = 4 * half + j;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const float4 v = *reinterpret_cast<const float4*>(tile + s * kRawSlab + r * 128 + ((cc ^ (r & 7)) << 4));
          ssq = fmaf(v.x, v.x, ssq); ssq = fmaf(v.y, v.y, ssq); ssq = fmaf(v.z, v.z, ssq); ssq = fmaf(v.w, v.w, ssq);
        }
      }
      float* ssb = s_ss + (i & 1) * 128;
      ssb[half * 64 + r] = ssq;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const float inv = rsqrtf(fmaxf(ssb[r] + ssb[64 + r], 1e-12f)) * kXScale;  // tf.nn.l2_normalize epsilon
      // pass 2: raw slabs (2j', 2j'+1) -> xh slab (at 2j') and xl slab (at 2j'+1), row r only: read the 256 bytes of
      // this row first, then overwrite them
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int s0 = 4 * half + 2 * pp;
        float4 v[16];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          v[cc] = *reinterpret_cast<const float4*>(tile + s0 * kRawSlab + r * 128 + ((cc ^ (r & 7)) << 4));
          v[8 + cc] = *reinterpret_cast<const float4*>(tile + (s0 + 1) * kRawSlab + r * 128 + ((cc ^ (r & 7)) << 4));
        }
        uint8_t* dh = tile + s0 * kRawSlab + r * 128;
        uint8_t* dl = tile + (s0 + 1) * kRawSlab + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {   // 16-byte chunk q of the fp16 rows = features 8q .. 8q+7 of this 64-feature slab
          uint4 hi, lo;
          nv_split8(v[2 * q], v[2 * q + 1], inv, hi, lo);
          const uint32_t off = (uint32_t)(q ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(dh + off) = hi;
          *reinterpret_cast<uint4*>(dl + off) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(&x_full[bb]);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ softmax + flush (128 threads)
    const int q = warp & 3;                 // TMEM lane quadrant
    // M = 64 accumulator: rows 16q .. 16q+15 sit in lanes 0..15 of the quadrant.  Lane L + 16 takes clusters 32..63
    // of row L (its copy of the upper 32 logits arrives by shuffle), so all 32 lanes share the exp / convert / store
    // work of the stage that paces the kernel.
    const int hsel = lane >> 4, kb = hsel * 32;
    const int row = 16 * q + (lane & 15);
    float mys[2] = {0.f, 0.f};              // S[k] partials of this warp, k = kb + 2 * (lane & 15) + j
    for (int f = 0; f < a.subslabs; ++f) {
      const int i0 = f * kNFlush, i1 = min(ntiles, i0 + kNFlush);
      for (int i = i0; i < i1; ++i) {
        mbar_wait(l_full, i & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tm_logits + ((uint32_t)(q * 32) << 16);
        DH3D_TMEM_LD_32X32(r0, taddr);
        DH3D_TMEM_LD_32X32(r1, taddr + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(l_free);                 // the logits of the next tile may overwrite the accumulator now
        float p[32];
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const uint32_t other = __shfl_xor_sync(0xffffffffu, hsel ? r0[k] : r1[k], 16);
          const uint32_t mine = hsel ? other : r0[k];
          p[k] = fmaf(__uint_as_float(mine), s_prm[kb + k], s_prm[64 + kb + k]);
          mx = fmaxf(mx, p[k]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) { p[k] = __expf(p[k] - mx); sum += p[k]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        const int n = (t_begin + i) * kNT + row;
        float scale = 0.f;
        if (n < a.N) scale = __ldg(a.att + (long long)b * a.N + n) / sum;
        if (i > 0) mbar_wait(a_free, (i - 1) & 1);   // the residual product of the previous tile has read the tile
        // assignment tile: B operand [64 k rows x 64 points], K-major, 128B swizzle: element (k, row)
        uint8_t* pah = smem + NvSmem2::ah + kb * 128;
        uint8_t* pal = smem + NvSmem2::al + kb * 128;
        const uint32_t col = ((row >> 3) << 4), sub = (row & 7) * 2;   // 16-byte chunk (8 points), byte inside it
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          p[k] *= scale;                          // assignment in [0, 1]
          const float as = p[k] * kAScale;
          const __half h = __float2half_rn(as);
          const __half l = __float2half_rn(as - __half2float(h));
          const uint32_t off = k * 128 + ((((col >> 4) ^ (k & 7))) << 4) + sub;   // (kb + k) & 7 == k & 7
          *reinterpret_cast<__half*>(pah + off) = h;
          *reinterpret_cast<__half*>(pal + off) = l;
        }
        fence_proxy_async();
        mbar_arrive(a_full);
        // S[k] += sum over the 16 rows of this warp: reduce-scatter over lane bits 3..0 (30 shuffles); lane L ends
        // with the sums of clusters kb + 2 (L & 15), +1
#pragma unroll
        for (int sft = 16; sft >= 2; sft >>= 1) {
          const int bit = sft >> 1;               // lane bit 8, 4, 2, 1
          const bool up = (lane & bit) != 0;
#pragma unroll
          for (int t = 0; t < sft; ++t) {
            const float send = up ? p[t] : p[t + sft];
            const float keep = up ? p[t + sft] : p[t];
            p[t] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
        mys[0] += p[0];
        mys[1] += p[1];
      }
      // ---- flush the sub-slab: V (TMEM, all 128 lanes) -> part_v, S -> part_s
      const int pidx = c * a.subslabs + f;
      float* pv = a.part_v + ((long long)(b * a.P + pidx) * kND) * kNK;
      if (i1 > i0) {
        mbar_wait(v_full, f & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int d = h * 128 + q * 32 + lane;
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t rv[32];
            DH3D_TMEM_LD_32X32(rv, tm_v + h * 64 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float4* dst = reinterpret_cast<float4*>(pv + (long long)d * kNK + c0);
            const float sc = 1.f / (kXScale * kAScale);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(rv[4 * j]) * sc, __uint_as_float(rv[4 * j + 1]) * sc,
                                   __uint_as_float(rv[4 * j + 2]) * sc, __uint_as_float(rv[4 * j + 3]) * sc);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(v_empty);
      } else {
        for (int e = threadIdx.x - 192; e < kND * kNK / 4; e += 128)
          reinterpret_cast<float4*>(pv)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) { s_sum[q * 64 + kb + 2 * (lane & 15) + j] = mys[j]; mys[j] = 0.f; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - 192;
      if (e < 64)
        a.part_s[(long long)(b * a.P + pidx) * kNK + e] = s_sum[e] + s_sum[64 + e] + s_sum[128 + e] + s_sum[192 + e];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// fp16 tensor map [rows, cols] row pitch ld; box = [box_rows x 64 cols] (128 bytes), 128B swizzle
static int nv_make_map_f16(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                           int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DH3D_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DH3D_OK : DH3D_ERR_UNSUPPORTED;
}

size_t linear_prepack16_bytes(int K, int N);
int linear_prepack16_launch(const float* w, int K, int N, void* packed, cudaStream_t st);

size_t netvlad_tc_workspace_bytes() { return linear_prepack16_bytes(kND, kNK); }

// partial-slab count per cloud for B clouds of N points (what part_v / part_s must hold)
int netvlad_tc_partials(int B, int N, int* ctas_per_cloud, int* tiles_per_cta, int* subslabs) {
  const int tiles = ceil_div(N, kNT);
  int C = num_sms() / B;
  if (C < 1) C = 1;
  if (C > tiles) C = tiles;
  int tpc = ceil_div(tiles, C);
  int nf = ceil_div(tpc, kNFlush);
  while (C * nf > 32 && C > 1) {   // netvlad.cu sizes its partial buffers for at most 32 slabs per cloud
    --C;
    tpc = ceil_div(tiles, C);
    nf = ceil_div(tpc, kNFlush);
  }
  if (ctas_per_cloud) *ctas_per_cloud = C;
  if (tiles_per_cta) *tiles_per_cta = tpc;
  if (subslabs) *subslabs = nf;
  return C * nf;
}

// features [B*N, 256] fp32, att [B, N]; writes P partial slabs per cloud into part_v / part_s; returns P in *P_out
int netvlad_tc_aggregate_launch(const float* features, const float* att, int B, int N, const float* cw,
                                const float* bn_scale, const float* bn_shift, float* part_v, float* part_s,
                                int* P_out, void* ws, unsigned int* zero_word, cudaStream_t st) {
  int C, tpc, nf;
  const int P = netvlad_tc_partials(B, N, &C, &tpc, &nf);
  if (P > 32) return DH3D_ERR_UNSUPPORTED;
  int rc = linear_prepack16_launch(cw, kND, kNK, ws, st);   // {W_h^T, W_l^T} [64, 256] fp16 + colscale[64]
  if (rc != DH3D_OK) return rc;
  const char* base = reinterpret_cast<const char*>(ws);
  const size_t plane = align_up((size_t)kND * kNK * 2, 256);
  CUtensorMap mx, mh, ml;
  if ((rc = make_map(&mx, features, (long long)B * N, kND, kND, kNT)) != DH3D_OK) return rc;
  if ((rc = nv_make_map_f16(&mh, base, kNK, kND, kND, kNK)) != DH3D_OK) return rc;
  if ((rc = nv_make_map_f16(&ml, base + plane, kNK, kND, kND, kNK)) != DH3D_OK) return rc;
  NvArgs a{att, reinterpret_cast<const float*>(base + 2 * plane), bn_scale, bn_shift, part_v, part_s, N, tpc, nf, P, zero_word};
  const int smem = (int)NvSmem2::total + 1024;
  cudaError_t e = cudaFuncSetAttribute(netvlad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  netvlad_tc2_kernel<<<dim3(C, B), kNThreads, smem, st>>>(mx, mh, ml, a);
  if (P_out) *P_out = P;
  return launch_status();
}

}  // namespace dh3d
