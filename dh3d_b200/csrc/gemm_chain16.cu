// Two chained 1x1 layers in one launch:  Y = act2( (act1((X @ W1) * s1 + b1) @ W2) * s2 + b2 )
// (detection_block's 128 -> 128 -> 256 stack in front of its 1024-wide head, core/backbones.py:132-147).
//
// As two launches of gemm_tc16_kernel the hidden [M,128] activation is written to and read back from HBM (268 MB of
// the 938 MB the detector chain moves per 32 x 8192 points; both kernels sit at 0.85-0.89 of the copy bandwidth, so
// only removing bytes helps).  Here the hidden tile never leaves the SM: per 128-row tile
//     GEMM1  X tile (<= 4 K slabs, fp32 -> xh | xl in place) x W1 -> accumulator 0 (128 TMEM columns, double-buffered)
//     E1     accumulator 0 -> BN / activation -> fp16 pair -> the H slabs in shared memory (A operand of GEMM2)
//     GEMM2  H slabs x W2 (two 128-column halves) -> accumulator 1 (256 TMEM columns)
//     E2     accumulator 1 -> BN / activation -> TMA store
// The per-tile dependency chain is what has to be kept short (the first version -- one epilogue warp group doing E1
// and E2 in turn, GEMM1(t + 1) issued after GEMM2(t) -- took 135 us against 119 us for the two separate launches; per
// tile: GEMM2 -> GEMM1 -> E1 at one warp per scheduler = 16 k cycles against the 8.6 k the HBM traffic allows):
//   * GEMM1 runs ONE TILE AHEAD (issue order G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) ...; accumulator 0 is double
//     buffered), so E1(t + 1) reads its accumulator while GEMM2(t) is still in the tensor pipe;
//   * GEMM2 runs half by half (128 output columns at a time, each half with its own full / empty barrier pair) and E2
//     is two warp groups, one per half: E2 of half 0 overlaps the MMAs of half 1, and
//     GEMM2(t + 1) may overwrite a half as soon as ITS group has read it (ncu: with one barrier pair for the whole
//     256-column accumulator the MMA thread spent the tile waiting for E2 to drain it);
//   * the X slabs of tile t + 1 are reloaded slab by slab as GEMM1(t)'s MMAs retire them.
// Both weight matrices stream through one ring of [128 x 32] (W_h | W_l) tiles in the order the MMA thread consumes them.
// Same fp16-pair split arithmetic and the same out-of-window guarantee as gemm_tc16.cu: rows whose X OR hidden
// activations leave the split's window are queued and recomputed through BOTH layers in fp32 after the tile loop.
// Shapes: K1 <= 128, N1 == 128, N2 <= 256.
#include "gemm_tc16.cuh"

namespace dh3d {

constexpr int kC16Threads = 608;   // warp 0: W producer, 1: MMA issuer, 2-5: split, 6-9: E1, 10-17: E2, 18: X producer
constexpr int kC16N1 = 128;        // hidden width == K of the second GEMM (4 slabs)
constexpr int kC16KB = 4;          // K slabs of either GEMM
constexpr int kC16WS = 3;          // W ring slots
constexpr int kC16MaxN2 = 256;

struct C16Cfg {
  static constexpr uint32_t kSlabBytes = kTcBM * kTcBK * 4;       // 16 KB: raw fp32, then xh | xl (or hh | hl)
  static constexpr uint32_t kABytes = kTcBM * kTcBK * 2;          // 8 KB
  static constexpr uint32_t kWBytes = 128 * kTcBK * 2;            // 8 KB: [128 x 32] fp16 tile of W_h^T (or W_l^T)
  static constexpr uint32_t kWSlotBytes = 2 * kWBytes;
  static constexpr uint32_t kParamBytes = (2 * kC16N1 + 2 * kC16MaxN2) * 4;
  static constexpr uint32_t kBarBytes = 512;
  static constexpr uint32_t kOutBytes = 2 * kTcStageOutBytes;     // E2: one [32 x 32] fp32 staging tile per warp (8 warps)
  static constexpr uint32_t kSmemBytes = 2 * kC16KB * kSlabBytes + kC16WS * kWSlotBytes + kOutBytes + kParamBytes +
                                         kBarBytes + kT16BadBytes + 1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

struct ChainArgs {
  const float* scale1; const float* shift1; const float* cs1; int act1;
  const float* scale2; const float* shift2; const float* cs2; int act2;
  const float* x; int ldx;
  const __half* w1h; const __half* w1l; int Kp1;
  const __half* w2h; const __half* w2l; int Kp2;
  float* y; int ldy;
  int M, K1, N2;
};

// fp32 recompute of up to RB rows through both layers by the whole CTA (out-of-window rows only).
// ring: [RB*K1 floats (x rows) | RB*128 floats (hidden rows)]
__device__ __forceinline__ void chain16_fixup_batch(const ChainArgs& a, const int* brow, int nrows, uint8_t* ring) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* xs = reinterpret_cast<float*>(ring);
  float* hs = xs + kT16RB * a.K1;
  __syncthreads();
  t16_stage_rows(xs, a.x, a.ldx, a.K1, brow, nrows);
  __syncthreads();
  for (int n = warp; n < kC16N1; n += nw) {
    float acc[kT16RB];
#pragma unroll
    for (int r = 0; r < kT16RB; ++r) acc[r] = 0.f;
    for (int k0 = lane * 8; k0 < a.K1; k0 += 256)
      t16_dot_rows(acc, xs, a.K1, k0, t16_load_w8(a.w1h + (long long)n * a.Kp1, a.w1l + (long long)n * a.Kp1, k0));
    const float cs = __ldg(a.cs1 + n) * kT16XScale;
    const float sc = a.scale1 ? __ldg(a.scale1 + n) : 1.f, sh = a.shift1 ? __ldg(a.shift1 + n) : 0.f;
#pragma unroll
    for (int r = 0; r < kT16RB; ++r) {
      const float v = tc_act(fmaf(warp_sum(acc[r]) * cs, sc, sh), a.act1);
      if (lane == 0) hs[r * kC16N1 + n] = v;
    }
  }
  __syncthreads();
  for (int n = warp; n < a.N2; n += nw) {
    float acc[kT16RB];
#pragma unroll
    for (int r = 0; r < kT16RB; ++r) acc[r] = 0.f;
    for (int k0 = lane * 8; k0 < kC16N1; k0 += 256)
      t16_dot_rows(acc, hs, kC16N1, k0, t16_load_w8(a.w2h + (long long)n * a.Kp2, a.w2l + (long long)n * a.Kp2, k0));
    const float cs = __ldg(a.cs2 + n) * kT16XScale;
    const float sc = a.scale2 ? __ldg(a.scale2 + n) : 1.f, sh = a.shift2 ? __ldg(a.shift2 + n) : 0.f;
#pragma unroll
    for (int r = 0; r < kT16RB; ++r) {
      const float v = tc_act(fmaf(warp_sum(acc[r]) * cs, sc, sh), a.act2);
      if (lane == 0 && r < nrows) a.y[(long long)brow[r] * a.ldy + n] = v;
    }
  }
}

__global__ void __launch_bounds__(kC16Threads, 1)
gemm_chain16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1h,
                    const __grid_constant__ CUtensorMap tmW1l, const __grid_constant__ CUtensorMap tmW2h,
                    const __grid_constant__ CUtensorMap tmW2l, const __grid_constant__ CUtensorMap tmY,
                    const ChainArgs a) {
  using Cfg = C16Cfg;
  constexpr int WS = kC16WS, N1 = kC16N1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* hbase = smem + kC16KB * Cfg::kSlabBytes;
  uint8_t* wring = hbase + kC16KB * Cfg::kSlabBytes;
  uint8_t* out_stage = wring + WS * Cfg::kWSlotBytes;                      // 8 warps x 4 KB, 1024-aligned
  float* prm1 = reinterpret_cast<float*>(out_stage + Cfg::kOutBytes);      // [scale * colscale | shift] x 128
  float* prm2 = prm1 + 2 * N1;                                             // [scale * colscale | shift] x 256
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(prm1) + Cfg::kParamBytes);
  uint64_t* x_full = bars;                 // [4] raw slab landed                        (count 1 + tx)
  uint64_t* x_ready = bars + 4;            // [4] xh | xl written                        (count 4, one per split warp)
  uint64_t* x_empty = bars + 8;            // [4] GEMM1's MMAs on the slab finished      (count 1, commit)
  uint64_t* w_full = bars + 12;            // [WS]                                       (count 1 + tx)
  uint64_t* w_empty = bars + 12 + WS;      // [WS]                                       (count 1, commit)
  uint64_t* acc0_full = bars + 12 + 2 * WS;       // [2]                                 (count 1, commit)
  uint64_t* acc0_empty = acc0_full + 2;           // [2]                                 (count 4, one per E1 warp)
  uint64_t* h_ready = acc0_full + 4;              // hidden tile written                 (count 4)
  uint64_t* h_empty = acc0_full + 5;              // GEMM2's MMAs finished               (count 1, commit)
  uint64_t* acc1_full = acc0_full + 6;            // [2] per 128-column half             (count 1, commit)
  uint64_t* acc1_empty = acc0_full + 8;           // [2]                                 (count 4, one per E2 warp of the half)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc0_full + 10);
  uint32_t* bad = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + Cfg::kBarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk1 = (a.K1 + kTcBK - 1) / kTcBK;
  const int num_mt = (a.M + kTcBM - 1) / kTcBM;
  const int nf = (a.N2 + 127) / 128;       // 128-column halves of the second GEMM

  auto xslab = [&](int kb) { return smem + kb * Cfg::kSlabBytes; };
  auto hslab = [&](int kb) { return hbase + kb * Cfg::kSlabBytes; };
  auto wslot = [&](int s) { return wring + s * Cfg::kWSlotBytes; };

  if (threadIdx.x == 0) {
    bad[0] = 0u;
    for (int k = 0; k < 4; ++k) {
      mbar_init(&x_full[k], 1);
      mbar_init(&x_ready[k], 4);
      mbar_init(&x_empty[k], 1);
    }
    for (int s = 0; s < WS; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc0_full[i], 1);
      mbar_init(&acc0_empty[i], 4);
    }
    mbar_init(h_ready, 4);
    mbar_init(h_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue constants of both layers, once (the layer widths are the tile widths)
  for (int c = threadIdx.x; c < N1; c += blockDim.x) {
    prm1[c] = (a.scale1 ? __ldg(a.scale1 + c) : 1.f) * __ldg(a.cs1 + c);
    prm1[N1 + c] = a.shift1 ? __ldg(a.shift1 + c) : 0.f;
  }
  for (int c = threadIdx.x; c < kC16MaxN2; c += blockDim.x) {
    const bool in = c < a.N2;
    prm2[c] = in ? (a.scale2 ? __ldg(a.scale2 + c) : 1.f) * __ldg(a.cs2 + c) : 0.f;
    prm2[kC16MaxN2 + c] = (in && a.shift2) ? __ldg(a.shift2 + c) : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc1 = tmem_base + 2 * N1;

  if (warp == 0) {
    // ------------------------------------------------------------------ W producer, in the MMA thread's order:
    // W1 slabs of the first two tiles, then per tile t its W2 tiles (half, kb) followed by the W1 slabs of tile t + 2
    if (lane == 0) {
      uint32_t it = 0;
      auto put = [&](const CUtensorMap* mh, const CUtensorMap* ml, int k0, int row0) {
        const int s = it % WS;
        mbar_wait(&w_empty[s], ((it / WS) & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[s], Cfg::kWSlotBytes);
        tma_load_2d(wslot(s), mh, k0, row0, &w_full[s]);
        tma_load_2d(wslot(s) + Cfg::kWBytes, ml, k0, row0, &w_full[s]);
        ++it;
      };
      const int G = (int)gridDim.x;
      auto put_w1 = [&]() { for (int kb = 0; kb < nk1; ++kb) put(&tmW1h, &tmW1l, kb * kTcBK, 0); };
      if ((int)blockIdx.x < num_mt) put_w1();
      if ((int)blockIdx.x + G < num_mt) put_w1();
      for (int mt = blockIdx.x; mt < num_mt; mt += G) {
        for (int f = 0; f < nf; ++f)
          for (int kb = 0; kb < kC16KB; ++kb) put(&tmW2h, &tmW2l, kb * kTcBK, f * 128);
        if (mt + 2 * G < num_mt) put_w1();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
      uint32_t wit = 0;
      auto mma_slab = [&](uint32_t tmem_d, const uint8_t* aslab, bool first_slab) {
        const int s = wit % WS;
        mbar_wait(&w_full[s], (wit / WS) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t a_h = umma_desc_sw64(smem_u32(aslab));
        const uint64_t a_l = umma_desc_sw64(smem_u32(aslab + Cfg::kABytes));
        const uint64_t b_h = umma_desc_sw64(smem_u32(wslot(s)));
        const uint64_t b_l = umma_desc_sw64(smem_u32(wslot(s) + Cfg::kWBytes));
#pragma unroll
        for (int k = 0; k < kTcBK / 16; ++k) {
          const uint64_t off = (uint64_t)(k * 16 * 2) >> 4;
          umma_f16(tmem_d, a_l + off, b_h + off, idesc, (first_slab && k == 0) ? 0u : 1u);
          umma_f16(tmem_d, a_h + off, b_l + off, idesc, 1u);
          umma_f16(tmem_d, a_h + off, b_h + off, idesc, 1u);
        }
        umma_commit(&w_empty[s]);
        ++wit;
      };
      auto gemm1 = [&](uint32_t tl) {
        const uint32_t acc = tl & 1;
        mbar_wait(&acc0_empty[acc], ((tl >> 1) & 1) ^ 1);
        for (int kb = 0; kb < nk1; ++kb) {
          mbar_wait(&x_ready[kb], tl & 1);
          mma_slab(tmem_base + acc * N1, xslab(kb), kb == 0);
          umma_commit(&x_empty[kb]);   // the slab may take the next tile's rows
        }
        umma_commit(&acc0_full[acc]);
      };
      auto gemm2 = [&](uint32_t tl) {
        mbar_wait(h_ready, tl & 1);
        for (int f = 0; f < nf; ++f) {
          mbar_wait(&acc1_empty[f], (tl & 1) ^ 1);
          for (int kb = 0; kb < kC16KB; ++kb) mma_slab(tm_acc1 + f * 128, hslab(kb), kb == 0);
          umma_commit(&acc1_full[f]);
        }
        umma_commit(h_empty);
      };
      const int G = (int)gridDim.x;
      uint32_t tl = 0;
      if ((int)blockIdx.x < num_mt) gemm1(0);
      if ((int)blockIdx.x + G < num_mt) gemm1(1);
      for (int mt = blockIdx.x; mt < num_mt; mt += G, ++tl) {
        gemm2(tl);
        if (mt + 2 * G < num_mt) gemm1(tl + 2);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ fp32 -> (xh | xl) in place (see gemm_tc16.cu)
    const int t = threadIdx.x - 64;
    const int c = t & 3;
    uint32_t tl = 0;
    uint32_t rmax[4] = {0u, 0u, 0u, 0u};
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tl) {
      for (int kb = 0; kb < nk1; ++kb) {
        mbar_wait(&x_full[kb], tl & 1);
        uint8_t* raw = xslab(kb);
        float4 v0[4], v1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (t >> 2) + 32 * i;
          v0[i] = *reinterpret_cast<const float4*>(raw + r * 128 + (((2 * c) ^ (r & 7)) << 4));
          v1[i] = *reinterpret_cast<const float4*>(raw + r * 128 + (((2 * c + 1) ^ (r & 7)) << 4));
          rmax[i] = t16_absmax8(rmax[i], v0[i], v1[i]);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");   // every split thread holds its part of the raw slab
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (t >> 2) + 32 * i;
          uint4 hi, lo;
          split8(v0[i], v1[i], hi, lo);
          const uint32_t off = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(raw + off) = hi;
          *reinterpret_cast<uint4*>(raw + Cfg::kABytes + off) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_ready[kb]);
      }
      t16_queue_bad_rows(rmax, t, mt * kTcBM, bad);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ E1 (warps 6..9): thread = tile row;
    // accumulator 0 -> hidden tile (fp16 pair, A operand of GEMM2)
    const int q = warp & 3;
    constexpr int cbeg = 0;
    const int row = q * 32 + lane;
    uint32_t tl = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1;
      mbar_wait(&acc0_full[acc], (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t hmax = 0u;
#pragma unroll 1
      for (int c0 = cbeg; c0 < N1; c0 += 32) {
        uint32_t r[32];
        DH3D_TMEM_LD_32X32(r, tmem_base + acc * N1 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 >= N1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc0_empty[acc]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), prm1[c0 + j], prm1[N1 + c0 + j]);
        tc_act32(v, a.act1);
        if (c0 == cbeg) mbar_wait(h_empty, (tl & 1) ^ 1);   // GEMM2 of the previous tile has read the hidden slabs
        uint8_t* hs = hslab(c0 >> 5);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 f0 = make_float4(v[8 * cc], v[8 * cc + 1], v[8 * cc + 2], v[8 * cc + 3]);
          const float4 f1 = make_float4(v[8 * cc + 4], v[8 * cc + 5], v[8 * cc + 6], v[8 * cc + 7]);
          hmax = t16_absmax8(hmax, f0, f1);
          uint4 hi, lo;
          split8(f0, f1, hi, lo);
          const uint32_t off = row * 64 + ((cc ^ ((row >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(hs + off) = hi;
          *reinterpret_cast<uint4*>(hs + Cfg::kABytes + off) = lo;
        }
      }
      // the hidden row left the split's window, or holds inf / NaN
      if (t16_row_out_of_window(hmax) && mt * kTcBM + row < a.M) {
        const uint32_t slot = atomicAdd(&bad[0], 1u);
        if (slot < (uint32_t)kT16BadCap) bad[4 + slot] = (uint32_t)(mt * kTcBM + row);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
    }
  } else if (warp < 18) {
    // ------------------------------------------------------------------ E2 (warps 10..17): group f = 128-column half f
    // of accumulator 1 -> output tile
    const int q = warp & 3;
    const int f = (warp - 10) >> 2;
    uint8_t* my_stage = out_stage + (warp - 10) * 4096;
    const int cend = min(a.N2, f * 128 + 128);
    uint32_t tl = 0;
    if (f < nf) {
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tl) {
        mbar_wait(&acc1_full[f], tl & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int col = f * 128; col < cend; col += 32) {
          uint32_t r[32];
          DH3D_TMEM_LD_32X32(r, tm_acc1 + ((uint32_t)(q * 32) << 16) + (uint32_t)col);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (col + 32 >= cend) {   // this half is fully read: GEMM2 of the next tile may overwrite it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc1_empty[f]);
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), prm2[col + j], prm2[kC16MaxN2 + col + j]);
          tc_act32(v, a.act2);
          uint8_t* st = my_stage;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(st + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmY, st, col, mt * kTcBM + q * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // ------------------------------------------------------------------ X producer (warp 18): raw slabs into place
    // A slab can only be reloaded once GEMM1 has retired it, so the load of tile t + 1 sits on the tile-to-tile
    // critical path (ncu: split warps and E1 each ~50 % of their time waiting for it): the rows of the tile after
    // next are pulled into L2 ahead of time, which turns that load's DRAM latency into L2 latency.
    if (lane == 0) {
      const int G = (int)gridDim.x;
      uint32_t tl = 0;
      if ((int)blockIdx.x + G < num_mt)
        for (int kb = 0; kb < nk1; ++kb) tma_prefetch_2d(&tmX, kb * kTcBK, ((int)blockIdx.x + G) * kTcBM);
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tl)
        for (int kb = 0; kb < nk1; ++kb) {
          if (mt + 2 * G < num_mt) tma_prefetch_2d(&tmX, kb * kTcBK, (mt + 2 * G) * kTcBM);
          mbar_wait(&x_empty[kb], (tl & 1) ^ 1);
          mbar_arrive_expect_tx(&x_full[kb], Cfg::kSlabBytes);
          tma_load_2d(xslab(kb), &tmX, kb * kTcBK, mt * kTcBM, &x_full[kb]);
        }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  const uint32_t nbad = bad[0];
  if (nbad != 0u) {
    int* brow = reinterpret_cast<int*>(prm1);   // the epilogue constants are dead now
    if (nbad <= (uint32_t)kT16BadCap) {
      for (uint32_t base = 0; base < nbad; base += kT16RB) {
        const int nrows = min((int)(nbad - base), kT16RB);
        __syncthreads();
        if ((int)threadIdx.x < nrows) brow[threadIdx.x] = (int)bad[4 + base + threadIdx.x];
        __syncthreads();
        chain16_fixup_batch(a, brow, nrows, smem);
      }
    } else {   // queue overflowed: redo every row this CTA owns
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x)
        for (int r0 = 0; r0 < kTcBM; r0 += kT16RB) {
          const int row0 = mt * kTcBM + r0;
          if (row0 >= a.M) break;
          const int nrows = min(a.M - row0, kT16RB);
          __syncthreads();
          if ((int)threadIdx.x < nrows) brow[threadIdx.x] = row0 + threadIdx.x;
          __syncthreads();
          chain16_fixup_batch(a, brow, nrows, smem);
        }
    }
  }
}

bool linear_chain16_applies(int M, int K1, int N1, int N2) {
  return M > 0 && K1 > 0 && K1 <= kC16KB * kTcBK && K1 % 4 == 0 && N1 == kC16N1 && N2 > 0 && N2 <= kC16MaxN2 &&
         N2 % 4 == 0;
}

// packed1 / packed2: linear_prepack16 buffers of W1 [K1,128] / W2 [128,N2]
int linear_chain_tc16_launch(const float* x, int ldx, const void* packed1, const float* scale1, const float* shift1,
                             int act1, const void* packed2, const float* scale2, const float* shift2, int act2,
                             float* y, int ldy, int M, int K1, int N1, int N2, cudaStream_t st) {
  if (!linear_chain16_applies(M, K1, N1, N2)) return DH3D_ERR_UNSUPPORTED;
  int rc = t16_check(x, ldx, packed1, M, K1, N1);
  if (rc != DH3D_OK) return rc;
  if (!packed2 || !y) return DH3D_ERR_NULL;
  if (ldy % 4 || ldy < N2) return DH3D_ERR_DIM;
  if ((((uintptr_t)y | (uintptr_t)packed2 | (uintptr_t)scale1 | (uintptr_t)shift1 | (uintptr_t)scale2 |
        (uintptr_t)shift2) & 15) != 0)
    return DH3D_ERR_ALIGN;
  const T16Packed p1 = t16_unpack(packed1, K1, N1), p2 = t16_unpack(packed2, N1, N2);
  const int Kp1 = t16_kp(K1), Kp2 = t16_kp(N1);
  CUtensorMap mx, m1h, m1l, m2h, m2l, my;
  if ((rc = make_map(&mx, x, M, K1, ldx, kTcBM)) != DH3D_OK) return rc;
  if ((rc = make_map_f16(&m1h, p1.wh, N1, Kp1, Kp1, 128)) != DH3D_OK) return rc;
  if ((rc = make_map_f16(&m1l, p1.wl, N1, Kp1, Kp1, 128)) != DH3D_OK) return rc;
  if ((rc = make_map_f16(&m2h, p2.wh, N2, Kp2, Kp2, 128)) != DH3D_OK) return rc;
  if ((rc = make_map_f16(&m2l, p2.wl, N2, Kp2, Kp2, 128)) != DH3D_OK) return rc;
  if ((rc = make_map(&my, y, M, N2, ldy, 32)) != DH3D_OK) return rc;
  ChainArgs a{scale1, shift1, p1.cs, act1, scale2, shift2, p2.cs, act2, x, ldx, p1.wh, p1.wl, Kp1,
              p2.wh, p2.wl, Kp2, y, ldy, M, K1, N2};
  cudaError_t e = cudaFuncSetAttribute(gemm_chain16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C16Cfg::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int num_mt = ceil_div(M, kTcBM);
  const int grid = num_mt < num_sms() ? num_mt : num_sms();
  gemm_chain16_kernel<<<grid, kC16Threads, C16Cfg::kSmemBytes, st>>>(mx, m1h, m1l, m2h, m2l, my, a);
  return launch_status();
}

}  // namespace dh3d
