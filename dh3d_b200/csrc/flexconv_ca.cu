// Fused FlexConv with asynchronously staged neighbours (default FlexConv path on B200):
//   neighbour rows -> shared memory by cp.async (LDGSTS, 16 B per thread, G items deep, no registers held)
//   -> 4*Din moments in registers -> swizzled smem -> tcgen05 contraction, one kernel.
//   (reference: user_ops/kernels/flex_conv_kernel_gpu.cu.cc:44-158; algebra in flexconv.cu)
//
//   out[n,:] = act( (A[n,:] @ Theta_ext) * scale + shift ),
//   A[n, p'*Din + c] = sum_k (1,dx,dy,dz)[p'] * f[nbr(n,k), c],   Theta_ext = [bias; theta_x; theta_y; theta_z]
//
// The gather is latency-bound: what matters is how many neighbour bytes are in flight per SM.  Three stagings were
// built and measured on B200 (profiles/flexconv_staging_ab_r2f.json, ncu, 64->64 x 262144 points / 128->128 x 65536):
//   per-thread LDG into registers   215 / 135 us: ~20-30 KB in flight (8 gather warps x registers), 22 % warps active
//   TMA tile::gather4 (4 rows/TMA)  240 / 127 us: no registers held, but 2.5x the instructions for the coordinate /
//                                   descriptor bookkeeping and a higher latency per stalled issue
//   cp.async ring (this file)       164 /  85 us -- kept; the other two kernels were deleted after that capture.
//   here             every consumer thread (point x 8 channels) copies exactly the 32 bytes it will read itself
//                    with two 16-byte cp.async, G-1 items (= neighbour slots) ahead of the one it is reducing:
//                    512 threads x (G-1) x 32 B in flight, a pure per-thread software pipeline
//                    (cp.async.commit_group / wait_group) with no barrier and no cross-thread visibility.
// Each 16-byte copy instruction of a warp covers whole 64-byte half rows (full 32-byte sectors); the stage
// layout is private to the thread that wrote it, chosen so that both its LDS.128 are conflict-free.
//
// Two consumer loops share the rest of the kernel: the generic one (any K, batches of 8 neighbour slots, dynamic
// ring / table cursor) and the K == 8 one used by every FlexConv of DH3D (template K8: fully unrolled schedule with
// static ring slots, (row = lane & 7, segment = lane >> 3) lane mapping that makes every shared-memory access
// conflict-free under the 128B swizzle, packed fp32x2 moment updates; 83 M -> 36 M warp instructions at 64 -> 64).
//
// 704 threads, one persistent CTA per SM:
//   warp 0      TMA producer of the Theta_ext^T hi/lo tiles      warp 1      tcgen05.mma issuer
//   warps 2-17  consumers: cp.async issue + moments               warps 18-21 epilogue (tcgen05.ld -> bias /
//                                                                             folded BN / ReLU -> TMA store)
#include <stdlib.h>

#include "tc_common.cuh"

namespace dh3d {

constexpr int kCAThreads = 704;
constexpr int kCAConsumerWarps = 16;
constexpr int kCAKB = 8;                        // neighbour slots per offset-table batch
constexpr uint32_t kCAStageBytes = 128 * 128;   // 128 rows x 32 channels fp32

// K8 = the K == 8 fast path (every FlexConv of DH3D): the neighbour loop is fully unrolled, so the ring slot, the
// table slot and the operand stage of every item are compile-time constants and the issue cursor advances at one
// static point per group (see the consumer branch).
// NB = K / 8 for the unrolled 8-slot schedule (K = 8, 16, 32), 0 = generic loop (any K).
// NF = output tiles fed from ONE set of moment slabs: for Dout = 2 * BN the gather used to run once per N tile; with
// NF = 2 every A slab is multiplied against both Theta tiles (two accumulators live, all 512 TMEM columns with the
// double buffering), so the neighbour rows are gathered once.
template <int BN, int NB, int NF = 1>
struct CACfg {
  static constexpr bool K8 = NB > 0;
  static constexpr int kStages = 2;                           // UMMA A-operand stages (moment slabs, hi + lo)
  // Theta tiles have their OWN ring, decoupled from the A stages: with the tiles inside the A stages (r1) the TMA of
  // slab i + 2 could only be issued once the MMAs of slab i had completed, so every slab paid an L2 round trip
  static constexpr int kBStages = NF > 1 ? 3 : (K8 ? (BN <= 64 ? 4 : 2) : 2);
  // gather stages (16 KB each): ring depth 6 / 3 / 2 measured within 10 % of each other (profiles/flexconv_*_r2s.txt)
  static constexpr int kGStages = NF > 1 ? 2 : (K8 ? 4 : (BN <= 64 ? 5 : 3));
  static constexpr int kOutRows = K8 ? 16 : 32;  // rows per epilogue TMA store (smem budget)
  static constexpr uint32_t kOutBytes = 4 * kOutRows * 32 * 4;
  static constexpr uint32_t kBBytes = BN * kTcBK * 4;
  static constexpr uint32_t kStageBytes = 2 * kTcABytes;
  static constexpr uint32_t kBStageBytes = 2 * kBBytes;
  static constexpr uint32_t kDeltaBytes = 128 * kCAKB * 16;   // float4 per (row, slot)
  static constexpr uint32_t kIdxBytes = 128 * kCAKB * 4;      // global feature row per (row, slot)
  static constexpr uint32_t kParamBytes = 2 * 2 * NF * BN * 4;   // double-buffered scale/shift slices
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBStages * kBStageBytes + kGStages * kCAStageBytes +
                                         kDeltaBytes + kIdxBytes + kOutBytes + kParamBytes + 256 /*barriers*/ +
                                         1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
  static constexpr uint32_t kTmemCols = 2 * NF * BN < 32 ? 32 : 2 * NF * BN;
  static_assert(kTmemCols <= 512 && (NF == 1 || NB > 0), "NF > 1: unrolled schedule only, 512 TMEM columns");
};

template <int BN, int NB, int NF = 1>
__global__ void __launch_bounds__(kCAThreads, 1)
flexconv_ca_kernel(const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmY,
                   const CAArgs a) {
  using Cfg = CACfg<BN, NB, NF>;
  constexpr bool K8 = NB > 0;
  constexpr int W = NF * BN;          // output columns per accumulator set
  constexpr int S = Cfg::kStages;
  constexpr int SB = Cfg::kBStages;
  constexpr int G = Cfg::kGStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bring = smem + S * Cfg::kStageBytes;                            // Theta ring, 1024-aligned
  uint8_t* gbase = bring + SB * Cfg::kBStageBytes;                         // gather stages, 1024-aligned
  float4* sdelta = reinterpret_cast<float4*>(gbase + G * kCAStageBytes);   // [128][kCAKB]
  int* sidx = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(sdelta) + Cfg::kDeltaBytes);  // [128][kCAKB]
  uint8_t* out_stage = reinterpret_cast<uint8_t*>(sidx) + Cfg::kIdxBytes;
  float* params = reinterpret_cast<float*>(out_stage + Cfg::kOutBytes);  // [2][2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(params) + Cfg::kParamBytes);
  uint64_t* afull = bars;               // [S]  A hi/lo slab written         (count 16, one per consumer warp)
  uint64_t* empty = bars + S;           // [S]  MMAs reading the A stage done (count 1, tcgen05.commit)
  uint64_t* bfull = bars + 2 * S;       // [SB] Theta tiles landed            (count 1 + tx)
  uint64_t* bempty = bars + 2 * S + SB; // [SB] MMAs reading the tiles done   (count 1, tcgen05.commit)
  uint64_t* tmem_full = bars + 2 * S + 2 * SB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_mt = (a.rows + kTcBM - 1) / kTcBM;
  const int num_nt = (a.Dout + W - 1) / W;   // accumulator sets (NF tiles of BN columns each) per row tile
  const int num_cg = a.Din / kTcBK;     // 32-channel groups; 4 K-slabs each
  const int num_kb = 4 * num_cg;
  const int num_batches = (a.K + kCAKB - 1) / kCAKB;

  auto stage_a = [&](int s) { return smem + s * Cfg::kStageBytes; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::kStageBytes + kTcABytes; };
  auto stage_bhi = [&](int s) { return bring + s * Cfg::kBStageBytes; };
  auto stage_blo = [&](int s) { return bring + s * Cfg::kBStageBytes + Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&afull[s], kCAConsumerWarps);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (Theta tiles)
    if (lane == 0) {
      uint32_t it = 0;
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x)
        for (int nt = 0; nt < num_nt; ++nt)
          for (int cg = 0; cg < num_cg; ++cg)
            for (int p = 0; p < 4; ++p)
              for (int f = 0; f < NF; ++f, ++it) {
                const int s = it % SB;
                const uint32_t ph = (it / SB) & 1;
                mbar_wait(&bempty[s], ph ^ 1);
                mbar_arrive_expect_tx(&bfull[s], 2 * Cfg::kBBytes);
                const int k0 = p * a.Din + cg * kTcBK;  // row block of Theta_ext == column block of Theta_ext^T
                tma_load_2d(stage_bhi(s), &tmBhi, k0, nt * W + f * BN, &bfull[s]);
                tma_load_2d(stage_blo(s), &tmBlo, k0, nt * W + f * BN, &bfull[s]);
              }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(kTcBM >> 4) << 24);
      uint32_t it = 0, itb = 0, tile = 0;
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x)
        for (int nt = 0; nt < num_nt; ++nt, ++tile) {
          const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
          mbar_wait(&tmem_empty[acc], aph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + acc * W;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % S;
            mbar_wait(&afull[s], (it / S) & 1);
            const uint64_t a_hi = umma_desc_sw128(smem_u32(stage_a(s)));
            const uint64_t a_lo = umma_desc_sw128(smem_u32(stage_alo(s)));
#pragma unroll
            for (int f = 0; f < NF; ++f, ++itb) {
              const int sb = itb % SB;
              mbar_wait(&bfull[sb], (itb / SB) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_bhi(sb)));
              const uint64_t b_lo = umma_desc_sw128(smem_u32(stage_blo(sb)));
#pragma unroll
              for (int k = 0; k < kTcBK / 8; ++k) {
                const uint64_t off = (uint64_t)(k * 8 * 4) >> 4;
                umma_tf32(tmem_d + f * BN, a_lo + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_tf32(tmem_d + f * BN, a_hi + off, b_lo + off, idesc, 1u);
                umma_tf32(tmem_d + f * BN, a_hi + off, b_hi + off, idesc, 1u);
              }
              umma_commit(&bempty[sb]);
            }
            umma_commit(&empty[s]);
          }
          umma_commit(&tmem_full[acc]);
        }
    }
  } else if (warp < 18) {
   if constexpr (K8) {
    // ------------------------------------------------------------------ consumers, K == 8 (16 warps x 8 rows)
    // lane = (row-in-warp rw = lane & 7, channel segment seg = lane >> 3): a quarter-warp is 8 rows x one
    // 16-byte chunk, so with the 128B-swizzle chunk index (chunk ^ rw) every LDS.128 / STS.128 below -- gather
    // ring, A slabs, offset table -- touches 8 distinct 16-byte columns: no bank conflicts (the (row, seg)
    // mapping of the generic path had 2-way conflicts on all three; ncu r1j: 10.8 M of 26 M wavefronts).
    // Item i of a group is neighbour slot k = i; the item issued while item k is being reduced is k + AH (next
    // group when k + AH >= 8), static under the unroll; the ring positions are two running byte offsets.
    constexpr int AH = G - 1;     // items in flight per thread
    constexpr uint32_t kRing = G * kCAStageBytes;
    const int wc = warp - 2;
    const int rw = lane & 7, seg = lane >> 3;
    const int r = 8 * wc + rw;
    const uint32_t o0 = r * 128 + ((seg ^ rw) << 4), o1 = r * 128 + (((seg + 4) ^ rw) << 4);
    float4* drow = sdelta + r * kCAKB;   // slot k at drow[k ^ rw]
    int* irow = sidx + r * kCAKB;        // slot k at irow[k ^ rw]
    const int groups = num_nt * num_cg;

    // ---- issue side: tile i_mt, group (i_nt, i_cg); entering a tile fills the index table of this warp's rows
    // and starts the loads the offset table of that tile will need three items later
    int i_mt = blockIdx.x, i_cg = 0, i_g = 0, i_b = 0;
    constexpr int nbatch = NB > 0 ? NB : 1;   // K == 8: one batch, the tables of a tile serve all its groups
    float nx[6], pc[3];
    auto enter_batch = [&](int mt, int b) {
      const int row = mt * kTcBM + r;
      const int rr = row < a.rows ? row : 0;   // tail rows gather row 0 (their outputs are clipped)
      const int cloud0 = (rr / a.n_per_cloud) * a.n_per_cloud;
      const int2 nb = __ldg(reinterpret_cast<const int2*>(a.nbr + (long long)rr * a.K + 8 * b + 2 * seg));
      const int v0 = cloud0 + nb.x, v1 = cloud0 + nb.y;
      __syncwarp();
      irow[(2 * seg) ^ rw] = v0;
      irow[(2 * seg + 1) ^ rw] = v1;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        nx[c] = __ldg(a.xyz + (long long)v0 * 3 + c);
        nx[3 + c] = __ldg(a.xyz + (long long)v1 * 3 + c);
        pc[c] = __ldg(a.xyz + (long long)rr * 3 + c);
      }
    };
    uint32_t i_off = 0, c_off = 0;   // ring byte offsets of the next item to issue / to consume
    auto issue = [&](int k) {
      if (i_mt < num_mt) {
        const float* src = a.feat + (long long)irow[k ^ rw] * a.Din + i_cg * kTcBK + seg * 4;
        uint8_t* dst = gbase + i_off;
        cp_async16(dst + o0, src);
        cp_async16(dst + o1, src + 16);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");  // one group per item, empty past the end
      i_off += kCAStageBytes;
      if (i_off == kRing) i_off = 0;
    };
    // the issue cursor crosses a batch boundary: next batch of the group, else next group / tile; K > 8 refills the
    // tables at every batch (loaded AH items before the consume side needs them), K == 8 once per tile
    auto next_batch = [&]() {
      bool new_tile = false;
      if (++i_b == nbatch) {
        i_b = 0;
        if (++i_cg == num_cg) i_cg = 0;
        if (++i_g == groups) {
          i_g = 0;
          i_mt += gridDim.x;
          new_tile = true;
        }
      }
      if (i_mt < num_mt && (nbatch > 1 || new_tile)) enter_batch(i_mt, i_b);
    };
    if (i_mt < num_mt) enter_batch(i_mt, 0);
#pragma unroll
    for (int i = 0; i < AH; ++i) issue(i);

    // ---- consume side
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
      for (int g = 0; g < groups; ++g) {
        unsigned long long m[4][4];  // [moment p'][channel pair]
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < 4; ++c) m[p][c] = 0ull;
        for (int b = 0; b < nbatch; ++b) {
          if (nbatch > 1 || g == 0) {
            // offset table of this batch from the coordinates loaded when the issue side entered it
            __syncwarp();
            drow[(2 * seg) ^ rw] = make_float4(nx[0] - pc[0], nx[1] - pc[1], nx[2] - pc[2], 0.f);
            drow[(2 * seg + 1) ^ rw] = make_float4(nx[3] - pc[0], nx[4] - pc[1], nx[5] - pc[2], 0.f);
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            asm volatile("cp.async.wait_group %0;" ::"n"(AH - 1) : "memory");  // this thread's item k landed
            const uint8_t* gs = gbase + c_off;
            c_off += kCAStageBytes;
            if (c_off == kRing) c_off = 0;
            const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(gs + o0);
            const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(gs + o1);
            const float4 d = drow[k ^ rw];
            const unsigned long long fv[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              fadd2(m[0][c], fv[c]);
              ffma2(m[1][c], fv[c], d.x);
              ffma2(m[2][c], fv[c], d.y);
              ffma2(m[3][c], fv[c], d.z);
            }
            if (k == 8 - AH) next_batch();
            issue((k + AH) & 7);
          }
        }
        // 4 K-slabs (p' = 1, x, y, z) -> operand stages p' & 1, swizzled K-major, hi (raw) + lo
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int s = p & 1;
          mbar_wait(&empty[s], (uint32_t)(p >> 1) ^ 1u);   // slab counter 4g + p: parity (p >> 1) for S == 2
          uint8_t* ah = stage_a(s);
          uint8_t* al = stage_alo(s);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t off = j ? o1 : o0;
            const float2 va = unpack2(m[p][2 * j]), vb = unpack2(m[p][2 * j + 1]);
            const float4 v = make_float4(va.x, va.y, vb.x, vb.y);
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(ah + off) = v;
            *reinterpret_cast<float4*>(al + off) = l;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&afull[s]);
        }
      }
    }
   } else {
    // ------------------------------------------------------------------ consumers, any K (16 warps x 8 rows)
    const int wc = warp - 2;      // consumer warp: rows 8*wc .. 8*wc+7 of the tile
    const int seg = lane & 3;     // this thread's channels: 4seg..4seg+3 and 16+4seg..16+4seg+3 of the group
    const int r = 8 * wc + (lane >> 2);
    const uint32_t roff = r * 128;
    // private stage layout: the two 16-byte pieces of row r sit at chunk seg ^ 4(r&1) and (seg+4) ^ 4(r&1),
    // so a quarter-warp (2 rows x 4 segs) touches 8 distinct 16-byte columns in either access
    const uint32_t g0 = roff + ((seg ^ ((r & 1) << 2)) << 4), g1 = roff + (((seg + 4) ^ ((r & 1) << 2)) << 4);
    // canonical 128B swizzle of the UMMA A slab for the same channels (16-byte chunks seg and seg+4)
    const uint32_t c0 = (seg ^ (r & 7)) << 4, c1 = ((seg + 4) ^ (r & 7)) << 4;

    // ---- issue side: a cursor over the item sequence (tile, nt, cg, batch, k), G-1 items ahead
    int i_mt = blockIdx.x, i_nt = 0, i_cg = 0, i_b = 0, i_k = 0;
    uint32_t iit = 0;
    auto issue_one = [&]() {
      if (i_mt < num_mt) {
        if (i_k == 0 && (num_batches > 1 || (i_nt == 0 && i_cg == 0))) {
          // global feature rows of this warp's 8 points x 8 slots -> its private part of the index table
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int e = lane * 2 + u, rl = e >> 3, kk = e & 7;
            const int row = i_mt * kTcBM + 8 * wc + rl;
            const int rr = row < a.rows ? row : 0;   // tail rows gather row 0 (their outputs are clipped)
            int v = 0;
            if (i_b * kCAKB + kk < a.K)
              v = (rr / a.n_per_cloud) * a.n_per_cloud + __ldg(a.nbr + (long long)rr * a.K + i_b * kCAKB + kk);
            sidx[(8 * wc + rl) * kCAKB + kk] = v;
          }
          __syncwarp();
        }
        const float* src = a.feat + (long long)sidx[r * kCAKB + i_k] * a.Din + i_cg * kTcBK + seg * 4;
        uint8_t* dst = gbase + (iit % G) * kCAStageBytes;
        cp_async16(dst + g0, src);
        cp_async16(dst + g1, src + 16);
        ++iit;
        if (++i_k == min(kCAKB, a.K - i_b * kCAKB)) {
          i_k = 0;
          if (++i_b == num_batches) {
            i_b = 0;
            if (++i_cg == num_cg) {
              i_cg = 0;
              if (++i_nt == num_nt) { i_nt = 0; i_mt += gridDim.x; }
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");  // one group per item, empty past the end
    };
    for (int i = 0; i < G - 1; ++i) issue_one();

    // ---- consume side
    uint32_t it = 0, git = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
      const int row = mt * kTcBM + r;
      const int rr = row < a.rows ? row : 0;
      const long long cloud0 = (long long)(rr / a.n_per_cloud) * a.n_per_cloud;
      const float px = __ldg(a.xyz + (long long)rr * 3), py = __ldg(a.xyz + (long long)rr * 3 + 1),
                  pz = __ldg(a.xyz + (long long)rr * 3 + 2);
      for (int nt = 0; nt < num_nt; ++nt)
        for (int cg = 0; cg < num_cg; ++cg) {
          float m[4][8];  // [moment p'][channel]
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) m[p][c] = 0.f;
          for (int b = 0; b < num_batches; ++b) {
            const int k0 = b * kCAKB;
            // offset table of this batch (once per tile when K <= 8): thread (r, seg) fills slots 2seg, 2seg+1
            // of its row; only the 4 threads of the row (same warp) read them
            if (num_batches > 1 || (nt == 0 && cg == 0)) {
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int k = 2 * seg + u;
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + k < a.K) {
                  const long long gs = cloud0 + __ldg(a.nbr + (long long)rr * a.K + k0 + k);
                  d.x = __ldg(a.xyz + gs * 3) - px;
                  d.y = __ldg(a.xyz + gs * 3 + 1) - py;
                  d.z = __ldg(a.xyz + gs * 3 + 2) - pz;
                }
                sdelta[r * kCAKB + k] = d;
              }
              __syncwarp();
            }
            const int kcnt = min(kCAKB, a.K - k0);
            for (int k = 0; k < kcnt; ++k, ++git) {
              asm volatile("cp.async.wait_group %0;" ::"n"(G - 2) : "memory");  // this thread's item git landed
              const uint8_t* gs = gbase + (git % G) * kCAStageBytes;
              const float4 f0 = *reinterpret_cast<const float4*>(gs + g0);
              const float4 f1 = *reinterpret_cast<const float4*>(gs + g1);
              const float4 d = sdelta[r * kCAKB + k];
              const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                m[0][c] += fv[c];
                m[1][c] = fmaf(d.x, fv[c], m[1][c]);
                m[2][c] = fmaf(d.y, fv[c], m[2][c]);
                m[3][c] = fmaf(d.z, fv[c], m[3][c]);
              }
              issue_one();
            }
          }
          // 4 K-slabs (p' = 1, x, y, z) -> consecutive UMMA stages, swizzled K-major, hi (raw) + lo
#pragma unroll
          for (int p = 0; p < 4; ++p, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* ah = stage_a(s) + roff;
            uint8_t* al = stage_alo(s) + roff;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint32_t off = j ? c1 : c0;
              const float4 v = make_float4(m[p][4 * j], m[p][4 * j + 1], m[p][4 * j + 2], m[p][4 * j + 3]);
              float4 l;
              l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
              l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
              l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
              l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
              *reinterpret_cast<float4*>(ah + off) = v;
              *reinterpret_cast<float4*>(al + off) = l;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[s]);
          }
        }
    }
   }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 18..21)
    const int q = warp & 3;
    const int et = threadIdx.x - 576;  // 0..127
    constexpr int OR = Cfg::kOutRows;                   // rows per TMA store: 32, or 16 in two passes
    uint8_t* my_stage = out_stage + (warp - 18) * (OR * 128);
    uint32_t tile = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x)
      for (int nt = 0; nt < num_nt; ++nt, ++tile) {
        const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
        float* prm = params + acc * 2 * W;
        for (int c = et; c < W; c += 128) {
          const int gc = nt * W + c;
          const bool in = gc < a.Dout;
          prm[c] = (in && a.scale) ? __ldg(a.scale + gc) : 1.f;
          prm[W + c] = (in && a.shift) ? __ldg(a.shift + gc) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tmem_full[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0e = 0; c0e < W; c0e += 32) {
          uint32_t rg[32];
          const uint32_t taddr = tmem_base + acc * W + ((uint32_t)(q * 32) << 16) + (uint32_t)c0e;
          DH3D_TMEM_LD_32X32(rg, taddr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0e + 32 >= W) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (nt * W + c0e < a.Dout) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(rg[j]), prm[c0e + j], prm[W + c0e + j]);
            tc_act32(v, a.act);
#pragma unroll
            for (int pass = 0; pass < 32 / OR; ++pass) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
              if (lane / OR == pass) {
                const int rl = lane % OR;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(my_stage + rl * 128 + ((j ^ (rl & 7)) << 4)) =
                      make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmY, my_stage, nt * W + c0e, mt * kTcBM + q * 32 + pass * OR);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

template <int BN, int NB, int NF = 1>
static int launch_ca(const CAArgs& a, const float* thi, const float* tlo, float* out, cudaStream_t st) {
  using Cfg = CACfg<BN, NB, NF>;
  CUtensorMap mh, ml, my;
  int rc;
  const int Kd = 4 * a.Din;
  if ((rc = make_map(&mh, thi, a.Dout, Kd, Kd, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&ml, tlo, a.Dout, Kd, Kd, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&my, out, a.rows, a.Dout, a.Dout, Cfg::kOutRows)) != DH3D_OK) return rc;
  auto kern = flexconv_ca_kernel<BN, NB, NF>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int num_mt = ceil_div(a.rows, kTcBM);
  const int grid = num_mt < num_sms() ? num_mt : num_sms();
  kern<<<grid, kCAThreads, Cfg::kSmemBytes, st>>>(mh, ml, my, a);
  return launch_status();
}

// theta_packed = {Theta_ext^T hi [Dout, 4*Din], lo [Dout, 4*Din]} (flexconv.cu theta_ext_packed_kernel)
int flexconv_ca_launch(const float* feat, const float* xyz, const int32_t* nbr, const void* theta_packed,
                       const float* scale, const float* shift, int act, float* out, int rows, int n_per_cloud,
                       int K, int Din, int Dout, cudaStream_t st) {
  if (Din % kTcBK != 0 || Dout % 4 != 0 || K < 1) return DH3D_ERR_UNSUPPORTED;
  CAArgs a{feat, xyz, nbr, scale, shift, act, rows, n_per_cloud, K, Din, Dout};
  const float* thi = reinterpret_cast<const float*>(theta_packed);
  const float* tlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(theta_packed) +
                                                    align_up((size_t)4 * Din * Dout * sizeof(float), 256));
  if (((uintptr_t)nbr & 7) == 0) {   // unrolled 8-slot schedule, K / 8 batches per group
    if (K == 8 && Dout > 128 && Dout <= 256) return launch_ca<128, 1, 2>(a, thi, tlo, out, st);   // one gather, two N tiles
    if (K == 8) return Dout <= 64 ? launch_ca<64, 1>(a, thi, tlo, out, st) : launch_ca<128, 1>(a, thi, tlo, out, st);
    if (K == 16) return Dout <= 64 ? launch_ca<64, 2>(a, thi, tlo, out, st) : launch_ca<128, 2>(a, thi, tlo, out, st);
    if (K == 32) return Dout <= 64 ? launch_ca<64, 4>(a, thi, tlo, out, st) : launch_ca<128, 4>(a, thi, tlo, out, st);
  }
  if (Dout <= 64) return launch_ca<64, 0>(a, thi, tlo, out, st);
  return launch_ca<128, 0>(a, thi, tlo, out, st);
}

}  // namespace dh3d
