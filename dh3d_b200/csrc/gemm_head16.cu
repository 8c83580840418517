// Attention heads  y[m] = act2( sum_n act((X @ W)[m,n] * scale[n] + shift[n]) * w2[n] + b2 )  for WIDE hidden layers
// (N > 256, K <= 256: the 256 -> 1024 -> 1 stacks of detection_block / globalatt_block, core/backbones.py:132-173) with
// the activation tile RESIDENT in shared memory.
//
// The streaming kernel (gemm_tc16.cu) re-loads and re-splits the X tile once per 256-column N pass: for the heads that is
// 4 x 128 KB of L2 -> SM traffic per 128 rows and 4 x the fp32 -> (xh, xl) conversion, and ncu's source view showed the
// four split warps busy ~100 % of the kernel's lifetime (40 % of all stall samples) -- the tensor pipe waited for them
// (75 % active), not for memory: tcgen05.mma 128 x 256 x 16 issues every 128 cycles whatever else the SM's shared
// memory is doing (scripts/ubench/mma_rate.cu).  Here the tile's fp16 pair (8 K slabs x [xh | xl] = 128 KB) stays in
// shared memory for all N passes:
//   * the raw fp32 slab kb of the NEXT tile is loaded by TMA straight into slab kb's 16 KB as soon as the last pass's
//     MMAs on that slab have completed (tcgen05.commit -> a_empty[kb]) and converted IN PLACE (every split thread holds
//     its part of the raw slab in registers before any of them writes, one named barrier per slab), so no extra
//     staging memory is needed and the conversion of tile i + 1 hides under the last pass of tile i and the first
//     pass of tile i + 1;
//   * only the W tiles stream: a 5-stage ring of this CTA's N half [128 x 32] of W_h^T / W_l^T (16 KB per stage);
//   * CTA pairs issue tcgen05.mma.cta_group::2 as in gemm_tc16.cu (leader's MMA thread; the peer's warp 1 relays its
//     "W slice landed" to the leader; split and epilogue warps of the peer arrive on the leader's barriers).
// X is read from L2 once (128 KB per 128 rows instead of 512 KB), the split runs once, L2 -> SM traffic per tile drops
// from 1 MB to 640 KB.  Out-of-window rows: the same queue + fp32 recompute as gemm_tc16.cu.
#include "gemm_tc16.cuh"

namespace dh3d {

constexpr int kH16Threads = 384;   // warp 0: W producer, 1: MMA issuer (leader) / relay (peer), 2: X producer, 3: idle,
                                   // 4-7: split, 8-11: epilogue
constexpr int kH16BN = 256;
constexpr int kH16MaxKB = 8;       // K <= 256
constexpr int kH16SB = 5;          // W ring stages

struct H16Cfg {
  static constexpr uint32_t kSlabBytes = kTcBM * kTcBK * 4;               // 16 KB: raw fp32, then xh | xl
  static constexpr uint32_t kABytes = kTcBM * kTcBK * 2;                  // 8 KB
  static constexpr uint32_t kBBytes = (kH16BN / 2) * kTcBK * 2;           // 8 KB: this CTA's N half of W_h^T (or W_l^T)
  static constexpr uint32_t kBStageBytes = 2 * kBBytes;
  static constexpr uint32_t kParamBytes = 2 * 3 * kH16BN * 4;
  static constexpr uint32_t kBarBytes = 512;
  static constexpr uint32_t kSmemBytes = kH16MaxKB * kSlabBytes + kH16SB * kBStageBytes + kParamBytes + kBarBytes +
                                         kT16BadBytes + 1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
  static_assert((3 * kH16MaxKB + 3 * kH16SB + 4) * 8 + 8 <= kBarBytes, "barrier block");
};

__global__ void __launch_bounds__(kH16Threads, 1)
gemm_head16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                   const __grid_constant__ CUtensorMap tmBl, const T16Epilogue ep, int M, int K, int N) {
  using Cfg = H16Cfg;
  constexpr int BN = kH16BN, SB = kH16SB, MC = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bring = smem + kH16MaxKB * Cfg::kSlabBytes;
  float* params = reinterpret_cast<float*>(bring + SB * Cfg::kBStageBytes);   // [2][3][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(params) + Cfg::kParamBytes);
  uint64_t* a_full = bars;                       // [8] raw slab landed (this CTA)               (count 1 + tx)
  uint64_t* a_ready = bars + kH16MaxKB;          // [8] xh | xl written, LEADER's               (count 4 * MC)
  uint64_t* a_empty = bars + 2 * kH16MaxKB;      // [8] last pass's MMAs on the slab finished     (count 1, commit)
  uint64_t* b_full = bars + 3 * kH16MaxKB;       // [SB] this CTA's W slices landed               (count 1 + tx)
  uint64_t* b_peer = b_full + SB;                // [SB] the peer's W slices landed, LEADER's     (count 1)
  uint64_t* b_empty = b_full + 2 * SB;           // [SB] MMAs reading the stage finished           (count 1, commit)
  uint64_t* tmem_full = b_full + 3 * SB;         // [2]                                            (count 1, commit)
  uint64_t* tmem_empty = tmem_full + 2;          // [2] LEADER's                                   (count 4 * MC)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* bad = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + Cfg::kBarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + kTcBK - 1) / kTcBK;
  const int num_mt = (M + kTcBM - 1) / kTcBM;
  const int num_nt = (N + BN - 1) / BN;
  const uint32_t crank = cluster_ctarank();
  const int mt_begin = (int)(blockIdx.x / MC) * MC;
  const int mt_stride = (int)gridDim.x;

  auto slab = [&](int kb) { return smem + kb * Cfg::kSlabBytes; };
  auto stage_bh = [&](int s) { return bring + s * Cfg::kBStageBytes; };
  auto stage_bl = [&](int s) { return bring + s * Cfg::kBStageBytes + Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    bad[0] = 0u;
    for (int k = 0; k < kH16MaxKB; ++k) {
      mbar_init(&a_full[k], 1);
      mbar_init(&a_ready[k], 4 * MC);
      mbar_init(&a_empty[k], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_peer[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4 * MC);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ W producer (this CTA's N half of every W tile)
    if (lane == 0) {
      uint32_t it = 0;
      for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride)
        for (int nt = 0; nt < num_nt; ++nt)
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % SB;
            const uint32_t ph = (it / SB) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&b_full[s], Cfg::kBStageBytes);
            tma_load_2d(stage_bh(s), &tmBh, kb * kTcBK, nt * BN + (int)crank * (BN / MC), &b_full[s]);
            tma_load_2d(stage_bl(s), &tmBl, kb * kTcBK, nt * BN + (int)crank * (BN / MC), &b_full[s]);
          }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader, for the pair)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((kTcBM * MC) >> 4) << 24);
      uint32_t it = 0, tile = 0, tl = 0;
      for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride, ++tl)
        for (int nt = 0; nt < num_nt; ++nt, ++tile) {
          const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
          mbar_wait(&tmem_empty[acc], aph ^ 1);
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % SB;
            const uint32_t ph = (it / SB) & 1;
            if (nt == 0) mbar_wait(&a_ready[kb], tl & 1);
            mbar_wait(&b_full[s], ph);
            mbar_wait(&b_peer[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t a_h = umma_desc_sw64(smem_u32(slab(kb)));
            const uint64_t a_l = umma_desc_sw64(smem_u32(slab(kb) + Cfg::kABytes));
            const uint64_t b_h = umma_desc_sw64(smem_u32(stage_bh(s)));
            const uint64_t b_l = umma_desc_sw64(smem_u32(stage_bl(s)));
#pragma unroll
            for (int k = 0; k < kTcBK / 16; ++k) {
              const uint64_t off = (uint64_t)(k * 16 * 2) >> 4;
              umma_f16<MC>(tmem_d, a_l + off, b_h + off, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_f16<MC>(tmem_d, a_h + off, b_l + off, idesc, 1u);
              umma_f16<MC>(tmem_d, a_h + off, b_h + off, idesc, 1u);
            }
            umma_commit_x<MC>(&b_empty[s]);
            if (nt == num_nt - 1) umma_commit_x<MC>(&a_empty[kb]);   // the slab may take the next tile's rows
          }
          umma_commit_x<MC>(&tmem_full[acc]);
        }
    } else if (lane == 0) {
      // ---------------------------------------------------------------- peer: "my W slices landed" -> leader
      uint32_t it = 0;
      for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride)
        for (int nt = 0; nt < num_nt; ++nt)
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % SB;
            mbar_wait(&b_full[s], (it / SB) & 1);
            mbar_arrive_x<MC>(&b_peer[s]);
          }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ X producer: raw slabs into the resident tile
    if (lane == 0) {
      uint32_t tl = 0;
      for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride, ++tl) {
        const int mt = mtb + (int)crank;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_empty[kb], (tl & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[kb], Cfg::kSlabBytes);
          tma_load_2d(slab(kb), &tmA, kb * kTcBK, mt * kTcBM, &a_full[kb]);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ fp32 -> (xh | xl) in place, once per tile
    const int t = threadIdx.x - 128;  // 0..127: chunk t&3 of rows (t>>2) + 32*i (see gemm_tc16.cu)
    const int c = t & 3;
    uint32_t tl = 0;
    uint32_t rmax[4] = {0u, 0u, 0u, 0u};
    for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride, ++tl) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&a_full[kb], tl & 1);
        uint8_t* raw = slab(kb);
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
        if (lane == 0) mbar_arrive_x<MC>(&a_ready[kb]);
      }
      t16_queue_bad_rows(rmax, t, (mtb + (int)crank) * kTcBM, bad);
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue (warps 8..11)
    const int q = warp & 3;              // TMEM lane quadrant of this warp
    const int et = threadIdx.x - 256;    // 0..127
    uint32_t tile = 0;
    for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride) {
      const int mt = mtb + (int)crank;
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
      for (int nt = 0; nt < num_nt; ++nt, ++tile) {
        const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
        float* prm = params + acc * 3 * BN;
        for (int cc = et; cc < BN; cc += 128) {
          const int gc = nt * BN + cc;
          const bool in = gc < N;
          const float cs = in ? __ldg(ep.colscale + gc) : 0.f;
          prm[cc] = ((in && ep.scale) ? __ldg(ep.scale + gc) : 1.f) * cs;
          prm[BN + cc] = (in && ep.shift) ? __ldg(ep.shift + gc) : 0.f;
          prm[2 * BN + cc] = in ? __ldg(ep.w2 + gc) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tmem_full[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          DH3D_TMEM_LD_32X32(r, taddr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 32 >= BN) {  // accumulator fully read: hand it back to the MMA thread
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_x<MC>(&tmem_empty[acc]);
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), prm[c0 + j], prm[BN + c0 + j]);
          tc_act32(v, ep.act);
#pragma unroll
          for (int j = 0; j < 32; ++j) dot[j & 3] = fmaf(v[j], prm[2 * BN + c0 + j], dot[j & 3]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // params[acc] may be rewritten two tiles later
      }
      const int row = mt * kTcBM + q * 32 + lane;
      if (row < M) ep.y2[row] = tc_act((dot[0] + dot[1]) + (dot[2] + dot[3]) + ep.b2, ep.act2);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while its peer can still signal its barriers or read its shared memory
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  if (bad[0] != 0u)
    t16_fixup_all<true>(ep, bad, M, K, N, smem, [&](int t) {
      const int mt = mt_begin + t * mt_stride;
      return mt < num_mt ? mt + (int)crank : -1;
    });
}

// Shapes the resident-tile kernel covers (the caller falls back to the streaming kernel otherwise)
bool linear_rowdot_head16_applies(int M, int K, int N) {
  return K <= kH16MaxKB * kTcBK && N > kH16BN && ceil_div(M, kTcBM) >= 2 * 16;
}

int linear_rowdot_head16_launch(const float* x, int ldx, const void* packed, const float* scale, const float* shift,
                                int act, const float* w2, float b2, int act2, float* y2, int M, int K, int N,
                                cudaStream_t st) {
  int rc = t16_check(x, ldx, packed, M, K, N);
  if (rc != DH3D_OK) return rc;
  if (!w2 || !y2) return DH3D_ERR_NULL;
  if (!linear_rowdot_head16_applies(M, K, N)) return DH3D_ERR_UNSUPPORTED;
  const T16Packed p = t16_unpack(packed, K, N);
  T16Epilogue ep{scale, shift, p.cs, act, w2, b2, act2, y2, x, ldx, p.wh, p.wl, t16_kp(K), nullptr, 0};
  CUtensorMap ma, mh, ml;
  const int Kp = t16_kp(K);
  if ((rc = make_map(&ma, x, M, K, ldx, kTcBM)) != DH3D_OK) return rc;
  if ((rc = make_map_f16(&mh, p.wh, N, Kp, Kp, kH16BN / 2)) != DH3D_OK) return rc;
  if ((rc = make_map_f16(&ml, p.wl, N, Kp, Kp, kH16BN / 2)) != DH3D_OK) return rc;
  cudaError_t e = cudaFuncSetAttribute(gemm_head16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)H16Cfg::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int num_mt = ceil_div(M, kTcBM);
  int grid = num_mt < num_sms() ? num_mt : num_sms();
  grid = ceil_div(grid, 2) * 2;
  if (grid > num_sms()) grid -= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kH16Threads);
  cfg.dynamicSmemBytes = H16Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, gemm_head16_kernel, ma, mh, ml, ep, M, K, N);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

}  // namespace dh3d
