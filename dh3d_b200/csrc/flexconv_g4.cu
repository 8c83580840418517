// Fused FlexConv with TMA-gathered neighbours (default FlexConv path on B200):
//   neighbour rows -> shared memory by cp.async.bulk.tensor ... tile::gather4 (SASS UTMALDG.2D.GATHER4)
//   -> 4*Din moments in registers -> swizzled smem -> tcgen05 contraction, one kernel.
//   (reference: user_ops/kernels/flex_conv_kernel_gpu.cu.cc:44-158; algebra in flexconv.cu)
//
//   out[n,:] = act( (A[n,:] @ Theta_ext) * scale + shift ),
//   A[n, p'*Din + c] = sum_k (1,dx,dy,dz)[p'] * f[nbr(n,k), c],   Theta_ext = [bias; theta_x; theta_y; theta_z]
//
// flexconv_tc.cu gathers with per-thread loads and is latency-bound (ncu r1f: long-scoreboard stalls,
// 30 % issue, 19 % L2): the bytes in flight are capped by the registers of 8 gather warps.  Here G x 16 KB
// of neighbour rows are in flight per SM with no registers behind them: for every item = (32-channel group,
// neighbour slot k) each of the 16 consumer warps issues two gather4 that pull the k-th neighbour rows of
// its own 8 points (2 x 4 x 128 B, row indices straight from the k-NN table) into a 128-row stage with the
// 128B swizzle, G-1 items ahead of the one it is consuming.  (A single producer warp does not work: UTMALDG
// takes its operands from uniform registers, so 32 per-lane gather4 serialise into a ~85-cycle waterfall
// each -- measured 1.4x SLOWER than the per-thread gather; two per warp across 16 warps hide it.)
// A consumer thread (point x 8 channels) reads its neighbour with two conflict-free LDS.128, takes the offset
// (dx,dy,dz) from a per-tile table in shared memory (built once per tile by the warp that owns the rows,
// reused by every channel group), accumulates the 4 moments and writes the hi/lo K-slabs for the MMA warp.
//
// 704 threads, one persistent CTA per SM:
//   warp 0      TMA producer of the Theta_ext^T hi/lo tiles      warp 1      tcgen05.mma issuer
//   warps 2-17  consumers: gather4 issue + moments                warps 18-21 epilogue (tcgen05.ld -> bias /
//                                                                             folded BN / ReLU -> TMA store)
#include <stdlib.h>

#include "tc_common.cuh"

namespace dh3d {

constexpr int kG4Threads = 704;
constexpr int kG4ConsumerWarps = 16;
constexpr int kG4KB = 8;                        // neighbour slots per offset-table batch
constexpr uint32_t kG4StageBytes = 128 * 128;   // 128 rows x 32 channels fp32

template <int BN>
struct G4Cfg {
  static constexpr int kStages = 2;                           // UMMA operand stages
  static constexpr int kGStages = BN <= 64 ? 5 : 3;           // gather stages (16 KB each)
  static constexpr uint32_t kBBytes = BN * kTcBK * 4;
  static constexpr uint32_t kStageBytes = 2 * kTcABytes + 2 * kBBytes;
  static constexpr uint32_t kDeltaBytes = 128 * kG4KB * 16;   // float4 per (row, slot)
  static constexpr uint32_t kIdxBytes = 128 * kG4KB * 4;      // global feature row per (row, slot)
  static constexpr uint32_t kParamBytes = 2 * 2 * BN * 4;     // double-buffered scale/shift slices
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kGStages * kG4StageBytes + kDeltaBytes +
                                         kIdxBytes + kTcStageOutBytes + kParamBytes + 256 /*barriers*/ + 1024 /*align*/;
  static constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
};

struct G4Args {
  const float* xyz;      // [rows, 3]
  const int32_t* nbr;    // [rows, K]  (indices within the cloud)
  const float* scale;    // [Dout] or null
  const float* shift;    // [Dout] or null (feature bias already folded in)
  int act;
  int rows, n_per_cloud, K, Din, Dout;
};

__device__ __forceinline__ void tma_gather4(void* smem_dst, const CUtensorMap* map, int col, int r0, int r1,
                                            int r2, int r3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(r0),
      "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kG4Threads, 1)
flexconv_g4_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmY,
                   const G4Args a) {
  using Cfg = G4Cfg<BN>;
  constexpr int S = Cfg::kStages;
  constexpr int G = Cfg::kGStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* gbase = smem + S * Cfg::kStageBytes;                            // gather stages, 1024-aligned
  float4* sdelta = reinterpret_cast<float4*>(gbase + G * kG4StageBytes);   // [128][kG4KB]
  int* sidx = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(sdelta) + Cfg::kDeltaBytes);  // [128][kG4KB]
  uint8_t* out_stage = reinterpret_cast<uint8_t*>(sidx) + Cfg::kIdxBytes;
  float* params = reinterpret_cast<float*>(out_stage + kTcStageOutBytes);  // [2][2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(params) + Cfg::kParamBytes);
  uint64_t* bfull = bars;               // [S] Theta tiles landed          (count 1 + tx)
  uint64_t* afull = bars + S;           // [S] A hi/lo slab written        (count 16, one per consumer warp)
  uint64_t* empty = bars + 2 * S;       // [S] MMAs reading the stage done (count 1, tcgen05.commit)
  uint64_t* gfull = bars + 3 * S;       // [G] gathered rows landed        (count 16 + tx)
  uint64_t* gempty = bars + 3 * S + G;  // [G] consumers done              (count 16)
  uint64_t* tmem_full = bars + 3 * S + 2 * G;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_mt = (a.rows + kTcBM - 1) / kTcBM;
  const int num_nt = (a.Dout + BN - 1) / BN;
  const int num_cg = a.Din / kTcBK;     // 32-channel groups; 4 K-slabs each
  const int num_kb = 4 * num_cg;
  const int num_batches = (a.K + kG4KB - 1) / kG4KB;

  auto stage_a = [&](int s) { return smem + s * Cfg::kStageBytes; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::kStageBytes + kTcABytes; };
  auto stage_bhi = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes; };
  auto stage_blo = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes + Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&afull[s], kG4ConsumerWarps);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < G; ++g) {
      mbar_init(&gfull[g], kG4ConsumerWarps);
      mbar_init(&gempty[g], kG4ConsumerWarps);
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
            for (int p = 0; p < 4; ++p, ++it) {
              const int s = it % S;
              const uint32_t ph = (it / S) & 1;
              mbar_wait(&empty[s], ph ^ 1);
              mbar_arrive_expect_tx(&bfull[s], 2 * Cfg::kBBytes);
              const int k0 = p * a.Din + cg * kTcBK;  // row block of Theta_ext == column block of Theta_ext^T
              tma_load_2d(stage_bhi(s), &tmBhi, k0, nt * BN, &bfull[s]);
              tma_load_2d(stage_blo(s), &tmBlo, k0, nt * BN, &bfull[s]);
            }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(kTcBM >> 4) << 24);
      uint32_t it = 0, tile = 0;
      for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x)
        for (int nt = 0; nt < num_nt; ++nt, ++tile) {
          const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
          mbar_wait(&tmem_empty[acc], aph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&bfull[s], ph);
            mbar_wait(&afull[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t a_hi = umma_desc_sw128(smem_u32(stage_a(s)));
            const uint64_t a_lo = umma_desc_sw128(smem_u32(stage_alo(s)));
            const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_bhi(s)));
            const uint64_t b_lo = umma_desc_sw128(smem_u32(stage_blo(s)));
#pragma unroll
            for (int k = 0; k < kTcBK / 8; ++k) {
              const uint64_t off = (uint64_t)(k * 8 * 4) >> 4;
              umma_tf32(tmem_d, a_lo + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_tf32(tmem_d, a_hi + off, b_lo + off, idesc, 1u);
              umma_tf32(tmem_d, a_hi + off, b_hi + off, idesc, 1u);
            }
            umma_commit(&empty[s]);
          }
          umma_commit(&tmem_full[acc]);
        }
    }
  } else if (warp < 18) {
    // ------------------------------------------------------------------ consumers (16 warps x 8 rows)
    const int wc = warp - 2;      // consumer warp: rows 8*wc .. 8*wc+7 of the tile
    const int seg = lane & 3;     // 8-channel segment inside the 32-channel group
    const int r = 8 * wc + (lane >> 2);
    const uint32_t roff = r * 128;
    const uint32_t c0 = ((2 * seg) ^ (r & 7)) << 4, c1 = ((2 * seg + 1) ^ (r & 7)) << 4;

    // ---- issue side: a cursor over the item sequence (tile, nt, cg, batch, k), G-1 items ahead
    int i_mt = blockIdx.x, i_nt = 0, i_cg = 0, i_b = 0, i_k = 0;
    uint32_t iit = 0;
    auto issue_one = [&]() {
      if (i_mt >= num_mt) return;
      if (i_k == 0 && (num_batches > 1 || (i_nt == 0 && i_cg == 0))) {
        // global feature rows of this warp's 8 points x 8 slots -> its private part of the index table
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = lane * 2 + u, rl = e >> 3, kk = e & 7;
          const int row = i_mt * kTcBM + 8 * wc + rl;
          const int rr = row < a.rows ? row : 0;   // tail rows gather row 0 (their outputs are clipped)
          int v = 0;
          if (i_b * kG4KB + kk < a.K)
            v = (rr / a.n_per_cloud) * a.n_per_cloud + __ldg(a.nbr + (long long)rr * a.K + i_b * kG4KB + kk);
          sidx[(8 * wc + rl) * kG4KB + kk] = v;
        }
        __syncwarp();
      }
      const int g = iit % G;
      const uint32_t ph = (iit / G) & 1;
      mbar_wait(&gempty[g], ph ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&gfull[g], 2 * 512);
      __syncwarp();
      if (lane < 2) {
        const int* si = sidx + (8 * wc + 4 * lane) * kG4KB + i_k;
        tma_gather4(gbase + g * kG4StageBytes + (8 * wc + 4 * lane) * 128, &tmF, i_cg * kTcBK, si[0],
                    si[kG4KB], si[2 * kG4KB], si[3 * kG4KB], &gfull[g]);
      }
      ++iit;
      if (++i_k == min(kG4KB, a.K - i_b * kG4KB)) {
        i_k = 0;
        if (++i_b == num_batches) {
          i_b = 0;
          if (++i_cg == num_cg) {
            i_cg = 0;
            if (++i_nt == num_nt) { i_nt = 0; i_mt += gridDim.x; }
          }
        }
      }
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
            const int k0 = b * kG4KB;
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
                sdelta[r * kG4KB + k] = d;
              }
              __syncwarp();
            }
            const int kcnt = min(kG4KB, a.K - k0);
            for (int k = 0; k < kcnt; ++k, ++git) {
              const int g = git % G;
              const uint32_t ph = (git / G) & 1;
              mbar_wait(&gfull[g], ph);
              const uint8_t* gs = gbase + g * kG4StageBytes + roff;
              const float4 f0 = *reinterpret_cast<const float4*>(gs + c0);
              const float4 f1 = *reinterpret_cast<const float4*>(gs + c1);
              const float4 d = sdelta[r * kG4KB + k];
              const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                m[0][c] += fv[c];
                m[1][c] = fmaf(d.x, fv[c], m[1][c]);
                m[2][c] = fmaf(d.y, fv[c], m[2][c]);
                m[3][c] = fmaf(d.z, fv[c], m[3][c]);
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&gempty[g]);
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
  } else {
    // ------------------------------------------------------------------ epilogue (warps 18..21)
    const int q = warp & 3;
    const int et = threadIdx.x - 576;  // 0..127
    uint8_t* my_stage = out_stage + (warp - 18) * 4096;
    uint32_t tile = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x)
      for (int nt = 0; nt < num_nt; ++nt, ++tile) {
        const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
        float* prm = params + acc * 2 * BN;
        for (int c = et; c < BN; c += 128) {
          const int gc = nt * BN + c;
          const bool in = gc < a.Dout;
          prm[c] = (in && a.scale) ? __ldg(a.scale + gc) : 1.f;
          prm[BN + c] = (in && a.shift) ? __ldg(a.shift + gc) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tmem_full[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0e = 0; c0e < BN; c0e += 32) {
          uint32_t rg[32];
          const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0e;
          DH3D_TMEM_LD_32X32(rg, taddr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0e + 32 >= BN) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (nt * BN + c0e < a.Dout) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(rg[j]), prm[c0e + j], prm[BN + c0e + j]);
            tc_act32(v, a.act);
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmY, my_stage, nt * BN + c0e, mt * kTcBM + q * 32);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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

// 2-D fp32 tensor [rows, cols]; box = [1 row x 32 cols] for tile::gather4 (4 rows per instruction), 128B swizzle
static int make_gather_map(CUtensorMap* m, const float* base, long long rows, long long cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DH3D_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DH3D_OK : DH3D_ERR_UNSUPPORTED;
}

template <int BN>
static int launch_g4(const float* feat, const G4Args& a, const float* thi, const float* tlo, float* out,
                     cudaStream_t st) {
  CUtensorMap mf, mh, ml, my;
  int rc;
  const int Kd = 4 * a.Din;
  if ((rc = make_gather_map(&mf, feat, a.rows, a.Din)) != DH3D_OK) return rc;
  if ((rc = make_map(&mh, thi, a.Dout, Kd, Kd, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&ml, tlo, a.Dout, Kd, Kd, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&my, out, a.rows, a.Dout, a.Dout, 32)) != DH3D_OK) return rc;
  auto kern = flexconv_g4_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)G4Cfg<BN>::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int num_mt = ceil_div(a.rows, kTcBM);
  const int grid = num_mt < num_sms() ? num_mt : num_sms();
  kern<<<grid, kG4Threads, G4Cfg<BN>::kSmemBytes, st>>>(mf, mh, ml, my, a);
  return launch_status();
}

// theta_packed = {Theta_ext^T hi [Dout, 4*Din], lo [Dout, 4*Din]} (flexconv.cu theta_ext_packed_kernel)
int flexconv_g4_launch(const float* feat, const float* xyz, const int32_t* nbr, const void* theta_packed,
                       const float* scale, const float* shift, int act, float* out, int rows, int n_per_cloud,
                       int K, int Din, int Dout, cudaStream_t st) {
  if (Din % kTcBK != 0 || Dout % 4 != 0 || K < 1) return DH3D_ERR_UNSUPPORTED;
  G4Args a{xyz, nbr, scale, shift, act, rows, n_per_cloud, K, Din, Dout};
  const float* thi = reinterpret_cast<const float*>(theta_packed);
  const float* tlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(theta_packed) +
                                                    align_up((size_t)4 * Din * Dout * sizeof(float), 256));
  if (Dout <= 64) return launch_g4<64>(feat, a, thi, tlo, out, st);
  return launch_g4<128>(feat, a, thi, tlo, out, st);
}

}  // namespace dh3d
