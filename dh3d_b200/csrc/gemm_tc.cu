// Tensor-core GEMM with fp32-grade accuracy:  Y = act((X @ W) * scale + shift)
//   optionally fused with a following 1-column layer:  y[m] = act2( sum_n Y[m,n] * w2[n] + b2 )
//
// 5th-gen tensor cores (tcgen05.mma, kind::tf32, accumulators in TMEM) with the 3xTF32 split
//     x = x_hi + x_lo,  w = w_hi + w_lo   (hi = tf32 part, lo = exact fp32 remainder)
//     x*w ~= x_lo*w_hi + x_hi*w_lo + x_hi*w_hi        (dropped x_lo*w_lo term ~ 2^-22 relative)
// Measured error vs fp64: ~3e-5 of the output rms at worst (the tensor core's truncating fp32
// accumulation dominates, not the split) -- inside north_star's 1e-4, where single-pass TF32
// (~1e-3) is not.  Used for the FlexConv contraction A[n,4Din] @ Theta_ext and the dense 1x1 stacks.
//
// Persistent, warp-specialised kernel: one CTA per SM loops over 128-row M tiles; for each M tile
// over the N tiles; for each over 32-wide K slabs (320 threads):
//   warp 0    : TMA producer: X tile [128 x 32] fp32, W_hi^T / W_lo^T tiles [BN x 32] (K-major,
//               pre-split once per weight by linear_prepack) -> smem ring, 128B swizzle.
//   warps 2-5 : x_lo = x - trunc_tf32(x) into a second buffer (element-wise on the swizzled tile;
//               the raw tile serves as x_hi because the tensor core ignores the low 13 bits),
//               fence to the async proxy, arrive on conv[stage].
//   warp 1    : one thread issues 4 (k) x 3 (split terms) tcgen05.mma 128 x BN x 8 per slab into one
//               of TWO TMEM accumulators; tcgen05.commit frees the smem stage / publishes the tile.
//   warps 6-9 : epilogue of tile i overlaps the MMAs of tile i+1: tcgen05.ld (32 lanes x 32 columns
//               per warp) -> scale/shift/act -> swizzled smem staging -> TMA store (coalesced, clips
//               the M/N tails), or the fused row-dot kept in a register per row across the N tiles.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dh3d {

constexpr int kTcThreads = 320;

template <int BN>
struct TcCfg {
  static constexpr int kStages = BN <= 64 ? 4 : 3;
  static constexpr uint32_t kBBytes = BN * kTcBK * 4;
  static constexpr uint32_t kStageBytes = 2 * kTcABytes + 2 * kBBytes;
  static constexpr uint32_t kParamBytes = 2 * 3 * BN * 4;  // double-buffered scale/shift/w2 slices
  static constexpr uint32_t kSmemBytes =
      kStages * kStageBytes + kTcStageOutBytes + kParamBytes + 256 /*barriers*/ + 1024 /*align*/;
  static constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
};

struct TcEpilogue {
  const float* scale;   // [N] or null
  const float* shift;   // [N] or null
  int act;
  const float* w2;      // [N]: fused row-dot (ROWDOT mode)
  float b2;
  int act2;
  float* y2;            // [M]   (ROWDOT mode)
};

// MC = CTAs per cluster sharing every W tile through TMA multicast (1 or 2): each CTA loads half of
// the tile and multicasts it to both, halving the L2->SM weight traffic that bounds this kernel.
template <int BN, bool ROWDOT, int MC>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmY,
               const TcEpilogue ep, int M, int K, int N) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment computed on the shared-window address so the pointer keeps its state space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* out_stage = smem + S * Cfg::kStageBytes;                       // 4 x 4 KB, 1024-aligned
  float* params = reinterpret_cast<float*>(out_stage + kTcStageOutBytes);  // [2][3][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(params) + Cfg::kParamBytes);
  uint64_t* full = bars;            // TMA bytes landed                 (count 1 + tx)
  uint64_t* conv = bars + S;        // x_lo written                     (count 128)
  uint64_t* empty = bars + 2 * S;   // MMAs reading the stage finished  (count 1, tcgen05.commit)
  uint64_t* tmem_full = bars + 3 * S;       // [2] accumulator ready    (count 1, tcgen05.commit)
  uint64_t* tmem_empty = bars + 3 * S + 2;  // [2] accumulator drained  (count 4, one per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + kTcBK - 1) / kTcBK;
  const int num_mt = (M + kTcBM - 1) / kTcBM;
  const int num_nt = (N + BN - 1) / BN;
  // M tiles are dealt out per cluster: cluster c takes tiles (c*MC + rank), stride gridDim.x; every CTA
  // of a cluster runs the same number of iterations (a tile index >= num_mt is all out-of-bounds:
  // TMA zero-fills the loads and clips the stores)
  const uint32_t crank = MC > 1 ? cluster_ctarank() : 0;
  const int mt_begin = (int)(blockIdx.x / MC) * MC;  // first tile of this cluster
  const int mt_stride = (int)gridDim.x;

  auto stage_a = [&](int s) { return smem + s * Cfg::kStageBytes; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::kStageBytes + kTcABytes; };
  auto stage_bhi = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes; };
  auto stage_blo = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes + Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 128);
      mbar_init(&empty[s], MC);  // one tcgen05.commit per CTA of the cluster (peers write this stage too)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
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
  if constexpr (MC > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride) {
        const int mt = mtb + (int)crank;
        for (int nt = 0; nt < num_nt; ++nt)
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], kTcABytes + 2 * Cfg::kBBytes);
            tma_load_2d(stage_a(s), &tmA, kb * kTcBK, mt * kTcBM, &full[s]);
            if constexpr (MC == 1) {
              tma_load_2d(stage_bhi(s), &tmBhi, kb * kTcBK, nt * BN, &full[s]);
              tma_load_2d(stage_blo(s), &tmBlo, kb * kTcBK, nt * BN, &full[s]);
            } else {
              // this CTA's half of the W tile (rows crank*BN/2 ..) lands in BOTH CTAs' stage s
              constexpr int HB = BN / MC;
              const uint32_t off = crank * HB * (kTcBK * 4);
              tma_load_2d_mc(stage_bhi(s) + off, &tmBhi, kb * kTcBK, nt * BN + (int)crank * HB, &full[s],
                             (uint16_t)((1u << MC) - 1));
              tma_load_2d_mc(stage_blo(s) + off, &tmBlo, kb * kTcBK, nt * BN + (int)crank * HB, &full[s],
                             (uint16_t)((1u << MC) - 1));
            }
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(kTcBM >> 4) << 24);
      uint32_t it = 0, tile = 0;
      for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride)
        for (int nt = 0; nt < num_nt; ++nt, ++tile) {
          const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
          mbar_wait(&tmem_empty[acc], aph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&conv[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t a_hi = umma_desc_sw128(smem_u32(stage_a(s)));
            const uint64_t a_lo = umma_desc_sw128(smem_u32(stage_alo(s)));
            const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_bhi(s)));
            const uint64_t b_lo = umma_desc_sw128(smem_u32(stage_blo(s)));
#pragma unroll
            for (int k = 0; k < kTcBK / 8; ++k) {
              const uint64_t off = (uint64_t)(k * 8 * 4) >> 4;  // 32 bytes per K=8 step, 16-byte units
              umma_tf32(tmem_d, a_lo + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_tf32(tmem_d, a_hi + off, b_lo + off, idesc, 1u);
              umma_tf32(tmem_d, a_hi + off, b_hi + off, idesc, 1u);
            }
            if constexpr (MC == 1) umma_commit(&empty[s]);
            else umma_commit_mc(&empty[s], (uint16_t)((1u << MC) - 1));  // frees the stage in every CTA
          }
          umma_commit(&tmem_full[acc]);
        }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ x_lo producers
    const int t = threadIdx.x - 64;  // 0..127
    uint32_t it = 0;
    for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride)
      for (int nt = 0; nt < num_nt; ++nt)
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full[s], ph);
          const float4* a = reinterpret_cast<const float4*>(stage_a(s));
          float4* lo = reinterpret_cast<float4*>(stage_alo(s));
#pragma unroll
          for (int j = 0; j < (int)(kTcABytes / 16 / 128); ++j) {
            const int i = t + j * 128;
            const float4 v = a[i];
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            lo[i] = l;
          }
          fence_proxy_async();
          mbar_arrive(&conv[s]);
        }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 6..9)
    const int q = warp & 3;              // TMEM lane quadrant of this warp
    const int et = threadIdx.x - 192;    // 0..127
    uint8_t* my_stage = out_stage + (warp - 6) * 4096;
    uint32_t tile = 0;
    for (int mtb = mt_begin; mtb < num_mt; mtb += mt_stride) {
      const int mt = mtb + (int)crank;
      float dot = 0.f;
      for (int nt = 0; nt < num_nt; ++nt, ++tile) {
        const uint32_t acc = tile & 1, aph = (tile >> 1) & 1;
        float* prm = params + acc * 3 * BN;
        // per-tile epilogue constants -> smem (double-buffered with the accumulator index)
        for (int c = et; c < BN; c += 128) {
          const int gc = nt * BN + c;
          const bool in = gc < N;
          prm[c] = (in && ep.scale) ? __ldg(ep.scale + gc) : 1.f;
          prm[BN + c] = (in && ep.shift) ? __ldg(ep.shift + gc) : 0.f;
          prm[2 * BN + c] = (ROWDOT && in) ? __ldg(ep.w2 + gc) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tmem_full[acc], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
              "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 32 >= BN) {  // accumulator fully read: hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), prm[c0 + j], prm[BN + c0 + j]);
          tc_act32(v, ep.act);
          if constexpr (ROWDOT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dot = fmaf(v[j], prm[2 * BN + c0 + j], dot);
          } else {
            if (nt * BN + c0 < N) {
              // staging tile [32 rows x 128 B], 128B-swizzled like the tensor map expects
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmY, my_stage, nt * BN + c0, mt * kTcBM + q * 32);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // params[acc] may be rewritten two tiles later
      }
      if constexpr (ROWDOT) {
        const int row = mt * kTcBM + q * 32 + lane;
        if (row < M) ep.y2[row] = tc_act(dot + ep.b2, ep.act2);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();  // no CTA leaves while a peer can still signal its barriers
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

// ---- weight pre-split: W [K,N] row-major -> packed = { W_hi^T [N,K] , W_lo^T [N,K] } ------------
__global__ void linear_prepack_kernel(const float* __restrict__ w, int K, int N,
                                      float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = (long long)K * N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e / K), k = (int)(e - (long long)n * K);
    const float v = w[(long long)k * N + n];
    const float h = tf32_rn(v);
    hi[e] = h;
    lo[e] = tf32_rn(v - h);
  }
}

size_t linear_prepack_bytes(int K, int N) {
  if (K <= 0 || N <= 0) return 0;
  return 2 * align_up((size_t)K * N * sizeof(float), 256);
}

int linear_prepack_launch(const float* w, int K, int N, void* packed, cudaStream_t st) {
  if (!w || !packed) return DH3D_ERR_NULL;
  if (K <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (K % 4 || N % 4) return DH3D_ERR_DIM;
  if (((uintptr_t)packed & 255) != 0) return DH3D_ERR_ALIGN;
  float* hi = reinterpret_cast<float*>(packed);
  float* lo = reinterpret_cast<float*>(reinterpret_cast<char*>(packed) +
                                       align_up((size_t)K * N * sizeof(float), 256));
  long long blocks = ((long long)K * N + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  linear_prepack_kernel<<<(int)blocks, 256, 0, st>>>(w, K, N, hi, lo);
  return launch_status();
}

static bool tc_use_multicast() {
  static const bool on = [] {
    const char* e = getenv("DH3D_GEMM_MULTICAST");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int BN, bool ROWDOT, int MC>
static int launch_tc_mc(const float* x, int ldx, const float* whi, const float* wlo, const TcEpilogue& ep,
                        float* y, int ldy, int M, int K, int N, cudaStream_t st) {
  CUtensorMap ma, mh, ml, my;
  int rc;
  if ((rc = make_map(&ma, x, M, K, ldx, kTcBM)) != DH3D_OK) return rc;
  if ((rc = make_map(&mh, whi, N, K, K, BN / MC)) != DH3D_OK) return rc;
  if ((rc = make_map(&ml, wlo, N, K, K, BN / MC)) != DH3D_OK) return rc;
  if (ROWDOT) my = ma;
  else if ((rc = make_map(&my, y, M, N, ldy, 32)) != DH3D_OK) return rc;
  auto kern = gemm_tc_kernel<BN, ROWDOT, MC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TcCfg<BN>::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int num_mt = ceil_div(M, kTcBM);
  int grid = num_mt < num_sms() ? num_mt : num_sms();
  grid = ceil_div(grid, MC) * MC;
  if (grid > num_sms()) grid -= MC;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = TcCfg<BN>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, ma, mh, ml, my, ep, M, K, N);
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}

template <int BN, bool ROWDOT>
static int launch_tc(const float* x, int ldx, const float* whi, const float* wlo, const TcEpilogue& ep,
                     float* y, int ldy, int M, int K, int N, cudaStream_t st) {
  // pairs of CTAs share each W tile when there are enough M tiles to pair up and W is re-streamed
  if (BN >= 64 && tc_use_multicast() && ceil_div(M, kTcBM) >= 2 * 16)
    return launch_tc_mc<BN, ROWDOT, 2>(x, ldx, whi, wlo, ep, y, ldy, M, K, N, st);
  return launch_tc_mc<BN, ROWDOT, 1>(x, ldx, whi, wlo, ep, y, ldy, M, K, N, st);
}

static int tc_check(const float* x, int ldx, const void* packed, int M, int K, int N) {
  if (!x || !packed) return DH3D_ERR_NULL;
  if (M <= 0 || K <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (K % 4 || N % 4 || ldx % 4 || ldx < K) return DH3D_ERR_DIM;
  if ((((uintptr_t)x | (uintptr_t)packed) & 15) != 0) return DH3D_ERR_ALIGN;
  return DH3D_OK;
}

int linear_tc_launch(const float* x, int ldx, const void* packed, const float* scale, const float* shift,
                     int act, float* y, int ldy, int M, int K, int N, cudaStream_t st) {
  int rc = tc_check(x, ldx, packed, M, K, N);
  if (rc != DH3D_OK) return rc;
  if (!y) return DH3D_ERR_NULL;
  if (ldy % 4 || ldy < N) return DH3D_ERR_DIM;
  if ((((uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift) & 15) != 0) return DH3D_ERR_ALIGN;
  const float* whi = reinterpret_cast<const float*>(packed);
  const float* wlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) +
                                                    align_up((size_t)K * N * sizeof(float), 256));
  TcEpilogue ep{scale, shift, act, nullptr, 0.f, 0, nullptr};
  if (N <= 32) return launch_tc<32, false>(x, ldx, whi, wlo, ep, y, ldy, M, K, N, st);
  if (N <= 64) return launch_tc<64, false>(x, ldx, whi, wlo, ep, y, ldy, M, K, N, st);
  return launch_tc<128, false>(x, ldx, whi, wlo, ep, y, ldy, M, K, N, st);
}

// y2[m] = act2( sum_n act((x @ W)[m,n] * scale[n] + shift[n]) * w2[n] + b2 ): the [M,N] activation
// of e.g. the detector's 256 -> 1024 -> 1 head never reaches HBM.
int linear_rowdot_tc_launch(const float* x, int ldx, const void* packed, const float* scale,
                            const float* shift, int act, const float* w2, float b2, int act2, float* y2,
                            int M, int K, int N, cudaStream_t st) {
  int rc = tc_check(x, ldx, packed, M, K, N);
  if (rc != DH3D_OK) return rc;
  if (!w2 || !y2) return DH3D_ERR_NULL;
  const float* whi = reinterpret_cast<const float*>(packed);
  const float* wlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) +
                                                    align_up((size_t)K * N * sizeof(float), 256));
  TcEpilogue ep{scale, shift, act, w2, b2, act2, y2};
  if (N <= 64) return launch_tc<64, true>(x, ldx, whi, wlo, ep, nullptr, 0, M, K, N, st);
  return launch_tc<128, true>(x, ldx, whi, wlo, ep, nullptr, 0, M, K, N, st);
}

}  // namespace dh3d
