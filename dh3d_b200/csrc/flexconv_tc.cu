// Fused FlexConv: neighbour gather -> 4*Din moments -> tensor-core contraction, one kernel.
//   (reference: user_ops/kernels/flex_conv_kernel_gpu.cu.cc:44-158; algebra in flexconv.cu)
//
//   out[n,:] = act( (A[n,:] @ Theta_ext) * scale + shift ),
//   A[n, p'*Din + c] = sum_k (1,dx,dy,dz)[p'] * f[nbr(n,k), c],   Theta_ext = [bias; theta_x; theta_y; theta_z]
//
// The moment matrix A (4x the size of the features) never reaches HBM: the gather warps build each
// 128-point x 32-channel slab of A directly in shared memory, in the 128B-swizzled K-major layout the
// UMMA descriptors read, already split into the 3xTF32 hi/lo pair.
//
// Persistent, warp-specialised, 448 threads, one CTA per SM, looping over 128-point tiles:
//   warp 0     : TMA producer for the Theta_ext^T hi/lo tiles [BN x 32] (K-major, 128B swizzle)
//   warp 1     : one thread issues 12 tcgen05.mma (kind::tf32) 128 x BN x 8 per slab into one of two
//                TMEM accumulators
//   warps 2-9  : gather: thread = (point, 8-channel segment); for each 32-channel group it reads the K
//                neighbour rows once (two 16-byte loads = one full 32-byte sector per neighbour),
//                accumulates the 4 moments in registers and writes the 4 K-slabs (p' = 1,x,y,z) into 4
//                consecutive pipeline stages
//   warps 10-13: epilogue (tcgen05.ld -> feature bias / folded BN / ReLU -> swizzled staging -> TMA store)
// Per 128-point tile the only HBM/L2 traffic is the neighbour gather (K*Din*4 B per point, L2-resident
// feature map), the index/xyz reads and the output tile.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dh3d {

constexpr int kFcThreads = 448;
constexpr int kFcGatherThreads = 256;

template <int BN>
struct FcCfg {
  static constexpr int kStages = BN <= 64 ? 4 : 3;
  static constexpr uint32_t kBBytes = BN * kTcBK * 4;
  static constexpr uint32_t kStageBytes = 2 * kTcABytes + 2 * kBBytes;
  static constexpr uint32_t kParamBytes = 2 * 2 * BN * 4;  // double-buffered scale/shift slices
  static constexpr uint32_t kSmemBytes =
      kStages * kStageBytes + kTcStageOutBytes + kParamBytes + 256 /*barriers*/ + 1024 /*align*/;
  static constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
};

struct FcArgs {
  const float* feat;     // [rows, Din]
  const float* xyz;      // [rows, 3]
  const int32_t* nbr;    // [rows, K]  (indices within the cloud)
  const float* scale;    // [Dout] or null
  const float* shift;    // [Dout] or null (feature bias already folded in)
  int act;
  int rows, n_per_cloud, K, Din, Dout;
};

template <int BN>
__global__ void __launch_bounds__(kFcThreads, 1)
flexconv_tc_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                   const __grid_constant__ CUtensorMap tmY, const FcArgs a) {
  using Cfg = FcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment computed on the shared-window address so the pointer keeps its state space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* out_stage = smem + S * Cfg::kStageBytes;
  float* params = reinterpret_cast<float*>(out_stage + kTcStageOutBytes);  // [2][2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(params) + Cfg::kParamBytes);
  uint64_t* bfull = bars;           // Theta tiles landed           (count 1 + tx)
  uint64_t* afull = bars + S;       // A hi/lo slab written         (count 256)
  uint64_t* empty = bars + 2 * S;   // MMAs reading the stage done  (count 1, tcgen05.commit)
  uint64_t* tmem_full = bars + 3 * S;
  uint64_t* tmem_empty = bars + 3 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_mt = (a.rows + kTcBM - 1) / kTcBM;
  const int num_nt = (a.Dout + BN - 1) / BN;
  const int num_cg = a.Din / kTcBK;     // 32-channel groups; 4 K-slabs each
  const int num_kb = 4 * num_cg;

  auto stage_a = [&](int s) { return smem + s * Cfg::kStageBytes; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::kStageBytes + kTcABytes; };
  auto stage_bhi = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes; };
  auto stage_blo = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes + Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&afull[s], kFcGatherThreads);
      mbar_init(&empty[s], 1);
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
  } else if (warp < 10) {
    // ------------------------------------------------------------------ gather + moments (256 threads)
    const int t = threadIdx.x - 64;
    const int seg = t & 3;        // 8-channel segment inside the 32-channel group
    const int r0 = t >> 2;        // rows r0 and r0 + 64 of the tile
    uint32_t it = 0;
    for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
      const int m0 = mt * kTcBM;
      for (int nt = 0; nt < num_nt; ++nt)
        for (int cg = 0; cg < num_cg; ++cg) {
          float m[2][4][8];  // [row][moment p'][channel]
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
              for (int c = 0; c < 8; ++c) m[h][p][c] = 0.f;
          // The gather is latency-bound (index -> neighbour row are dependent L2 round trips), so the
          // indices of 8 neighbours x 2 rows are fetched first, then neighbour rows are loaded two
          // neighbours x two rows at a time (20 independent loads in flight per thread).
          const int coff = cg * kTcBK + seg * 8;
          long long cloud0[2];
          float pxyz[2][3];
          bool rv[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = m0 + r0 + h * 64;
            rv[h] = row < a.rows;
            const int rr = rv[h] ? row : 0;
            cloud0[h] = (long long)(rr / a.n_per_cloud) * a.n_per_cloud;
            pxyz[h][0] = __ldg(a.xyz + (long long)rr * 3);
            pxyz[h][1] = __ldg(a.xyz + (long long)rr * 3 + 1);
            pxyz[h][2] = __ldg(a.xyz + (long long)rr * 3 + 2);
          }
          for (int k8 = 0; k8 < a.K; k8 += 8) {
            int g[2][8];  // global row of the neighbour (rows < 2^31), -1 = none
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int rr = rv[h] ? (m0 + r0 + h * 64) : 0;
              const int32_t* nb = a.nbr + (long long)rr * a.K + k8;
#pragma unroll
              for (int k = 0; k < 8; ++k) g[h][k] = (k8 + k < a.K) ? (int)cloud0[h] + __ldg(nb + k) : -1;
            }
#pragma unroll
            for (int kk = 0; kk < 8; kk += 2) {
              float4 f0[2][2], f1[2][2];
              float d[2][2][3];
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int gg = g[h][kk + u];
                  const bool ok = rv[h] && gg >= 0;
                  const long long gs = ok ? gg : 0;
                  const float* f = a.feat + gs * a.Din + coff;
                  f0[h][u] = ldg4(f);
                  f1[h][u] = ldg4(f + 4);
                  d[h][u][0] = __ldg(a.xyz + gs * 3) - pxyz[h][0];
                  d[h][u][1] = __ldg(a.xyz + gs * 3 + 1) - pxyz[h][1];
                  d[h][u][2] = __ldg(a.xyz + gs * 3 + 2) - pxyz[h][2];
                  if (!ok) { f0[h][u] = make_float4(0.f, 0.f, 0.f, 0.f); f1[h][u] = f0[h][u]; }
                }
#pragma unroll
              for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const float fv[8] = {f0[h][u].x, f0[h][u].y, f0[h][u].z, f0[h][u].w,
                                       f1[h][u].x, f1[h][u].y, f1[h][u].z, f1[h][u].w};
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    m[h][0][c] += fv[c];
                    m[h][1][c] = fmaf(d[h][u][0], fv[c], m[h][1][c]);
                    m[h][2][c] = fmaf(d[h][u][1], fv[c], m[h][2][c]);
                    m[h][3][c] = fmaf(d[h][u][2], fv[c], m[h][3][c]);
                  }
                }
            }
          }
          // 4 K-slabs (p' = 1, x, y, z) -> 4 consecutive stages, swizzled K-major, hi (raw) + lo
#pragma unroll
          for (int p = 0; p < 4; ++p, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* ah = stage_a(s);
            uint8_t* al = stage_alo(s);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = r0 + h * 64;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int chunk = seg * 2 + j;
                const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
                const float4 v = make_float4(m[h][p][4 * j], m[h][p][4 * j + 1], m[h][p][4 * j + 2],
                                             m[h][p][4 * j + 3]);
                float4 l;
                l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                *reinterpret_cast<float4*>(ah + off) = v;
                *reinterpret_cast<float4*>(al + off) = l;
              }
            }
            fence_proxy_async();
            mbar_arrive(&afull[s]);
          }
        }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 10..13)
    const int q = warp & 3;
    const int et = threadIdx.x - 320;  // 0..127
    uint8_t* my_stage = out_stage + (warp - 10) * 4096;
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
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          DH3D_TMEM_LD_32X32(r, taddr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 32 >= BN) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (nt * BN + c0 < a.Dout) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), prm[c0 + j], prm[BN + c0 + j]);
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
              tma_store_2d(&tmY, my_stage, nt * BN + c0, mt * kTcBM + q * 32);
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

template <int BN>
static int launch_fc(const FcArgs& a, const float* thi, const float* tlo, float* out, cudaStream_t st) {
  CUtensorMap mh, ml, my;
  int rc;
  const int Kd = 4 * a.Din;
  if ((rc = make_map(&mh, thi, a.Dout, Kd, Kd, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&ml, tlo, a.Dout, Kd, Kd, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&my, out, a.rows, a.Dout, a.Dout, 32)) != DH3D_OK) return rc;
  auto kern = flexconv_tc_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)FcCfg<BN>::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int num_mt = ceil_div(a.rows, kTcBM);
  const int grid = num_mt < num_sms() ? num_mt : num_sms();
  kern<<<grid, kFcThreads, FcCfg<BN>::kSmemBytes, st>>>(mh, ml, my, a);
  return launch_status();
}

// theta_packed = {Theta_ext^T hi [Dout, 4*Din], lo [Dout, 4*Din]} (flexconv.cu theta_ext_packed_kernel)
int flexconv_fused_launch(const float* feat, const float* xyz, const int32_t* nbr, const void* theta_packed,
                          const float* scale, const float* shift, int act, float* out, int rows,
                          int n_per_cloud, int K, int Din, int Dout, cudaStream_t st) {
  if (Din % kTcBK != 0 || Dout % 4 != 0 || K < 1) return DH3D_ERR_UNSUPPORTED;
  FcArgs a{feat, xyz, nbr, scale, shift, act, rows, n_per_cloud, K, Din, Dout};
  const float* thi = reinterpret_cast<const float*>(theta_packed);
  const float* tlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(theta_packed) +
                                                    align_up((size_t)4 * Din * Dout * sizeof(float), 256));
  if (Dout <= 64) return launch_fc<64>(a, thi, tlo, out, st);
  return launch_fc<128>(a, thi, tlo, out, st);
}

bool flexconv_fused_supported(int Din, int Dout) { return Din % kTcBK == 0 && Dout % 4 == 0; }

}  // namespace dh3d
