// Tensor-core GEMM with fp32-grade accuracy:  Y = act((X @ W) * scale + shift)
//
// 5th-gen tensor cores (tcgen05.mma, kind::tf32, accumulators in TMEM) with the 3xTF32 split
//     x = x_hi + x_lo,  w = w_hi + w_lo   (hi = tf32-representable part, lo = exact fp32 remainder)
//     x*w ~= x_lo*w_hi + x_hi*w_lo + x_hi*w_hi        (dropped x_lo*w_lo term ~ 2^-22 relative)
// which keeps the contraction within ~1e-6 of an fp32 FFMA result -- inside north_star's 1e-4 --
// where single-pass TF32 (~1e-3) is not.  Used for the FlexConv contraction A[n,4Din] @ Theta_ext
// and for the dense 1x1 stacks.
//
// Layout / pipeline (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   : TMA producer.  Per 32-wide K slab: X tile [128 x 32] fp32, W_hi^T and W_lo^T tiles
//              [BN x 32] (K-major, pre-split once per weight by linear_prepack) -> smem, 128B swizzle.
//   warps 2-5: split X in shared memory (hi in place, lo to a second buffer; element-wise, so the
//              swizzle pattern is irrelevant), fence to the async proxy, arrive on conv[stage].
//   warp 1   : one thread issues 4 (k) x 3 (split terms) tcgen05.mma 128 x BN x 8 per slab into a
//              TMEM accumulator, tcgen05.commit frees the stage / signals the epilogue.
//   warps 2-5: epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> scale/shift/act -> global.
#include <cuda.h>

#include "common.cuh"

namespace dh3d {

constexpr int kTcBM = 128;
constexpr int kTcBK = 32;  // fp32 elements per K slab = 128 bytes = one swizzle-128B row
constexpr int kTcThreads = 192;
constexpr uint32_t kTcABytes = kTcBM * kTcBK * 4;

template <int BN>
struct TcCfg {
  static constexpr int kStages = BN <= 64 ? 4 : (BN <= 128 ? 3 : 2);
  static constexpr uint32_t kBBytes = BN * kTcBK * 4;
  static constexpr uint32_t kStageBytes = 2 * kTcABytes + 2 * kBBytes;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// fp32 -> nearest tf32-representable fp32 (low 13 mantissa bits zero)
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v));
  return __uint_as_float(b & 0xFFFFE000u);
}

__device__ __forceinline__ float tc_act(float v, int act) {
  if (act == DH3D_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DH3D_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const float* __restrict__ scale,
               const float* __restrict__ shift, int act, float* __restrict__ Y, int ldy, int M, int K,
               int N) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;           // TMA bytes landed            (count 1 + tx)
  uint64_t* conv = bars + S;       // X split done                (count 128)
  uint64_t* empty = bars + 2 * S;  // MMAs reading the stage done (count 1, tcgen05.commit)
  uint64_t* tmem_full = bars + 3 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * kTcBM;
  const int num_kb = (K + kTcBK - 1) / kTcBK;

  auto stage_a = [&](int s) { return smem + s * Cfg::kStageBytes; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::kStageBytes + kTcABytes; };
  auto stage_bhi = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes; };
  auto stage_blo = [&](int s) { return smem + s * Cfg::kStageBytes + 2 * kTcABytes + Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], kTcABytes + 2 * Cfg::kBBytes);
        tma_load_2d(stage_a(s), &tmA, kb * kTcBK, m0, &full[s]);
        tma_load_2d(stage_bhi(s), &tmBhi, kb * kTcBK, n0, &full[s]);
        tma_load_2d(stage_blo(s), &tmBlo, kb * kTcBK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(kTcBM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t a_hi = umma_desc_sw128(smem_u32(stage_a(s)));
        const uint64_t a_lo = umma_desc_sw128(smem_u32(stage_alo(s)));
        const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_bhi(s)));
        const uint64_t b_lo = umma_desc_sw128(smem_u32(stage_blo(s)));
#pragma unroll
        for (int k = 0; k < kTcBK / 8; ++k) {
          const uint64_t off = (uint64_t)(k * 8 * 4) >> 4;  // 32 bytes per K=8 step, in 16-byte units
          umma_tf32(tmem_base, a_lo + off, b_hi + off, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_tf32(tmem_base, a_hi + off, b_lo + off, idesc, 1u);
          umma_tf32(tmem_base, a_hi + off, b_hi + off, idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int t = threadIdx.x - 64;  // 0..127
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (kb / S) & 1;
      mbar_wait(&full[s], ph);
      float4* a = reinterpret_cast<float4*>(stage_a(s));
      float4* lo = reinterpret_cast<float4*>(stage_alo(s));
#pragma unroll
      for (int j = 0; j < (int)(kTcABytes / 16 / 128); ++j) {
        const int i = t + j * 128;
        const float4 v = a[i];
        // round-to-nearest split (a truncating split biases every term the same way and the
        // error then grows like K instead of sqrt(K))
        float4 h, l;
        h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
        l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y);
        l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
        a[i] = h;
        lo[i] = l;
      }
      fence_proxy_async();
      mbar_arrive(&conv[s]);
    }
    // ---- epilogue: TMEM lanes [32*(warp%4), +32) belong to this warp ----
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
      if (row < M) {
        float* yrow = Y + (long long)row * ldy;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = n0 + c0 + j;
          if (c < N) {
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (scale) sc = ldg4(scale + c);
            if (shift) sh = ldg4(shift + c);
            float4 o;
            o.x = tc_act(fmaf(__uint_as_float(r[j + 0]), sc.x, sh.x), act);
            o.y = tc_act(fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y), act);
            o.z = tc_act(fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z), act);
            o.w = tc_act(fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w), act);
            *reinterpret_cast<float4*>(yrow + c) = o;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
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

// ---- host: tensor maps + launch -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (elements); box = [box_rows x 32 cols], 128B swizzle.
static int make_map(CUtensorMap* m, const float* base, int rows, int cols, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DH3D_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DH3D_OK : DH3D_ERR_UNSUPPORTED;
}

template <int BN>
static int launch_tc(const float* x, int ldx, const float* whi, const float* wlo, const float* scale,
                     const float* shift, int act, float* y, int ldy, int M, int K, int N,
                     cudaStream_t st) {
  CUtensorMap ma, mh, ml;
  int rc;
  if ((rc = make_map(&ma, x, M, K, ldx, kTcBM)) != DH3D_OK) return rc;
  if ((rc = make_map(&mh, whi, N, K, K, BN)) != DH3D_OK) return rc;
  if ((rc = make_map(&ml, wlo, N, K, K, BN)) != DH3D_OK) return rc;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TcCfg<BN>::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(ceil_div(N, BN), ceil_div(M, kTcBM));
  gemm_tc_kernel<BN><<<grid, kTcThreads, TcCfg<BN>::kSmemBytes, st>>>(ma, mh, ml, scale, shift, act, y, ldy,
                                                                     M, K, N);
  return launch_status();
}

int linear_tc_launch(const float* x, int ldx, const void* packed, const float* scale, const float* shift,
                     int act, float* y, int ldy, int M, int K, int N, cudaStream_t st) {
  if (!x || !packed || !y) return DH3D_ERR_NULL;
  if (M <= 0 || K <= 0 || N <= 0) return DH3D_ERR_DIM;
  if (K % 4 || N % 4 || ldx % 4 || ldy % 4 || ldx < K || ldy < N) return DH3D_ERR_DIM;
  if ((((uintptr_t)x | (uintptr_t)packed | (uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift) & 15) != 0)
    return DH3D_ERR_ALIGN;
  if (ceil_div(M, kTcBM) > 65535) return DH3D_ERR_UNSUPPORTED;
  const float* whi = reinterpret_cast<const float*>(packed);
  const float* wlo = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) +
                                                    align_up((size_t)K * N * sizeof(float), 256));
  if (N <= 32) return launch_tc<32>(x, ldx, whi, wlo, scale, shift, act, y, ldy, M, K, N, st);
  if (N <= 64) return launch_tc<64>(x, ldx, whi, wlo, scale, shift, act, y, ldy, M, K, N, st);
  return launch_tc<128>(x, ldx, whi, wlo, scale, shift, act, y, ldy, M, K, N, st);
}

}  // namespace dh3d
