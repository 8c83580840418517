// Micro-benchmark: issue rate of tcgen05.mma kind::f16 128 x N x 16 per CTA, cta_group::1 vs ::2, SW64 vs SW128 operands,
// operands in shared memory (contents irrelevant).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, int sw128) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sw128 ? 1024 : 512) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw128 ? 2 : 4) << 61;
  return d;
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(
                   smem_u32(bar)), "r"(parity) : "memory");
}

// LOADERS > 0: that many extra warps hammer shared memory with 16-byte loads + stores (the split warps' traffic)
template <int CG, int SW128, int N>
__global__ void __launch_bounds__(256, 1) k(int iters, int loaders, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  // 4 "stages": A hi/lo 8 KB each (16 KB for SW128: 128 rows x 128 B) + B hi/lo
  const uint32_t abytes = SW128 ? 16384 : 8192, bbytes = (N / CG) * (SW128 ? 128 : 64);
  const uint32_t stage = 2 * abytes + 2 * bbytes;
  const int nst = (int)((180u * 1024u) / stage) < 4 ? (int)((180u * 1024u) / stage) : 4;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0 && crank == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t s = smem_u32(smem) + (uint32_t)(it % nst) * stage;
      const uint64_t ah = desc(s, SW128), al = desc(s + abytes, SW128), bh = desc(s + 2 * abytes, SW128),
                     bl = desc(s + 2 * abytes + bbytes, SW128);
      const uint32_t d = tb + ((it >> 3) & 1) * N;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        mma<CG>(d, al + 2 * kk, bh + 2 * kk, idesc, 1u);
        mma<CG>(d, ah + 2 * kk, bl + 2 * kk, idesc, 1u);
        mma<CG>(d, ah + 2 * kk, bh + 2 * kk, idesc, 1u);
      }
    }
    commit<CG>(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (warp >= 4 && warp < 4 + loaders) {
    // traffic generator: read 16 B + write 16 B per thread per step into a scratch region after the stages
    uint8_t* scr = smem + 184 * 1024 + (warp - 4) * 4096;
    uint4 v = make_uint4(lane, 1, 2, 3);
    for (int it = 0; it < iters * 6; ++it) {
      uint4 a;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                   : "r"(smem_u32(scr + lane * 16 + ((it & 7) << 9))));
      v.x ^= a.x; v.y += a.y;
      asm volatile("st.shared.v4.u32 [%4], {%0,%1,%2,%3};" ::"r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
                   "r"(smem_u32(scr + lane * 16 + (((it + 3) & 7) << 9))) : "memory");
    }
    if (v.x == 0x12345678) out[1000 + blockIdx.x] = v.y;
  }
  if (CG == 2 && crank == 1 && warp == 0 && lane == 0) mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

template <int CG, int SW128, int N>
void run(const char* name, int grid, int loaders) {
  long long* out;
  cudaMalloc(&out, 4096 * sizeof(long long));
  cudaMemset(out, 0, 4096 * sizeof(long long));
  auto kern = k<CG, SW128, N>;
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, loaders, out);
    if (e != cudaSuccess) { printf("%s: launch %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: sync %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[4096];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; int n = 0; double sum = 0;
  for (int i = 0; i < grid; ++i) if (h[i]) { sum += h[i]; ++n; if (h[i] > mx) mx = h[i]; }
  printf("%-40s grid %3d loaders %d: %.1f cycles per MMA (128 rows/CTA x N=%d x 16) avg, %.1f max\n", name, grid, loaders,
         sum / n / (iters * 6.0), N, mx / (iters * 6.0));
  cudaFree(out);
}

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0, loaders = argc > 2 ? atoi(argv[2]) : 0, grid = argc > 3 ? atoi(argv[3]) : 148;
  switch (which) {
    case 0: run<1, 0, 256>("cta_group::1 SW64  N=256", grid, loaders); break;
    case 1: run<1, 1, 256>("cta_group::1 SW128 N=256", grid, loaders); break;
    case 2: run<2, 0, 256>("cta_group::2 SW64  N=256", grid, loaders); break;
    case 3: run<2, 1, 256>("cta_group::2 SW128 N=256", grid, loaders); break;
    case 4: run<1, 0, 128>("cta_group::1 SW64  N=128", grid, loaders); break;
    case 5: run<2, 0, 128>("cta_group::2 SW64  N=128", grid, loaders); break;
  }
  return 0;
}
