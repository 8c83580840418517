// Stand-in for tensorflow/core/util/cuda_kernel_helper.h (see op_kernel.h stub): launch-config
// helper and atomic add used by the reference's backward kernels.  No reference code.
#ifndef DH3D_REF_STUB_CUDA_KERNEL_HELPER_H_
#define DH3D_REF_STUB_CUDA_KERNEL_HELPER_H_
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
struct CudaLaunchConfig {
  int virtual_thread_count = 0;
  int thread_per_block = 0;
  int block_count = 0;
};
inline CudaLaunchConfig GetCudaLaunchConfig(int work, const Eigen::GpuDevice&) {
  CudaLaunchConfig c;
  c.virtual_thread_count = work;
  c.thread_per_block = 1024;
  int blocks = (work + 1023) / 1024;
  c.block_count = blocks < 1 ? 1 : (blocks > 148 * 2 ? 148 * 2 : blocks);
  return c;
}
#if defined(__CUDACC__)
template <typename T>
__device__ inline T CudaAtomicAdd(T* p, T v) { return atomicAdd(p, v); }
#endif
}  // namespace tensorflow
#endif
