// Minimal stand-in for the two TensorFlow headers the reference's user_ops kernels include, so
// that the UNMODIFIED reference sources under /root/reference/user_ops/kernels compile without
// TensorFlow (oracle/build_ref.py).  Written for this repo; contains no reference code.
// Provides exactly what those translation units touch: tensorflow::Tensor (dim_size, flat<T>().data(),
// tensor<T,N>() with operator()/setZero/setConstant), OpKernelContext (eigen_gpu_device().ok(),
// eigen_device<D>(), SetStatus), errors::Internal, Eigen::{GpuDevice,ThreadPoolDevice,NumTraits}.
#ifndef DH3D_REF_STUB_OP_KERNEL_H_
#define DH3D_REF_STUB_OP_KERNEL_H_

#include <cstdint>
#include <cstring>
#include <limits>
#include <string>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

namespace Eigen {
struct GpuDevice {
  bool ok() const {
#if defined(__CUDACC__)
    return cudaPeekAtLastError() == cudaSuccess;
#else
    return true;
#endif
  }
};
struct ThreadPoolDevice {};
template <typename T>
struct NumTraits {
  static T lowest() { return std::numeric_limits<T>::lowest(); }
};
}  // namespace Eigen

namespace tensorflow {

typedef long long int64;

struct Status {
  bool failed = false;
  std::string msg;
};

namespace errors {
inline Status Internal(const char* m) {
  Status s;
  s.failed = true;
  s.msg = m;
  return s;
}
}  // namespace errors

template <typename T>
struct FlatView {
  T* ptr;
  T* data() const { return ptr; }
};

template <typename T, int N>
struct TensorView {
  T* ptr;
  long long d[4];
  T* data() const { return ptr; }
  long long size() const {
    long long s = 1;
    for (int i = 0; i < N; ++i) s *= d[i];
    return s;
  }
  T& operator()(long long i) const { return ptr[i]; }
  T& operator()(long long i, long long j) const { return ptr[i * d[1] + j]; }
  T& operator()(long long i, long long j, long long k) const { return ptr[(i * d[1] + j) * d[2] + k]; }
  T& operator()(long long i, long long j, long long k, long long l) const {
    return ptr[((i * d[1] + j) * d[2] + k) * d[3] + l];
  }
  void setZero() const { std::memset((void*)ptr, 0, sizeof(T) * size()); }
  void setConstant(T v) const {
    long long s = size();
    for (long long i = 0; i < s; ++i) ptr[i] = v;
  }
};

class Tensor {
 public:
  Tensor() : buf_(nullptr), nd_(0) { dims_[0] = dims_[1] = dims_[2] = dims_[3] = 1; }
  Tensor(void* buf, int nd, const long long* dims) : buf_(buf), nd_(nd) {
    for (int i = 0; i < 4; ++i) dims_[i] = i < nd ? dims[i] : 1;
  }
  long long dim_size(int i) const { return dims_[i]; }
  int dims() const { return nd_; }
  long long NumElements() const { return dims_[0] * dims_[1] * dims_[2] * dims_[3]; }
  template <typename T>
  FlatView<T> flat() const { return FlatView<T>{reinterpret_cast<T*>(buf_)}; }
  template <typename T, int N>
  TensorView<T, N> tensor() const {
    TensorView<T, N> v;
    v.ptr = reinterpret_cast<T*>(buf_);
    for (int i = 0; i < 4; ++i) v.d[i] = dims_[i];
    return v;
  }

 private:
  void* buf_;
  int nd_;
  long long dims_[4];
};

class OpKernelContext {
 public:
  const Eigen::GpuDevice& eigen_gpu_device() const { return gpu_; }
  template <typename D>
  const D& eigen_device() const {
    static D d;
    return d;
  }
  void SetStatus(const Status& s) { status = s; }
  Status status;

 private:
  Eigen::GpuDevice gpu_;
};

}  // namespace tensorflow

#endif  // DH3D_REF_STUB_OP_KERNEL_H_
