// extern "C" entry points into the REFERENCE's own CUDA kernels (compiled unmodified from
// /root/reference by oracle/build_ref.py into oracle/_ref/libdh3d_ref_cuda.so).  Test/bench
// infrastructure: pins the oracle against the real reference on the GPU box and serves as the
// "reference CUDA build on the same box" comparator.  All pointers are device pointers; kernels
// run on the legacy default stream exactly as the reference launches them.
#include <cuda_runtime.h>

#include "flex_conv_op.h"
#include "flex_pool_op.h"
#include "conv_pointset_op.h"
#include "knn_bruteforce_op.h"
#include "flex_deconv_op.h"

// tf_ops launchers (tf_ops/sampling/tf_sampling_g.cu:194-211, tf_ops/grouping/tf_grouping_g.cu:179-199)
void farthestpointsamplingLauncher(int b, int n, int m, const float* inp, float* temp, int* out);
void gatherpointLauncher(int b, int n, int m, const float* inp, const int* idx, float* out);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float* xyz1,
                            const float* xyz2, int* idx, int* pts_cnt);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float* points, const int* idx,
                        float* out);

void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                            float* grad_points);
void scatteraddpointLauncher(int b, int n, int m, const float* out_g, const int* idx, float* inp_g);

using tensorflow::Tensor;
typedef Eigen::GpuDevice GPU;

static Tensor T(const void* p, long long a, long long b = 1, long long c = 1, int nd = 3) {
  long long d[4] = {a, b, c, 1};
  return Tensor(const_cast<void*>(p), nd, d);
}
static int finish(tensorflow::OpKernelContext& ctx) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  return ctx.status.failed ? -1 : 0;
}

extern "C" {
#define REF_API __attribute__((visibility("default")))

REF_API int ref_fps(int b, int n, int m, const float* inp, float* temp /*[32,n]*/, int* out) {
  farthestpointsamplingLauncher(b, n, m, inp, temp, out);
  return (int)cudaDeviceSynchronize();
}
REF_API int ref_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out) {
  gatherpointLauncher(b, n, m, inp, idx, out);
  return (int)cudaDeviceSynchronize();
}
REF_API int ref_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1,
                                 const float* xyz2, int* idx, int* cnt) {
  queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, cnt);
  return (int)cudaDeviceSynchronize();
}
REF_API int ref_group_point(int b, int n, int c, int m, int nsample, const float* points,
                            const int* idx, float* out) {
  groupPointLauncher(b, n, c, m, nsample, points, idx, out);
  return (int)cudaDeviceSynchronize();
}
REF_API int ref_knn(int B, int Dp, int N, int K, const float* pos, int* ids, float* dists) {
  tensorflow::OpKernelContext ctx;
  Tensor p = T(pos, B, Dp, N), i = T(ids, B, N, K), d = T(dists, B, N, K);
  tensorflow::functor::KnnBruteforceFunctor<GPU, float, int>()(&ctx, p, &i, &d);
  return finish(ctx);
}
REF_API int ref_flex_conv(int B, int N, int K, int Din, int Dout, const float* feat,
                          const float* theta, const float* bias, const int* nbr, const float* pos,
                          float* out) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, 3, Din, Dout), bi = T(bias, Din, Dout, 1, 2),
         nb = T(nbr, B, K, N), p = T(pos, B, 3, N), o = T(out, B, Dout, N);
  tensorflow::functor::FlexConvFunctor<GPU, float>()(&ctx, f, th, bi, nb, p, &o);
  return finish(ctx);
}
REF_API int ref_flex_pool(int B, int N, int K, int D, const float* feat, const int* nbr, float* out,
                          int* argmax) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, D, N), nb = T(nbr, B, K, N), o = T(out, B, D, N), a = T(argmax, B, D, N);
  tensorflow::functor::FlexPoolFunctor<GPU, float>()(&ctx, f, nb, &o, &a);
  return finish(ctx);
}
REF_API int ref_conv_pointset(int B, int N, int K, int Din, int Dout, const float* feat,
                              const float* theta, const float* bias, const int* nbr, float* out) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, Din, Dout, 1, 2), bi = T(bias, Dout, 1, 1, 1),
         nb = T(nbr, B, K, N), o = T(out, B, Dout, N);
  tensorflow::functor::ConvPointsetFunctor<GPU, float>()(&ctx, f, th, bi, nb, &o);
  return finish(ctx);
}
// ---- backward passes / FlexDeconv (outputs that the reference OpKernels memset are zeroed here) ----
REF_API int ref_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                                 float* grad_points) {
  cudaMemset(grad_points, 0, sizeof(float) * (size_t)b * n * c);  // tf_grouping.cpp:270
  groupPointGradLauncher(b, n, c, m, nsample, grad_out, idx, grad_points);
  return (int)cudaDeviceSynchronize();
}
REF_API int ref_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g) {
  cudaMemset(inp_g, 0, (size_t)b * n * 3 * 4);  // tf_sampling.cpp:174
  scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
  return (int)cudaDeviceSynchronize();
}
REF_API int ref_flex_conv_grad(int B, int N, int K, int Din, int Dout, const float* feat, const float* theta,
                               const float* bias, const int* nbr, const float* pos, const float* top, float* gf,
                               float* gtheta, float* gbias) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, 3, Din, Dout), bi = T(bias, Din, Dout, 1, 2), nb = T(nbr, B, K, N),
         p = T(pos, B, 3, N), t = T(top, B, Dout, N), o1 = T(gf, B, Din, N), o2 = T(gtheta, 3, Din, Dout),
         o3 = T(gbias, Din, Dout, 1, 2);
  tensorflow::functor::FlexConvGrad<GPU, float>()(&ctx, f, th, bi, nb, p, t, &o1, &o2, &o3);
  return finish(ctx);
}
REF_API int ref_flex_pool_grad(int B, int N, int K, int D, const float* feat, const int* nbr, const float* top,
                               const int* argmax, float* gf) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, D, N), nb = T(nbr, B, K, N), t = T(top, B, D, N), a = T(argmax, B, D, N), o = T(gf, B, D, N);
  tensorflow::functor::FlexPoolGrad<GPU, float>()(&ctx, f, nb, t, a, &o);
  return finish(ctx);
}
REF_API int ref_conv_pointset_grad(int B, int N, int K, int Din, int Dout, const float* feat, const float* theta,
                                   const float* bias, const int* nbr, const float* top, float* gf, float* gtheta,
                                   float* gbias) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, Din, Dout, 1, 2), bi = T(bias, Dout, 1, 1, 1), nb = T(nbr, B, K, N),
         t = T(top, B, Dout, N), o1 = T(gf, B, Din, N), o2 = T(gtheta, Din, Dout, 1, 2), o3 = T(gbias, Dout, 1, 1, 1);
  tensorflow::functor::ConvPointsetGrad<GPU, float>()(&ctx, f, th, bi, nb, t, &o1, &o2, &o3);
  return finish(ctx);
}
REF_API int ref_flex_deconv(int B, int N, int K, int Din, int Dout, const float* feat, const float* theta,
                            const float* bias, const int* nbr, const float* pos, float* out) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, 3, Din, Dout), bi = T(bias, Din, Dout, 1, 2), nb = T(nbr, B, K, N),
         p = T(pos, B, 3, N), o = T(out, B, Dout, N);
  tensorflow::functor::FlexDeconvFunctor<GPU, float>()(&ctx, f, th, bi, nb, p, &o);
  return finish(ctx);
}
}  // extern "C"
