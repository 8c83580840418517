// extern "C" entry points into the REFERENCE's own CPU code (compiled unmodified from
// /root/reference by oracle/build_ref.py into oracle/_ref/libdh3d_ref_cpu.so): the user_ops CPU
// functors and the plain-C loops of tf_ops/interpolation/tf_interpolate.cpp:60-127.  Host pointers.
#include "flex_conv_op.h"
#include "flex_pool_op.h"
#include "conv_pointset_op.h"
#include "flex_deconv_op.h"

void threenn_cpu(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx);
void threeinterpolate_cpu(int b, int m, int c, int n, const float* points, const int* idx,
                          const float* weight, float* out);

void threeinterpolate_grad_cpu(int b, int n, int c, int m, const float* grad_out, const int* idx,
                               const float* weight, float* grad_points);

using tensorflow::Tensor;
typedef Eigen::ThreadPoolDevice CPU;

static Tensor T(const void* p, long long a, long long b = 1, long long c = 1, int nd = 3) {
  long long d[4] = {a, b, c, 1};
  return Tensor(const_cast<void*>(p), nd, d);
}

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API void ref_cpu_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                              int* idx) {
  threenn_cpu(b, n, m, xyz1, xyz2, dist, idx);
}
REF_API void ref_cpu_three_interpolate(int b, int m, int c, int n, const float* points,
                                       const int* idx, const float* weight, float* out) {
  threeinterpolate_cpu(b, m, c, n, points, idx, weight, out);
}
REF_API void ref_cpu_flex_conv(int B, int N, int K, int Din, int Dout, const float* feat,
                               const float* theta, const float* bias, const int* nbr,
                               const float* pos, float* out) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, 3, Din, Dout), bi = T(bias, Din, Dout, 1, 2),
         nb = T(nbr, B, K, N), p = T(pos, B, 3, N), o = T(out, B, Dout, N);
  tensorflow::functor::FlexConvFunctor<CPU, float>()(&ctx, f, th, bi, nb, p, &o);
}
REF_API void ref_cpu_flex_pool(int B, int N, int K, int D, const float* feat, const int* nbr,
                               float* out, int* argmax) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, D, N), nb = T(nbr, B, K, N), o = T(out, B, D, N), a = T(argmax, B, D, N);
  tensorflow::functor::FlexPoolFunctor<CPU, float>()(&ctx, f, nb, &o, &a);
}
REF_API void ref_cpu_conv_pointset(int B, int N, int K, int Din, int Dout, const float* feat,
                                   const float* theta, const float* bias, const int* nbr,
                                   float* out) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, Din, Dout, 1, 2), bi = T(bias, Dout, 1, 1, 1),
         nb = T(nbr, B, K, N), o = T(out, B, Dout, N);
  tensorflow::functor::ConvPointsetFunctor<CPU, float>()(&ctx, f, th, bi, nb, &o);
}
// ---- backward passes / FlexDeconv (the reference's CPU functors; outputs are zeroed by the functors) ----
REF_API void ref_cpu_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx,
                                            const float* weight, float* grad_points /* pre-zeroed */) {
  threeinterpolate_grad_cpu(b, n, c, m, grad_out, idx, weight, grad_points);
}
REF_API void ref_cpu_flex_conv_grad(int B, int N, int K, int Din, int Dout, const float* feat, const float* theta,
                                    const float* bias, const int* nbr, const float* pos, const float* top,
                                    float* gf, float* gtheta, float* gbias) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, 3, Din, Dout), bi = T(bias, Din, Dout, 1, 2), nb = T(nbr, B, K, N),
         p = T(pos, B, 3, N), t = T(top, B, Dout, N), o1 = T(gf, B, Din, N), o2 = T(gtheta, 3, Din, Dout),
         o3 = T(gbias, Din, Dout, 1, 2);
  tensorflow::functor::FlexConvGrad<CPU, float>()(&ctx, f, th, bi, nb, p, t, &o1, &o2, &o3);
}
REF_API void ref_cpu_flex_pool_grad(int B, int N, int K, int D, const float* feat, const int* nbr, const float* top,
                                    const int* argmax, float* gf) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, D, N), nb = T(nbr, B, K, N), t = T(top, B, D, N), a = T(argmax, B, D, N), o = T(gf, B, D, N);
  tensorflow::functor::FlexPoolGrad<CPU, float>()(&ctx, f, nb, t, a, &o);
}
REF_API void ref_cpu_conv_pointset_grad(int B, int N, int K, int Din, int Dout, const float* feat, const float* theta,
                                        const float* bias, const int* nbr, const float* top, float* gf, float* gtheta,
                                        float* gbias) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, Din, Dout, 1, 2), bi = T(bias, Dout, 1, 1, 1), nb = T(nbr, B, K, N),
         t = T(top, B, Dout, N), o1 = T(gf, B, Din, N), o2 = T(gtheta, Din, Dout, 1, 2), o3 = T(gbias, Dout, 1, 1, 1);
  tensorflow::functor::ConvPointsetGrad<CPU, float>()(&ctx, f, th, bi, nb, t, &o1, &o2, &o3);
}
REF_API void ref_cpu_flex_deconv(int B, int N, int K, int Din, int Dout, const float* feat, const float* theta,
                                 const float* bias, const int* nbr, const float* pos, float* out) {
  tensorflow::OpKernelContext ctx;
  Tensor f = T(feat, B, Din, N), th = T(theta, 3, Din, Dout), bi = T(bias, Din, Dout, 1, 2), nb = T(nbr, B, K, N),
         p = T(pos, B, 3, N), o = T(out, B, Dout, N);
  tensorflow::functor::FlexDeconvFunctor<CPU, float>()(&ctx, f, th, bi, nb, p, &o);
}
}  // extern "C"
