// extern "C" surface of libdh3d_b200.so -- see include/dh3d_b200.h for the contract and the
// reference interface (file:line) each entry point replaces.
#include <stdlib.h>

#include "common.cuh"

namespace dh3d {
// knn.cu
size_t knn_workspace_bytes(int B, int N);
int knn_sort_launch(const float* pos, int B, int N, long long sb, int sp, int sd, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);
int knn_query_sorted_launch(const void* workspace, int B, int N, int K, int32_t* ids, float* dists, cudaStream_t st);
int fps_presorted_launch(int b, int n, int m, const void* knn_workspace, int32_t* out, cudaStream_t st);
int knn_launch(const float* pos, int B, int N, int K, long long sb, int sp, int sd, int32_t* ids,
               float* dists, void* workspace, size_t workspace_bytes, cudaStream_t st);
// fps.cu
int fps_launch(int b, int n, int m, const float* inp, int32_t* out, cudaStream_t st);
// gather.cu
int group_point_launch(int b, int n, int c, int m, int s, const float* points, const int32_t* idx,
                       float* out, cudaStream_t st);
int group_point_ld_launch(int b, int n, int c, int m, int s, const float* points, int ldp, const int32_t* idx,
                          float* out, cudaStream_t st);
int three_interpolate_ld_launch(int b, int m, int c, int n, const float* points, const int32_t* idx,
                                const float* wsrc, float* out, int ldo, bool from_dist, cudaStream_t st);
int add_l2norm_rows_launch(const float* a, const float* b, float* sum, float* y, int M, int C, float eps,
                           cudaStream_t st);
int flex_pool_pm_launch(const float* feat, const int32_t* nbr, float* out, int32_t* argmax, int B,
                        int N, int K, int D, cudaStream_t st);
int flex_pool_cm_launch(const float* feat, const int32_t* nbr, float* out, int32_t* argmax, int B,
                        int N, int K, int D, cudaStream_t st);
int conv_pointset_pm_launch(const float* feat, const float* theta, const float* bias,
                            const int32_t* nbr, float* out, int B, int N, int K, int Din, int Dout,
                            const float* scale, const float* shift, int act, cudaStream_t st);
int conv_pointset_cm_launch(const float* feat, const float* theta, const float* bias,
                            const int32_t* nbr, float* out, int B, int N, int K, int Din, int Dout,
                            cudaStream_t st);
int three_nn_launch(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                    int32_t* idx, cudaStream_t st);
size_t three_nn_workspace_bytes(int b, int n, int m);
int three_nn_presorted_launch(int b, int n, int m, const void* knn_workspace_of_xyz1, const float* xyz2, float* dist,
                              int32_t* idx, void* workspace, size_t workspace_bytes, cudaStream_t st);
int three_nn_presorted2_launch(int b, int n, int m, const void* knn_workspace_of_xyz1, const void* knn_workspace_of_xyz2,
                               float* dist, int32_t* idx, cudaStream_t st);
int three_nn_pruned_launch(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                           int32_t* idx, void* workspace, size_t workspace_bytes, cudaStream_t st);
int three_interpolate_launch(int b, int m, int c, int n, const float* points, const int32_t* idx,
                             const float* wsrc, float* out, bool from_dist, cudaStream_t st);
size_t query_ball_workspace_bytes(int b, int m);
int query_ball_launch(int b, int n, int m, float radius, int nsample, const float* xyz1,
                      const float* xyz2, int32_t* idx, int32_t* pts_cnt, void* ws, size_t ws_bytes,
                      cudaStream_t st);
int transpose_launch(const void* src, void* dst, int B, int R, int C, cudaStream_t st);
int se_excite_launch(const float* x, const float* g, float* y, size_t count, cudaStream_t st);
// se_fused.cu
int se_pool_excite_launch(const float* x, const int32_t* nbr, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* out, int B, int N, int K, int C, int H, cudaStream_t st);
int add_launch(const float* a, const float* b, float* y, size_t count, cudaStream_t st);
int copy_cols_launch(const float* src, int lds, float* dst, int ldd, int M, int C, cudaStream_t st);
int l2norm_rows_launch(const float* x, int ldx, float* y, int ldy, int M, int C, float eps,
                       cudaStream_t st);
int rowdot_launch(const float* x, int ldx, const float* w, float bias, int act, float* y, int M,
                  int K, cudaStream_t st);
// gemm_simt.cu
int linear_simt_launch(const float* x, int ldx, const float* w, const float* scale,
                       const float* shift, int act, float* y, int ldy, int M, int K, int N,
                       cudaStream_t st);
size_t linear_prepack16_bytes(int K, int N);
int linear_prepack16_launch(const float* w, int K, int N, void* packed, cudaStream_t st);
int linear_tc16_launch(const float* x, int ldx, const void* packed, const float* scale, const float* shift,
                       int act, float* y, int ldy, int M, int K, int N, cudaStream_t st);
int linear_rowdot_tc16_launch(const float* x, int ldx, const void* packed, const float* scale,
                              const float* shift, int act, const float* w2, float b2, int act2, float* y2,
                              int M, int K, int N, cudaStream_t st);
int linear_chain_tc16_launch(const float* x, int ldx, const void* packed1, const float* scale1, const float* shift1,
                             int act1, const void* packed2, const float* scale2, const float* shift2, int act2,
                             float* y, int ldy, int M, int K1, int N1, int N2, cudaStream_t st);
int linear_join_tc16_launch(const float* xa, int ldxa, const void* packed_a, const float* scale_a,
                            const float* shift_a, int act_a, const float* xb, int ldxb, const void* packed_b,
                            const float* scale_b, const float* shift_b, int act_b, float* y, int ldy, float* yn,
                            int ldn, float eps, int M, int Ka, int Kb, int N, cudaStream_t st);
// flexconv.cu
size_t flex_conv_prepack_bytes(int Din, int Dout);
int flex_conv_prepack(const float* theta, const float* bias, const float* feature_bias, const float* scale,
                      const float* shift, int Din, int Dout, void* packed, cudaStream_t st);
size_t flex_conv_pm_packed_workspace_bytes(int B, int N, int K, int Din, int Dout);
int flex_conv_pm_packed(const float* feat, const void* packed, const int32_t* nbr, const float* xyz, float* out,
                        int B, int N, int K, int Din, int Dout, const float* scale, int act, void* ws,
                        size_t ws_bytes, cudaStream_t st);
size_t flex_conv_pm_total_workspace_bytes(int B, int N, int K, int Din, int Dout);
size_t flex_conv_cm_workspace_bytes(int B, int N, int K, int Din, int Dout);
int flex_conv_pm(const float* feat, const float* theta, const float* bias, const int32_t* nbr,
                 const float* xyz, float* out, int B, int N, int K, int Din, int Dout,
                 const float* feature_bias, const float* scale, const float* shift, int act,
                 void* ws, size_t ws_bytes, cudaStream_t st);
int flex_conv_cm(const float* feat_cm, const float* theta, const float* bias, const int32_t* nbr_cm,
                 const float* pos_cm, float* out_cm, int B, int N, int K, int Din, int Dout, void* ws,
                 size_t ws_bytes, cudaStream_t st);
// backward.cu
size_t flex_conv_grad_pm_workspace_bytes(int B, int N, int K, int Din, int Dout);
int flex_conv_grad_pm(const float* feat, const float* theta, const float* bias, const int32_t* nbr, const float* xyz,
                      const float* topdiff, float* grad_feat, float* grad_theta, float* grad_bias, int B, int N, int K,
                      int Din, int Dout, void* ws, size_t ws_bytes, cudaStream_t st);
size_t flex_conv_grad_cm_workspace_bytes(int B, int N, int K, int Din, int Dout);
int flex_conv_grad_cm(const float* feat_cm, const float* theta, const float* bias, const int32_t* nbr_cm,
                      const float* pos_cm, const float* topdiff_cm, float* grad_feat_cm, float* grad_theta,
                      float* grad_bias, int B, int N, int K, int Din, int Dout, void* ws, size_t ws_bytes,
                      cudaStream_t st);
size_t flex_deconv_cm_workspace_bytes(int B, int N, int K, int Din, int Dout);
int flex_deconv_cm(const float* feat_cm, const float* theta, const float* bias, const int32_t* nbr_cm,
                   const float* pos_cm, float* out_cm, int B, int N, int K, int Din, int Dout, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int flex_pool_grad_cm(const float* topdiff, const int32_t* argmax, float* grad_feat, int B, int N, int D,
                      cudaStream_t st);
size_t conv_pointset_grad_cm_workspace_bytes(int B, int N, int K, int Din, int Dout);
int conv_pointset_grad_cm(const float* feat_cm, const float* theta, const int32_t* nbr_cm, const float* topdiff_cm,
                          float* grad_feat_cm, float* grad_theta, float* grad_bias, int B, int N, int K, int Din,
                          int Dout, void* ws, size_t ws_bytes, cudaStream_t st);
int group_point_grad_launch(int b, int n, int c, int m, int nsample, const float* grad_out, const int32_t* idx,
                            float* grad_points, cudaStream_t st);
int three_interpolate_grad_launch(int b, int n, int c, int m, const float* grad_out, const int32_t* idx,
                                  const float* weight, float* grad_points, cudaStream_t st);
// nms.cu
int affine_launch(const float* x, float a, float b, float* y, size_t n, cudaStream_t st);
int gather_rows_launch(int b, int n, int c, int m, const float* src, const int32_t* idx, float* out, int ldo,
                       cudaStream_t st);
size_t keypoint_nms_workspace_bytes(int B, int N);
int keypoint_nms_launch(const float* xyz, const float* attention, int B, int N, float nms_radius,
                        float min_response_ratio, int max_keypoints, int remove_noise, int32_t* out_idx,
                        int32_t* out_cnt, void* ws, size_t ws_bytes, cudaStream_t st);
// topk.cu
int topk_l2_exact_launch(const float* gram, int ldg, const float* qn, const float* rn, const float* qd,
                         const float* rd, int Q, int R, int D, int K, int32_t* idx, float* val, int32_t* cand,
                         float* cand_val, cudaStream_t st);
int topk_l2_launch(const float* gram, int ldg, const float* qn, const float* rn, int Q, int R, int K,
                   int32_t* idx, float* val, cudaStream_t st);
// netvlad.cu
size_t netvlad_workspace_bytes(int B, int N, int D, int Kc, int out_dim);
int netvlad_launch(const float* features, const float* att, int B, int N, int D, int Kc, int out_dim,
                   const float* cw, const float* cbn_scale, const float* cbn_shift, const float* cw2,
                   const float* hw, const float* bn_scale, const float* bn_shift, const float* gw,
                   const float* gbn_scale, const float* gbn_shift, int final_l2norm, float* out,
                   void* ws, size_t ws_bytes, cudaStream_t st);

// The ONE process-wide switch of this library: DH3D_EXACT_FP32=1 selects the exact-fp32 debug path -- FFMA GEMM
// stages inside FlexConv (two-kernel form) and the FFMA NetVLAD aggregation -- instead of the tcgen05 kernels.
// (The Python layer reads the same variable to skip dh3d_linear_prepack and call dh3d_linear.)  No buffer layout
// depends on it: prepacked weights always carry every form a kernel may read.
bool exact_fp32() {
  static const bool on = [] {
    const char* e = getenv("DH3D_EXACT_FP32");
    return e && e[0] != '\0' && e[0] != '0';
  }();
  return on;
}

// GEMM dispatch (one place to switch the dense path)
int linear_launch(const float* x, int ldx, const float* w, const float* scale, const float* shift,
                  int act, float* y, int ldy, int M, int K, int N, cudaStream_t st) {
  return linear_simt_launch(x, ldx, w, scale, shift, act, y, ldy, M, K, N, st);
}
}  // namespace dh3d

using namespace dh3d;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int dh3d_version(void) { return 100; }

const char* dh3d_error_string(int code) {
  switch (code) {
    case DH3D_OK: return "ok";
    case DH3D_ERR_NULL: return "a required pointer is NULL";
    case DH3D_ERR_DIM: return "a dimension is <= 0 or inconsistent";
    case DH3D_ERR_UNSUPPORTED: return "size or attribute outside the supported range";
    case DH3D_ERR_WORKSPACE: return "workspace missing or too small";
    case DH3D_ERR_ALIGN: return "pointer not 16-byte aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown dh3d error";
  }
}

size_t dh3d_knn_workspace_bytes(int B, int N) { return knn_workspace_bytes(B, N); }

int dh3d_knn_bruteforce(const float* positions_cm, int B, int Dp, int N, int K, int32_t* ids,
                        float* dists, void* workspace, size_t workspace_bytes, void* stream) {
  if (Dp != 3) return Dp <= 0 ? DH3D_ERR_DIM : DH3D_ERR_UNSUPPORTED;
  return knn_launch(positions_cm, B, N, K, 3LL * N, 1, N, ids, dists, workspace, workspace_bytes,
                    S(stream));
}

int dh3d_knn_bruteforce_pm(const float* xyz_pm, int B, int N, int K, int32_t* ids, float* dists,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return knn_launch(xyz_pm, B, N, K, 3LL * N, 3, 1, ids, dists, workspace, workspace_bytes,
                    S(stream));
}

size_t dh3d_flex_conv_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return flex_conv_cm_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_flex_conv(const float* features_cm, const float* theta, const float* bias,
                   const int32_t* neighborhood_cm, const float* positions_cm, float* out_cm, int B,
                   int N, int K, int Din, int Dout, void* workspace, size_t workspace_bytes,
                   void* stream) {
  return flex_conv_cm(features_cm, theta, bias, neighborhood_cm, positions_cm, out_cm, B, N, K, Din,
                      Dout, workspace, workspace_bytes, S(stream));
}
size_t dh3d_flex_conv_prepack_bytes(int Din, int Dout) { return flex_conv_prepack_bytes(Din, Dout); }
int dh3d_flex_conv_prepack(const float* theta, const float* bias, const float* feature_bias, const float* scale,
                           const float* shift, int Din, int Dout, void* packed, void* stream) {
  return flex_conv_prepack(theta, bias, feature_bias, scale, shift, Din, Dout, packed, S(stream));
}
size_t dh3d_flex_conv_pm_packed_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return flex_conv_pm_packed_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_flex_conv_pm_packed(const float* features_pm, const void* packed, const int32_t* neighborhood_pm,
                             const float* xyz_pm, float* out_pm, int B, int N, int K, int Din, int Dout,
                             const float* scale, int act, void* workspace, size_t workspace_bytes,
                             void* stream) {
  return flex_conv_pm_packed(features_pm, packed, neighborhood_pm, xyz_pm, out_pm, B, N, K, Din, Dout, scale,
                             act, workspace, workspace_bytes, S(stream));
}
size_t dh3d_flex_conv_pm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return flex_conv_pm_total_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_flex_conv_pm(const float* features_pm, const float* theta, const float* bias,
                      const int32_t* neighborhood_pm, const float* xyz_pm, float* out_pm, int B,
                      int N, int K, int Din, int Dout, const float* feature_bias,
                      const float* scale, const float* shift, int act, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return flex_conv_pm(features_pm, theta, bias, neighborhood_pm, xyz_pm, out_pm, B, N, K, Din, Dout,
                      feature_bias, scale, shift, act, workspace, workspace_bytes, S(stream));
}

int dh3d_flex_pool(const float* features_cm, const int32_t* neighborhood_cm, float* out_cm,
                   int32_t* argmax_cm, int B, int N, int K, int D, void* stream) {
  return flex_pool_cm_launch(features_cm, neighborhood_cm, out_cm, argmax_cm, B, N, K, D, S(stream));
}
int dh3d_flex_pool_pm(const float* features_pm, const int32_t* neighborhood_pm, float* out_pm,
                      int32_t* argmax_pm, int B, int N, int K, int D, void* stream) {
  return flex_pool_pm_launch(features_pm, neighborhood_pm, out_pm, argmax_pm, B, N, K, D, S(stream));
}

int dh3d_conv_pointset(const float* features_cm, const float* theta, const float* bias,
                       const int32_t* neighborhood_cm, float* out_cm, int B, int N, int K, int Din,
                       int Dout, void* stream) {
  return conv_pointset_cm_launch(features_cm, theta, bias, neighborhood_cm, out_cm, B, N, K, Din,
                                 Dout, S(stream));
}
int dh3d_conv_pointset_pm(const float* features_pm, const float* theta, const float* bias,
                          const int32_t* neighborhood_pm, float* out_pm, int B, int N, int K,
                          int Din, int Dout, const float* scale, const float* shift, int act,
                          void* stream) {
  return conv_pointset_pm_launch(features_pm, theta, bias, neighborhood_pm, out_pm, B, N, K, Din,
                                 Dout, scale, shift, act, S(stream));
}

int dh3d_knn_sort_pm(const float* xyz_pm, int B, int N, void* workspace, size_t workspace_bytes, void* stream) {
  return knn_sort_launch(xyz_pm, B, N, 3LL * N, 3, 1, workspace, workspace_bytes, S(stream));
}
int dh3d_knn_query_sorted(const void* workspace, int B, int N, int K, int32_t* ids, float* dists, void* stream) {
  return knn_query_sorted_launch(workspace, B, N, K, ids, dists, S(stream));
}
int dh3d_farthest_point_sample_presorted(int b, int n, int m, const void* knn_workspace_of_inp, int32_t* out,
                                         void* stream) {
  return fps_presorted_launch(b, n, m, knn_workspace_of_inp, out, S(stream));
}
int dh3d_farthest_point_sample(int b, int n, int m, const float* inp, int32_t* out, void* stream) {
  return fps_launch(b, n, m, inp, out, S(stream));
}
int dh3d_gather_point(int b, int n, int m, const float* inp, const int32_t* idx, float* out,
                      void* stream) {
  return group_point_launch(b, n, 3, m, 1, inp, idx, out, S(stream));
}
int dh3d_group_point(int b, int n, int c, int m, int nsample, const float* points,
                     const int32_t* idx, float* out, void* stream) {
  return group_point_launch(b, n, c, m, nsample, points, idx, out, S(stream));
}
size_t dh3d_query_ball_point_workspace_bytes(int b, int m) { return query_ball_workspace_bytes(b, m); }
int dh3d_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1,
                          const float* xyz2, int32_t* idx, int32_t* pts_cnt, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return query_ball_launch(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, workspace,
                           workspace_bytes, S(stream));
}
int dh3d_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                  int32_t* idx, void* stream) {
  return three_nn_launch(b, n, m, xyz1, xyz2, dist, idx, S(stream));
}
size_t dh3d_three_nn_workspace_bytes(int b, int n, int m) { return three_nn_workspace_bytes(b, n, m); }
int dh3d_three_nn_presorted2(int b, int n, int m, const void* knn_workspace_of_xyz1, const void* knn_workspace_of_xyz2,
                             float* dist, int32_t* idx, void* stream) {
  return three_nn_presorted2_launch(b, n, m, knn_workspace_of_xyz1, knn_workspace_of_xyz2, dist, idx, S(stream));
}
int dh3d_three_nn_ws_presorted(int b, int n, int m, const void* knn_workspace_of_xyz1, const float* xyz2, float* dist,
                               int32_t* idx, void* workspace, size_t workspace_bytes, void* stream) {
  return three_nn_presorted_launch(b, n, m, knn_workspace_of_xyz1, xyz2, dist, idx, workspace, workspace_bytes,
                                   S(stream));
}
int dh3d_three_nn_ws(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                     int32_t* idx, void* workspace, size_t workspace_bytes, void* stream) {
  return three_nn_pruned_launch(b, n, m, xyz1, xyz2, dist, idx, workspace, workspace_bytes, S(stream));
}
int dh3d_three_interpolate(int b, int m, int c, int n, const float* points, const int32_t* idx,
                           const float* weight, float* out, void* stream) {
  return three_interpolate_launch(b, m, c, n, points, idx, weight, out, false, S(stream));
}
int dh3d_three_interpolate_from_dist(int b, int m, int c, int n, const float* points,
                                     const int32_t* idx, const float* dist2, float* out,
                                     void* stream) {
  return three_interpolate_launch(b, m, c, n, points, idx, dist2, out, true, S(stream));
}

int dh3d_linear(const float* x, int ldx, const float* w, const float* scale, const float* shift,
                int act, float* y, int ldy, int M, int K, int N, void* stream) {
  return linear_launch(x, ldx, w, scale, shift, act, y, ldy, M, K, N, S(stream));
}
size_t dh3d_linear_prepack_bytes(int K, int N) { return linear_prepack16_bytes(K, N); }
int dh3d_linear_prepack(const float* w, int K, int N, void* packed, void* stream) {
  return linear_prepack16_launch(w, K, N, packed, S(stream));
}
int dh3d_linear_packed(const float* x, int ldx, const void* packed_w, const float* scale,
                       const float* shift, int act, float* y, int ldy, int M, int K, int N,
                       void* stream) {
  return linear_tc16_launch(x, ldx, packed_w, scale, shift, act, y, ldy, M, K, N, S(stream));
}
int dh3d_linear_rowdot_packed(const float* x, int ldx, const void* packed_w, const float* scale,
                              const float* shift, int act, const float* w2, float b2, int act2,
                              float* y, int M, int K, int N, void* stream) {
  return linear_rowdot_tc16_launch(x, ldx, packed_w, scale, shift, act, w2, b2, act2, y, M, K, N, S(stream));
}
int dh3d_linear_chain_packed(const float* x, int ldx, const void* packed_w1, const float* scale1, const float* shift1,
                             int act1, const void* packed_w2, const float* scale2, const float* shift2, int act2,
                             float* y, int ldy, int M, int K1, int N1, int N2, void* stream) {
  return linear_chain_tc16_launch(x, ldx, packed_w1, scale1, shift1, act1, packed_w2, scale2, shift2, act2, y, ldy, M,
                                  K1, N1, N2, S(stream));
}

int dh3d_linear_join_packed(const float* xa, int ldxa, const void* packed_wa, const float* scale_a,
                            const float* shift_a, int act_a, const float* xb, int ldxb, const void* packed_wb,
                            const float* scale_b, const float* shift_b, int act_b, float* y, int ldy,
                            float* y_normalized, int ldn, float eps, int M, int Ka, int Kb, int N, void* stream) {
  return linear_join_tc16_launch(xa, ldxa, packed_wa, scale_a, shift_a, act_a, xb, ldxb, packed_wb, scale_b,
                                 shift_b, act_b, y, ldy, y_normalized, ldn, eps, M, Ka, Kb, N, S(stream));
}
int dh3d_rowdot(const float* x, int ldx, const float* w, float bias, int act, float* y, int M,
                int K, void* stream) {
  return rowdot_launch(x, ldx, w, bias, act, y, M, K, S(stream));
}
int dh3d_se_excite(const float* x, const float* gate, float* y, size_t count, void* stream) {
  return se_excite_launch(x, gate, y, count, S(stream));
}
int dh3d_se_pool_excite(const float* x, const int32_t* neighborhood, const float* w1, const float* b1,
                        const float* w2, const float* b2, float* y, int B, int N, int K, int C, int H,
                        void* stream) {
  return se_pool_excite_launch(x, neighborhood, w1, b1, w2, b2, y, B, N, K, C, H, S(stream));
}
int dh3d_l2_normalize_rows(const float* x, int ldx, float* y, int ldy, int M, int C, float eps,
                           void* stream) {
  return l2norm_rows_launch(x, ldx, y, ldy, M, C, eps, S(stream));
}
int dh3d_add(const float* a, const float* b, float* y, size_t count, void* stream) {
  return add_launch(a, b, y, count, S(stream));
}
int dh3d_copy_cols(const float* src, int lds, float* dst, int ldd, int M, int C, void* stream) {
  return copy_cols_launch(src, lds, dst, ldd, M, C, S(stream));
}
int dh3d_transpose_cm_to_pm(const void* src_cm, void* dst_pm, int B, int C, int N, void* stream) {
  return transpose_launch(src_cm, dst_pm, B, C, N, S(stream));
}
int dh3d_transpose_pm_to_cm(const void* src_pm, void* dst_cm, int B, int N, int C, void* stream) {
  return transpose_launch(src_pm, dst_cm, B, N, C, S(stream));
}

int dh3d_topk_l2_exact(const float* gram, int ldg, const float* qn, const float* rn, const float* query,
                       const float* ref, int Q, int R, int D, int K, int32_t* idx, float* val, int32_t* cand,
                       float* cand_val, void* stream) {
  return topk_l2_exact_launch(gram, ldg, qn, rn, query, ref, Q, R, D, K, idx, val, cand, cand_val, S(stream));
}
int dh3d_topk_l2(const float* gram, int ldg, const float* qn, const float* rn, int Q, int R, int K,
                 int32_t* idx, float* val, void* stream) {
  return topk_l2_launch(gram, ldg, qn, rn, Q, R, K, idx, val, S(stream));
}

size_t dh3d_netvlad_workspace_bytes(int B, int N, int D, int Kc, int out_dim) {
  return netvlad_workspace_bytes(B, N, D, Kc, out_dim);
}
int dh3d_netvlad(const float* features, const float* att, int B, int N, int D, int Kc, int out_dim,
                 const float* cluster_weights, const float* cluster_bn_scale,
                 const float* cluster_bn_shift, const float* cluster_weights2,
                 const float* hidden1_weights, const float* bn_scale, const float* bn_shift,
                 const float* gating_weights, const float* gating_bn_scale,
                 const float* gating_bn_shift, int final_l2norm, float* out, void* workspace,
                 size_t workspace_bytes, void* stream) {
  return netvlad_launch(features, att, B, N, D, Kc, out_dim, cluster_weights, cluster_bn_scale,
                        cluster_bn_shift, cluster_weights2, hidden1_weights, bn_scale, bn_shift,
                        gating_weights, gating_bn_scale, gating_bn_shift, final_l2norm, out,
                        workspace, workspace_bytes, S(stream));
}

size_t dh3d_flex_conv_grad_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return flex_conv_grad_cm_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_flex_conv_grad(const float* features_cm, const float* theta, const float* bias,
                        const int32_t* neighborhood_cm, const float* positions_cm, const float* topdiff_cm,
                        float* grad_features_cm, float* grad_theta, float* grad_bias, int B, int N, int K,
                        int Din, int Dout, void* workspace, size_t workspace_bytes, void* stream) {
  return flex_conv_grad_cm(features_cm, theta, bias, neighborhood_cm, positions_cm, topdiff_cm, grad_features_cm,
                           grad_theta, grad_bias, B, N, K, Din, Dout, workspace, workspace_bytes, S(stream));
}
size_t dh3d_flex_conv_grad_pm_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return flex_conv_grad_pm_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_flex_conv_grad_pm(const float* features_pm, const float* theta, const float* bias,
                           const int32_t* neighborhood_pm, const float* xyz_pm, const float* topdiff_pm,
                           float* grad_features_pm, float* grad_theta, float* grad_bias, int B, int N, int K,
                           int Din, int Dout, void* workspace, size_t workspace_bytes, void* stream) {
  return flex_conv_grad_pm(features_pm, theta, bias, neighborhood_pm, xyz_pm, topdiff_pm, grad_features_pm, grad_theta,
                           grad_bias, B, N, K, Din, Dout, workspace, workspace_bytes, S(stream));
}
int dh3d_flex_pool_grad(const float* topdiff_cm, const int32_t* argmax_cm, float* grad_features_cm, int B, int N,
                        int D, void* stream) {
  return flex_pool_grad_cm(topdiff_cm, argmax_cm, grad_features_cm, B, N, D, S(stream));
}
size_t dh3d_conv_pointset_grad_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return conv_pointset_grad_cm_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_conv_pointset_grad(const float* features_cm, const float* theta, const int32_t* neighborhood_cm,
                            const float* topdiff_cm, float* grad_features_cm, float* grad_theta,
                            float* grad_bias, int B, int N, int K, int Din, int Dout, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return conv_pointset_grad_cm(features_cm, theta, neighborhood_cm, topdiff_cm, grad_features_cm, grad_theta, grad_bias,
                               B, N, K, Din, Dout, workspace, workspace_bytes, S(stream));
}
size_t dh3d_flex_deconv_workspace_bytes(int B, int N, int K, int Din, int Dout) {
  return flex_deconv_cm_workspace_bytes(B, N, K, Din, Dout);
}
int dh3d_flex_deconv(const float* features_cm, const float* theta, const float* bias,
                     const int32_t* neighborhood_cm, const float* positions_cm, float* out_cm, int B, int N,
                     int K, int Din, int Dout, void* workspace, size_t workspace_bytes, void* stream) {
  return flex_deconv_cm(features_cm, theta, bias, neighborhood_cm, positions_cm, out_cm, B, N, K, Din, Dout, workspace,
                        workspace_bytes, S(stream));
}
int dh3d_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int32_t* idx,
                          float* grad_points, void* stream) {
  return group_point_grad_launch(b, n, c, m, nsample, grad_out, idx, grad_points, S(stream));
}
int dh3d_gather_point_grad(int b, int n, int m, const float* out_g, const int32_t* idx, float* inp_g,
                           void* stream) {
  return group_point_grad_launch(b, n, 3, m, 1, out_g, idx, inp_g, S(stream));
}
int dh3d_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int32_t* idx,
                                const float* weight, float* grad_points, void* stream) {
  return three_interpolate_grad_launch(b, n, c, m, grad_out, idx, weight, grad_points, S(stream));
}

int dh3d_affine(const float* x, float a, float b, float* y, size_t count, void* stream) {
  return affine_launch(x, a, b, y, count, S(stream));
}
int dh3d_gather_rows(int b, int n, int c, int m, const float* src, const int32_t* idx, float* out, int ldo,
                     void* stream) {
  return gather_rows_launch(b, n, c, m, src, idx, out, ldo, S(stream));
}
size_t dh3d_keypoint_nms_workspace_bytes(int B, int N) { return keypoint_nms_workspace_bytes(B, N); }
int dh3d_keypoint_nms(const float* xyz_pm, const float* attention, int B, int N, float nms_radius,
                      float min_response_ratio, int max_keypoints, int remove_noise, int32_t* out_idx,
                      int32_t* out_cnt, void* workspace, size_t workspace_bytes, void* stream) {
  return keypoint_nms_launch(xyz_pm, attention, B, N, nms_radius, min_response_ratio, max_keypoints, remove_noise,
                             out_idx, out_cnt, workspace, workspace_bytes, S(stream));
}

int dh3d_group_point_ld(int b, int n, int c, int m, int nsample, const float* points, int ldp, const int32_t* idx,
                        float* out, void* stream) {
  return group_point_ld_launch(b, n, c, m, nsample, points, ldp, idx, out, S(stream));
}
int dh3d_three_interpolate_ld(int b, int m, int c, int n, const float* points, const int32_t* idx,
                              const float* weight_or_dist2, int weight_is_dist2, float* out, int ldo, void* stream) {
  return three_interpolate_ld_launch(b, m, c, n, points, idx, weight_or_dist2, out, ldo, weight_is_dist2 != 0,
                                     S(stream));
}
int dh3d_add_l2_normalize_rows(const float* a, const float* b, float* sum, float* normalized, int M, int C,
                               float eps, void* stream) {
  return add_l2norm_rows_launch(a, b, sum, normalized, M, C, eps, S(stream));
}

}  // extern "C"
