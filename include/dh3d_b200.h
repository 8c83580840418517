/*
 * dh3d_b200.h -- C ABI of libdh3d_b200.so: the DH3D point-cloud feature-extraction hot path
 * as hand-written sm_100a CUDA.  This is the drop-in boundary: every entry point replaces one
 * TensorFlow custom op (or one TF-library block) of the reference and takes exactly what the
 * reference's OpKernel::Compute / *Launcher gets -- raw device pointers, dimensions, a stream.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer to a dense, contiguous fp32 / int32 buffer owned by the
 *     caller (the reference: `ctx->allocate_output`, e.g. user_ops/kernels/flex_conv_op.cc:47-49);
 *     inputs are never written; outputs are fully overwritten;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); launches are
 *     asynchronous, nothing here synchronises (the reference's kNN does a cudaDeviceSynchronize,
 *     knn_bruteforce_kernel_gpu.cu.cc:222 -- deliberately not inherited);
 *   - re-entrant, no global mutable state; `workspace` buffers are caller-owned scratch whose
 *     minimum size comes from the matching *_workspace_bytes() query;
 *   - return value: DH3D_OK (0), a negative DH3D_ERR_* argument-validation code, or a positive
 *     cudaError_t from the launch (the reference: TF Status via OP_REQUIRES / errors::Internal,
 *     flex_conv_kernel_gpu.cu.cc:436-439);
 *   - layouts: "cm" = channel-major [B,C,N] (the reference's user_ops layout), "pm" = point-major
 *     [B,N,C] (the reference's tf_ops layout and this library's native layout).
 */
#ifndef DH3D_B200_H_
#define DH3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DH3D_API __attribute__((visibility("default")))
#else
#define DH3D_API
#endif

#define DH3D_OK 0
#define DH3D_ERR_NULL (-1)        /* a required pointer is NULL */
#define DH3D_ERR_DIM (-2)         /* a dimension is <= 0 or inconsistent */
#define DH3D_ERR_UNSUPPORTED (-3) /* size / attribute outside what the kernels support */
#define DH3D_ERR_WORKSPACE (-4)   /* workspace missing or too small */
#define DH3D_ERR_ALIGN (-5)       /* pointer not aligned for vector access (16 bytes) */

#define DH3D_ACT_NONE 0
#define DH3D_ACT_RELU 1
#define DH3D_ACT_SIGMOID 2

DH3D_API int dh3d_version(void);                 /* 100 * major + minor */
DH3D_API const char* dh3d_error_string(int code); /* static string for a DH3D_ERR_* / cudaError_t code */

/* ---------------------------------------------------------------------------------------------
 * k-NN  -- replaces op KnnBruteforce
 *   reference: user_ops/ops/knn_bruteforce.cc:11-35 (schema), kernels/knn_bruteforce_op.cc:30-60,
 *   kernels/knn_bruteforce_kernel_gpu.cu.cc:45-134,162-228; python user_ops/__init__.py:50.
 *   positions [B,3,N] cm (or [B,N,3] pm for the _pm entry) -> ids [B,N,K] i32 (self included,
 *   ascending Euclidean distance, ties in the reference's BlockRadixSort blocked order), dists
 *   [B,N,K] f32 (sqrt'ed).  K in [1,64]; N in [1, 65536]; Dp must be 3 (the only value DH3D uses).
 *   Unlike the reference there is no N <= 8192 cap (kernel_gpu.cu.cc:213-221); for N > 8192 the
 *   tie order is plain index order.
 * ------------------------------------------------------------------------------------------- */
DH3D_API size_t dh3d_knn_workspace_bytes(int B, int N);
DH3D_API int dh3d_knn_bruteforce(const float* positions_cm, int B, int Dp, int N, int K, int32_t* ids,
                        float* dists, void* workspace, size_t workspace_bytes, void* stream);
DH3D_API int dh3d_knn_bruteforce_pm(const float* xyz_pm, int B, int N, int K, int32_t* ids, float* dists,
                           void* workspace, size_t workspace_bytes, void* stream);
/* dh3d_knn_bruteforce_pm in its two halves, for callers that share the cell-sorted copy of the cloud between ops
 * (dh3d_farthest_point_sample_presorted, dh3d_three_nn_ws_presorted): sort xyz_pm [B,N,3] into the workspace
 * (dh3d_knn_workspace_bytes(B, N)), then answer the K-NN query from it.  sort + query == dh3d_knn_bruteforce_pm. */
DH3D_API int dh3d_knn_sort_pm(const float* xyz_pm, int B, int N, void* workspace, size_t workspace_bytes, void* stream);
DH3D_API int dh3d_knn_query_sorted(const void* workspace, int B, int N, int K, int32_t* ids, float* dists, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FlexConv forward -- replaces op FlexConv
 *   reference: user_ops/ops/flex_conv.cc:25-100, kernels/flex_conv_op.cc:35-53,
 *   kernels/flex_conv_kernel_gpu.cu.cc:44-158,402-441; python user_ops/__init__.py:63-89
 *   (note the op's input order features, theta, bias, neighborhood, position).
 *   out[b,o,n] = sum_k sum_c (bias[c,o] + sum_dp theta[dp,c,o]*(p[nbr_k]-p[n])[dp]) * f[c,nbr_k]
 *   Centre = p[n] (the CUDA kernel's rule, :75-79,109).
 *
 *   dh3d_flex_conv: reference layouts -- features [B,Din,N], theta [3,Din,Dout], bias [Din,Dout],
 *   neighborhood [B,K,N] i32, positions [B,3,N] -> out [B,Dout,N].
 *   dh3d_flex_conv_pm: native layouts -- features [B,N,Din], neighborhood [B,N,K], xyz [B,N,3]
 *   -> out [B,N,Dout] with an optional fused epilogue
 *       y = act( (x + feature_bias[o]) * scale[o] + shift[o] )
 *   (feature_bias: core/layers.py:330-331; scale/shift: folded inference BatchNorm,
 *   core/tf_utils.py:58-63).  Any of feature_bias/scale/shift may be NULL.
 *   Din % 4 == 0, Dout % 4 == 0, K in [1,64].
 * ------------------------------------------------------------------------------------------- */
DH3D_API size_t dh3d_flex_conv_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_flex_conv(const float* features_cm, const float* theta, const float* bias,
                   const int32_t* neighborhood_cm, const float* positions_cm, float* out_cm, int B,
                   int N, int K, int Din, int Dout, void* workspace, size_t workspace_bytes,
                   void* stream);
DH3D_API size_t dh3d_flex_conv_pm_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_flex_conv_pm(const float* features_pm, const float* theta, const float* bias,
                      const int32_t* neighborhood_pm, const float* xyz_pm, float* out_pm, int B,
                      int N, int K, int Din, int Dout, const float* feature_bias,
                      const float* scale, const float* shift, int act, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Inference form with the weight-only work hoisted out of the forward (the Keras layer of the reference owns
 * position_theta / position_bias / feature_bias, core/layers.py:252-288, and the BatchNorm that follows it,
 * core/tf_utils.py:58-63): dh3d_flex_conv_prepack derives, once per layer, the contraction operand
 * [position_bias; theta_x; theta_y; theta_z] in the kernel's layout plus the folded shift
 * feature_bias*scale + shift; dh3d_flex_conv_pm_packed is dh3d_flex_conv_pm on that buffer (pass the SAME
 * scale).  packed: dh3d_flex_conv_prepack_bytes(Din, Dout) bytes, 256-byte aligned.  The buffer holds the operand
 * in every form a kernel may read (plain fp32 for the two-kernel FFMA form, {hi^T, lo^T} tf32 pairs for the fused
 * tcgen05 kernel), so it does not depend on any setting of the process that made it. */
DH3D_API size_t dh3d_flex_conv_prepack_bytes(int Din, int Dout);
DH3D_API int dh3d_flex_conv_prepack(const float* theta, const float* bias, const float* feature_bias,
                           const float* scale, const float* shift, int Din, int Dout, void* packed,
                           void* stream);
DH3D_API size_t dh3d_flex_conv_pm_packed_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_flex_conv_pm_packed(const float* features_pm, const void* packed,
                             const int32_t* neighborhood_pm, const float* xyz_pm, float* out_pm, int B,
                             int N, int K, int Din, int Dout, const float* scale, int act,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FlexPool forward -- replaces op FlexPool
 *   reference: user_ops/ops/flex_pool.cc:25-69, kernels/flex_pool_kernel_gpu.cu.cc:30-63,100-125;
 *   python user_ops/__init__.py:115-135.  max over the K neighbours per channel; argmax is the
 *   GLOBAL point id; first neighbour reaching the max wins.  argmax may be NULL.
 * ------------------------------------------------------------------------------------------- */
DH3D_API int dh3d_flex_pool(const float* features_cm, const int32_t* neighborhood_cm, float* out_cm,
                   int32_t* argmax_cm, int B, int N, int K, int D, void* stream);
DH3D_API int dh3d_flex_pool_pm(const float* features_pm, const int32_t* neighborhood_pm, float* out_pm,
                      int32_t* argmax_pm, int B, int N, int K, int D, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ConvPointset forward ("conv_relative") -- replaces op ConvPointset
 *   reference: user_ops/ops/conv_pointset.cc:26-87, kernels/conv_pointset_kernel_gpu.cu.cc:45-147,
 *   364-402; python user_ops/__init__.py:205-225.
 *   out[b,o,n] = bias[o] + sum_k sum_c theta[c,o]*(f[c,nbr_k]-f[c,nbr_0]);  theta [Din,Dout].
 *   Din <= 64 (one Din chunk of the reference kernel; DH3D uses Din = 3).  The _pm entry has the
 *   same optional epilogue as dh3d_flex_conv_pm (scale/shift/act).
 * ------------------------------------------------------------------------------------------- */
DH3D_API int dh3d_conv_pointset(const float* features_cm, const float* theta, const float* bias,
                       const int32_t* neighborhood_cm, float* out_cm, int B, int N, int K, int Din,
                       int Dout, void* stream);
DH3D_API int dh3d_conv_pointset_pm(const float* features_pm, const float* theta, const float* bias,
                          const int32_t* neighborhood_pm, float* out_pm, int B, int N, int K,
                          int Din, int Dout, const float* scale, const float* shift, int act,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * PointNet++ ops -- replace the raw-pointer launchers of tf_ops (already a C-like layer there)
 *   farthestpointsamplingLauncher(b,n,m,inp,temp,out)   tf_ops/sampling/tf_sampling_g.cu:105-170,203
 *   gatherpointLauncher(b,n,m,inp,idx,out)              tf_sampling_g.cu:172-181,206
 *   groupPointLauncher(b,n,c,m,nsample,points,idx,out)  tf_ops/grouping/tf_grouping_g.cu:94-111,191
 *   queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt) tf_grouping_g.cu:3-52,179
 *   threenn_cpu / threeinterpolate_cpu                  tf_ops/interpolation/tf_interpolate.cpp:60-127
 *   All point-major.  FPS needs no `temp` scratch (state lives in registers / shared memory).
 *   FPS: n <= 65536; the sample order reproduces the reference's 512-thread tie rule.
 * ------------------------------------------------------------------------------------------- */
DH3D_API int dh3d_farthest_point_sample(int b, int n, int m, const float* inp, int32_t* out, void* stream);
/* The same sampling on a cloud that is already cell-sorted: knn_workspace_of_inp = the workspace of a
 * dh3d_knn_bruteforce_pm / dh3d_knn_sort_pm call on the same inp [b,n,3] (n <= 8192).  Identical indices in identical
 * order (same distance arithmetic, same tie rule); a round only revisits the 32-point chunks whose bounding box can still
 * hold a point whose running min-distance changes, so it is several times less work than the exhaustive rounds.
 * DH3D's graph runs KnnBruteforce on the dense cloud before anything else (core/model.py:157), so the sort exists. */
DH3D_API int dh3d_farthest_point_sample_presorted(int b, int n, int m, const void* knn_workspace_of_inp, int32_t* out,
                                         void* stream);
DH3D_API int dh3d_gather_point(int b, int n, int m, const float* inp, const int32_t* idx, float* out,
                      void* stream);
DH3D_API int dh3d_group_point(int b, int n, int c, int m, int nsample, const float* points,
                     const int32_t* idx, float* out, void* stream);
DH3D_API size_t dh3d_query_ball_point_workspace_bytes(int b, int m);
DH3D_API int dh3d_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1,
                          const float* xyz2, int32_t* idx, int32_t* pts_cnt, void* workspace,
                          size_t workspace_bytes, void* stream);
DH3D_API int dh3d_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                  int32_t* idx, void* stream);
/* Same results through the k-NN engine (Morton-sorted known points with chunk bounding boxes, exact
 * box skipping with the reference's un-fused arithmetic): needs a caller workspace. */
DH3D_API size_t dh3d_three_nn_workspace_bytes(int b, int n, int m);
DH3D_API int dh3d_three_nn_ws(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                     int32_t* idx, void* workspace, size_t workspace_bytes, void* stream);
/* dh3d_three_nn_ws with the QUERY cloud already sorted: knn_workspace_of_xyz1 = the workspace a dh3d_knn_bruteforce_pm
 * call on the same xyz1 [b,n,3] left behind (it starts with the cell-sorted (x,y,z,index) array this scan walks), so
 * the query sort is skipped -- DH3D's forward runs the k-NN of the dense cloud anyway (core/model.py:157).  Same
 * results as dh3d_three_nn / dh3d_three_nn_ws; the workspace of THIS call as for dh3d_three_nn_ws. */
DH3D_API int dh3d_three_nn_ws_presorted(int b, int n, int m, const void* knn_workspace_of_xyz1, const float* xyz2,
                               float* dist, int32_t* idx, void* workspace, size_t workspace_bytes, void* stream);
/* ... and with BOTH clouds sorted (the k-NN workspaces of xyz1 [b,n,3] and of xyz2 [b,m,3]: DH3D also runs the k-NN of
 * the sampled points before the interpolation, core/backbones.py:64-70): one launch, no workspace. */
DH3D_API int dh3d_three_nn_presorted2(int b, int n, int m, const void* knn_workspace_of_xyz1,
                             const void* knn_workspace_of_xyz2, float* dist, int32_t* idx, void* stream);
DH3D_API int dh3d_three_interpolate(int b, int m, int c, int n, const float* points, const int32_t* idx,
                           const float* weight, float* out, void* stream);
/* fused caller-side step of core/backbones.py:91-96: weight = (1/max(d,1e-10))/sum(...) computed
 * in-kernel from the squared distances of dh3d_three_nn, then interpolated. */
DH3D_API int dh3d_three_interpolate_from_dist(int b, int m, int c, int n, const float* points,
                                     const int32_t* idx, const float* dist2, float* out,
                                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense 1x1 layers (tensorpack Conv2D kernel_shape=1 + BatchNorm + activation on [B,N,1,C];
 * call sites core/tf_utils.py:99-109, core/backbones.py:132-173)
 *   y[m, :] = act( (x[m, :] @ W[K,N]) * scale[:] + shift[:] )      x [M,K] row-major, ldx >= K
 *   scale NULL -> 1, shift NULL -> 0 (conv bias and folded BN are both carried by scale/shift).
 *   K % 4 == 0, N % 4 == 0, ldx % 4 == 0, ldy % 4 == 0.
 * dh3d_rowdot: y[m] = act( sum_k x[m,k]*w[k] + bias )  -- the final 1024->1 / sigmoid layers.
 * dh3d_se_excite: y = relu(x + x*gate)  (core/backbones.py:45-55, se_res_bottleneck).
 * dh3d_l2_normalize_rows: y = x / sqrt(max(sum x^2, eps))  (tf.nn.l2_normalize, model.py:177,205)
 * dh3d_add: y = a + b   (residual join, backbones.py:123)
 * ------------------------------------------------------------------------------------------- */
DH3D_API int dh3d_linear(const float* x, int ldx, const float* w, const float* scale, const float* shift,
                int act, float* y, int ldy, int M, int K, int N, void* stream);
/* Tensor-core path of dh3d_linear: tcgen05.mma kind::f16 on 2-term fp16 splits (22 mantissa bits, 3 MMAs per
 * product; ~3e-5 * rms against fp64, inside the 1e-4 bar).  The weight is pre-split ONCE into packed =
 * {W_hi^T, W_lo^T, per-column scale} ([N,K] K-major, what the UMMA descriptors want); activations are split on the
 * fly in shared memory with a fixed 2^4 scale, and rows whose largest |x| leaves [2^-11, 3750] (or hold inf / NaN)
 * are recomputed in fp32 by the same launch: any finite activation magnitude is accurate (csrc/gemm_tc16.cu). */
DH3D_API size_t dh3d_linear_prepack_bytes(int K, int N);
DH3D_API int dh3d_linear_prepack(const float* w, int K, int N, void* packed, void* stream);
DH3D_API int dh3d_linear_packed(const float* x, int ldx, const void* packed_w, const float* scale,
                       const float* shift, int act, float* y, int ldy, int M, int K, int N,
                       void* stream);
/* Fused two-layer head: y[m] = act2( sum_n act((x @ W)[m,n]*scale[n] + shift[n]) * w2[n] + b2 ).
 * The [M,N] hidden activation (the detector's / global attention's [B*N,1024], core/backbones.py:
 * 141-147,168-172) stays in TMEM/registers and never reaches HBM. */
DH3D_API int dh3d_linear_rowdot_packed(const float* x, int ldx, const void* packed_w, const float* scale,
                              const float* shift, int act, const float* w2, float b2, int act2,
                              float* y, int M, int K, int N, void* stream);
/* Two-branch join of backbone_local_dilate (core/backbones.py:121-123) with the descriptor normalisation of
 * core/model.py:177-181 in one launch:
 *     y  = act_a((xa @ Wa)*scale_a + shift_a) + act_b((xb @ Wb)*scale_b + shift_b)          [M,N]
 *     yn = y / sqrt(max(sum_n y^2, eps))      (y_normalized may be NULL)
 * packed_wa / packed_wb come from dh3d_linear_prepack.  N must be 128 (the row norm needs the whole output row
 * in one tile); otherwise DH3D_ERR_UNSUPPORTED and the caller composes dh3d_linear_packed x2 +
 * dh3d_add_l2_normalize_rows. */
DH3D_API int dh3d_linear_join_packed(const float* xa, int ldxa, const void* packed_wa, const float* scale_a,
                            const float* shift_a, int act_a, const float* xb, int ldxb,
                            const void* packed_wb, const float* scale_b, const float* shift_b, int act_b,
                            float* y, int ldy, float* y_normalized, int ldn, float eps, int M, int Ka,
                            int Kb, int N, void* stream);
/* Two chained 1x1 layers in one launch (detection_block's 128 -> 128 -> 256 stack, core/backbones.py:132-147):
 *     y = act2( (act1((x @ W1)*scale1 + shift1) @ W2)*scale2 + shift2 )        x [M,K1] -> y [M,N2]
 * the hidden [M,N1] activation stays on the SM.  packed_w1 / packed_w2 = dh3d_linear_prepack of W1 [K1,N1] /
 * W2 [N1,N2].  Built for K1 <= 128, N1 == 128, N2 <= 256 (multiples of 4); other shapes return
 * DH3D_ERR_UNSUPPORTED and the caller composes dh3d_linear_packed x2.  Same accuracy and out-of-window-row
 * guarantee as dh3d_linear_packed (rows whose input OR hidden activations leave the fp16-pair window are
 * recomputed through both layers in fp32). */
DH3D_API int dh3d_linear_chain_packed(const float* x, int ldx, const void* packed_w1, const float* scale1,
                             const float* shift1, int act1, const void* packed_w2, const float* scale2,
                             const float* shift2, int act2, float* y, int ldy, int M, int K1, int N1, int N2,
                             void* stream);
DH3D_API int dh3d_rowdot(const float* x, int ldx, const float* w, float bias, int act, float* y, int M,
                int K, void* stream);
DH3D_API int dh3d_se_excite(const float* x, const float* gate, float* y, size_t count, void* stream);
/* Whole se_res_bottleneck of flex_conv_dilate in one launch (core/backbones.py:45-55 called at :84-88 with
 * add_se='max_pool'): y = relu(x + x * sigmoid(W2^T relu(W1^T flex_pool(x, nbr) + b1) + b2)).
 *   x, y [B,N,C] point-major; neighborhood [B,N,K] int32 (ids within the cloud); w1 [C,H], b1 [H], w2 [H,C],
 *   b2 [C] (the two feature_conv1d_1 layers, no BN).  C in {64,128} and H == C/4 (DH3D's shapes), otherwise
 *   DH3D_ERR_UNSUPPORTED and the caller composes dh3d_flex_pool_pm / dh3d_linear / dh3d_se_excite. */
DH3D_API int dh3d_se_pool_excite(const float* x, const int32_t* neighborhood, const float* w1, const float* b1,
                        const float* w2, const float* b2, float* y, int B, int N, int K, int C, int H,
                        void* stream);
DH3D_API int dh3d_l2_normalize_rows(const float* x, int ldx, float* y, int ldy, int M, int C, float eps,
                           void* stream);
DH3D_API int dh3d_add(const float* a, const float* b, float* y, size_t count, void* stream);
/* strided 2-D copy: dst[m, 0:C] = src[m, 0:C] (concat / slice helper; C % 4 == 0, lds % 4 == 0) */
DH3D_API int dh3d_copy_cols(const float* src, int lds, float* dst, int ldd, int M, int C, void* stream);
/* [B,C,N] <-> [B,N,C] for fp32 / int32 payloads (bit copies) */
DH3D_API int dh3d_transpose_cm_to_pm(const void* src_cm, void* dst_pm, int B, int C, int N, void* stream);
DH3D_API int dh3d_transpose_pm_to_cm(const void* src_pm, void* dst_cm, int B, int N, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention-weighted NetVLAD -- replaces the TF-library graph of
 *   core/backbones.py:202-279 (global_netvald_block) + :282-320 (context_gating)
 *   features [B,N,D] pm, att [B,N] -> out [B,out_dim] (pre final l2-normalise; model.py:205 is
 *   applied when `final_l2norm` != 0, eps 1e-8).
 *   cluster_weights [D,Kc], cluster_weights2 [D,Kc] (the reference's [1,D,Kc]),
 *   hidden1_weights [D*Kc, out_dim] (row index d*Kc + k: feature-major, cluster-minor, :258-260),
 *   gating_weights [out_dim,out_dim]; the three BatchNorms arrive folded to scale/shift:
 *   cluster_bn (slim, eps 1e-3) [Kc], bn (contrib, eps 1e-3) [out_dim], gating_bn (slim) [out_dim].
 *   D == 256, Kc == 64, out_dim == 256 (the shipped configuration).
 * ------------------------------------------------------------------------------------------- */
DH3D_API size_t dh3d_netvlad_workspace_bytes(int B, int N, int D, int Kc, int out_dim);
DH3D_API int dh3d_netvlad(const float* features, const float* att, int B, int N, int D, int Kc, int out_dim,
                 const float* cluster_weights, const float* cluster_bn_scale,
                 const float* cluster_bn_shift, const float* cluster_weights2,
                 const float* hidden1_weights, const float* bn_scale, const float* bn_shift,
                 const float* gating_weights, const float* gating_bn_scale,
                 const float* gating_bn_shift, int final_l2norm, float* out, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Retrieval after the descriptor all-gather -- the step downstream of the one collective
 *   reference: evaluate/global_eval/evaluation_retrieval.py:37-40 (host cKDTree, k = 25)
 *   gram [Q,R] (row stride ldg >= R) = query @ ref^T (e.g. from dh3d_linear), qn [Q] / rn [R] squared norms
 *   -> idx [Q,K] i32, val [Q,K] squared L2 distances, ascending (ties: smaller index).  K <= 32.
 * ------------------------------------------------------------------------------------------- */
DH3D_API int dh3d_topk_l2(const float* gram, int ldg, const float* qn, const float* rn, int Q, int R, int K,
                 int32_t* idx, float* val, void* stream);
/* Same, with the selected neighbours re-ranked by their EXACT distances sum_d (q_d - r_d)^2 (the Gram form loses
 * the order of near-identical descriptors to cancellation; a cKDTree computes the differences directly):
 *   query [Q,D], ref [R,D] row-major descriptors; cand [Q,32] i32 and cand_val [Q,32] f32 are caller scratch.
 *   The Gram pass selects min(K + 8, 32, R) candidates.  K <= 24. */
DH3D_API int dh3d_topk_l2_exact(const float* gram, int ldg, const float* qn, const float* rn, const float* query,
                       const float* ref, int Q, int R, int D, int K, int32_t* idx, float* val, int32_t* cand,
                       float* cand_val, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward passes and the transposed FlexConv (training side of boundary A/B)
 *
 * dh3d_flex_conv_grad -- replaces op FlexConvGrad
 *   reference: user_ops/ops/flex_conv.cc:102-150, kernels/flex_conv_op.cc:57-100,
 *   kernels/flex_conv_kernel.cc:75-163 (CPU), flex_conv_kernel_gpu.cu.cc:160-400,443-520 (GPU);
 *   python user_ops/__init__.py:95-111.  Inputs as dh3d_flex_conv plus topdiff [B,Dout,N];
 *   outputs grad_features [B,Din,N], grad_theta [3,Din,Dout], grad_bias [Din,Dout].  The offset is
 *   taken from p[nbr(0,n)] (both reference backward kernels, :134 / :196-202), not from p[n].
 *   Parameter gradients are reduced in a fixed order (deterministic); grad_features uses fp32
 *   atomics like the reference (CudaAtomicAdd, :372).  The _pm entry takes the native layouts
 *   (features [B,N,Din], neighborhood [B,N,K], xyz [B,N,3], topdiff [B,N,Dout]; Din, Dout % 4 == 0).
 *
 * dh3d_flex_pool_grad -- replaces op FlexPoolGrad (flex_pool_kernel.cc:63-95,
 *   flex_pool_kernel_gpu.cu.cc:65-98; python :141-151): topdiff, argmax [B,D,N] -> grad_features
 *   [B,D,N] with grad[b,d,argmax[b,d,n]] += topdiff[b,d,n].
 *
 * dh3d_conv_pointset_grad -- replaces op ConvPointsetGrad (conv_pointset_kernel.cc:72-147; python
 *   :231-246): features [B,Din,N], theta [Din,Dout], neighborhood [B,K,N], topdiff [B,Dout,N] ->
 *   grad_features [B,Din,N], grad_theta [Din,Dout], grad_bias [Dout].
 *
 * dh3d_flex_deconv -- replaces op FlexDeconv forward (flex_deconv_kernel.cc:25-70; python
 *   flex_convolution_transpose :155-179): out[b,:,nbr(k,n)] += W(p[nbr(k,n)] - p[nbr(0,n)]) . f[b,:,nbr(0,n)].
 *
 * dh3d_group_point_grad / dh3d_gather_point_grad / dh3d_three_interpolate_grad -- replace
 *   group_point_grad_gpu (tf_ops/grouping/tf_grouping_g.cu:114-133, launcher :195-198),
 *   scatteraddpointKernel (tf_ops/sampling/tf_sampling_g.cu:183-192, launcher :209-211) and
 *   threeinterpolate_grad_cpu (tf_ops/interpolation/tf_interpolate.cpp:131-153): same argument
 *   order as the reference launchers; the output is zeroed here (the reference ops memset it in
 *   their OpKernel, e.g. tf_grouping.cpp:188).
 * ------------------------------------------------------------------------------------------- */
DH3D_API size_t dh3d_flex_conv_grad_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_flex_conv_grad(const float* features_cm, const float* theta, const float* bias,
                        const int32_t* neighborhood_cm, const float* positions_cm, const float* topdiff_cm,
                        float* grad_features_cm, float* grad_theta, float* grad_bias, int B, int N, int K,
                        int Din, int Dout, void* workspace, size_t workspace_bytes, void* stream);
DH3D_API size_t dh3d_flex_conv_grad_pm_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_flex_conv_grad_pm(const float* features_pm, const float* theta, const float* bias,
                           const int32_t* neighborhood_pm, const float* xyz_pm, const float* topdiff_pm,
                           float* grad_features_pm, float* grad_theta, float* grad_bias, int B, int N, int K,
                           int Din, int Dout, void* workspace, size_t workspace_bytes, void* stream);
DH3D_API int dh3d_flex_pool_grad(const float* topdiff_cm, const int32_t* argmax_cm, float* grad_features_cm,
                        int B, int N, int D, void* stream);
DH3D_API size_t dh3d_conv_pointset_grad_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_conv_pointset_grad(const float* features_cm, const float* theta, const int32_t* neighborhood_cm,
                            const float* topdiff_cm, float* grad_features_cm, float* grad_theta,
                            float* grad_bias, int B, int N, int K, int Din, int Dout, void* workspace,
                            size_t workspace_bytes, void* stream);
DH3D_API size_t dh3d_flex_deconv_workspace_bytes(int B, int N, int K, int Din, int Dout);
DH3D_API int dh3d_flex_deconv(const float* features_cm, const float* theta, const float* bias,
                     const int32_t* neighborhood_cm, const float* positions_cm, float* out_cm, int B, int N,
                     int K, int Din, int Dout, void* workspace, size_t workspace_bytes, void* stream);
DH3D_API int dh3d_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out,
                          const int32_t* idx, float* grad_points, void* stream);
DH3D_API int dh3d_gather_point_grad(int b, int n, int m, const float* out_g, const int32_t* idx, float* inp_g,
                           void* stream);
DH3D_API int dh3d_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int32_t* idx,
                                const float* weight, float* grad_points, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Keypoint NMS on the detector attention -- replaces core/utils.py:15-43 `single_nms` (host numpy +
 *   sklearn ball tree; caller evaluate/local_eval/localdesc_extract.py:92-98), batched over clouds.
 *   xyz [B,N,3] pm, attention [B,N] -> out_idx [B,max_keypoints] i32 (descending (attention, index)
 *   order, padded with -1), out_cnt [B] i32 = min(#local maxima above the response threshold, max_keypoints).
 *   50-NN lists from the k-NN engine; radius / noise thresholds are evaluated on fp64 pair distances
 *   (sklearn's arithmetic).  remove_noise != 0 zeroes the attention of points whose 8th neighbour is
 *   farther than 2.0 (:19-22).  N >= 50.  `attention` is not modified (the reference mutates it).
 * ------------------------------------------------------------------------------------------- */
/* The two array expressions around single_nms in the reference's --perform_nms output mode
 * (evaluate/local_eval/localdesc_extract.py:92-104), so that only keypoint rows leave the GPU:
 *   dh3d_affine:      y = a * x + b                        (`attention = 1 - res[:, -1]`, :95)
 *   dh3d_gather_rows: out[b, j, 0:c] = src[b, idx[b,j], 0:c], zero row where idx[b,j] < 0
 *                     (`res[max_indices, :]`, :99; out rows are ldo >= c floats apart: a column block) */
DH3D_API int dh3d_affine(const float* x, float a, float b, float* y, size_t count, void* stream);
DH3D_API int dh3d_gather_rows(int b, int n, int c, int m, const float* src, const int32_t* idx, float* out, int ldo,
                     void* stream);
DH3D_API size_t dh3d_keypoint_nms_workspace_bytes(int B, int N);
DH3D_API int dh3d_keypoint_nms(const float* xyz_pm, const float* attention, int B, int N, float nms_radius,
                      float min_response_ratio, int max_keypoints, int remove_noise, int32_t* out_idx,
                      int32_t* out_cnt, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Strided / fused variants used by the forward pass to avoid materialising the channel concat of
 * flex_conv_dilate (core/backbones.py:96-99) and the separate add + l2-normalise (backbone_local_dilate
 * :127 + model.py:177-181).  Results are bit-identical to the unfused sequence.
 *   dh3d_group_point_ld: `points` rows have stride ldp >= c floats (a column block of a wider tensor).
 *   dh3d_three_interpolate_ld: `out` rows have stride ldo >= c floats; weight_is_dist2 != 0 derives the
 *     inverse-distance weights from the squared distances in-kernel (backbones.py:92-95).
 *   dh3d_add_l2_normalize_rows: sum = a + b and normalized = sum / sqrt(max(|sum|^2, eps)), [M,C] each.
 * ------------------------------------------------------------------------------------------- */
DH3D_API int dh3d_group_point_ld(int b, int n, int c, int m, int nsample, const float* points, int ldp,
                        const int32_t* idx, float* out, void* stream);
DH3D_API int dh3d_three_interpolate_ld(int b, int m, int c, int n, const float* points, const int32_t* idx,
                              const float* weight_or_dist2, int weight_is_dist2, float* out, int ldo,
                              void* stream);
DH3D_API int dh3d_add_l2_normalize_rows(const float* a, const float* b, float* sum, float* normalized, int M, int C,
                               float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DH3D_B200_H_ */
