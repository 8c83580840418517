"""Network blocks of the DH3D forward pass, point-major, inference only.

Mirrors ``core/backbones.py`` + the glue of ``core/tf_utils.py`` of the reference; module and
parameter names follow the reference's TF variable scopes with '/' -> '.':

    subsample                 core/tf_utils.py:86-96      FPS + group_point
    feature_conv1d_1          core/tf_utils.py:99-109     -> FeatureConv1d ('<name>/tfconv0/{W,b,bn}')
    flexconv_withBatchnorm    core/tf_utils.py:48-64      -> FlexConvolution + '<name>_bn', fused ReLU
    se_res_bottleneck         core/backbones.py:45-55     -> SEBlock
    flex_conv_dilate          core/backbones.py:58-101    -> FlexConvDilate
    backbone_local_dilate     core/backbones.py:104-127   -> LocalBackbone
    detection_block           core/backbones.py:132-151   -> DetectionBlock
    globalatt_block           core/backbones.py:156-173   -> GlobalAttBlock
    global_netvald_block      core/backbones.py:202-279   -> GlobalNetVLADBlock (+ context_gating :282-320)

One structural difference from the reference graph: FPS, the k-NN of the sampled points and the
3-NN of the dense points depend only on xyz, and ``stage2`` and ``global_before_assemble`` both
use dilate 8 on the same xyz, so they are computed ONCE per cloud (``DilateGeometry``) and shared
(the reference recomputes them).
"""
import torch
from torch import nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID
from ._lib import Dh3dError
from .layers import (SLIM_BN_EPS, BatchNorm, Conv1x1, ConvolutionPointset, Flex_Avg, FlexConvolution,
                     FlexPooling, FoldedModule, default_store)

# The fused block kernels (dh3d_se_pool_excite, dh3d_linear_join_packed) cover DH3D's shapes; other shapes compose the
# per-op kernels.  Tests flip this module attribute to run the composed form on the SAME shapes (no env switch).
USE_FUSED_BLOCKS = True


class DilateGeometry(object):
    """Everything ``flex_conv_dilate(dilate > 1)`` derives from xyz alone."""

    ready = None   # event recorded on the stream that computed the geometry (DH3D.forward's side stream)

    def join(self):
        """Makes the current stream wait for the geometry (once); a no-op when it was computed on this stream."""
        if self.ready is not None:
            torch.cuda.current_stream(self.kp_indices.device).wait_event(self.ready)
            self.ready = None

    def __init__(self, xyz, npoint, knn, sorted_xyz=None, sorted_ready=None):
        """sorted_xyz: the k-NN engine's cell-sorted copy of ``xyz`` (``ops.knn_sort`` / ``ops.knn_points(...,
        keep_workspace=True)``): FPS then runs its box-pruned kernel on it and the 3-NN skips its own sort of the dense
        cloud.  sorted_ready: the event after which that copy may be read, if it was written on another stream."""
        if sorted_ready is not None:
            torch.cuda.current_stream(xyz.device).wait_event(sorted_ready)
        self.kp_indices = ops.farthest_point_sample(npoint, xyz, sorted_ws=sorted_xyz)          # [B,M]
        self.points_sampled = ops.gather_point(xyz, self.kp_indices)          # [B,M,3]
        self.knn_indices, _, sorted_s = ops.knn_points(self.points_sampled, knn, keep_workspace=True)    # [B,M,K]
        self.nn_dist, self.nn_idx = ops.three_nn(xyz, self.points_sampled, sorted1=sorted_xyz,
                                                 sorted2=sorted_s if sorted_xyz is not None else None)   # [B,N,3] x2


def subsample(points, feat, targetnum, kp_idx=None):
    """core/tf_utils.py:86-96 -> (xyz_sampled [B,M,3], feat_sampled [B,M,C], kp_indices [B,M,1])."""
    if kp_idx is None:
        kp_idx = ops.farthest_point_sample(targetnum, points).unsqueeze(2)
    kp_idx = kp_idx.contiguous()
    feat_sampled = ops.group_point(feat, kp_idx).squeeze(2)
    xyz_sampled = ops.group_point(points, kp_idx).squeeze(2)
    return xyz_sampled, feat_sampled, kp_idx


class FeatureConv1d(nn.Module):
    """feature_conv1d_1(feat, dim, name, ac_func): variables under '<name>/tfconv0'."""

    def __init__(self, cin, cout, bn=True, act=ACT_RELU):
        super().__init__()
        self.tfconv0 = Conv1x1(cin, cout, bn=bn, act=act)

    def forward(self, x, out=None, out_col=0):
        return self.tfconv0(x, out=out, out_col=out_col)


class SEBlock(nn.Module):
    """se_res_bottleneck: per-point squeeze/excite computed from the POOLED features (no global
    pooling): gate = sigmoid(W2 relu(W1 pool + b1) + b2); out = relu(x + x*gate)."""

    def __init__(self, ch):
        super().__init__()
        self.f1 = FeatureConv1d(ch, ch // 4, bn=False, act=ACT_RELU)
        self.f2 = FeatureConv1d(ch // 4, ch, bn=False, act=ACT_SIGMOID)

    def forward(self, x, pooled):
        return ops.se_excite(x, self.f2(self.f1(pooled)))

    def forward_fused(self, x, nbr):
        """flex_pool + both 1x1 layers + excite in one launch (``dh3d_se_pool_excite``); None when the
        shape is not one the fused kernel covers (the caller then composes the separate ops)."""
        C = x.shape[2]
        if C not in (64, 128) or not USE_FUSED_BLOCKS:
            return None
        w1, _, b1, _ = self.f1.tfconv0.folded()
        w2, _, b2, _ = self.f2.tfconv0.folded()
        return ops.se_pool_excite(x, nbr, w1, b1, w2, b2)


class FlexConvDilate(nn.Module):
    def __init__(self, cin, outdims, dilate, knn=8, concat=True, add_se="max_pool", upsample=True):
        super().__init__()
        if add_se not in ("max_pool", "avg_pool", "", None):
            raise Dh3dError("flex_conv_dilate: add_se must be 'max_pool', 'avg_pool' or '' (got %r)" % (add_se,))
        self.dilate, self.knn, self.upsample, self.add_se = dilate, knn, upsample, add_se or ""
        self.outdims = list(outdims)
        c = cin
        for i, d in enumerate(outdims):
            setattr(self, "flexconv_%d" % i, FlexConvolution(c, d))
            setattr(self, "flexconv_%d_bn" % i, BatchNorm(d))
            c = d
        self.se = SEBlock(c) if self.add_se else None
        # add_se='avg_pool' (core/backbones.py:79-82): x_pool = flex_avg(x, ...) * (1/knn); 'se_avgpool' owns a
        # zero, non-trainable position_theta variable
        self.se_avgpool = Flex_Avg(c) if self.add_se == "avg_pool" else None
        self.concat_conv1d = FeatureConv1d(c + cin, c) if concat else None

    def forward(self, xyz, feat, knn_indices=None, geometry=None, cat=None, defer_concat=False):
        """xyz [B,N,3], feat [B,N,C] -> new_feat [B,N,outdims[-1]].
        ``defer_concat`` (with ``cat``): stop before ``concat_conv1d`` and return the filled concat buffer, so the
        caller can fuse that 1x1 layer with what follows it (LocalBackbone's join).

        ``cat`` (dilate > 1 with concat only): a [B,N,outdims[-1]+C] buffer whose LAST C columns already hold
        ``feat`` (written there by the producing 1x1 layer); the up-sampled features are interpolated straight
        into its first columns, so the channel concat of core/backbones.py:96-99 is never copied."""
        if self.dilate > 1:
            g = geometry or DilateGeometry(xyz, xyz.shape[1] // self.dilate, self.knn)
            g.join()
            pts = g.points_sampled
            if cat is not None:
                cin = cat.shape[2] - self.outdims[-1]
                x = ops.group_point_cols(cat, self.outdims[-1], cin, g.kp_indices.unsqueeze(2)).squeeze(2)
            else:
                x = ops.group_point(feat, g.kp_indices.unsqueeze(2)).squeeze(2)
            nbr = g.knn_indices
        else:
            g, pts, x = None, xyz, feat
            nbr = knn_indices if knn_indices is not None else ops.knn_points(xyz, self.knn)[0]
        for i in range(len(self.outdims)):
            x = getattr(self, "flexconv_%d" % i).forward_pm(
                x, pts, nbr, bn=getattr(self, "flexconv_%d_bn" % i), act=ACT_RELU)
        if self.se_avgpool is not None:
            # the 1/knn scaling rides in the FlexConv epilogue (scale = 1/knn, no shift)
            inv_k = torch.full((x.shape[2],), 1.0 / self.knn, dtype=x.dtype, device=x.device)
            pooled = ops.flex_conv(x, self.se_avgpool.position_theta, self.se_avgpool.position_bias, nbr, pts,
                                   scale=inv_k, shift=torch.zeros_like(inv_k))
            x = self.se(x, pooled)
        elif self.se is not None:
            y = self.se.forward_fused(x, nbr)
            x = y if y is not None else self.se(x, ops.flex_pool(x, nbr))
        if self.upsample and self.dilate > 1:
            if cat is not None and self.concat_conv1d is not None:
                ops.three_interpolate(x, g.nn_idx, g.nn_dist, weight_is_dist2=True, out=cat, out_col=0)
                return cat if defer_concat else self.concat_conv1d(cat)
            x = ops.three_interpolate(x, g.nn_idx, g.nn_dist, weight_is_dist2=True)
        if self.concat_conv1d is not None:
            B, N, C = x.shape
            cat = torch.empty((B, N, C + feat.shape[2]), dtype=x.dtype, device=x.device)
            ops.copy_cols(x, cat, 0)
            ops.copy_cols(feat, cat, C)
            x = self.concat_conv1d(cat)
        return x


class LocalBackbone(nn.Module):
    """backbone_local_dilate (core/backbones.py:104-127); featdim < 128 appends 'final_fc' (:125-126)."""

    def __init__(self, init_feat_dim=32, featdim=128, dilate2=8, knn=8, add_se="max_pool"):
        super().__init__()
        if not 0 < featdim <= 128:
            raise Dh3dError("backbone_local_dilate: featdim must be in 1..128 (got %d)" % featdim)
        self.knn, self.featdim = knn, featdim
        self.initconv = ConvolutionPointset(3, init_feat_dim)
        self.initconv_bn = BatchNorm(init_feat_dim)
        self.stage1 = FlexConvDilate(init_feat_dim, [64, 64], dilate=1, knn=knn, concat=False, add_se=add_se)
        self.before_stage2_conv1d = FeatureConv1d(64, 64)
        self.stage2 = FlexConvDilate(64, [128, 128], dilate=dilate2, knn=knn, concat=True, add_se=add_se)
        self.local_stage1_shortcut = FeatureConv1d(64, 128)
        self.final_fc = FeatureConv1d(128, featdim) if featdim < 128 else None

    def forward(self, points, knn_ind, geometry=None, with_desc=False, desc_out=None):
        """-> feat [B,N,128]; with_desc=True: (feat, l2-normalised feat) from one fused pass (model.py:177-181).
        ``desc_out``: caller-owned [B,N,featdim] tensor the normalised descriptors are written into."""
        nn_8 = knn_ind if knn_ind.shape[2] == 8 else knn_ind[:, :, :8].contiguous()
        f = self.initconv.forward_pm(points, nn_8, bn=self.initconv_bn, act=ACT_RELU)
        f = ops.flex_pool(f, nn_8)
        x1 = self.stage1(points, f, knn_indices=nn_8)
        B, N, _ = x1.shape
        # stage2's concat buffer: [up-sampled 128 | before_stage2 64]; the 1x1 layer writes its block in place
        cat = torch.empty((B, N, self.stage2.outdims[-1] + 64), dtype=x1.dtype, device=x1.device)
        self.before_stage2_conv1d(x1, out=cat, out_col=self.stage2.outdims[-1])
        # stage-2 concat layer + stage-1 shortcut + residual add (+ descriptor l2-norm): one launch
        # (dh3d_linear_join_packed) when both weights are in the fp16-pair layout, else the separate ops
        la, lb = self.stage2.concat_conv1d.tfconv0, self.local_stage1_shortcut.tfconv0
        (_, sa, ba, pa), (_, sb, bb, pb) = la.folded(), lb.folded()
        fused = pa is not None and pb is not None and la.W.shape[3] == 128 and USE_FUSED_BLOCKS
        want_desc = with_desc and self.final_fc is None
        if fused:
            cat = self.stage2(points, None, geometry=geometry, cat=cat, defer_concat=True)
            y = ops.linear_join(cat, pa, sa, ba, la.act, x1, pb, sb, bb, lb.act, eps=1e-8 if want_desc else None,
                                out_norm=desc_out if want_desc else None)
        else:
            x2 = self.stage2(points, None, geometry=geometry, cat=cat)
            sc = self.local_stage1_shortcut(x1)
            y = ops.add_l2_normalize_rows(sc, x2, 1e-8) if want_desc else ops.add(sc, x2)
            if want_desc and desc_out is not None:
                ops.copy_cols(y[1], desc_out, 0)
                y = (y[0], desc_out)
        if self.final_fc is None:
            return y
        y = self.final_fc(y)   # featdim < 128 (core/backbones.py:125-126): Conv2D + BN + ReLU, then the l2 norm
        return (y, ops.l2_normalize_rows(y, 1e-8, out=desc_out)) if with_desc else y


class AttentionHead(FoldedModule):
    """1x1 stack -> 1 logit -> sigmoid.  detection_block: 128->128->256->1024->1;
    globalatt_block: 256->1024->1."""

    def __init__(self, cin, conv_dims):
        super().__init__()
        c = cin
        self.n = len(conv_dims)
        for i, d in enumerate(conv_dims):
            setattr(self, "detec_conv%d" % i, Conv1x1(c, d, bn=True, act=ACT_RELU))
            c = d
        self.detec_conv_fc = Conv1x1(c, 1, bn=False, act=ACT_NONE)
        self._folded = None

    def forward(self, x, out=None):
        """``out``: caller-owned [B,N,1] tensor for the attention."""
        if self._folded is None:
            fc = self.detec_conv_fc
            self._folded = (fc.W.reshape(-1).contiguous(), float(fc.b.reshape(-1)[0].item()))
        w2, b2 = self._folded
        i0 = 0
        if self.n == 3 and USE_FUSED_BLOCKS:
            # detection_block's 128 -> 128 -> 256 layers in one launch (dh3d_linear_chain_packed): the hidden
            # [B*N,128] activation stays on the SM
            c0, c1 = self.detec_conv0, self.detec_conv1
            (_, s0, b0, p0), (_, s1, b1, p1) = c0.folded(), c1.folded()
            if (p0 is not None and p1 is not None and
                    ops.linear_chain_supported(c0.W.shape[2], c0.W.shape[3], c1.W.shape[3])):
                x = ops.linear_chain(x, p0, s0, b0, c0.act, p1, s1, b1, c1.act)
                i0 = 2
        for i in range(i0, self.n - 1):
            x = getattr(self, "detec_conv%d" % i)(x)
        last = getattr(self, "detec_conv%d" % (self.n - 1))
        w, scale, shift, packed = last.folded()
        if packed is not None:  # fused: the [B,N,1024] hidden layer never reaches HBM
            y = ops.linear_rowdot(x, packed, scale, shift, last.act, w2, b2, ACT_SIGMOID,
                                  out=None if out is None else out.view(out.shape[:-1]))
            return y.unsqueeze(-1)
        y = ops.rowdot(last(x), w2, bias=b2, act=ACT_SIGMOID).unsqueeze(-1)  # [B,N,1]
        if out is not None:
            ops.copy_cols(y, out, 0)
            return out
        return y


class DetectionBlock(AttentionHead):
    def __init__(self, cin=128, conv_dims=(128, 256, 1024)):
        super().__init__(cin, conv_dims)


class GlobalAttBlock(AttentionHead):
    def __init__(self, cin=256):
        super().__init__(cin, (256, 1024) if cin > 256 else (1024,))


class GlobalNetVLADBlock(FoldedModule):
    """global_netvald_block(xyz, features, att, is_training, cluster_size=64, output_dim=256,
    add_batch_norm=True, gating=True) -- variables live at the root scope in the checkpoint."""

    def __init__(self, feature_size=256, cluster_size=64, output_dim=256):
        super().__init__()
        z = lambda *s: nn.Parameter(torch.zeros(*s), requires_grad=False)
        self.cluster_weights = z(feature_size, cluster_size)
        self.cluster_weights2 = z(1, feature_size, cluster_size)
        self.cluster_bn = BatchNorm(cluster_size, eps=SLIM_BN_EPS)
        self.hidden1_weights = z(cluster_size * feature_size, output_dim)
        self.bn = BatchNorm(output_dim, eps=SLIM_BN_EPS)
        self.gating_weights = z(output_dim, output_dim)
        self.gating_bn = BatchNorm(output_dim, eps=SLIM_BN_EPS)
        self._folded = None

    def forward(self, xyz, features, att, final_l2norm=True, out=None):
        if self._folded is None:
            self._folded = (self.cluster_bn.fold(), self.bn.fold(), self.gating_bn.fold(),
                            self.cluster_weights2.reshape(self.cluster_weights.shape).contiguous())
        cbn, bn, gbn, cw2 = self._folded
        return ops.netvlad(features, att, self.cluster_weights, cbn, cw2, self.hidden1_weights, bn,
                           self.gating_weights, gbn, final_l2norm=final_l2norm, out=out)


def global_netvald_block(xyz, features, att, is_training=False, cluster_size=64, output_dim=256, add_batch_norm=True,
                         gating=True, store=None, scope="netvlad", **unused_kwargs):
    """core/backbones.py:202-279, same argument order: xyz [B,N,3] (unused by the reference too), features
    [B,N,D], att [B,N,1] -> 'final_global' [B,output_dim] (NOT l2-normalised; core/model.py:205 does that).
    The variables (cluster_weights, cluster_weights2, cluster_bn, hidden1_weights, bn, gating_weights, gating_bn)
    live in ``store[scope]``.  Inference only; the shipped add_batch_norm=True / gating=True form only."""
    if is_training:
        raise Dh3dError("global_netvald_block: inference only (is_training must be False)")
    if not add_batch_norm or not gating:
        raise Dh3dError("global_netvald_block: only add_batch_norm=True, gating=True (the shipped configuration) is built")
    D = features.shape[2]
    block = (store if store is not None else default_store).layer(
        scope, lambda: GlobalNetVLADBlock(D, cluster_size, output_dim).to(features.device))
    return block(xyz, features, att, final_l2norm=False)
