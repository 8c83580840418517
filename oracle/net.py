"""numpy restatement of the TF-library part of the DH3D forward (TEST INFRASTRUCTURE, see
oracle/__init__.py) and the composition of the full forward from the oracle ops.

Follows the reference graph literally -- channel-major Flex ops with explicit transposes,
FPS / kNN / three_nn recomputed per block exactly where ``core/backbones.py`` calls them:

    feature_conv1d_1 / Conv2D+BN+act       core/tf_utils.py:99-109, tensorpack Conv2D (bias AND BN)
    flexconv_withBatchnorm                 core/tf_utils.py:48-64
    convolution_pointset_withBatchnorm     core/tf_utils.py:67-83
    se_res_bottleneck                      core/backbones.py:45-55
    flex_conv_dilate                       core/backbones.py:58-101
    backbone_local_dilate                  core/backbones.py:104-127
    detection_block / globalatt_block      core/backbones.py:132-173
    global_netvald_block + context_gating  core/backbones.py:202-320
    DH3D.build_graph inference branch      core/model.py:135-206

Dense math runs in float64 ("truth"); the custom ops run through the C oracle in fp32 exactly as
the reference kernels would.  Library constants that no reference test pins (PARITY UNPINNED,
SURVEY 8c): tensorpack BatchNorm eps 1e-5, slim/contrib batch_norm eps 1e-3, tf.nn.l2_normalize
eps 1e-12 unless passed.

``params`` is a flat dict {dotted-name: ndarray} with the names of ``dh3d_b200.model.DH3D``.
"""
import numpy as np

import oracle

TP_EPS, SLIM_EPS = 1e-5, 1e-3

# reference_cpu mode (bench.py CPU arm): evaluate the custom ops the way the reference's CPU
# functors do -- fp32 FlexConv loops (flex_conv_kernel.cc) and a full sort per query for k-NN
# (knn_bruteforce_kernel.cc:41-69) -- instead of the fp64-truth / fast-selection forms.
_MODE = {"reference_cpu": False}


def _bn(x, p, prefix, eps):
    g, b = p[prefix + ".gamma"], p[prefix + ".beta"]
    m, v = p[prefix + ".mean_ema"], p[prefix + ".variance_ema"]
    return (x - m) / np.sqrt(v + eps) * g + b


def _act(x, act):
    if act == "relu":
        return np.maximum(x, 0.0)
    if act == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    return x


def conv1x1(x, p, prefix, bn=True, act="relu"):
    """x [..., Cin] float64 -> [..., Cout]; prefix names a Conv1x1 ('<scope>.tfconv0' / 'detec_conv0')."""
    W = p[prefix + ".W"].astype(np.float64)
    W = W.reshape(W.shape[2], W.shape[3])
    y = x @ W + p[prefix + ".b"].astype(np.float64)
    if bn:
        y = _bn(y, p, prefix + ".bn", TP_EPS)
    return _act(y, act)


def l2_normalize(x, axis, eps=1e-12):
    ss = np.sum(x * x, axis=axis, keepdims=True)
    return x / np.sqrt(np.maximum(ss, eps))


def flexconv_bn_relu(feat_pm, xyz_pm, nbr_pm, p, prefix):
    """feat [B,N,Din] f64/f32, nbr [B,N,K] -> relu(BN(flexconv + feature_bias)) [B,N,Dout] f64.
    The op itself is evaluated with the fp64 literal loop (flex_conv_kernel semantics)."""
    out = oracle.flex_convolution(
        np.transpose(feat_pm, (0, 2, 1)), np.transpose(xyz_pm, (0, 2, 1)),
        np.transpose(nbr_pm, (0, 2, 1)), p[prefix + ".position_theta"], p[prefix + ".position_bias"],
        centre_is_self=True, f64=not _MODE["reference_cpu"])
    out = out.astype(np.float64) + p[prefix + ".feature_bias"].reshape(1, -1, 1)
    out = np.transpose(out, (0, 2, 1))
    return np.maximum(_bn(out, p, prefix + "_bn", TP_EPS), 0.0)


def flex_pool_pm(feat_pm, nbr_pm):
    out, _ = oracle.flex_pooling(np.transpose(feat_pm, (0, 2, 1)), np.transpose(nbr_pm, (0, 2, 1)))
    return np.transpose(out, (0, 2, 1)).astype(np.float64)


def se_block(x, pooled, p, prefix):
    s = conv1x1(pooled, p, prefix + ".f1.tfconv0", bn=False, act="relu")
    g = conv1x1(s, p, prefix + ".f2.tfconv0", bn=False, act="sigmoid")
    return np.maximum(x + x * g, 0.0)


def knn_pm(xyz_pm, k):
    ids, _ = oracle.knn_bruteforce(np.transpose(xyz_pm, (0, 2, 1)), k, literal=_MODE["reference_cpu"])
    return ids  # [B,N,K]


def flex_avg_pm(feat_pm, xyz_pm, nbr_pm, p, prefix):
    """Flex_Avg (core/layers.py:342-436): FlexConv with the stored (zero) theta and bias = eye -> neighbour sum."""
    C = feat_pm.shape[2]
    out = oracle.flex_convolution(np.transpose(feat_pm, (0, 2, 1)), np.transpose(xyz_pm, (0, 2, 1)),
                                  np.transpose(nbr_pm, (0, 2, 1)), p[prefix + ".position_theta"],
                                  np.eye(C, dtype=np.float32), centre_is_self=True, f64=not _MODE["reference_cpu"])
    return np.transpose(out.astype(np.float64), (0, 2, 1))


def flex_conv_dilate(xyz, feat, p, prefix, dilate, outdims, knn=8, knn_indices=None, concat=True,
                     add_se=True, upsample=True):
    xyz32 = xyz.astype(np.float32)
    N = xyz.shape[1]
    if dilate > 1:
        kp = oracle.farthest_point_sample(N // dilate, xyz32)
        pts = oracle.group_point(xyz32, kp[:, :, None])[:, :, 0, :]
        x = np.take_along_axis(feat, kp[:, :, None].astype(np.int64), axis=1)
        knn_indices = None
    else:
        pts, x = xyz32, feat
    if knn_indices is None:
        knn_indices = knn_pm(pts, knn)
    for i, _ in enumerate(outdims):
        x = flexconv_bn_relu(x, pts, knn_indices, p, "%s.flexconv_%d" % (prefix, i))
    if add_se == "avg_pool":   # core/backbones.py:79-82
        x_pool = flex_avg_pm(x.astype(np.float32), pts, knn_indices, p, prefix + ".se_avgpool") * (1.0 / knn)
        x = se_block(x, x_pool, p, prefix + ".se")
    elif add_se:
        x = se_block(x, flex_pool_pm(x.astype(np.float32), knn_indices), p, prefix + ".se")
    if upsample and dilate > 1:
        dist, idx = oracle.three_nn(xyz32, pts)
        w = oracle.three_nn_weights(dist).astype(np.float64)
        g = x[np.arange(x.shape[0])[:, None, None], idx]           # [B,N,3,C]
        x = (g * w[..., None]).sum(axis=2)
    if concat:
        x = conv1x1(np.concatenate([x, feat], axis=2), p, prefix + ".concat_conv1d.tfconv0")
    return x


def backbone_local_dilate(points, p, knn_ind, prefix="local", add_se="max_pool"):
    pts32 = points.astype(np.float32)
    nn8 = knn_ind[:, :, :8]
    f = oracle.convolution_pointset(np.transpose(pts32, (0, 2, 1)), np.transpose(nn8, (0, 2, 1)),
                                    p[prefix + ".initconv.position_theta"],
                                    p[prefix + ".initconv.position_bias"])
    f = np.transpose(f, (0, 2, 1)).astype(np.float64)
    f = np.maximum(_bn(f, p, prefix + ".initconv_bn", TP_EPS), 0.0)
    f = flex_pool_pm(f.astype(np.float32), nn8)
    x1 = flex_conv_dilate(points, f, p, prefix + ".stage1", 1, [64, 64], knn_indices=nn8, concat=False,
                          add_se=add_se)
    x2 = conv1x1(x1, p, prefix + ".before_stage2_conv1d.tfconv0")
    x2 = flex_conv_dilate(points, x2, p, prefix + ".stage2", 8, [128, 128], concat=True, add_se=add_se)
    feat = conv1x1(x1, p, prefix + ".local_stage1_shortcut.tfconv0") + x2
    if prefix + ".final_fc.tfconv0.W" in p:     # featdim < 128 (core/backbones.py:125-126)
        feat = conv1x1(feat, p, prefix + ".final_fc.tfconv0")
    return feat


def attention_head(x, p, prefix, n):
    for i in range(n):
        x = conv1x1(x, p, "%s.detec_conv%d" % (prefix, i))
    return conv1x1(x, p, prefix + ".detec_conv_fc", bn=False, act="sigmoid")


def netvlad(features, att, p, prefix="netvlad", final_l2norm=True):
    """features [B,N,D], att [B,N,1] -> [B,out]; literal restatement of backbones.py:202-320."""
    B, N, D = features.shape
    cw = p[prefix + ".cluster_weights"].astype(np.float64)
    Kc = cw.shape[1]
    x = l2_normalize(features.reshape(-1, D).astype(np.float64), 1)
    a = _bn(x @ cw, p, prefix + ".cluster_bn", SLIM_EPS)
    a = a - a.max(axis=1, keepdims=True)
    a = np.exp(a)
    a = a / a.sum(axis=1, keepdims=True)
    a = a * att.reshape(-1, 1)
    a = a.reshape(B, N, Kc)
    a_sum = a.sum(axis=-2, keepdims=True)
    cw2 = p[prefix + ".cluster_weights2"].astype(np.float64).reshape(1, D, Kc)
    a2 = a_sum * cw2
    vlad = np.transpose(a, (0, 2, 1)) @ x.reshape(B, N, D)   # [B,Kc,D]
    vlad = np.transpose(vlad, (0, 2, 1)) - a2                  # [B,D,Kc]
    vlad = l2_normalize(vlad, 1)
    vlad = l2_normalize(vlad.reshape(B, D * Kc), 1)
    vlad = vlad @ p[prefix + ".hidden1_weights"].astype(np.float64)
    vlad = _bn(vlad, p, prefix + ".bn", SLIM_EPS)
    gates = _bn(vlad @ p[prefix + ".gating_weights"].astype(np.float64), p, prefix + ".gating_bn",
                SLIM_EPS)
    vlad = vlad * _act(gates, "sigmoid")
    return l2_normalize(vlad, -1, 1e-8) if final_l2norm else vlad


def forward(points, p, detection=True, extract_global=True, knn_num=8, reference_cpu=False, add_se="max_pool",
            gl_dims=(256,), global_subsample=-1):
    """The inference branch of DH3D.build_graph.  points [B,N,3].  If ``p`` carries a second backbone under
    'global_local.' (the global checkpoint's own copy), the global branch runs on it -- the reference's two
    separately trained networks evaluated on the same cloud."""
    _MODE["reference_cpu"] = bool(reference_cpu)
    points = np.asarray(points, np.float32)
    knn = knn_pm(points, knn_num)
    pts64 = points.astype(np.float64)
    feat = backbone_local_dilate(pts64, p, knn, add_se=add_se)
    out = {"feat": feat, "local_desc": l2_normalize(feat, 2, 1e-8), "knn": knn}
    if detection:
        out["attention"] = attention_head(feat, p, "detection_block_reliable", 3)
    if extract_global:
        gfeat = feat
        if "global_local.initconv.position_theta" in p:
            gfeat = backbone_local_dilate(pts64, p, knn, prefix="global_local", add_se=add_se)
        fg = flex_conv_dilate(pts64, gfeat, p, "global_before_assemble", 8, list(gl_dims), concat=False,
                              add_se=False)
        if global_subsample > 0:    # core/model.py:118-121
            kp = oracle.farthest_point_sample(global_subsample, points)
            fg = np.take_along_axis(fg, kp[:, :, None].astype(np.int64), axis=1)
        att = attention_head(fg, p, "globalatt", 1)
        out["forglobal"], out["global_att"] = fg, att
        out["globaldesc"] = netvlad(fg, att, p)
    return out
