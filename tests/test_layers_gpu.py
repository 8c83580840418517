"""GPU tests of boundary C -- the layer / model level API with the reference's signatures
(core/layers.py:439-480,686-707, core/tf_utils.py:48-109, core/backbones.py:202) -- and of the reference forward
configurations beyond the two shipped ones (add_se='avg_pool', featdim < 128, global_subsample > 0,
separate backbones with the real shipped weights)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import net
from conftest import make_cloud

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_err(a, e):
    a = a.detach().cpu().numpy().astype(np.float64)
    e = np.asarray(e, np.float64)
    return np.abs(a - e).max() / (np.sqrt(np.mean(e ** 2)) + 1e-30)


def _case(rng, B=2, N=600, K=8, Din=16):
    pts = make_cloud(rng, B, N, extent=8.0)
    pos = np.ascontiguousarray(pts.transpose(0, 2, 1))
    nb, _ = oracle.knn_bruteforce(pos, K)
    f = rng.randn(B, Din, N).astype(np.float32)
    return f, pos, np.ascontiguousarray(nb.transpose(0, 2, 1))


def test_function_forms_have_the_reference_signatures():
    """flex_convolution(features, positions, neighborhoods, filters, activation=..., name=...) etc.: positional
    order of core/layers.py, parameters owned by the named scope and re-used by a second call."""
    from dh3d_b200 import layers
    rng = np.random.RandomState(0)
    f, pos, nb = _case(rng)
    store = layers.VariableStore()
    y = layers.flex_convolution(cu(f), cu(pos), cu(nb), 24, activation=None, name="fc0", store=store)
    lay = store["fc0"]
    assert tuple(lay.position_theta.shape) == (3, 16, 24) and tuple(lay.position_bias.shape) == (16, 24)
    assert tuple(lay.feature_bias.shape) == (24, 1)
    with torch.no_grad():
        lay.position_bias.normal_()
        lay.feature_bias.normal_()
    y = layers.flex_convolution(cu(f), cu(pos), cu(nb), 24, activation="relu", name="fc0", store=store)
    exp = oracle.flex_convolution(f, pos, nb, lay.position_theta.cpu().numpy(), lay.position_bias.cpu().numpy(),
                                  f64=True) + lay.feature_bias.cpu().numpy().reshape(1, -1, 1)
    assert y.shape == (2, 24, 600) and rel_err(y, np.maximum(exp, 0)) < 1e-4
    assert len(store) == 1

    z = layers.convolution_pointset(cu(pos), cu(nb), 32, name="init", store=store)
    cp = store["init"]
    assert tuple(cp.position_theta.shape) == (3, 32) and cp.feature_bias is None
    ez = oracle.convolution_pointset(pos, nb, cp.position_theta.cpu().numpy(), cp.position_bias.cpu().numpy())
    assert np.array_equal(z.cpu().numpy(), ez)

    nn_, dist = layers.knn_bruteforce(cu(pos), 8)
    assert nn_.shape == (2, 8, 600) and np.array_equal(nn_.cpu().numpy(), nb)
    pooled = layers.flex_pooling(cu(f), cu(nb))
    assert np.array_equal(pooled.cpu().numpy(), oracle.flex_pooling(f, nb)[0])
    with pytest.raises(ValueError):
        layers.flex_convolution(cu(f), cu(pos), cu(nb), 24)          # no scope name -> nowhere to own variables


def test_flex_avg_is_the_neighbour_sum():
    """core/layers.py:342-436,464-480: theta = 0 (stored, non-trainable), bias = eye -> sum over the K neighbours."""
    from dh3d_b200 import layers
    rng = np.random.RandomState(1)
    f, pos, nb = _case(rng, Din=64)
    store = layers.VariableStore()
    y = layers.flex_avg(cu(f), cu(pos), cu(nb), 64, name="se_avgpool", store=store)
    lay = store["se_avgpool"]
    assert [n for n, _ in lay.named_parameters()] == ["position_theta"] and float(lay.position_theta.abs().max()) == 0
    exp = np.take_along_axis(f[:, :, None, :].astype(np.float64), nb[:, None, :, :].astype(np.int64), axis=3).sum(axis=2)
    assert rel_err(y, exp) < 1e-4
    with pytest.raises(ValueError):
        layers.flex_avg(cu(f), cu(pos), cu(nb), 32, name="bad", store=store)


def test_tf_utils_forms():
    from dh3d_b200 import layers, tf_utils
    rng = np.random.RandomState(2)
    f, pos, nb = _case(rng, Din=32)
    store = layers.VariableStore()
    y = tf_utils.flexconv_withBatchnorm(cu(f), cu(pos), cu(nb), 64, name="flexconv_0", store=store)
    conv, bn = store["flexconv_0"], store["flexconv_0_bn"]
    with torch.no_grad():
        conv.position_theta.normal_(0, 0.1)
        conv.position_bias.normal_(0, 0.1)
        bn.gamma.uniform_(0.5, 1.5)
        bn.mean_ema.normal_()
    layers.invalidate_folded(store)
    y = tf_utils.flexconv_withBatchnorm(cu(f), cu(pos), cu(nb), 64, name="flexconv_0", store=store)
    exp = oracle.flex_convolution(f, pos, nb, conv.position_theta.cpu().numpy(), conv.position_bias.cpu().numpy(), f64=True)
    g, b, m, v = (t.cpu().numpy().reshape(1, -1, 1) for t in (bn.gamma, bn.beta, bn.mean_ema, bn.variance_ema))
    exp = np.maximum((exp - m) / np.sqrt(v + 1e-5) * g + b, 0)
    assert y.shape == (2, 64, 600) and rel_err(y, exp) < 1e-4

    x = rng.randn(2, 600, 48).astype(np.float32)
    h = tf_utils.feature_conv1d_1(cu(x), 16, "f1", ac_func="relu", store=store)
    lay = store["f1"].tfconv0
    assert lay.bn is None and h.shape == (2, 600, 16)
    pts = np.ascontiguousarray(pos.transpose(0, 2, 1))
    xyz_s, feat_s, kp = tf_utils.subsample(cu(pts), cu(x), 75, None)
    ekp = oracle.farthest_point_sample(75, pts)
    assert kp.shape == (2, 75, 1) and np.array_equal(kp[:, :, 0].cpu().numpy(), ekp)
    assert np.array_equal(feat_s.cpu().numpy(), np.take_along_axis(x, ekp[:, :, None].astype(np.int64), axis=1))
    assert np.array_equal(xyz_s.cpu().numpy(), np.take_along_axis(pts, ekp[:, :, None].astype(np.int64), axis=1))


def test_global_netvald_block_function_form():
    from dh3d_b200 import layers
    from dh3d_b200.backbones import global_netvald_block
    from dh3d_b200.model import init_random_
    rng = np.random.RandomState(3)
    feats = rng.randn(2, 500, 256).astype(np.float32)
    att = rng.rand(2, 500, 1).astype(np.float32)
    xyz = make_cloud(rng, 2, 500)
    store = layers.VariableStore()
    global_netvald_block(cu(xyz), cu(feats), cu(att), False, store=store)        # creates the scope
    init_random_(store["netvlad"], seed=3)
    y = global_netvald_block(cu(xyz), cu(feats), cu(att), False, cluster_size=64, output_dim=256, add_batch_norm=True,
                             gating=True, store=store)
    p = {"netvlad." + k: v.detach().cpu().numpy() for k, v in store["netvlad"].named_parameters()}
    exp = net.netvlad(feats, att, p, final_l2norm=False)
    assert y.shape == (2, 256) and rel_err(y, exp) < 1e-4
    from dh3d_b200._lib import Dh3dError
    with pytest.raises(Dh3dError):
        global_netvald_block(cu(xyz), cu(feats), cu(att), True, store=store)


@pytest.mark.parametrize("kw", [dict(add_se="avg_pool"), dict(featdim=64), dict(global_subsample=256),
                                dict(add_se="", featdim=32)])
def test_other_reference_configurations_match_oracle(kw):
    """Forward configurations the reference code offers beyond the two shipped ones: add_se='avg_pool'
    (core/backbones.py:79-82), featdim < 128 with final_fc (:125-126), global_subsample > 0 (core/model.py:118-121)."""
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D, init_random_
    cfg = full_config(**kw)
    model = init_random_(DH3D(cfg), seed=11)
    params = {k: v.detach().numpy().copy() for k, v in model.named_parameters()}
    pts = make_cloud(np.random.RandomState(11), 2, 1024, extent=10.0)
    out = model.cuda()(cu(pts))
    exp = net.forward(pts, params, add_se=cfg.add_se, global_subsample=cfg.global_subsample)
    assert out["local_desc"].shape == (2, 1024, cfg.featdim)
    for k in ("feat", "local_desc", "attention", "globaldesc"):
        assert rel_err(out[k], exp[k]) < 5e-4, (k, rel_err(out[k], exp[k]))


def test_unsupported_configurations_raise_clearly():
    from dh3d_b200._lib import Dh3dError
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D
    with pytest.raises(Dh3dError, match="gl_dims"):
        DH3D(full_config(gl_dims=[128]))
    with pytest.raises(Dh3dError, match="add_se"):
        DH3D(full_config(add_se="median"))
    DH3D(full_config(gl_dims=[128, 256]))


def test_device_move_and_state_dict_refresh_cached_operands():
    """folded BN / prepacked weights / the b2 scalar must follow .cuda(), load_state_dict() and
    invalidate_folded -- never stale."""
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D, init_random_
    pts = cu(make_cloud(np.random.RandomState(4), 1, 1024, extent=10.0))
    a = init_random_(DH3D(full_config()), seed=1).cuda()
    b = init_random_(DH3D(full_config()), seed=2).cuda()
    ya, yb = a(pts), b(pts)
    assert not torch.equal(ya["globaldesc"], yb["globaldesc"])
    b.load_state_dict(a.state_dict())        # b has warm caches from its own weights
    yb2 = b(pts)
    for k in ("feat", "attention", "globaldesc"):
        assert torch.equal(ya[k], yb2[k]), k
    c = init_random_(DH3D(full_config()), seed=1)
    c = c.cuda()
    assert torch.equal(c(pts)["globaldesc"], ya["globaldesc"])


@pytest.mark.timeout(900)
def test_real_weights_both_networks_on_demo_clouds():
    """The reference's two shipped networks (local + detector, global) evaluated in ONE pass with separate
    backbones on four of its own demo clouds (real Oxford LiDAR; two padded with duplicated points), against the
    fp64 oracle composition with the same weights.  Weights are staged under oracle/_ref by build() in the build
    container (the GPU box has no /root/reference)."""
    from oracle import build_ref
    from dh3d_b200.checkpoint import load_reference_checkpoint
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D
    staged = build_ref.load_staged_weights()
    if staged is None:
        pytest.skip("oracle/_ref/dh3d_weights.npz not staged")
    z = np.load(os.path.join(GOLD, "demo_clouds.npz"))
    clouds = z["clouds"]
    assert clouds.shape == (4, 8192, 3) and int((z["ori_num"] < 8192).sum()) == 2
    model = DH3D(full_config(), separate_global_backbone=True)
    load_reference_checkpoint(model, staged[0], staged[1])
    params = {k: v.detach().numpy().copy() for k, v in model.named_parameters()}
    out = model.cuda()(cu(clouds))
    exp = net.forward(clouds, params)
    errs = {k: rel_err(out[k], exp[k]) for k in ("feat", "local_desc", "attention", "globaldesc")}
    print("real weights, demo clouds, GPU vs fp64 oracle:", errs)
    # the golden descriptors of round 1 were made with the global checkpoint's backbone: same numbers here
    gold = np.load(os.path.join(GOLD, "demo_globaldesc.npz"))
    assert np.abs(out["globaldesc"].cpu().numpy() - gold["oracle_globaldesc"][z["index"]]).max() < 5e-5
    for k, e in errs.items():
        assert e < 5e-4, (k, errs)
