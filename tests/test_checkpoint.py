"""CPU tests of the TensorBundle reader / name mapping against the reference's shipped checkpoints
(skipped where /root/reference is absent, e.g. on the GPU box)."""
import os

import numpy as np
import pytest

REF = "/root/reference/models"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkpoints not present")


def test_read_bundle_and_fill_every_model_parameter():
    from dh3d_b200.checkpoint import load_reference_checkpoint, read_tensor_bundle, tf_name_to_param
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D
    g = read_tensor_bundle(os.path.join(REF, "global", "globalmodel"))
    assert g["hidden1_weights"].shape == (16384, 256) and g["cluster_weights2"].shape == (1, 256, 64)
    assert g["stage1/flexconv_0/position_theta"].shape == (3, 32, 64)
    assert np.isfinite(g["hidden1_weights"]).all() and abs(float(g["hidden1_weights"].std())) > 1e-4
    assert tf_name_to_param("stage1/flexconv_0_bn/mean/EMA") == "local.stage1.flexconv_0_bn.mean_ema"
    assert tf_name_to_param("cluster_bn/moving_variance") == "netvlad.cluster_bn.variance_ema"
    # optimizer slots map to names no model has; the loader filters on the model, not on name heuristics
    assert tf_name_to_param("stage1/flexconv_0/position_theta/Adam_1") == "local.stage1.flexconv_0.position_theta.Adam_1"
    assert tf_name_to_param("stage1/flexconv_0/position_theta", "global_local") == \
        "global_local.stage1.flexconv_0.position_theta"
    model = DH3D(full_config(), separate_global_backbone=True)
    loaded, missing = load_reference_checkpoint(model, os.path.join(REF, "local", "localmodel"),
                                                os.path.join(REF, "global", "globalmodel"))
    assert missing == [] and len(loaded) == len(list(model.named_parameters()))
    p = dict(model.named_parameters())
    assert np.array_equal(p["netvlad.gating_weights"].numpy(), g["gating_weights"])
    # the detector exists only in the local checkpoint
    l = read_tensor_bundle(os.path.join(REF, "local", "localmodel"))
    assert np.array_equal(p["detection_block_reliable.detec_conv_fc.W"].numpy(),
                          l["detection_block_reliable/detec_conv_fc/W"])


def test_two_checkpoints_never_silently_share_one_backbone():
    """the shipped local and global checkpoints carry DIFFERENT backbones; loading both into a
    one-backbone model must not silently evaluate the detector on the global run's backbone."""
    from dh3d_b200.checkpoint import (CheckpointError, checkpoint_model, load_reference_checkpoint,
                                      read_tensor_bundle)
    from dh3d_b200.configs import detection_config, full_config, global_config
    from dh3d_b200.model import DH3D
    lp, gp = os.path.join(REF, "local", "localmodel"), os.path.join(REF, "global", "globalmodel")
    l, g = read_tensor_bundle(lp), read_tensor_bundle(gp)
    k = "stage1/flexconv_0/position_theta"
    assert np.abs(l[k] - g[k]).max() > 1e-2          # they really differ
    with pytest.raises(CheckpointError, match="different backbones"):
        load_reference_checkpoint(DH3D(full_config()), lp, gp)
    for which, src in (("local", l), ("global", g)):
        m = DH3D(full_config())
        load_reference_checkpoint(m, lp, gp, shared_backbone=which)
        p = dict(m.named_parameters())
        assert np.array_equal(p["local.stage1.flexconv_0.position_theta"].numpy(), src[k])
        assert np.array_equal(p["netvlad.gating_weights"].numpy(), g["gating_weights"])
        assert np.array_equal(p["detection_block_reliable.detec_conv_fc.W"].numpy(),
                              l["detection_block_reliable/detec_conv_fc/W"])
    # separate backbones: each network keeps its own
    m = checkpoint_model(lp, gp)
    p = dict(m.named_parameters())
    assert np.array_equal(p["local.stage1.flexconv_0.position_theta"].numpy(), l[k])
    assert np.array_equal(p["global_local.stage1.flexconv_0.position_theta"].numpy(), g[k])
    # strict: a branch the checkpoint does not carry raises instead of staying at its zero init
    with pytest.raises(CheckpointError, match="neither checkpoint"):
        load_reference_checkpoint(DH3D(full_config()), lp, None)
    with pytest.raises(CheckpointError, match="neither checkpoint"):
        load_reference_checkpoint(DH3D(detection_config()), None, gp)
    loaded, missing = load_reference_checkpoint(DH3D(full_config()), lp, None, strict=False)
    assert any(n.startswith("netvlad.") for n in missing)
    assert load_reference_checkpoint(DH3D(global_config()), None, gp)[1] == []
    assert load_reference_checkpoint(DH3D(detection_config()), lp, None)[1] == []


def test_staged_weights_equal_the_checkpoints():
    """oracle/_ref/dh3d_weights.npz (what the GPU-box tests read) == the shipped TensorBundles."""
    from dh3d_b200.checkpoint import read_tensor_bundle
    from oracle import build_ref
    staged = build_ref.load_staged_weights()
    if staged is None:
        pytest.skip("weights not staged")
    for tag, st in zip(("local", "global"), staged):
        full = read_tensor_bundle(os.path.join(REF, tag, tag + "model"))
        assert all(np.array_equal(v, full[k]) for k, v in st.items())
        assert "hidden1_weights" in st or tag == "local"


def test_data_helpers_on_demo_cloud():
    from dh3d_b200.data import get_fixednum_pcd, load_single_pcfile
    demo = "/root/reference/evaluate/global_eval/demo_data/2015-03-10-14-18-10"
    f = sorted(os.listdir(demo))[0]
    pc = load_single_pcfile(os.path.join(demo, f))
    assert pc.ndim == 2 and pc.shape[1] == 3 and pc.shape[0] > 4096
    fixed, ori = get_fixednum_pcd(pc, 8192, rng=np.random.RandomState(0))
    assert fixed.shape == (8192, 3) and fixed.dtype == np.float32 and 0 < ori <= 8192


def test_real_weights_demo_descriptors_match_gpu_golden():
    """tests/golden/demo_globaldesc.npz holds the global descriptors this repo's CUDA forward produced
    on a B200 for the reference's 100 demo clouds with its SHIPPED checkpoints
    (scripts/eval_demo_retrieval.py --backend gpu) next to the fp64 oracle's.  Re-derive three of
    them with the oracle here and check all 100 stored pairs: real LiDAR data incl. 15 clouds padded
    with duplicated points (exact k-NN / FPS ties)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from eval_demo_retrieval import prepare
    from dh3d_b200.checkpoint import load_reference_checkpoint
    from dh3d_b200.configs import full_config
    from dh3d_b200.model import DH3D
    from oracle import net
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo_globaldesc.npz"))
    gpu, orc = gold["gpu_globaldesc"], gold["oracle_globaldesc"]
    assert gpu.shape == (100, 256)
    assert np.abs(gpu - orc).max() < 5e-5 and np.allclose(np.linalg.norm(gpu, axis=1), 1, atol=1e-5)
    assert gold["recall"][:, 0].min() >= 0.7 and gold["recall"][:, 1].min() >= 0.9   # recall@1 / @5
    names, clouds, ori = prepare("/root/reference")
    assert list(names) == list(gold["names"]) and int((ori < 8192).sum()) == 15
    model = DH3D(full_config())   # the stored descriptors were made on the global checkpoint's backbone
    load_reference_checkpoint(model, os.path.join(REF, "local", "localmodel"), os.path.join(REF, "global", "globalmodel"),
                              shared_backbone="global")
    params = {k: v.detach().numpy() for k, v in model.named_parameters()}
    padded = int(np.nonzero(ori < 8192)[0][0])
    for i in (0, padded, 99):
        d = net.forward(clouds[i:i + 1], params)["globaldesc"][0]
        assert np.abs(d - gpu[i]).max() < 5e-5, i
