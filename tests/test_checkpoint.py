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
    assert tf_name_to_param("stage1/flexconv_0/position_theta/Adam_1") is None
    model = DH3D(full_config())
    loaded, missing = load_reference_checkpoint(model, os.path.join(REF, "local", "localmodel"),
                                                os.path.join(REF, "global", "globalmodel"))
    assert missing == [] and len(loaded) == len(list(model.named_parameters()))
    p = dict(model.named_parameters())
    assert np.array_equal(p["netvlad.gating_weights"].numpy(), g["gating_weights"])
    # the detector exists only in the local checkpoint
    l = read_tensor_bundle(os.path.join(REF, "local", "localmodel"))
    assert np.array_equal(p["detection_block_reliable.detec_conv_fc.W"].numpy(),
                          l["detection_block_reliable/detec_conv_fc/W"])


def test_data_helpers_on_demo_cloud():
    from dh3d_b200.data import get_fixednum_pcd, load_single_pcfile
    demo = "/root/reference/evaluate/global_eval/demo_data/2015-03-10-14-18-10"
    f = sorted(os.listdir(demo))[0]
    pc = load_single_pcfile(os.path.join(demo, f))
    assert pc.ndim == 2 and pc.shape[1] == 3 and pc.shape[0] > 4096
    fixed, ori = get_fixednum_pcd(pc, 8192, rng=np.random.RandomState(0))
    assert fixed.shape == (8192, 3) and fixed.dtype == np.float32 and 0 < ori <= 8192
