"""CPU replay of tests/golden/refcuda_*.npz -- outputs of the REFERENCE'S OWN CUDA kernels recorded
on a B200 by tests/golden/make_refcuda_golden.py -- against the CPU oracle.  Keeps the oracle
pinned to the real reference (k-NN tie order, FPS selection rule, ball-query nearest leak, Flex ops)
in runs without a GPU.  The -m gpu suite replays the same files against this repo's kernels."""
import glob
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _files(prefix):
    fs = sorted(glob.glob(os.path.join(GOLD, prefix + "*.npz")))
    assert fs, "golden fixtures missing: " + prefix
    return fs


@pytest.mark.parametrize("path", _files("refcuda_knn"))
def test_oracle_knn_equals_reference_cuda(path):
    g = np.load(path)
    ids, d = oracle.knn_bruteforce(g["positions"], g["ids"].shape[2])
    assert np.array_equal(ids, g["ids"]) and np.array_equal(d, g["dists"])
    lids, ld = oracle.knn_bruteforce(g["positions"], g["ids"].shape[2], literal=True)
    assert np.array_equal(lids, g["ids"]) and np.array_equal(ld, g["dists"])


@pytest.mark.parametrize("path", _files("refcuda_fps"))
def test_oracle_fps_equals_reference_cuda(path):
    g = np.load(path)
    assert np.array_equal(oracle.farthest_point_sample(g["idx"].shape[1], g["xyz"]), g["idx"])


def test_oracle_ball_query_equals_reference_cuda():
    g = np.load(os.path.join(GOLD, "refcuda_ball.npz"))
    idx, cnt = oracle.query_ball_point(float(g["radius"]), int(g["nsample"]), g["xyz1"], g["xyz2"])
    assert (g["cnt"] == 0).any(), "fixture should contain no-hit queries (nearest fallback)"
    assert np.array_equal(cnt, g["cnt"]) and np.array_equal(idx, g["idx"])


def test_oracle_flex_ops_equal_reference_cuda():
    g = np.load(os.path.join(GOLD, "refcuda_flex.npz"))
    fc = oracle.flex_convolution(g["features"], g["position"], g["neighborhood"], g["theta"], g["bias"])
    assert np.abs(fc - g["flex_conv"]).max() <= 1e-5 * np.sqrt((g["flex_conv"] ** 2).mean())
    po, pa = oracle.flex_pooling(g["features"], g["neighborhood"])
    assert np.array_equal(po, g["pool"]) and np.array_equal(pa, g["argmax"])
    cp = oracle.convolution_pointset(g["features"], g["neighborhood"], g["theta_rel"], g["bias_rel"])
    assert np.array_equal(cp, g["conv_pointset"])


@pytest.mark.gpu
def test_kernels_equal_reference_cuda_golden():
    import torch
    from dh3d_b200 import tf_ops, user_ops
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for path in _files("refcuda_knn"):
        g = np.load(path)
        ids, d = user_ops.knn_bruteforce(cu(g["positions"]), g["ids"].shape[2])
        assert np.array_equal(ids.cpu().numpy(), g["ids"]) and np.array_equal(d.cpu().numpy(), g["dists"])
    for path in _files("refcuda_fps"):
        g = np.load(path)
        assert np.array_equal(tf_ops.farthest_point_sample(g["idx"].shape[1], cu(g["xyz"])).cpu().numpy(), g["idx"])
    g = np.load(os.path.join(GOLD, "refcuda_ball.npz"))
    idx, cnt = tf_ops.query_ball_point(float(g["radius"]), int(g["nsample"]), cu(g["xyz1"]), cu(g["xyz2"]))
    assert np.array_equal(idx.cpu().numpy(), g["idx"]) and np.array_equal(cnt.cpu().numpy(), g["cnt"])
    g = np.load(os.path.join(GOLD, "refcuda_flex.npz"))
    fc = user_ops.flex_convolution(cu(g["features"]), cu(g["position"]), cu(g["neighborhood"]), cu(g["theta"]),
                                   cu(g["bias"])).cpu().numpy()
    assert np.abs(fc - g["flex_conv"]).max() <= 1e-4 * np.sqrt((g["flex_conv"] ** 2).mean())
    po, pa = user_ops.flex_pooling(cu(g["features"]), cu(g["neighborhood"]))
    assert np.array_equal(po.cpu().numpy(), g["pool"]) and np.array_equal(pa.cpu().numpy(), g["argmax"])
    cp = user_ops.convolution_pointset(cu(g["features"]), cu(g["neighborhood"]), cu(g["theta_rel"]), cu(g["bias_rel"]))
    assert np.array_equal(cp.cpu().numpy(), g["conv_pointset"])
