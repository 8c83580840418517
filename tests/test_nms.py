"""Keypoint NMS (core/utils.py:15-43): the numpy oracle against golden outputs of the reference function itself
(tests/golden/nms_*.npz, made by tests/golden/make_nms_golden.py with scikit-learn), and -- on the GPU -- this
repo's kernels against both."""
import glob
import os

import numpy as np
import pytest

from oracle import nms as oracle_nms

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nms_*.npz")))


def _load(path):
    g = np.load(path)
    rad, ratio, kp, noise = g["params"]
    return g["xyz"], g["attention"], float(rad), float(ratio), int(kp), bool(noise), int(g["num_keypoints"]), g["max_indices"]


def test_golden_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_function(path):
    xyz, att, rad, ratio, kp, noise, num, idx = _load(path)
    n, i = oracle_nms.single_nms(xyz, att, rad, ratio, kp, remove_noise=noise)
    assert n == num and np.array_equal(i, idx)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_matches_reference_function(path):
    import torch
    from dh3d_b200 import utils
    xyz, att, rad, ratio, kp, noise, num, idx = _load(path)
    a = torch.from_numpy(att).cuda()
    n, i = utils.single_nms(torch.from_numpy(xyz).cuda(), a, rad, ratio, kp, remove_noise=noise)
    assert n == num and np.array_equal(i.cpu().numpy(), idx)
    assert np.array_equal(a.cpu().numpy(), att)          # the input is not modified


@pytest.mark.gpu
def test_gpu_batched_matches_oracle_and_truncates():
    import torch
    from dh3d_b200 import utils
    rng = np.random.RandomState(5)
    B, N = 3, 3000
    xyz = rng.uniform([0, 0, 0], [10, 10, 1.0], (B, N, 3)).astype(np.float32)
    att = rng.rand(B, N).astype(np.float32)
    att[1, :100] = att[1, 100:200]                        # equal responses: (attention, index) ordering
    out, cnt = utils.batched_nms(torch.from_numpy(xyz).cuda(), torch.from_numpy(att).cuda(), 0.6, 0.05, 128)
    out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        n, i = oracle_nms.single_nms(xyz[b], att[b], 0.6, 0.05, 128)
        assert cnt[b] == n and np.array_equal(out[b, :n], i) and np.all(out[b, n:] == -1)
    with pytest.raises(Exception):
        utils.single_nms(torch.from_numpy(xyz[0, :40]).cuda(), torch.from_numpy(att[0, :40]).cuda(), 0.5, 0.01, 8)
