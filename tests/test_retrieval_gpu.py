"""GPU tests of the retrieval step (top-k after the descriptor all-gather) and of the checkpoint
loader on synthetic bundles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Q,R,K", [(24, 50, 25), (100, 4096, 25), (7, 33, 5), (512, 1000, 32)])
def test_retrieve_topk_matches_bruteforce(Q, R, K):
    from dh3d_b200.retrieval import retrieve_topk
    rng = np.random.RandomState(Q + R)
    ref = rng.randn(R, 256).astype(np.float32)
    ref /= np.linalg.norm(ref, axis=1, keepdims=True)
    qry = ref[rng.randint(0, R, Q)] + 0.3 * rng.randn(Q, 256).astype(np.float32) / 16
    idx, d2 = retrieve_topk(torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda(), K)
    full = ((qry[:, None, :].astype(np.float64) - ref[None]) ** 2).sum(-1)
    order = np.argsort(full, axis=1, kind="stable")[:, :K]
    got = idx.cpu().numpy()
    # identical neighbour sets; order may differ only where fp32 distances tie within rounding
    exp_d = np.take_along_axis(full, order, 1)
    got_d = np.take_along_axis(full, got.astype(np.int64), 1)
    assert np.allclose(got_d, exp_d, rtol=1e-4, atol=1e-5)
    assert np.all(np.diff(d2.cpu().numpy(), axis=1) >= -1e-6)


def test_retrieve_topk_near_identical_descriptors_find_themselves():
    """Descriptors that differ by ~1e-3 (what a random-weight network emits for noise clouds; real descriptors of
    revisited places are close too): the Gram form ||q||^2 + ||r||^2 - 2 q.r cancels to rounding noise at that
    scale, the exact re-rank (dh3d_topk_l2_exact) must put every query's own descriptor first and reproduce the
    fp64 order of the rest."""
    from dh3d_b200.retrieval import retrieve_topk
    rng = np.random.RandomState(3)
    base = rng.randn(256).astype(np.float32)
    base /= np.linalg.norm(base)
    ref = (base[None, :] + 1e-3 * rng.randn(2048, 256)).astype(np.float32)
    ref /= np.linalg.norm(ref, axis=1, keepdims=True)
    d = torch.from_numpy(ref).cuda()
    idx, d2 = retrieve_topk(d, d, 10)
    got = idx.cpu().numpy()
    assert np.array_equal(got[:, 0], np.arange(2048))
    assert float(d2[:, 0].abs().max()) == 0.0
    full = ((ref[:, None, :].astype(np.float64) - ref[None]) ** 2).sum(-1)
    exp_d = np.sort(full, axis=1)[:, :10]
    got_d = np.take_along_axis(full, got.astype(np.int64), 1)
    assert np.allclose(got_d[:, :6], exp_d[:, :6], rtol=2e-3, atol=0)    # inside the k + 8 candidate margin
