"""CPU tests: pin the oracle against the reference's own fixtures / known answers and its
internal consistency (literal vs fast forms).  No GPU, no CUDA calls."""
import numpy as np
import pytest

import oracle
from oracle import fixtures
from conftest import lattice_cloud, make_cloud


def test_knn_matches_reference_numpy_oracle_on_reference_fixtures():
    # user_ops/test_knn_bruteforce.py:47-62 -- both module-level cases, k=4, rtol=atol=1e-6
    for case in fixtures.reference_test_cases():
        exp_ids, exp_d = fixtures.python_bruteforce(case.position, 4)
        ids, d = oracle.knn_bruteforce(case.position, 4)
        assert np.array_equal(ids, exp_ids)
        assert np.allclose(d, exp_d, rtol=1e-6, atol=1e-6)


def test_knn_config1_1024_points_k8():
    # BASELINE.json configs[0]: 1024 random 3-D points, K=8, numpy reference path
    rng = np.random.RandomState(1234)
    pos = rng.randn(1, 3, 1024).astype(np.float32)
    exp_ids, exp_d = fixtures.python_bruteforce(pos.astype(np.float64), 8)
    ids, d = oracle.knn_bruteforce(pos, 8)
    assert np.array_equal(ids, exp_ids)
    assert np.allclose(d, exp_d, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("N", [4, 33, 100, 257, 600, 1500])
def test_knn_fast_equals_literal_with_ties(N):
    rng = np.random.RandomState(N)
    pts = lattice_cloud(rng, 2, N)
    pos = np.ascontiguousarray(pts.transpose(0, 2, 1))
    K = min(8, N)
    a = oracle.knn_bruteforce(pos, K)
    b = oracle.knn_bruteforce(pos, K, literal=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_knn_tie_order_is_blocked_rank_not_index():
    # 300 identical points: N<=512 -> T=128,V=4, rank(x) = (x%128)*4 + x//128, so the first
    # neighbours are x = 0, 128, 256, 1, 129, 257, 2, ...  
    pos = np.zeros((1, 3, 300), np.float32)
    ids, d = oracle.knn_bruteforce(pos, 8, literal=True)
    assert ids[0, 0].tolist() == [0, 128, 256, 1, 129, 257, 2, 130]
    assert np.all(d == 0)


def test_knn_k_larger_than_n_pads_like_reference():
    pos = np.random.RandomState(0).randn(1, 3, 5).astype(np.float32)
    ids, d = oracle.knn_bruteforce(pos, 8, literal=True)
    assert np.all(ids[:, :, 5:] == -1) and np.all(d[:, :, 5:] == np.float32(3.4028235e38))


def test_flexpool_four_point_case():
    # user_ops/test_flex_pooling.py:76-98: every ring neighbourhood contains the 5 at index 2
    x, n = fixtures.flexpool_four_point_case()
    out, arg = oracle.flex_pooling(x, n)
    assert np.all(out == 5.0) and np.all(arg == 2)


def test_flex_ops_on_seed42_fixture_fp32_vs_fp64():
    # test_flex_convolution.py:42-50 compares fp32 against fp64 at rtol 1e-4 on this fixture
    case, _ = fixtures.reference_test_cases()
    a = oracle.flex_convolution(case.features, case.position, case.neighborhood, case.theta, case.bias)
    b = oracle.flex_convolution(case.features, case.position, case.neighborhood, case.theta, case.bias,
                                f64=True)
    assert np.allclose(a, b, rtol=1e-4, atol=1e-5)
    # self is the first neighbour on this fixture, so the CPU (nbr_0) and GPU (n) centres agree
    c = oracle.flex_convolution(case.features, case.position, case.neighborhood, case.theta, case.bias,
                                centre_is_self=False)
    assert np.allclose(a, c, rtol=1e-5, atol=1e-6)


def test_flex_conv_factored_form_is_exact():
    # out = A @ Theta_ext with A the 4*Din neighbour moments
    case, _ = fixtures.reference_test_cases()
    f, p, nb = case.features, case.position, case.neighborhood
    B, Din, N = f.shape
    ref = oracle.flex_convolution(f, p, nb, case.theta, case.bias, f64=True)
    out = np.zeros_like(ref)
    for b in range(B):
        for n in range(N):
            ids = nb[b, :, n]
            d = p[b][:, ids] - p[b][:, n:n + 1]                     # [3,K]
            dext = np.concatenate([np.ones((1, len(ids))), d], 0)    # [4,K]
            A = dext @ f[b][:, ids].T                                # [4,Din]
            theta_ext = np.concatenate([case.bias[None], case.theta], 0)  # [4,Din,Dout]
            out[b, :, n] = np.einsum("pc,pco->o", A, theta_ext)
    assert np.abs(out - ref).max() < 1e-12


def test_conv_pointset_matches_direct_numpy():
    case, _ = fixtures.reference_test_cases()
    f, nb = case.features, case.neighborhood
    out = oracle.convolution_pointset(f, nb, case.theta_rel, case.bias_rel)
    B, Din, N = f.shape
    ref = np.zeros_like(out, dtype=np.float64)
    for b in range(B):
        for n in range(N):
            ids = nb[b, :, n]
            delta = f[b][:, ids] - f[b][:, ids[0]:ids[0] + 1]
            ref[b, :, n] = case.bias_rel + (case.theta_rel.T @ delta).sum(1)
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_fps_basic_properties_and_tie_rule():
    rng = np.random.RandomState(3)
    pts = make_cloud(rng, 2, 2000)
    idx = oracle.farthest_point_sample(64, pts)
    assert idx.shape == (2, 64) and np.all(idx[:, 0] == 0)
    for b in range(2):
        assert len(set(idx[b].tolist())) == 64
        # greedy property: each pick maximises the distance to the picks so far
        d = np.full(2000, np.inf)
        for j in range(1, 64):
            d = np.minimum(d, ((pts[b] - pts[b, idx[b, j - 1]]) ** 2).sum(1))
            assert d[idx[b, j]] >= d.max() * (1 - 1e-5)
    # all-duplicate cloud: every distance is 0, strict '>' from best=-1 picks tid 0's first point
    dup = np.ones((1, 1500, 3), np.float32)
    assert np.all(oracle.farthest_point_sample(16, dup) == 0)
    # two-cluster tie: points 1..N-1 identical => the (k mod 512, k) rule picks k=512, not k=1
    tie = np.zeros((1, 1100, 3), np.float32)
    tie[0, 1:] = 1.0
    assert oracle.farthest_point_sample(2, tie)[0, 1] == 512


def test_three_nn_against_bruteforce_and_tie_rule():
    rng = np.random.RandomState(5)
    a, b = make_cloud(rng, 2, 300), make_cloud(rng, 2, 50)
    dist, idx = oracle.three_nn(a, b)
    d = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :]) ** 2).sum(-1)
    order = np.argsort(d, axis=2, kind="stable")[:, :, :3]
    assert np.array_equal(idx, order)
    assert np.allclose(dist, np.take_along_axis(d, order, 2), rtol=1e-5)
    # ties: earlier index wins; fewer than 3 candidates leave 1e40 -> inf and index 0
    dist, idx = oracle.three_nn(np.zeros((1, 2, 3), np.float32), np.zeros((1, 2, 3), np.float32))
    assert idx[0, 0].tolist() == [0, 1, 0] and np.isinf(dist[0, 0, 2])


def test_three_interpolate_and_weights():
    rng = np.random.RandomState(6)
    pts = rng.randn(2, 20, 8).astype(np.float32)
    idx = rng.randint(0, 20, (2, 33, 3)).astype(np.int32)
    dist = rng.rand(2, 33, 3).astype(np.float32)
    dist[0, 0] = 0  # exact hit -> clamped to 1e-10
    w = oracle.three_nn_weights(dist)
    assert np.allclose(w.sum(-1), 1, atol=1e-6)
    out = oracle.three_interpolate(pts, idx, w)
    ref = (pts[np.arange(2)[:, None, None], idx] * w[..., None]).sum(2)
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_query_ball_point_semantics_and_leak():
    rng = np.random.RandomState(7)
    xyz1, xyz2 = make_cloud(rng, 2, 400, extent=2.0), make_cloud(rng, 2, 300, extent=2.0)
    idx, cnt = oracle.query_ball_point(0.8, 16, xyz1, xyz2)
    d = np.sqrt(((xyz2[:, :, None, :].astype(np.float64) - xyz1[:, None, :, :]) ** 2).sum(-1))
    for b in range(2):
        for j in range(300):
            inball = np.nonzero(d[b, j] < 0.8)[0]
            if len(inball):
                exp = inball[:16]
                assert cnt[b, j] == len(exp)
                assert idx[b, j, :len(exp)].tolist() == exp.tolist()
                assert np.all(idx[b, j, len(exp):] == exp[0])
    # leak (tf_grouping_g.cu:13-14): nearest_d persists across the queries of one thread
    # (j, j+256 share a thread).  Query 0 sits on dataset point 5; query 256 is far from everything
    # and closest to dataset point 9; no ball hits -> query 256 inherits nearest_k=5 from query 0.
    x1 = np.zeros((1, 10, 3), np.float32); x1[0, :, 0] = np.arange(10) * 10.0
    x2 = np.full((1, 257, 3), 1000.0, np.float32)
    x2[0, 0] = x1[0, 5]; x2[0, 0, 1] = 3.0      # 3 m off point 5, radius 1 -> no hit
    x2[0, 256] = (95.0, 50.0, 0.0)               # nearest would be 9 at ~50.2 m
    idx, cnt = oracle.query_ball_point(1.0, 4, x1, x2)
    assert cnt[0, 0] == 0 and cnt[0, 256] == 0
    assert np.all(idx[0, 0] == 5) and np.all(idx[0, 256] == 5)


def test_group_and_gather():
    rng = np.random.RandomState(8)
    pts = rng.randn(2, 50, 7).astype(np.float32)
    idx = rng.randint(0, 50, (2, 9, 4)).astype(np.int32)
    assert np.array_equal(oracle.group_point(pts, idx), pts[np.arange(2)[:, None, None], idx])
    xyz = rng.randn(2, 50, 3).astype(np.float32)
    gi = rng.randint(0, 50, (2, 11)).astype(np.int32)
    assert np.array_equal(oracle.gather_point(xyz, gi), xyz[np.arange(2)[:, None], gi])


def test_net_oracle_netvlad_shapes_and_norm():
    from oracle import net
    rng = np.random.RandomState(9)
    D, K = 256, 64
    p = {"netvlad.cluster_weights": rng.randn(D, K) / 16, "netvlad.cluster_weights2": rng.randn(1, D, K) / 16,
         "netvlad.hidden1_weights": rng.randn(D * K, 256) / 8, "netvlad.gating_weights": rng.randn(256, 256) / 16}
    for bn, c in (("cluster_bn", K), ("bn", 256), ("gating_bn", 256)):
        p["netvlad.%s.gamma" % bn] = np.ones(c); p["netvlad.%s.beta" % bn] = np.zeros(c)
        p["netvlad.%s.mean_ema" % bn] = np.zeros(c); p["netvlad.%s.variance_ema" % bn] = np.ones(c)
    out = net.netvlad(rng.randn(2, 100, D), rng.rand(2, 100, 1), p)
    assert out.shape == (2, 256) and np.allclose((out ** 2).sum(1), 1)
