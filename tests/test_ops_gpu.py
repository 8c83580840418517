"""GPU parity tests: every CUDA op, called through the C ABI (via the reference-signature python
mirrors), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact for k-NN / FPS / grouping / ball-query / 3-NN indices and
for pure-select float outputs; FlexConv / ConvPointset / interpolation / NetVLAD within 1e-4
relative fp32 (stated per test as rtol with an atol tied to the output scale).
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import fixtures
from conftest import lattice_cloud, make_cloud

pytestmark = pytest.mark.gpu

REL = 1e-4  # north_star tolerance for floating-point features


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(actual, expected, rel=REL):
    """|a-e| <= rel*|e| + rel*rms(e): 1e-4 relative, with the absolute floor scaled to the output."""
    a = actual.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(actual) else np.asarray(actual, np.float64)
    e = np.asarray(expected, np.float64)
    scale = np.sqrt(np.mean(e ** 2)) + 1e-30
    err = np.abs(a - e)
    ok = err <= rel * np.abs(e) + rel * scale
    assert ok.all(), "max err %.3e (rms %.3e) at %d / %d elements" % (err.max(), scale, (~ok).sum(), ok.size)


# ------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("B,N,K", [(1, 4, 4), (2, 32, 4), (1, 1024, 8), (2, 2000, 8), (2, 4096, 16),
                                   (2, 8192, 8), (1, 8192, 32), (3, 100, 5), (1, 700, 1), (2, 3000, 50), (1, 400, 64)])
def test_knn_bitexact_random(B, N, K):
    from dh3d_b200 import user_ops
    rng = np.random.RandomState(B * 131 + N + K)
    pos = np.ascontiguousarray(make_cloud(rng, B, N).transpose(0, 2, 1))
    ids, d = user_ops.knn_bruteforce(cu(pos), K)
    eids, ed = oracle.knn_bruteforce(pos, K)
    assert np.array_equal(ids.cpu().numpy(), eids)
    assert np.array_equal(d.cpu().numpy(), ed)  # same fma chain + sqrt.rn -> identical bits


def test_knn_reference_fixture_known_answer():
    # user_ops/test_knn_bruteforce.py: B=1,N=4 case with k=4 against the numpy oracle, 1e-6
    from dh3d_b200 import user_ops
    for case in fixtures.reference_test_cases():
        ids, d = user_ops.knn_bruteforce(cu(case.position.astype(np.float32)), 4)
        eids, ed = fixtures.python_bruteforce(case.position, 4)
        assert np.array_equal(ids.cpu().numpy(), eids)
        assert np.allclose(d.cpu().numpy(), ed, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("N", [300, 1000, 3000, 8192])
def test_knn_bitexact_with_ties(N):
    """Lattice points and duplicated points: masses of equal keys -> the BlockRadixSort blocked
    rank order decides."""
    from dh3d_b200 import ops
    rng = np.random.RandomState(N)
    pts = lattice_cloud(rng, 2, N)
    pts[1] = make_cloud(rng, 1, N, duplicates=N // 5)[0]
    ids, d = ops.knn_points(cu(pts), 8)
    eids, ed = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), 8)
    assert np.array_equal(ids.cpu().numpy(), eids)
    assert np.array_equal(d.cpu().numpy(), ed)


def test_knn_k_larger_than_n_and_all_identical():
    from dh3d_b200 import user_ops
    pos = np.random.RandomState(0).randn(1, 3, 5).astype(np.float32)
    ids, d = user_ops.knn_bruteforce(cu(pos), 8)
    eids, ed = oracle.knn_bruteforce(pos, 8, literal=True)
    assert np.array_equal(ids.cpu().numpy(), eids) and np.array_equal(d.cpu().numpy(), ed)
    same = np.zeros((1, 3, 300), np.float32)
    ids, _ = user_ops.knn_bruteforce(cu(same), 8)
    assert ids[0, 0].tolist() == [0, 128, 256, 1, 129, 257, 2, 130]


def test_knn_above_reference_cap_properties():
    """N > 8192 (the reference refuses): sorted, self first, matches an fp64 brute force."""
    from dh3d_b200 import ops
    rng = np.random.RandomState(11)
    pts = make_cloud(rng, 1, 16384)
    ids, d = ops.knn_points(cu(pts), 8)
    ids, d = ids.cpu().numpy(), d.cpu().numpy()
    assert np.all(ids[0, :, 0] == np.arange(16384)) and np.all(np.diff(d, axis=2) >= 0)
    eids, ed = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), 8)
    assert np.array_equal(ids, eids) and np.array_equal(d, ed)


# ------------------------------------------------------------------------------------------ FPS
@pytest.mark.parametrize("B,N,M", [(2, 8192, 1024), (3, 1024, 128), (2, 700, 70), (1, 5000, 333),
                                   (2, 3000, 512), (1, 100, 100), (1, 16384, 256), (4, 2048, 256)])
def test_fps_bitexact(B, N, M):
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(N + M)
    pts = make_cloud(rng, B, N)
    idx = tf_ops.farthest_point_sample(M, cu(pts))
    assert np.array_equal(idx.cpu().numpy(), oracle.farthest_point_sample(M, pts))


def test_fps_ties_duplicates_and_lattice():
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(4)
    pts = lattice_cloud(rng, 3, 4096, step=1.0, side=6)      # only 216 distinct positions
    pts[2] = make_cloud(rng, 1, 4096, duplicates=3000)[0]
    idx = tf_ops.farthest_point_sample(512, cu(pts))
    assert np.array_equal(idx.cpu().numpy(), oracle.farthest_point_sample(512, pts))
    tie = np.zeros((1, 1100, 3), np.float32)
    tie[0, 1:] = 1.0
    assert tf_ops.farthest_point_sample(2, cu(tie))[0, 1].item() == 512  # (k mod 512, k) rule


# ---------------------------------------------------------------------------- gather / group / pool
def test_group_and_gather_point_exact():
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(8)
    for C in (3, 64, 128, 1, 7):
        pts = rng.randn(2, 500, C).astype(np.float32)
        idx = rng.randint(0, 500, (2, 77, 3)).astype(np.int32)
        out = tf_ops.group_point(cu(pts), cu(idx))
        assert np.array_equal(out.cpu().numpy(), oracle.group_point(pts, idx))
    xyz = rng.randn(2, 500, 3).astype(np.float32)
    gi = rng.randint(0, 500, (2, 91)).astype(np.int32)
    assert np.array_equal(tf_ops.gather_point(cu(xyz), cu(gi)).cpu().numpy(), oracle.gather_point(xyz, gi))


@pytest.mark.parametrize("B,N,K,D", [(2, 32, 4, 2), (2, 1000, 8, 32), (1, 8192, 8, 64), (2, 1024, 8, 128)])
def test_flex_pool_exact_both_layouts(B, N, K, D):
    from dh3d_b200 import ops, user_ops
    rng = np.random.RandomState(N + D)
    f = rng.randn(B, D, N).astype(np.float32)
    f[:, :, ::3] = np.round(f[:, :, ::3])  # repeated values -> first-max-wins matters
    nb = rng.randint(0, N, (B, K, N)).astype(np.int32)
    eo, ea = oracle.flex_pooling(f, nb)
    o, a = user_ops.flex_pooling(cu(f), cu(nb))
    assert np.array_equal(o.cpu().numpy(), eo) and np.array_equal(a.cpu().numpy(), ea)
    o, a = ops.flex_pool(cu(f.transpose(0, 2, 1)), cu(nb.transpose(0, 2, 1)), with_argmax=True)
    assert np.array_equal(o.cpu().numpy().transpose(0, 2, 1), eo)
    assert np.array_equal(a.cpu().numpy().transpose(0, 2, 1), ea)


def test_flex_pool_reference_four_point_case():
    from dh3d_b200 import user_ops
    x, n = fixtures.flexpool_four_point_case()
    o, a = user_ops.flex_pooling(cu(x), cu(n))
    assert torch.all(o == 5) and torch.all(a == 2)


# ------------------------------------------------------------------------------------ conv_pointset
def test_conv_pointset_reference_fixture_and_dh3d_shape():
    from dh3d_b200 import ops, user_ops
    case, _ = fixtures.reference_test_cases()
    f32 = lambda a: a.astype(np.float32)
    out = user_ops.convolution_pointset(cu(f32(case.features)), cu(case.neighborhood),
                                        cu(f32(case.theta_rel)), cu(f32(case.bias_rel)))
    close(out, oracle.convolution_pointset(case.features, case.neighborhood, case.theta_rel, case.bias_rel))
    rng = np.random.RandomState(12)
    pts = make_cloud(rng, 2, 4096)
    nb, _ = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), 8)
    th, bi = rng.randn(3, 32).astype(np.float32), rng.randn(32).astype(np.float32)
    exp = oracle.convolution_pointset(pts.transpose(0, 2, 1), nb.transpose(0, 2, 1), th, bi)
    out = ops.conv_pointset(cu(pts), cu(th), cu(bi), cu(nb))
    assert np.array_equal(out.cpu().numpy().transpose(0, 2, 1), exp)  # same FMA order -> same bits
    sc, sh = rng.rand(32).astype(np.float32) + 0.5, rng.randn(32).astype(np.float32)
    out = ops.conv_pointset(cu(pts), cu(th), cu(bi), cu(nb), scale=cu(sc), shift=cu(sh), act=1)
    close(out, np.maximum(exp.transpose(0, 2, 1) * sc + sh, 0))
    # ragged row count (tail of the 8-lanes-per-point fast path) and the generic kernel (Dout != 32)
    pts = make_cloud(rng, 3, 1001)
    nb, _ = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), 8)
    for dout in (32, 40):
        th, bi = rng.randn(3, dout).astype(np.float32), rng.randn(dout).astype(np.float32)
        exp = oracle.convolution_pointset(pts.transpose(0, 2, 1), nb.transpose(0, 2, 1), th, bi)
        out = ops.conv_pointset(cu(pts), cu(th), cu(bi), cu(nb))
        assert np.array_equal(out.cpu().numpy().transpose(0, 2, 1), exp)


# ---------------------------------------------------------------------------------------- FlexConv
def _flexconv_case(rng, B, N, K, Din, Dout, extent=25.0):
    pts = make_cloud(rng, B, N, extent=extent)
    nb, _ = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), K)
    f = rng.randn(B, N, Din).astype(np.float32)
    th = (rng.randn(3, Din, Dout) / np.sqrt(Din)).astype(np.float32)
    bi = (rng.randn(Din, Dout) / np.sqrt(Din)).astype(np.float32)
    return pts, nb, f, th, bi


def test_flex_conv_reference_fixture_odd_dims():
    """The reference's own fixture (Din=2, Dout=6, K=4, N=32): fp32 vs fp64 at rtol 1e-4
    (user_ops/test_flex_convolution.py:42-50)."""
    from dh3d_b200 import user_ops
    case, _ = fixtures.reference_test_cases()
    f32 = lambda a: cu(a.astype(np.float32))
    out = user_ops.flex_convolution(f32(case.features), f32(case.position), cu(case.neighborhood),
                                    f32(case.theta), f32(case.bias))
    exp = oracle.flex_convolution(case.features, case.position, case.neighborhood, case.theta, case.bias,
                                  f64=True)
    assert out.shape == (2, 6, 32)
    close(out, exp)


@pytest.mark.parametrize("B,N,K,Din,Dout", [(2, 1024, 8, 32, 64), (1, 8192, 8, 64, 64), (2, 1024, 8, 64, 128),
                                            (2, 1024, 8, 128, 128), (1, 1024, 8, 128, 256), (1, 2048, 16, 128, 128),
                                            (1, 512, 32, 128, 128), (1, 300, 5, 8, 12), (3, 333, 8, 32, 64),
                                            (2, 777, 11, 64, 72), (3, 333, 8, 64, 64), (2, 777, 8, 128, 200),
                                            (5, 8192, 8, 64, 64), (3, 8192, 8, 64, 128)])
def test_flex_conv_pm_vs_fp64_truth(B, N, K, Din, Dout):
    from dh3d_b200 import ops
    rng = np.random.RandomState(N + Din + Dout + K)
    pts, nb, f, th, bi = _flexconv_case(rng, B, N, K, Din, Dout)
    exp = oracle.flex_convolution(f.transpose(0, 2, 1), pts.transpose(0, 2, 1), nb.transpose(0, 2, 1), th, bi,
                                  f64=True).transpose(0, 2, 1)
    out = ops.flex_conv(cu(f), cu(th), cu(bi), cu(nb), cu(pts))
    close(out, exp)
    # the fp32 reference-order restatement is itself this close to the truth (sanity of the bar)
    if N <= 2048:
        ref32 = oracle.flex_convolution(f.transpose(0, 2, 1), pts.transpose(0, 2, 1), nb.transpose(0, 2, 1),
                                        th, bi).transpose(0, 2, 1)
        close(ref32, exp)


def test_flex_conv_fused_epilogue_and_cm_entry():
    from dh3d_b200 import ops, user_ops
    rng = np.random.RandomState(21)
    pts, nb, f, th, bi = _flexconv_case(rng, 2, 1024, 8, 64, 128)
    exp = oracle.flex_convolution(f.transpose(0, 2, 1), pts.transpose(0, 2, 1), nb.transpose(0, 2, 1), th, bi,
                                  f64=True)
    fb = rng.randn(128).astype(np.float32)
    sc, sh = (rng.rand(128) + 0.5).astype(np.float32), rng.randn(128).astype(np.float32)
    out = ops.flex_conv(cu(f), cu(th), cu(bi), cu(nb), cu(pts), feature_bias=cu(fb), scale=cu(sc),
                        shift=cu(sh), act=1)
    close(out, np.maximum((exp.transpose(0, 2, 1) + fb) * sc + sh, 0))
    out_cm = user_ops.flex_convolution(cu(f.transpose(0, 2, 1)), cu(pts.transpose(0, 2, 1)),
                                       cu(nb.transpose(0, 2, 1)), cu(th), cu(bi))
    close(out_cm, exp)


def test_flex_conv_prepacked_is_bit_identical_to_the_per_call_form():
    # dh3d_flex_conv_prepack + dh3d_flex_conv_pm_packed == dh3d_flex_conv_pm (same kernels, weights prepared once)
    from dh3d_b200 import ops
    rng = np.random.RandomState(23)
    for (B, N, K, Din, Dout) in [(2, 1024, 8, 64, 128), (1, 777, 5, 32, 64), (1, 300, 8, 8, 12)]:
        pts, nb, f, th, bi = _flexconv_case(rng, B, N, K, Din, Dout)
        fb = rng.randn(Dout).astype(np.float32)
        sc, sh = (rng.rand(Dout) + 0.5).astype(np.float32), rng.randn(Dout).astype(np.float32)
        a = ops.flex_conv(cu(f), cu(th), cu(bi), cu(nb), cu(pts), feature_bias=cu(fb), scale=cu(sc), shift=cu(sh), act=1)
        packed = ops.flex_conv_prepack(cu(th), cu(bi), cu(fb), cu(sc), cu(sh))
        b = ops.flex_conv_packed(cu(f), packed, cu(nb), cu(pts), scale=cu(sc), act=1)
        assert torch.equal(a, b)
        a0 = ops.flex_conv(cu(f), cu(th), cu(bi), cu(nb), cu(pts))
        b0 = ops.flex_conv_packed(cu(f), ops.flex_conv_prepack(cu(th), cu(bi)), cu(nb), cu(pts))
        assert torch.equal(a0, b0)


def test_flex_conv_centre_is_the_point_itself_on_duplicates():
    """Duplicated points: nbr(0,n) may differ from n; the CUDA reference centres on p[n]
    (flex_conv_kernel_gpu.cu.cc:75-79), and p[nbr0] == p[n] anyway for exact duplicates."""
    from dh3d_b200 import ops
    rng = np.random.RandomState(22)
    pts = make_cloud(rng, 1, 1000, duplicates=300)
    nb, _ = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), 8)
    f = rng.randn(1, 1000, 16).astype(np.float32)
    th, bi = rng.randn(3, 16, 8).astype(np.float32), rng.randn(16, 8).astype(np.float32)
    exp = oracle.flex_convolution(f.transpose(0, 2, 1), pts.transpose(0, 2, 1), nb.transpose(0, 2, 1), th, bi,
                                  f64=True)
    close(ops.flex_conv(cu(f), cu(th), cu(bi), cu(nb), cu(pts)), exp.transpose(0, 2, 1))


# ------------------------------------------------------------------------- 3-NN / interpolation
@pytest.mark.parametrize("B,n,m", [(2, 8192, 1024), (1, 1000, 3), (2, 333, 2), (1, 2048, 2500)])
def test_three_nn_bitexact(B, n, m):
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(n + m)
    a, b = make_cloud(rng, B, n), make_cloud(rng, B, m)
    if m >= 100:
        b[:, 50:60] = b[:, 40:50]  # duplicated known points -> strict '<' keeps the earlier index
        h = min(n, m) // 2
        a[:, :h] = b[:, :h]  # exact hits -> dist 0
    from dh3d_b200 import ops
    ed, ei = oracle.three_nn(a, b)
    for dist, idx in (tf_ops.three_nn(cu(a), cu(b)), ops.three_nn(cu(a), cu(b), exhaustive=True)):
        assert np.array_equal(idx.cpu().numpy(), ei)
        assert np.array_equal(dist.cpu().numpy(), ed)


def test_three_interpolate_bitexact_and_fused_weights():
    from dh3d_b200 import ops, tf_ops
    rng = np.random.RandomState(31)
    for C in (128, 256, 5, 12, 192):
        pts = rng.randn(2, 100, C).astype(np.float32)
        idx = rng.randint(0, 100, (2, 999, 3)).astype(np.int32)
        dist = (rng.rand(2, 999, 3) * 4).astype(np.float32)
        dist[0, :10] = 0
        w = oracle.three_nn_weights(dist)
        out = tf_ops.three_interpolate(cu(pts), cu(idx), cu(w))
        assert np.array_equal(out.cpu().numpy(), oracle.three_interpolate(pts, idx, w))
        fused = ops.three_interpolate(cu(pts), cu(idx), cu(dist), weight_is_dist2=True)
        close(fused, oracle.three_interpolate(pts, idx, w), rel=1e-5)


# ------------------------------------------------------------------------------------ ball query
@pytest.mark.parametrize("B,n,m,r,ns", [(2, 400, 300, 0.8, 16), (1, 3000, 1000, 0.3, 32), (2, 512, 128, 0.05, 8),
                                        (1, 100, 600, 10.0, 4)])
def test_query_ball_point_bitexact(B, n, m, r, ns):
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(n + m)
    xyz1, xyz2 = make_cloud(rng, B, n, extent=2.0), make_cloud(rng, B, m, extent=2.0)
    idx, cnt = tf_ops.query_ball_point(r, ns, cu(xyz1), cu(xyz2))
    eidx, ecnt = oracle.query_ball_point(r, ns, xyz1, xyz2)
    assert np.array_equal(cnt.cpu().numpy(), ecnt)
    assert np.array_equal(idx.cpu().numpy(), eidx)


def test_query_ball_point_leaked_nearest_quirk():
    from dh3d_b200 import tf_ops
    x1 = np.zeros((1, 10, 3), np.float32); x1[0, :, 0] = np.arange(10) * 10.0
    x2 = np.full((1, 257, 3), 1000.0, np.float32)
    x2[0, 0] = x1[0, 5]; x2[0, 0, 1] = 3.0
    x2[0, 256] = (95.0, 50.0, 0.0)
    idx, cnt = tf_ops.query_ball_point(1.0, 4, cu(x1), cu(x2))
    eidx, ecnt = oracle.query_ball_point(1.0, 4, x1, x2)
    assert np.array_equal(idx.cpu().numpy(), eidx) and np.array_equal(cnt.cpu().numpy(), ecnt)
    assert torch.all(idx[0, 256] == 5)


# --------------------------------------------------------------------------------- dense + NetVLAD
@pytest.mark.parametrize("M,K,N,act", [(1000, 64, 64, 1), (4096, 192, 128, 1), (777, 256, 1024, 1), (512, 64, 16, 1),
                                       (512, 16, 64, 2), (129, 128, 128, 0)])
def test_linear_vs_fp64(M, K, N, act):
    from dh3d_b200 import ops
    rng = np.random.RandomState(M + K + N)
    x = rng.randn(M, K).astype(np.float32)
    w = (rng.randn(K, N) / np.sqrt(K)).astype(np.float32)
    sc, sh = (rng.rand(N) + 0.5).astype(np.float32), rng.randn(N).astype(np.float32)
    y = x.astype(np.float64) @ w.astype(np.float64) * sc + sh
    y = np.maximum(y, 0) if act == 1 else (1 / (1 + np.exp(-y)) if act == 2 else y)
    close(ops.linear(cu(x), cu(w), scale=cu(sc), shift=cu(sh), act=act), y)


def test_small_glue_ops():
    from dh3d_b200 import ops
    rng = np.random.RandomState(41)
    x, g = rng.randn(3, 100, 64).astype(np.float32), rng.rand(3, 100, 64).astype(np.float32)
    close(ops.se_excite(cu(x), cu(g)), np.maximum(x + x * g, 0), rel=1e-6)
    close(ops.add(cu(x), cu(g)), x + g, rel=1e-6)
    n = ops.l2_normalize_rows(cu(x), 1e-8)
    close(n, x / np.sqrt(np.maximum((x.astype(np.float64) ** 2).sum(-1, keepdims=True), 1e-8)), rel=1e-5)
    w = rng.randn(64).astype(np.float32)
    close(ops.rowdot(cu(x), cu(w), bias=0.125, act=2), 1 / (1 + np.exp(-(x.astype(np.float64) @ w + 0.125))), rel=1e-5)
    cat = torch.zeros(3, 100, 128, device="cuda")
    ops.copy_cols(cu(x), cat, 0); ops.copy_cols(cu(g), cat, 64)
    assert np.array_equal(cat.cpu().numpy(), np.concatenate([x, g], -1))
    t = rng.randn(2, 5, 77).astype(np.float32)
    assert np.array_equal(ops.transpose_cm_to_pm(cu(t)).cpu().numpy(), t.transpose(0, 2, 1))
    assert np.array_equal(ops.transpose_pm_to_cm(cu(t)).cpu().numpy(), t.transpose(0, 2, 1))


@pytest.mark.parametrize("B,N,K,C", [(2, 1000, 8, 64), (3, 333, 8, 64), (1, 77, 5, 64), (2, 500, 8, 128),
                                     (1, 129, 3, 128), (2, 8192, 8, 64)])
def test_se_pool_excite_fused_vs_fp64_and_unfused(B, N, K, C):
    # se_res_bottleneck (core/backbones.py:45-55): fused launch vs the fp64 composition and vs the separate ops
    from dh3d_b200 import ops
    rng = np.random.RandomState(B * 7 + N + K + C)
    H = C // 4
    x = rng.randn(B, N, C).astype(np.float32)
    nbr = rng.randint(0, N, (B, N, K)).astype(np.int32)
    w1, b1 = (rng.randn(C, H) / np.sqrt(C)).astype(np.float32), rng.randn(H).astype(np.float32) * 0.1
    w2, b2 = (rng.randn(H, C) / np.sqrt(H)).astype(np.float32), rng.randn(C).astype(np.float32) * 0.1
    b1, b2 = b1.astype(np.float32), b2.astype(np.float32)
    pooled = np.stack([x[b][nbr[b]].max(axis=1) for b in range(B)]).astype(np.float64)
    hid = np.maximum(pooled @ w1.astype(np.float64) + b1, 0)
    gate = 1 / (1 + np.exp(-(hid @ w2.astype(np.float64) + b2)))
    want = np.maximum(x + x * gate, 0)
    got = ops.se_pool_excite(cu(x), cu(nbr), cu(w1), cu(b1), cu(w2), cu(b2))
    close(got, want, rel=2e-5)
    mx = ops.flex_pool(cu(x), cu(nbr))
    mx = mx[0] if isinstance(mx, tuple) else mx
    assert np.array_equal(mx.cpu().numpy(), pooled.astype(np.float32))
    unf = ops.se_excite(cu(x), ops.linear(ops.linear(mx, cu(w1), shift=cu(b1), act=1), cu(w2), shift=cu(b2), act=2))
    close(got, unf.cpu().numpy(), rel=2e-5)


@pytest.mark.parametrize("B,N", [(2, 8192), (3, 1000), (1, 37)])
def test_netvlad_vs_fp64(B, N):
    from dh3d_b200.backbones import GlobalNetVLADBlock
    from dh3d_b200.model import init_random_
    from oracle import net
    blk = init_random_(GlobalNetVLADBlock(), seed=5)
    p = {"netvlad." + k: v.detach().numpy() for k, v in blk.named_parameters()}
    blk = blk.cuda()
    rng = np.random.RandomState(N)
    feat = np.maximum(rng.randn(B, N, 256), 0).astype(np.float32) * 3
    att = rng.rand(B, N, 1).astype(np.float32)
    out = blk(None, cu(feat), cu(att), final_l2norm=True)
    close(out, net.netvlad(feat, att, p, final_l2norm=True))
    raw = blk(None, cu(feat), cu(att), final_l2norm=False)
    close(raw, net.netvlad(feat, att, p, final_l2norm=False))


def test_netvlad_tiny_and_huge_feature_norms():
    """Rows are l2-normalised before the fp16 split: feature magnitude must not matter (1e-6 .. 1e4), and an
    all-zero row contributes the reference's zero vector (epsilon 1e-12 on the squared norm)."""
    from dh3d_b200.backbones import GlobalNetVLADBlock
    from dh3d_b200.model import init_random_
    from oracle import net
    blk = init_random_(GlobalNetVLADBlock(), seed=9)
    p = {"netvlad." + k: v.detach().numpy() for k, v in blk.named_parameters()}
    blk = blk.cuda()
    rng = np.random.RandomState(77)
    feat = rng.randn(2, 1000, 256).astype(np.float32)
    feat *= (10.0 ** rng.uniform(-6, 4, (2, 1000, 1))).astype(np.float32)
    feat[0, 5] = 0
    att = rng.rand(2, 1000, 1).astype(np.float32)
    out = blk(None, cu(feat), cu(att), final_l2norm=True)
    close(out, net.netvlad(feat, att, p, final_l2norm=True))


def test_netvlad_more_than_32_clouds():
    """B > 32: two projection groups in the tail (phase-B items of both groups wait for every phase-A item)."""
    from dh3d_b200.backbones import GlobalNetVLADBlock
    from dh3d_b200.model import init_random_
    from oracle import net
    blk = init_random_(GlobalNetVLADBlock(), seed=11)
    p = {"netvlad." + k: v.detach().numpy() for k, v in blk.named_parameters()}
    blk = blk.cuda()
    rng = np.random.RandomState(40)
    feat = np.maximum(rng.randn(40, 96, 256), 0).astype(np.float32)
    att = rng.rand(40, 96, 1).astype(np.float32)
    close(blk(None, cu(feat), cu(att), final_l2norm=True), net.netvlad(feat, att, p, final_l2norm=True))


@pytest.mark.timeout(120)
def test_netvlad_calls_on_concurrent_streams_complete_and_agree():
    """The tail's three phases hand over through counters inside one launch.  CTAs take tickets and only wait for lower
    tickets, so any number of calls may overlap on other streams (a spinning grid barrier would deadlock here: four
    544-CTA launches over-subscribe the machine's resident slots).  Results: bit-identical to the serial calls."""
    from dh3d_b200.backbones import GlobalNetVLADBlock
    from dh3d_b200.model import init_random_
    blk = init_random_(GlobalNetVLADBlock(), seed=13).cuda()
    rng = np.random.RandomState(3)
    feats = [cu(np.maximum(rng.randn(32, 256, 256), 0).astype(np.float32)) for _ in range(4)]
    atts = [cu(rng.rand(32, 256, 1).astype(np.float32)) for _ in range(4)]
    want = [blk(None, f, a, final_l2norm=True).clone() for f, a in zip(feats, atts)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    got = [[] for _ in range(4)]
    for rep in range(25):
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                got[k].append(blk(None, feats[k], atts[k], final_l2norm=True))
    torch.cuda.synchronize()
    for k in range(4):
        for g in got[k]:
            assert torch.equal(g, want[k])


# ------------------------------------------------------------------------ fused concat / add+l2norm
def test_strided_group_interp_and_add_l2norm_are_bit_identical_to_the_unfused_ops():
    from dh3d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    B, N, M = 2, 3000, 256
    wide = torch.randn((B, N, 192), device="cuda", generator=g)
    idx = torch.randint(0, N, (B, M, 1), device="cuda", generator=g, dtype=torch.int32)
    a = ops.group_point_cols(wide, 128, 64, idx)
    b = ops.group_point(wide[:, :, 128:].contiguous(), idx)
    assert torch.equal(a, b)
    known = torch.randn((B, M, 128), device="cuda", generator=g)
    i3 = torch.randint(0, M, (B, N, 3), device="cuda", generator=g, dtype=torch.int32)
    d3 = torch.rand((B, N, 3), device="cuda", generator=g)
    before = wide.clone()
    ops.three_interpolate(known, i3, d3, weight_is_dist2=True, out=wide, out_col=0)
    assert torch.equal(wide[:, :, :128], ops.three_interpolate(known, i3, d3, weight_is_dist2=True))
    assert torch.equal(wide[:, :, 128:], before[:, :, 128:])          # the other column block is untouched
    x, y = torch.randn((B, N, 128), device="cuda", generator=g), torch.randn((B, N, 128), device="cuda", generator=g)
    s, nrm = ops.add_l2_normalize_rows(x, y, 1e-8)
    assert torch.equal(s, ops.add(x, y))
    assert torch.equal(nrm, ops.l2_normalize_rows(ops.add(x, y), 1e-8))


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,m", [(2, 8192, 1024), (3, 1000, 77), (1, 4097, 512)])
def test_three_nn_presorted_queries_bit_exact(B, n, m):
    """dh3d_three_nn_ws_presorted: the query cloud's sort is taken from the workspace a k-NN call on the same points left
    behind (what DH3D.forward does); distances and indices identical to the exhaustive scan and to the self-sorting form."""
    from dh3d_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + n)
    xyz1 = (torch.rand((B, n, 3), device="cuda", generator=g) * 40 - 20).contiguous()
    xyz1[0, : n // 8] = xyz1[0, n // 8: 2 * (n // 8)]          # duplicated points (distance ties)
    xyz2 = xyz1[:, torch.randperm(n, device="cuda", generator=g)[:m]].contiguous()
    _, _, ws = ops.knn_points(xyz1, 8, keep_workspace=True)
    d0, i0 = ops.three_nn(xyz1, xyz2, exhaustive=True)
    d1, i1 = ops.three_nn(xyz1, xyz2)
    d2, i2 = ops.three_nn(xyz1, xyz2, sorted1=ws)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)
    assert torch.equal(i0, i2) and torch.equal(d0, d2)
    # both clouds from k-NN workspaces (the candidates' boxes carry the smallest index next to the k-NN tie rank)
    _, _, ws2 = ops.knn_points(xyz2, min(8, m), keep_workspace=True)
    d3, i3 = ops.three_nn(xyz1, xyz2, sorted1=ws, sorted2=ws2)
    assert torch.equal(i0, i3) and torch.equal(d0, d3)
    z1, z2 = torch.zeros_like(xyz1), torch.zeros_like(xyz2)          # all-zero padding clouds: every distance ties
    dz0, iz0 = ops.three_nn(z1, z2, exhaustive=True)
    dz3, iz3 = ops.three_nn(z1, z2, sorted1=ops.knn_sort(z1), sorted2=ops.knn_sort(z2))
    assert torch.equal(iz0, iz3) and torch.equal(dz0, dz3)
    with pytest.raises(Exception):
        ops.three_nn(xyz1[:, :-1].contiguous(), xyz2, sorted1=ws)      # workspace of another shape


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,M", [(2, 8192, 1024), (3, 1000, 300), (1, 4097, 512), (2, 700, 700), (4, 8192, 64), (1, 33, 33)])
def test_fps_presorted_bit_exact(B, N, M):
    """dh3d_farthest_point_sample_presorted (box-pruned rounds on the k-NN engine's cell-sorted copy of the cloud) selects
    the same points in the same order as the exhaustive kernel and as the oracle's restatement of the reference."""
    from dh3d_b200 import ops
    rng = np.random.RandomState(N + M)
    pts = make_cloud(rng, B, N, duplicates=N // 5)
    d = cu(pts)
    ws = ops.knn_sort(d)
    a = ops.farthest_point_sample(M, d)
    b = ops.farthest_point_sample(M, d, sorted_ws=ws)
    assert torch.equal(a, b)
    if N * M <= 8192 * 300:
        assert np.array_equal(b.cpu().numpy(), oracle.farthest_point_sample(M, pts))


@pytest.mark.gpu
def test_fps_presorted_ties_degenerate_and_outlier_clouds():
    """Lattice clouds (masses of equal distances: the (k mod 512, k) tie rule decides), an all-zero padding cloud, a cloud
    with 1/8 of its points at 1e5 m (the reference's randsample=False padding) and the 1100-point tie case."""
    from dh3d_b200 import ops
    rng = np.random.RandomState(14)
    pts = lattice_cloud(rng, 4, 8192, step=1.0, side=6)
    pts[1] = 0.0
    pts[2] = make_cloud(rng, 1, 8192)[0]
    pts[2, ::8] = 1.0e5
    pts[3] = make_cloud(rng, 1, 8192, duplicates=6000)[0]
    d = cu(pts)
    for M in (2, 512, 1024):
        a = ops.farthest_point_sample(M, d)
        b = ops.farthest_point_sample(M, d, sorted_ws=ops.knn_sort(d))
        assert torch.equal(a, b), M
    tie = np.zeros((1, 1100, 3), np.float32)
    tie[0, 1:] = 1.0
    t = cu(tie)
    assert ops.farthest_point_sample(2, t, sorted_ws=ops.knn_sort(t))[0, 1].item() == 512
    with pytest.raises(Exception):
        ops.farthest_point_sample(8, d[:, :100].contiguous(), sorted_ws=ops.knn_sort(d))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,K", [(2, 8192, 8), (3, 1000, 16), (1, 300, 50)])
def test_knn_sort_plus_query_equals_knn_points(B, N, K):
    from dh3d_b200 import ops
    rng = np.random.RandomState(N + K)
    d = cu(make_cloud(rng, B, N, duplicates=N // 7))
    i0, d0 = ops.knn_points(d, K)
    i1, d1 = ops.knn_query_sorted(ops.knn_sort(d), K)
    assert torch.equal(i0, i1) and torch.equal(d0, d1)
