"""CPU tests: the oracle against the REFERENCE'S OWN CPU code (oracle/_ref/libdh3d_ref_cpu.so,
compiled unmodified from /root/reference by oracle/build_ref.py).  Skipped where that build is
absent."""
import numpy as np
import pytest

import oracle
from oracle import fixtures, ref
from conftest import make_cloud

pytestmark = pytest.mark.skipif(not ref.have_cpu(), reason="oracle/_ref/libdh3d_ref_cpu.so not built")


def test_three_nn_bitexact_vs_reference_cpu():
    rng = np.random.RandomState(1)
    for (B, n, m) in ((2, 500, 64), (1, 2000, 1024), (3, 17, 2)):
        a, b = make_cloud(rng, B, n), make_cloud(rng, B, m)
        if m > 20:
            b[:, 10:20] = b[:, 0:10]
            a[:, :10] = b[:, :10]
        rd, ri = ref.cpu_three_nn(a, b)
        od, oi = oracle.three_nn(a, b)
        assert np.array_equal(ri, oi) and np.array_equal(rd, od)


def test_three_interpolate_bitexact_vs_reference_cpu():
    rng = np.random.RandomState(2)
    pts = rng.randn(2, 64, 128).astype(np.float32)
    idx = rng.randint(0, 64, (2, 700, 3)).astype(np.int32)
    w = oracle.three_nn_weights(rng.rand(2, 700, 3).astype(np.float32))
    assert np.array_equal(ref.cpu_three_interpolate(pts, idx, w), oracle.three_interpolate(pts, idx, w))


def test_flex_ops_vs_reference_cpu_functors():
    case, _ = fixtures.reference_test_cases()
    # FlexConv: the CPU functor centres on nbr(0,n) and adds bias first; same values to fp32 rounding
    r = ref.cpu_flex_conv(case.features, case.position, case.neighborhood, case.theta, case.bias)
    o = oracle.flex_convolution(case.features, case.position, case.neighborhood, case.theta, case.bias,
                                centre_is_self=False, f64=True)
    assert np.allclose(r, o, rtol=1e-4, atol=1e-5)
    ro, ra = ref.cpu_flex_pool(case.features, case.neighborhood)
    oo, oa = oracle.flex_pooling(case.features, case.neighborhood)
    assert np.array_equal(ro, oo) and np.array_equal(ra, oa)
    r = ref.cpu_conv_pointset(case.features, case.neighborhood, case.theta_rel, case.bias_rel)
    o = oracle.convolution_pointset(case.features, case.neighborhood, case.theta_rel, case.bias_rel)
    assert np.allclose(r, o, rtol=1e-5, atol=1e-6)
    # a DH3D-sized layer: Din=32 -> Dout=64 on 1024 points, random neighbourhoods incl. non-self first
    rng = np.random.RandomState(3)
    f = rng.randn(1, 32, 1024).astype(np.float32)
    p = rng.randn(1, 3, 1024).astype(np.float32)
    nb = rng.randint(0, 1024, (1, 8, 1024)).astype(np.int32)
    th, bi = (rng.randn(3, 32, 64) / 6).astype(np.float32), (rng.randn(32, 64) / 6).astype(np.float32)
    r = ref.cpu_flex_conv(f, p, nb, th, bi)
    o = oracle.flex_convolution(f, p, nb, th, bi, centre_is_self=False, f64=True)
    assert np.abs(r - o).max() <= 1e-4 * np.sqrt((o ** 2).mean())
    x, n = fixtures.flexpool_four_point_case()
    ro, ra = ref.cpu_flex_pool(x, n)
    assert np.all(ro == 5) and np.all(ra == 2)


# ---- backward passes / FlexDeconv: oracle (fp64 loops) vs the reference's own CPU Grad functors ------
def _rel(a, b):
    b = np.asarray(b, np.float64)
    return np.abs(np.asarray(a, np.float64) - b).max() / max(np.sqrt((b ** 2).mean()), 1e-30)


def _grad_case(rng, B, N, K, Din, Dout, self_first=True):
    f = rng.randn(B, Din, N).astype(np.float32)
    p = rng.randn(B, 3, N).astype(np.float32)
    nb = rng.randint(0, N, (B, K, N)).astype(np.int32)
    if self_first:
        nb[:, 0, :] = np.arange(N, dtype=np.int32)[None]
    th = (rng.randn(3, Din, Dout) / 4).astype(np.float32)
    bi = (rng.randn(Din, Dout) / 4).astype(np.float32)
    top = rng.randn(B, Dout, N).astype(np.float32)
    return f, p, nb, th, bi, top


@pytest.mark.parametrize("B,N,K,Din,Dout,self_first", [(2, 32, 4, 2, 6, True), (1, 300, 8, 16, 24, False),
                                                       (2, 257, 5, 7, 3, True)])
def test_flex_conv_grad_vs_reference_cpu(B, N, K, Din, Dout, self_first):
    f, p, nb, th, bi, top = _grad_case(np.random.RandomState(B * 100 + N), B, N, K, Din, Dout, self_first)
    rf, rt, rb = ref.cpu_flex_conv_grad(f, th, bi, nb, p, top)
    of, ot, ob = oracle.flex_convolution_grad(f, th, bi, nb, p, top)
    assert _rel(rf, of) < 2e-5 and _rel(rt, ot) < 2e-5 and _rel(rb, ob) < 2e-5


def test_flex_pool_grad_vs_reference_cpu():
    rng = np.random.RandomState(7)
    f, _, nb, _, _, _ = _grad_case(rng, 2, 200, 6, 9, 9)
    _, arg = oracle.flex_pooling(f, nb)
    top = rng.randn(2, 9, 200).astype(np.float32)
    r = ref.cpu_flex_pool_grad(f, nb, top, arg)
    assert _rel(r, oracle.flex_pooling_grad(top, arg)) < 1e-6
    # the reference's 4-point case (user_ops/test_flex_pooling.py:76-98): all gradient lands on point 2
    x, n = fixtures.flexpool_four_point_case()
    _, a = oracle.flex_pooling(x, n)
    g = oracle.flex_pooling_grad(np.ones_like(x), a)
    assert np.all(g[:, :, 2] == x.shape[2]) and g.sum() == x.size


@pytest.mark.parametrize("B,N,K,Din,Dout", [(2, 128, 8, 3, 32), (1, 77, 4, 5, 6)])
def test_conv_pointset_grad_vs_reference_cpu(B, N, K, Din, Dout):
    rng = np.random.RandomState(N)
    f, _, nb, _, _, _ = _grad_case(rng, B, N, K, Din, Dout, self_first=False)
    th = (rng.randn(Din, Dout) / 3).astype(np.float32)
    bi = rng.randn(Dout).astype(np.float32)
    top = rng.randn(B, Dout, N).astype(np.float32)
    rf, rt, rb = ref.cpu_conv_pointset_grad(f, th, bi, nb, top)
    of, ot, ob = oracle.convolution_pointset_grad(f, th, nb, top)
    assert _rel(rf, of) < 2e-5 and _rel(rt, ot) < 2e-5 and _rel(rb, ob) < 2e-5


def test_flex_deconv_vs_reference_cpu():
    f, p, nb, th, bi, _ = _grad_case(np.random.RandomState(11), 2, 150, 6, 8, 12, self_first=False)
    r = ref.cpu_flex_deconv(f, p, nb, th, bi)
    assert _rel(r, oracle.flex_convolution_transpose(f, p, nb, th, bi)) < 2e-5


def test_three_interpolate_grad_vs_reference_cpu():
    rng = np.random.RandomState(12)
    g = rng.randn(2, 300, 16).astype(np.float32)
    idx = rng.randint(0, 40, (2, 300, 3)).astype(np.int32)
    w = oracle.three_nn_weights(rng.rand(2, 300, 3).astype(np.float32))
    assert _rel(ref.cpu_three_interpolate_grad(40, g, idx, w), oracle.three_interpolate_grad(40, g, idx, w)) < 2e-6


def test_backward_is_adjoint_of_forward():
    """<J x, y> == <x, J^T y>: ties the backward restatements to the (separately pinned) forward oracle."""
    rng = np.random.RandomState(13)
    f, p, nb, th, bi, top = _grad_case(rng, 1, 64, 4, 5, 7, self_first=True)   # nbr(0,n) == n: both centre rules agree
    out = oracle.flex_convolution(f, p, nb, th, bi, centre_is_self=False, f64=True)
    gf, gt, gb = oracle.flex_convolution_grad(f, th, bi, nb, p, top)
    lhs = float((out * top).sum())
    assert abs(lhs - float((gf * f).sum())) < 1e-6 * abs(lhs)          # linear in f
    assert abs(lhs - float((gt * th).sum() + (gb * bi).sum())) < 1e-6 * abs(lhs)   # and in (theta, bias)
    pts = rng.randn(2, 30, 6).astype(np.float32)
    idx = rng.randint(0, 30, (2, 9, 4)).astype(np.int32)
    go = rng.randn(2, 9, 4, 6).astype(np.float32)
    lhs = float((oracle.group_point(pts, idx).astype(np.float64) * go).sum())
    assert abs(lhs - float((oracle.group_point_grad(30, go, idx) * pts).sum())) < 1e-6 * abs(lhs)
