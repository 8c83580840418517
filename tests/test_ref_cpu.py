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
