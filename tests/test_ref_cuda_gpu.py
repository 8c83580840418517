"""GPU tests against the REFERENCE'S OWN CUDA kernels (oracle/_ref/libdh3d_ref_cuda.so: the
unmodified reference sources compiled for sm_100a, run on this GPU).  Three-way: reference CUDA
== CPU oracle == this repo's kernels, bit for bit on every index output -- this is what pins the
oracle's tie rules (CUB blocked-order k-NN, the 512-thread FPS rule, the ball-query nearest leak)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref
from conftest import lattice_cloud, make_cloud

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.have_cuda(), reason="oracle/_ref/libdh3d_ref_cuda.so not built")]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("B,N,K", [(1, 4, 4), (2, 32, 4), (2, 100, 8), (2, 300, 8), (1, 1024, 8), (2, 2000, 8),
                                   (1, 4096, 16), (2, 8192, 8), (1, 8192, 32)])
def test_knn_three_way(B, N, K):
    from dh3d_b200 import user_ops
    rng = np.random.RandomState(N + K)
    pts = make_cloud(rng, B, N)
    if N >= 100:
        pts[0] = lattice_cloud(rng, 1, N)[0]                      # ties everywhere
        pts[-1, N - N // 4:] = pts[-1, :N // 4]                    # duplicated points
    pos = np.ascontiguousarray(pts.transpose(0, 2, 1))
    rid, rd = ref.cuda_knn(cu(pos), K)
    oid, od = oracle.knn_bruteforce(pos, K)
    mid, md = user_ops.knn_bruteforce(cu(pos), K)
    assert np.array_equal(rid.cpu().numpy(), oid), "oracle tie order != reference CUDA"
    assert np.array_equal(rd.cpu().numpy(), od)
    assert torch.equal(mid, rid) and torch.equal(md, rd)


@pytest.mark.parametrize("B,N,M", [(2, 8192, 1024), (3, 1024, 128), (2, 700, 70), (1, 5000, 333), (40, 2048, 64)])
def test_fps_three_way(B, N, M):
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(N + M)
    pts = make_cloud(rng, B, N)
    pts[0] = lattice_cloud(rng, 1, N, step=1.0, side=5)[0]
    if B > 1:
        pts[1, N // 2:] = pts[1, :N - N // 2]
    r = ref.cuda_fps(M, cu(pts))
    assert np.array_equal(r.cpu().numpy(), oracle.farthest_point_sample(M, pts)), "oracle FPS rule != reference"
    assert torch.equal(tf_ops.farthest_point_sample(M, cu(pts)), r)


def test_group_gather_three_way():
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(5)
    pts = rng.randn(3, 400, 64).astype(np.float32)
    idx = rng.randint(0, 400, (3, 50, 4)).astype(np.int32)
    r = ref.cuda_group_point(cu(pts), cu(idx))
    assert np.array_equal(r.cpu().numpy(), oracle.group_point(pts, idx))
    assert torch.equal(tf_ops.group_point(cu(pts), cu(idx)), r)
    xyz = rng.randn(3, 400, 3).astype(np.float32)
    gi = rng.randint(0, 400, (3, 77)).astype(np.int32)
    r = ref.cuda_gather_point(cu(xyz), cu(gi))
    assert torch.equal(tf_ops.gather_point(cu(xyz), cu(gi)), r)


@pytest.mark.parametrize("B,n,m,r,ns", [(2, 400, 300, 0.8, 16), (1, 3000, 1000, 0.3, 32), (2, 512, 700, 0.05, 8)])
def test_query_ball_point_three_way(B, n, m, r, ns):
    from dh3d_b200 import tf_ops
    rng = np.random.RandomState(n + m)
    xyz1, xyz2 = make_cloud(rng, B, n, extent=2.0), make_cloud(rng, B, m, extent=2.0)
    ri, rc = ref.cuda_query_ball_point(r, ns, cu(xyz1), cu(xyz2))
    oi, oc = oracle.query_ball_point(r, ns, xyz1, xyz2)
    assert np.array_equal(rc.cpu().numpy(), oc) and np.array_equal(ri.cpu().numpy(), oi)
    mi, mc = tf_ops.query_ball_point(r, ns, cu(xyz1), cu(xyz2))
    assert torch.equal(mi, ri) and torch.equal(mc, rc)


def test_flex_ops_vs_reference_cuda():
    from dh3d_b200 import user_ops
    rng = np.random.RandomState(7)
    for (B, N, K, Din, Dout) in ((2, 1024, 8, 32, 64), (1, 8192, 8, 64, 64), (1, 1024, 8, 128, 256)):
        pts = make_cloud(rng, B, N)
        nb, _ = oracle.knn_bruteforce(np.ascontiguousarray(pts.transpose(0, 2, 1)), K)
        f = rng.randn(B, Din, N).astype(np.float32)
        th = (rng.randn(3, Din, Dout) / np.sqrt(Din)).astype(np.float32)
        bi = (rng.randn(Din, Dout) / np.sqrt(Din)).astype(np.float32)
        pos, nbc = cu(pts.transpose(0, 2, 1)), cu(nb.transpose(0, 2, 1))
        r = ref.cuda_flex_conv(cu(f), pos, nbc, cu(th), cu(bi))
        mine = user_ops.flex_convolution(cu(f), pos, nbc, cu(th), cu(bi))
        scale = r.pow(2).mean().sqrt()
        assert (mine - r).abs().max() <= 1e-4 * scale, "FlexConv vs reference CUDA beyond 1e-4"
        if N <= 1024:   # the oracle's fp32 restatement follows the reference kernel's FMA order
            o = oracle.flex_convolution(f, pts.transpose(0, 2, 1), nb.transpose(0, 2, 1), th, bi)
            assert np.abs(o - r.cpu().numpy()).max() <= 1e-5 * float(scale)
        ro, ra = ref.cuda_flex_pool(cu(f), nbc)
        mo, ma = user_ops.flex_pooling(cu(f), nbc)
        assert torch.equal(mo, ro) and torch.equal(ma, ra)
    p3 = make_cloud(rng, 2, 4096)
    nb, _ = oracle.knn_bruteforce(np.ascontiguousarray(p3.transpose(0, 2, 1)), 8)
    th, bi = rng.randn(3, 32).astype(np.float32), rng.randn(32).astype(np.float32)
    r = ref.cuda_conv_pointset(cu(p3.transpose(0, 2, 1)), cu(nb.transpose(0, 2, 1)), cu(th), cu(bi))
    m = user_ops.convolution_pointset(cu(p3.transpose(0, 2, 1)), cu(nb.transpose(0, 2, 1)), cu(th), cu(bi))
    assert torch.equal(m, r)   # same FMA order as the reference kernel
