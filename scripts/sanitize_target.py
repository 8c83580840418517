"""One small launch of every hand-synchronised kernel (compute-sanitizer target):
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_target.py [names...]
fps (DSMEM spin protocol, 4-CTA clusters), flexconv (cp.async ring + mbarrier stages + tcgen05), netvlad (TMA ring,
in-place split, TMEM hand-offs, grid-barrier tail), gemm heads / join (CTA pairs with remote mbarrier arrives, in-place
split, resident-activation head, out-of-window row queue), knn / three_nn (staged chunks, warp-synchronous flushes), se_pool_excite, gather family."""
import sys

import torch

sys.path.insert(0, ".")
from dh3d_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
which = set(sys.argv[1:])


def want(name):
    return not which or name in which


def rnd(*shape):
    return torch.randn(shape, device="cuda", generator=g)


pts = (torch.rand((2, 2048, 3), device="cuda", generator=g) * 20 - 10).contiguous()
if want("knn"):
    for k in (8, 16, 50):
        ops.knn_points(pts, k)
    ops.knn_points(torch.zeros((1, 1024, 3), device="cuda"), 8)
    print("knn ok")
nbr, _ = ops.knn_points(pts, 8)
if want("fps"):
    ops.farthest_point_sample(256, pts)
    ops.farthest_point_sample(64, pts[:, :700].contiguous())
    ops.farthest_point_sample(256, pts, sorted_ws=ops.knn_sort(pts))       # box-pruned kernel on the sorted cloud
    z = torch.zeros((1, 1024, 3), device="cuda")
    ops.farthest_point_sample(64, z, sorted_ws=ops.knn_sort(z))
    print("fps ok")
if want("three_nn"):
    kp = ops.farthest_point_sample(256, pts)
    m = ops.gather_point(pts, kp)
    d, i = ops.three_nn(pts, m)
    _, _, ws = ops.knn_points(pts, 8, keep_workspace=True)
    ops.three_nn(pts, m, sorted1=ws)                    # query sort taken from the k-NN workspace
    ops.three_nn(pts, m, sorted1=ws, sorted2=ops.knn_sort(m))
    ops.three_interpolate(rnd(2, 256, 128), i, d, weight_is_dist2=True)
    print("three_nn ok")
if want("flexconv"):
    for (ci, co) in ((32, 64), (64, 64), (128, 256)):
        ops.flex_conv(rnd(2, 2048, ci), rnd(3, ci, co) / 8, rnd(ci, co) / 8, nbr, pts)
    nbr16, _ = ops.knn_points(pts, 16)
    ops.flex_conv(rnd(2, 2048, 128), rnd(3, 128, 128) / 8, rnd(128, 128) / 8, nbr16, pts)
    nbr32, _ = ops.knn_points(pts, 32)
    ops.flex_conv(rnd(2, 2048, 128), rnd(3, 128, 128) / 8, rnd(128, 128) / 8, nbr32, pts)   # K = 32: four batches per group
    nbr12, _ = ops.knn_points(pts, 12)
    ops.flex_conv(rnd(2, 2048, 64), rnd(3, 64, 64) / 8, rnd(64, 64) / 8, nbr12, pts)        # generic loop (K % 8 != 0)
    ops.flex_conv(rnd(2, 2048, 8), rnd(3, 8, 12) / 8, rnd(8, 12) / 8, nbr, pts)
    print("flexconv ok")
if want("gather"):
    f = rnd(2, 2048, 64)
    ops.flex_pool(f, nbr, with_argmax=True)
    ops.conv_pointset(pts, rnd(3, 32), rnd(32), nbr)
    ops.se_pool_excite(f, nbr, rnd(64, 16) / 8, rnd(16), rnd(16, 64) / 4, rnd(64))
    print("gather ok")
if want("gemm"):
    x = rnd(4200, 256)    # 33 row tiles: CTA pairs (tcgen05.mma.cta_group::2) and the resident-activation head kernel
    x[17] *= 1e6          # one out-of-window row: exercises the queue + fp32 recompute
    w = rnd(256, 1024) / 16
    p = ops.linear_prepack(w)
    ops.linear(x, w, packed=p, act=1)
    ops.linear_rowdot(x, p, None, None, 1, rnd(1024) / 32, 0.1, 2)
    ops.linear_rowdot(x[:3000].contiguous(), p, None, None, 1, rnd(1024) / 32, 0.1, 2)   # single-CTA streaming kernel
    xa, xb = rnd(3000, 192), rnd(3000, 64)
    xa[5] *= 1e5
    ops.linear_join(xa, ops.linear_prepack(rnd(192, 128) / 14), None, None, 1, xb, ops.linear_prepack(rnd(64, 128) / 8),
                    None, None, 1, eps=1e-8)
    xc = rnd(4200, 128)
    xc[9] *= 1e6          # out-of-window input row; column scale 1e5 below pushes hidden rows out of the window too
    sc = torch.ones(128, device="cuda")
    sc[::5] = 1e5
    ops.linear_chain(xc, ops.linear_prepack(rnd(128, 128) / 11), None, None, 1, ops.linear_prepack(rnd(128, 256) / 11),
                     None, None, 1)
    ops.linear_chain(xc[:700].contiguous(), ops.linear_prepack(rnd(128, 128) / 11), sc, None, 1,
                     ops.linear_prepack(rnd(128, 256) / 11), None, None, 0)
    ops.linear(rnd(1000, 64), rnd(64, 16) / 8)     # FFMA kernel
    print("gemm ok")
if want("netvlad"):
    one = torch.ones(64, device="cuda")
    one256 = torch.ones(256, device="cuda")
    ops.netvlad(rnd(2, 2048, 256), torch.rand((2, 2048), device="cuda", generator=g), rnd(256, 64) / 16,
                (one, one * 0), rnd(256, 64) / 16, rnd(16384, 256) / 8, (one256, one256 * 0), rnd(256, 256) / 16,
                (one256, one256 * 0))
    print("netvlad ok")
torch.cuda.synchronize()
print("sanitize target done")
