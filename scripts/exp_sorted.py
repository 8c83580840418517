"""Feasibility: FlexConv / SE / flex_pool on a cloud given in random point order vs Morton-sorted order (tiles of
consecutive indices are then spatially compact).  python scripts/exp_sorted.py"""
import sys, os
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops

def morton_order(pts):
    q = ((pts - pts.amin(1, keepdim=True)) / (pts.amax(1, keepdim=True) - pts.amin(1, keepdim=True) + 1e-9) * 1023).long().clamp(0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[..., 0]) | (spread(q[..., 1]) << 1) | (spread(q[..., 2]) << 2)
    return code.argsort(dim=1)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

B, N, K = 32, 8192, 8
g = torch.Generator(device="cuda").manual_seed(0)
pts = (torch.rand((B, N, 3), device="cuda", generator=g) * 50 - 25).contiguous()
order = morton_order(pts)
pts_s = torch.gather(pts, 1, order.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
for name, P in (("random", pts), ("sorted", pts_s)):
    nb, _ = ops.knn_points(P, K)
    for Ci, Co in ((32, 64), (64, 64)):
        f = torch.randn((B, N, Ci), device="cuda", generator=g)
        th = torch.randn((3, Ci, Co), device="cuda", generator=g) / Ci ** 0.5
        bi = torch.randn((Ci, Co), device="cuda", generator=g) / Ci ** 0.5
        pk = ops.flex_conv_prepack(th, bi)
        ms = timeit(lambda: ops.flex_conv_packed(f, pk, nb, P))
        print("%s flexconv %d->%d: %.4f ms" % (name, Ci, Co, ms))
    x = torch.randn((B, N, 64), device="cuda", generator=g)
    w1 = torch.randn((64, 16), device="cuda", generator=g) / 8; b1 = torch.zeros(16, device="cuda")
    w2 = torch.randn((16, 64), device="cuda", generator=g) / 4; b2 = torch.zeros(64, device="cuda")
    print("%s se_pool_excite<64>: %.4f ms" % (name, timeit(lambda: ops.se_pool_excite(x, nb, w1, b1, w2, b2))))
    x32 = torch.randn((B, N, 32), device="cuda", generator=g)
    print("%s flex_pool<32>: %.4f ms" % (name, timeit(lambda: ops.flex_pool(x32, nb))))
    print("%s knn: %.4f ms" % (name, timeit(lambda: ops.knn_points(P, K))))
