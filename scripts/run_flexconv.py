"""Runs one FlexConv shape a few times (ncu target / timing): python scripts/run_flexconv.py B N K Cin Cout [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops
B, N, K, Ci, Co = (int(a) for a in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
g = torch.Generator(device="cuda").manual_seed(0)
pts = (torch.rand((B, N, 3), device="cuda", generator=g) * 50 - 25).contiguous()
nb, _ = ops.knn_points(pts, K)
f = torch.randn((B, N, Ci), device="cuda", generator=g)
th = torch.randn((3, Ci, Co), device="cuda", generator=g) / Ci ** 0.5
bi = torch.randn((Ci, Co), device="cuda", generator=g) / Ci ** 0.5
for _ in range(2):
    y = ops.flex_conv(f, th, bi, nb, pts)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    y = ops.flex_conv(f, th, bi, nb, pts)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
n = B * N
print("flex_conv B=%d N=%d K=%d %d->%d: %.3f ms  gathered %.0f GB/s  algorithmic %.0f GB/s" % (
    B, N, K, Ci, Co, ms, 4.0 * n * K * Ci / ms / 1e6, 4.0 * (n * Ci + n * Co + n * K + 3 * n + 4 * Ci * Co) / ms / 1e6))
