"""Runs the fused two-branch join (dh3d_linear_join_packed) at the DH3D shape a few times (ncu target / timing):
    python scripts/run_join.py [M] [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator(device="cuda").manual_seed(0)
xa = torch.randn((M, 192), device="cuda", generator=g)
xb = torch.randn((M, 64), device="cuda", generator=g)
wa = torch.randn((192, 128), device="cuda", generator=g) / 192 ** 0.5
wb = torch.randn((64, 128), device="cuda", generator=g) / 8.0
sa = torch.rand(128, device="cuda", generator=g) + 0.5
ba = torch.randn(128, device="cuda", generator=g)
pa, pb = ops.linear_prepack(wa), ops.linear_prepack(wb)
def fused():
    return ops.linear_join(xa, pa, sa, ba, 1, xb, pb, sa, ba, 1, eps=1e-8)
def unfused():
    u = ops.linear(xa, wa, scale=sa, shift=ba, act=1, packed=pa)
    v = ops.linear(xb, wb, scale=sa, shift=ba, act=1, packed=pb)
    return ops.add_l2_normalize_rows(u, v, 1e-8)
for name, fn in (("fused", fused), ("unfused", unfused)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    print("%s: %.3f ms  (%.0f GB/s of the fused form's 4*M*(192+64+256) bytes)" % (
        name, a.elapsed_time(b) / reps, 4.0 * M * 512 / (a.elapsed_time(b) / reps) / 1e6))
