"""Runs one big packed GEMM a few times (ncu target): python scripts/run_gemm.py M K N [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops
M, K, N = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x = torch.randn(M, K, device="cuda")
w = torch.randn(K, N, device="cuda") / K ** 0.5
sc, sh = torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
p = ops.linear_prepack(w)
for _ in range(reps):
    y = ops.linear(x, w, scale=sc, shift=sh, act=1, packed=p)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    y = ops.linear(x, w, scale=sc, shift=sh, act=1, packed=p)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("M=%d K=%d N=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent 2MKN)" % (M, K, N, ms, 2.0 * M * K * N / ms / 1e9))
