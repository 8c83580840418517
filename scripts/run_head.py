"""Times the fused 256 -> 1024 -> 1 head (dh3d_linear_rowdot_packed) alone: python scripts/run_head.py [M K N reps]"""
import sys
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops
M, K, N = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (262144, 256, 1024)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
x = torch.randn(M, K, device="cuda")
w = torch.randn(K, N, device="cuda") / K ** 0.5
sc, sh = torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
w2 = torch.randn(N, device="cuda") / N ** 0.5
p = ops.linear_prepack(w)
for _ in range(3):
    y = ops.linear_rowdot(x, p, sc, sh, 1, w2, 0.1, 2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    y = ops.linear_rowdot(x, p, sc, sh, 1, w2, 0.1, 2)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print("head M=%d K=%d N=%d: %.4f ms  %.1f TFLOP/s (fp32-equivalent 2MKN)" % (M, K, N, ms, 2.0 * M * K * N / ms / 1e9))
