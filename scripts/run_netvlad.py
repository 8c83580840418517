"""Times dh3d_netvlad (32 x 8192 x 256, the benchmark shape) alone: python scripts/run_netvlad.py [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = torch.Generator(device="cuda").manual_seed(0)
B, N, D, K, O = 32, 8192, 256, 64, 256
rnd = lambda *s: torch.randn(s, device="cuda", generator=g)
f, att = rnd(B, N, D), torch.rand((B, N), device="cuda", generator=g)
one, one256 = torch.ones(K, device="cuda"), torch.ones(O, device="cuda")
args = (f, att, rnd(D, K) / 16, (one, one * 0), rnd(D, K) / 16, rnd(D * K, O) / 128, (one256, one256 * 0), rnd(O, O) / 16,
        (one256, one256 * 0))
for _ in range(3):
    ops.netvlad(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ops.netvlad(*args)
b.record()
torch.cuda.synchronize()
print("netvlad B=%d N=%d: %.4f ms" % (B, N, a.elapsed_time(b) / reps))
