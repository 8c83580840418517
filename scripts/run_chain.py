"""Times the chained 128 -> 128 -> 256 layers (dh3d_linear_chain_packed) against the two separate launches:
python scripts/run_chain.py [M reps]"""
import sys
import torch
sys.path.insert(0, ".")
from dh3d_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
x = torch.randn(M, 128, device="cuda")
w1 = torch.randn(128, 128, device="cuda") / 128 ** 0.5
w2 = torch.randn(128, 256, device="cuda") / 128 ** 0.5
s1, b1 = torch.rand(128, device="cuda") + 0.5, torch.randn(128, device="cuda") * 0.1
s2, b2 = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda") * 0.1
p1, p2 = ops.linear_prepack(w1), ops.linear_prepack(w2)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_chain = timeit(lambda: ops.linear_chain(x, p1, s1, b1, 1, p2, s2, b2, 1))
t_sep = timeit(lambda: ops.linear(ops.linear(x, w1, scale=s1, shift=b1, act=1, packed=p1), w2, scale=s2, shift=b2, act=1,
                                  packed=p2))
print("chain %.4f ms (%.0f GB/s of x + y)   separate %.4f ms" % (t_chain, 4.0 * M * (128 + 256) / t_chain / 1e6, t_sep))
