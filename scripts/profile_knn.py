#!/usr/bin/env python
"""k-NN alone on the benchmark clouds (32 x 8192, K = 8): timing with CUDA events, or one pass for ncu.
    python scripts/profile_knn.py            # prints ms (median of 20)
    ncu ... --profile-from-start off python scripts/profile_knn.py ncu
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import synth_clouds  # noqa: E402
from dh3d_b200 import ops  # noqa: E402


def main():
    torch.cuda.set_device(0)
    pts = synth_clouds(32, 8192, 0).cuda()
    for _ in range(3):
        ops.knn_points(pts, 8)
    torch.cuda.synchronize()
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":
        torch.cuda.profiler.start()
        ops.knn_points(pts, 8)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.knn_points(pts, 8)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("knn 32x8192 K=8: median %.4f ms  min %.4f ms" % (ts[len(ts) // 2], ts[0]))


if __name__ == "__main__":
    main()
