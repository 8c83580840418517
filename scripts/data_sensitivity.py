"""Per-dataset timing of the data-dependent kernels (k-NN, FPS, 3-NN) and of the whole forward, 32 x 8192 points:
uniform U(-25,25)^3, LiDAR-like synthetic (dh3d_b200.data.synth_lidar_clouds), the reference's own demo clouds
(tests/golden/demo_clouds.npz tiled), all-zero padding clouds, and clouds padded with far outliers
(get_fixednum_pcd(randsample=False): 1e5 points).  Run on the GPU box:  python scripts/data_sensitivity.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dh3d_b200 import ops  # noqa: E402
from dh3d_b200.configs import full_config  # noqa: E402
from dh3d_b200.data import synth_lidar_clouds  # noqa: E402
from dh3d_b200.model import DH3D, init_random_  # noqa: E402


def time_ms(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def datasets(B=32, N=8192):
    g = torch.Generator().manual_seed(1234)
    uni = (torch.rand((B, N, 3), generator=g) * 50 - 25).float()
    d = {"uniform": uni, "lidar_like": torch.from_numpy(synth_lidar_clouds(B, N, 0))}
    gold = os.path.join(ROOT, "tests", "golden", "demo_clouds.npz")
    if os.path.exists(gold):
        c = np.load(gold)["clouds"]
        d["oxford_demo"] = torch.from_numpy(np.concatenate([c] * (B // len(c)), 0)[:B].copy())
    d["all_zero"] = torch.zeros((B, N, 3))
    out = uni.clone()
    out[:, N - N // 8:] = 100000.0       # get_fixednum_pcd(randsample=False) padding
    d["far_outlier_padding"] = out
    return d


def main():
    model = init_random_(DH3D(full_config()), seed=0).cuda()
    res = {}
    for name, pts in datasets().items():
        p = pts.cuda()
        m = p[:, :1024].contiguous()
        r = {"knn_ms": time_ms(lambda: ops.knn_points(p, 8)),
             "fps_ms": time_ms(lambda: ops.farthest_point_sample(1024, p)),
             "three_nn_ms": time_ms(lambda: ops.three_nn(p, m)),
             "forward_ms": time_ms(lambda: model(p), reps=5)}
        r["clouds_per_s"] = 32 / r["forward_ms"] * 1e3
        res[name] = {k: round(v, 4) for k, v in r.items()}
        print(name, res[name], flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
