"""Per-op device time of one eager forward on a chosen input set:  python scripts/op_times.py <dataset> [B]
dataset in uniform | lidar_like | oxford_demo | all_zero | far_outlier_padding (scripts/data_sensitivity.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from data_sensitivity import datasets  # noqa: E402
from dh3d_b200 import _lib  # noqa: E402
from dh3d_b200.configs import full_config  # noqa: E402
from dh3d_b200.model import DH3D, init_random_  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "uniform"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pts = datasets(B)[name].cuda()
model = init_random_(DH3D(full_config()), seed=0).cuda()
for _ in range(2):
    out = model(pts, overlap=False)
torch.cuda.synchronize()
_lib.stats.reset()
_lib.stats.timing_filter = "all"
out = model(pts, overlap=False)
torch.cuda.synchronize()
t = _lib.stats.op_times_ms()
for k, v in sorted(t.items(), key=lambda kv: -kv[1][0]):
    print("%-70s %8.3f ms  x%d" % (k, v[0], v[1]))
for k in ("feat", "local_desc", "attention", "globaldesc"):
    v = out[k]
    print(k, "finite" if torch.isfinite(v).all() else "NON-FINITE", float(v.abs().max()))
