#!/usr/bin/env python
"""One eager full forward (BASELINE configs[2]: 32 clouds x 8192 points) for ncu.

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof python scripts/profile_forward.py

Two warm-up forwards run with the profiler off; the third is bracketed by cudaProfilerStart/Stop.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import N_POINTS, synth_clouds  # noqa: E402
from dh3d_b200.configs import full_config  # noqa: E402
from dh3d_b200.model import DH3D, init_random_  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = init_random_(DH3D(full_config()), seed=0).to(dev)
    outputs = ("local_desc", "attention", "globaldesc")
    clouds = [synth_clouds(batch, N_POINTS, i).to(dev) for i in range(3)]
    for i in range(2):
        model(clouds[i], outputs=outputs)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(clouds[2], outputs=outputs)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
