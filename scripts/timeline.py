#!/usr/bin/env python
"""Kernel timeline of one CUDA-graph replay of the full forward (BASELINE configs[2]), from torch.profiler's CUPTI
activity records: per kernel its stream, start offset, duration and the idle gap in front of it on its stream.

    python scripts/timeline.py [out.txt] [batch]

This is the gap / overlap view the ncu launch list cannot give (ncu serialises the launches).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import N_POINTS, synth_clouds  # noqa: E402
from dh3d_b200.configs import full_config  # noqa: E402
from dh3d_b200.model import DH3D, GraphedForward, init_random_  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = init_random_(DH3D(full_config()), seed=0).to(dev)
    clouds = [synth_clouds(batch, N_POINTS, i).to(dev) for i in range(4)]
    fwd = GraphedForward(model, clouds[0])
    for i in range(3):
        fwd(clouds[i])
    torch.cuda.synchronize()
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(reps):
            fwd(clouds[i])
        torch.cuda.synchronize()
    import json
    import tempfile
    tmp = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(tmp)
    with open(tmp) as f:
        trace = json.load(f)
    os.unlink(tmp)
    evs = [e for e in trace["traceEvents"] if e.get("cat") == "kernel"]
    evs.sort(key=lambda e: e["ts"])
    n = len(evs) // reps
    evs = evs[(reps - 1) * n:]          # the last replay
    t0 = evs[0]["ts"]
    lines = []
    last_end = {}
    busy_end = t0
    idle_all = 0.0
    for e in evs:
        s, d = e["ts"] - t0, e["dur"]
        stream = e.get("args", {}).get("stream", "?")
        gap = s - last_end.get(stream, 0.0)
        last_end[stream] = s + d
        if e["ts"] > busy_end:
            idle_all += e["ts"] - busy_end
        busy_end = max(busy_end, e["ts"] + d)
        lines.append("%9.1f us  +%7.1f us  gap %6.1f  stream %-4s %s" % (s, d, gap, stream, e["name"][:90]))
    total = busy_end - t0
    lines.append("replay: %.1f us from the first kernel's start to the last kernel's end; %.1f us with NO kernel running; "
                 "%d kernels" % (total, idle_all, len(evs)))
    text = "\n".join(lines)
    print(text)
    if out_path:
        with open(out_path, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
