#!/usr/bin/env python
"""Summarises ncu CSV exports (run here, no GPU needed).
    python scripts/ncu_summary.py launches gpurun_out/launches_X.csv     per-kernel device time + share
    python scripts/ncu_summary.py raw gpurun_out/prof_X_raw.csv          key metrics of a --set full capture
"""
import collections
import csv
import re
import sys


def rows_of(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    return list(csv.reader(lines))


def launches(path):
    rows = [r for r in rows_of(path) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        agg.setdefault((name, r[7], r[8]), []).append(float(r[-1]) / 1e3)
    total = sum(sum(v) for v in agg.values())
    print("%-46s %-14s %-14s %4s %10s %7s" % ("kernel", "block", "grid", "n", "us/launch", "share"))
    for (name, blk, grd), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-46s %-14s %-14s %4d %10.1f %6.1f%%" % (name[:46], blk, grd, len(v), sum(v) / len(v),
                                                        100 * sum(v) / total))
    print("total %.1f us over %d launches" % (total, len(rows)))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.avg.per_cycle_active",
        "l1tex__t_bytes.sum", "smsp__cycles_active.avg"]


def raw(path):
    rows = rows_of(path)
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [k for k in KEYS if k in idx]
    missing = [k for k in KEYS if k not in idx]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        print("== %s  grid %s block %s" % (name[:70], r[idx["Grid Size"]], r[idx["Block Size"]]))
        for k in want:
            print("   %-66s %s %s" % (k, r[idx[k]], units[idx[k]]))
    if missing:
        print("metrics not in this capture:", missing)


def traffic(path):
    """JSON list {kernel, grid, block, time_us, dram_bytes} per captured launch (bench.py's roofline.traffic)."""
    import json
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[idx[k]]) * mult[units[idx[k]]]
        t = float(r[idx["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[idx["gpu__time_duration.sum"]]]
        out.append({"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", ""),
                    "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]], "time_us": t, "dram_bytes": b})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2])
