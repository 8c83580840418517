#!/usr/bin/env python
"""Top stalled SASS instructions of an `ncu --page source --csv` export (run here, no GPU needed).
    python scripts/ncu_src_top.py gpurun_out/prof_X_src.csv [kernel-substring] [launch-index] [top-n]
"""
import csv
import sys

path = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]) - 2:
        cur["rows"].append(r)
sel = [b for b in blocks if sub in b["name"]]
print("kernels:", [b["name"][:50] for b in blocks])
b = sel[which]
h = b["hdr"]
ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_")]
tot = sum(int(r[ix["# Samples"]] or 0) for r in b["rows"])
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in b["rows"])
print(b["name"][:90], "samples", tot, "warp-instructions", tot_inst)
agg = {n: 0 for n in stall_cols}
for r in b["rows"]:
    for n in stall_cols:
        try:
            agg[n] += int(r[ix[n]] or 0)
        except (ValueError, IndexError):
            pass
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][ix["# Samples"]] or 0))[:topn]
for i in sorted(order):
    r = b["rows"][i]
    st = {n[6:]: int(r[ix[n]] or 0) for n in stall_cols if len(r) > ix[n] and (r[ix[n]] or "0") != "0"}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d %6s %5.1f%% exec %8s  %-58s %s" % (i, r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1),
                                              r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:58], top))
