#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --workload sweep > $out/bench_r3j_sweep.json 2> $out/bench_r3j_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3j_sweep.json') if l.startswith('{')][0]
for r in d['sweep']: print(r['N'], r['K'], 'flexconv %.4f ms frac %.3f gathered %.0f GB/s | dropin %.4f | knn %.4f ms' % (r['flexconv_ms'], r['flexconv_frac_hbm'], r['flexconv_gathered_l2_gbs'], r['flexconv_dropin_cm_ms'], r['knn_ms']))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 | cut -c1-400
