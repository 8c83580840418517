#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x -k "knn or three_nn or nms or ref_cuda or sweep or degenerate or identical or model or retriev" > $out/pytest_r2y.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2y.log
timeout 200 python scripts/exp_sorted.py 2>&1 | grep -E "random knn" | tee $out/knn_r2y.txt
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-modes --op-table $out/op_table_r2y.json > $out/bench_r2y.json 2> $out/bench_r2y.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2y.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for r in d['op_roofline']:
    if 'knn' in r['op'] or 'three_nn' in r['op'] or 'farthest' in r['op']: print('  %-60s %8.4f ms' % (r['op'], r['ms']))
print(json.dumps(d.get('data_sensitivity'))[:1500])
PY
timeout 300 python scripts/timeline.py $out/timeline_r2y.txt > $out/timeline_r2y.log 2>&1; head -16 $out/timeline_r2y.txt | cut -c1-100; tail -1 $out/timeline_r2y.txt
