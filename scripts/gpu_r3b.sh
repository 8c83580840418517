#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -k "chain" > $out/pytest_r3b.log 2>&1; echo "chain tests rc=$?"; tail -15 $out/pytest_r3b.log
timeout 900 python -m pytest tests -m gpu -q -x -k "model or layers or checkpoint or cabi" > $out/pytest_r3b2.log 2>&1; echo "model tests rc=$?"; tail -3 $out/pytest_r3b2.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r3b.json > $out/bench_r3b.json 2> $out/bench_r3b.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r3b.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for r in d['op_roofline']:
    if 'chain' in r['op'] or 'linear_packed' in r['op']: print('  %-60s %8.4f ms' % (r['op'], r['ms']))
PY
timeout 300 python scripts/timeline.py $out/timeline_r3b.txt > $out/timeline_r3b.log 2>&1; grep -E "gemm_|replay" $out/timeline_r3b.txt | cut -c1-100
