#!/bin/bash
# Runs on the GPU box (under gpurun): the whole GPU suite under a timeout, then a default bench run.
# Usage: bash scripts/gpu_check.sh <tag> [bench args ...]
tag=${1:-chk}; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi_$tag.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x -rP > $out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?" | tee $out/summary_$tag.txt
grep -E "passed|failed|error|^FAILED|^ERROR" $out/pytest_gpu_$tag.log | tail -n 6
timeout 900 python bench.py --steps 30 --warmup 3 --op-table $out/op_table_$tag.json "$@" > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/summary_$tag.txt
tail -3 $out/bench_$tag.err
python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.0f  ms/step %.4f  e2e %.0f (ceiling frac %.3f)  sustained %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['frac_of_ceiling'], (d.get('sustained') or {}).get('value', 0)))
for r in d['op_roofline'][:16]: print('  %-60s %8.4f ms  frac_hbm %s' % (r['op'], r['ms'], r.get('frac_hbm')))
if d.get('data_sensitivity'):
    for k,v in d['data_sensitivity'].items(): print('  sens', k, v)
PY
