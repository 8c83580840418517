#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_r2r.log 2>&1; echo "tests rc=$?"; tail -5 $out/pytest_r2r.log
timeout 300 python scripts/timeline.py $out/timeline_r2r.txt > $out/timeline_r2r.log 2>&1; tail -6 $out/timeline_r2r.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2r.json > $out/bench_r2r.json 2> $out/bench_r2r.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2r.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
