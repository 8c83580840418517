#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "three_nn or model or retriev or cabi" > $out/pytest_r3e.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r3e.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r3e.json > $out/bench_r3e.json 2> $out/bench_r3e.err; echo "bench rc=$?"; tail -3 $out/bench_r3e.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r3e.json'))
print('value %.0f  ms/step %.4f e2e %.0f launches/step %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']/d['steps']))
PY
timeout 300 python scripts/timeline.py $out/timeline_r3e.txt > $out/timeline_r3e.log 2>&1; sed -n 6,17p $out/timeline_r3e.txt | cut -c1-100; tail -1 $out/timeline_r3e.txt
