#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_retrieval_gpu.py -m gpu -q -x -k "netvlad or model or forward or retriev or oracle" > $out/pytest_r2q.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2q.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2q.json > $out/bench_r2q.json 2> $out/bench_r2q.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2q.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
timeout 300 python scripts/timeline.py $out/timeline_r2q.txt > $out/timeline_r2q.log 2>&1; tail -8 $out/timeline_r2q.log
