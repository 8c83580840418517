#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2o.json > $out/bench_r2o.json 2> $out/bench_r2o.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2o.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x > $out/pytest_r2o.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2o.log
