#!/bin/bash
out=gpurun_out; mkdir -p $out
DH3D_T16_PREFETCH=0 timeout 120 python scripts/run_head.py > $out/head_r2m.txt 2>&1; DH3D_T16_PREFETCH=1 timeout 120 python scripts/run_head.py >> $out/head_r2m.txt 2>&1; cat $out/head_r2m.txt
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x > $out/pytest_r2m.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2m.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2m.json > $out/bench_r2m.json 2> $out/bench_r2m.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2m.json'))
print('value %.0f  ms/step %.4f' % (d['value'], d['ms_per_step']))
for r in d['op_roofline'][:8]: print('  %-60s %8.4f ms' % (r['op'], r['ms']))
PY
