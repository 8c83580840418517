#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 120 python scripts/run_head.py > $out/head_r2n.txt 2>&1; echo "head rc=$?"; cat $out/head_r2n.txt
timeout 400 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x > $out/pytest_r2n.log 2>&1; echo "tests rc=$?"; tail -5 $out/pytest_r2n.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2n.json > $out/bench_r2n.json 2> $out/bench_r2n.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2n.json'))
print('value %.0f  ms/step %.4f' % (d['value'], d['ms_per_step']))
for r in d['op_roofline'][:8]: print('  %-60s %8.4f ms' % (r['op'], r['ms']))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_head16 -s 3 -c 1 -o /tmp/prof_head python scripts/run_head.py 262144 256 1024 3 > $out/ncu_head.log 2>&1
ncu -i /tmp/prof_head.ncu-rep --page raw --csv > $out/prof_head_raw.csv 2>> $out/ncu_head.log
ncu -i /tmp/prof_head.ncu-rep --page source --csv > $out/prof_head_src.csv 2>> $out/ncu_head.log
tail -2 $out/ncu_head.log
