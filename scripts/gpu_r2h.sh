#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -k "rowdot or out_of_window" > $out/pytest_r2h_a.log 2>&1; echo "head tests rc=$?"; tail -4 $out/pytest_r2h_a.log
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_sweep_gpu.py tests/test_ref_cuda_gpu.py -m gpu -q -x -k "flex_conv or sweep or flex_ops" > $out/pytest_r2h_b.log 2>&1; echo "flexconv tests rc=$?"; tail -4 $out/pytest_r2h_b.log
for shape in "32 8192 8 64 64" "32 8192 8 32 64" "8 8192 8 128 128" "32 1024 8 128 256"; do timeout 120 python scripts/run_flexconv.py $shape 20; done
timeout 120 python scripts/run_head.py 2>&1 | tail -5
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2h.json > $out/bench_r2h.json 2> $out/bench_r2h.err; echo "bench rc=$?"; tail -3 $out/bench_r2h.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2h.json'))
print('value %.0f  ms/step %.4f  e2e %.0f  sustained %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('sustained') or {}).get('value', 0)))
for r in d['op_roofline'][:12]: print('  %-60s %8.4f ms  frac_hbm %s %s' % (r['op'], r['ms'], r.get('frac_hbm'), r.get('frac_bf16_burst','')))
print(json.dumps(d['roofline'])[:400])
PY
