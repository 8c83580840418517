#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_nms.py -m gpu -q -x > $out/pytest_r2e.log 2>&1; echo "gemm tests rc=$?"; tail -3 $out/pytest_r2e.log
timeout 300 python scripts/op_times.py far_outlier_padding > $out/op_times_outlier_r2e.log 2>&1; head -6 $out/op_times_outlier_r2e.log
timeout 900 python bench.py --steps 20 --warmup 3 --op-table $out/op_table_r2e.json > $out/bench_r2e.json 2> $out/bench_r2e.err; echo "bench rc=$?"; tail -5 $out/bench_r2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
for k in ('value','ms_per_step','e2e','e2e_modes','sustained','step_hbm','cpu_baseline','data_sensitivity'):
    print(k, json.dumps(d.get(k))[:900])
print('roofline', json.dumps(d['roofline'])[:600])
for r in d['op_roofline'][:14]: print(r)
print('ref_cuda', json.dumps(d['ref_cuda'])[:1500])
PY
timeout 600 python bench.py --workload retrieval > $out/bench_retrieval_r2e.json 2> $out/bench_retrieval_r2e.err; echo "retrieval rc=$?"; tail -3 $out/bench_retrieval_r2e.err; cut -c1-1500 $out/bench_retrieval_r2e.json
timeout 600 python bench.py --workload sweep > $out/bench_sweep_r2e.json 2> $out/bench_sweep_r2e.err; echo "sweep rc=$?"; tail -3 $out/bench_sweep_r2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_sweep_r2e.json'))
for r in d['sweep']: print(r)
PY
