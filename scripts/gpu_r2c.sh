#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_sweep_gpu.py tests/test_ops_gpu.py tests/test_ref_cuda_gpu.py tests/test_nms.py tests/test_model_gpu.py -m gpu -q -x -rP > $out/pytest_r2c.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|error|identical-cloud|Error" $out/pytest_r2c.log | tail -n 8
timeout 600 python scripts/data_sensitivity.py $out/data_sensitivity_r2c.json > $out/data_sensitivity_r2c.log 2>&1; echo "sens rc=$?"
cat $out/data_sensitivity_r2c.log | tail -8
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --op-table $out/op_table_r2c.json > $out/bench_r2c.json 2> $out/bench_r2c.err; echo "bench rc=$?"
cut -c1-400 $out/bench_r2c.json
