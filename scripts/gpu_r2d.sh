#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python scripts/op_times.py far_outlier_padding > $out/op_times_outlier_r2d.log 2>&1; head -12 $out/op_times_outlier_r2d.log; tail -4 $out/op_times_outlier_r2d.log
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_sweep_gpu.py tests/test_ops_gpu.py tests/test_ref_cuda_gpu.py tests/test_nms.py tests/test_model_gpu.py tests/test_layers_gpu.py -m gpu -q -rP > $out/pytest_r2d.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|error|identical-cloud|^FAILED|^ERROR" $out/pytest_r2d.log | tail -n 12
timeout 600 python scripts/data_sensitivity.py $out/data_sensitivity_r2d.json > $out/data_sensitivity_r2d.log 2>&1; echo "sens rc=$?"
cat $out/data_sensitivity_r2d.log | tail -8
