#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/data_sensitivity.py $out/data_sensitivity_r2b.json > $out/data_sensitivity_r2b.log 2>&1; echo "sens rc=$?"
cat $out/data_sensitivity_r2b.log | tail -8
timeout 900 python -m pytest tests/test_sweep_gpu.py tests/test_ops_gpu.py tests/test_ref_cuda_gpu.py tests/test_nms.py -m gpu -q -x -k "knn or three_nn or identical or nms or sweep" -rP > $out/pytest_knn_r2b.log 2>&1; echo "knn tests rc=$?"
grep -E "passed|failed|error|identical-cloud" $out/pytest_knn_r2b.log | tail -n 8
