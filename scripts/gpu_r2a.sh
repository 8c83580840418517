#!/bin/bash
# round-2 first GPU pass: the new parity tests (with their printed error levels), then the old suite
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi_r2a.txt 2>&1
nproc > $out/nproc_r2a.txt; numactl -H >> $out/nproc_r2a.txt 2>&1; nvidia-smi topo -m >> $out/nproc_r2a.txt 2>&1
timeout 1500 python -m pytest tests/test_sweep_gpu.py tests/test_layers_gpu.py tests/test_model_gpu.py -m gpu -q -rP > $out/pytest_new_r2a.log 2>&1; echo "new tests rc=$?" | tee $out/summary_r2a.txt
grep -E "passed|failed|error|vs oracle|identical-cloud|real weights" $out/pytest_new_r2a.log | tail -n 30
