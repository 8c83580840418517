#!/bin/bash
# Runs on the GPU box (under gpurun): the whole GPU suite under a timeout, then a short bench.
# Usage: bash scripts/gpu_check.sh <tag> [extra bench env assignments ...]
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi_$tag.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?" | tee $out/summary_$tag.txt
tail -n 4 $out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --op-table $out/op_table_$tag.json > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/summary_$tag.txt
cut -c1-260 $out/bench_$tag.json
