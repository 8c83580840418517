#!/bin/bash
# Runs on the GPU box (under gpurun): new/risky tests first under a short timeout, then the whole GPU suite,
# then a short bench.  Usage: bash scripts/gpu_check.sh <tag>
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi_$tag.txt 2>&1
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "fps" > $out/pytest_fps_$tag.log 2>&1; echo "fps tests rc=$?" | tee -a $out/summary_$tag.txt
timeout 600 python -m pytest tests/test_backward_gpu.py -q > $out/pytest_bwd_$tag.log 2>&1; echo "backward tests rc=$?" | tee -a $out/summary_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1; echo "gpu suite rc=$?" | tee -a $out/summary_$tag.txt
tail -5 $out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --op-table $out/op_table_$tag.json > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?" | tee -a $out/summary_$tag.txt
DH3D_FPS=cta timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_fpscta.json 2>> $out/bench_$tag.err
cat $out/bench_$tag.json | head -c 3000
tail -3 $out/pytest_fps_$tag.log $out/pytest_bwd_$tag.log
