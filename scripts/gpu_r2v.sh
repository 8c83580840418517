#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "flex or sweep" > $out/pytest_r2v.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2v.log
for s in "8 8192 8 128 128" "8 8192 16 128 128" "8 8192 32 128 128" "32 8192 8 64 64"; do timeout 100 python scripts/run_flexconv.py $s 10 2>&1 | tail -1; done | tee $out/flexconv_r2v.txt
timeout 600 python bench.py --workload sweep --steps 5 --warmup 3 > $out/bench_sweep_r2v.json 2> $out/bench_sweep_r2v.err; echo "sweep rc=$?"; head -c 3000 $out/bench_sweep_r2v.json
