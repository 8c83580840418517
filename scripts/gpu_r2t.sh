#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 200 python scripts/exp_sorted.py 2>&1 | grep -v Warn > $out/exp_sorted_r2t.txt; cat $out/exp_sorted_r2t.txt
for s in "32 1024 8 64 128" "32 1024 8 128 128" "32 1024 8 128 256" "8 8192 8 128 128"; do timeout 100 python scripts/run_flexconv.py $s 10 2>&1 | tail -1; done | tee $out/flexconv_r2t.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "flex or sweep or model" > $out/pytest_r2t.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2t.log
