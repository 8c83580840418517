#!/bin/bash
out=gpurun_out; mkdir -p $out
: > $out/mma_rate.txt
for l in 0 4; do for w in 0 1 2 3 4 5; do timeout 60 scripts/ubench/mma_rate $w $l >> $out/mma_rate.txt 2>&1; done; done
timeout 60 scripts/ubench/mma_rate 0 0 2 >> $out/mma_rate.txt 2>&1; timeout 60 scripts/ubench/mma_rate 2 0 2 >> $out/mma_rate.txt 2>&1
cat $out/mma_rate.txt
timeout 120 python scripts/run_head.py > $out/head_r2l.txt 2>&1; echo "head rc=$?"; cat $out/head_r2l.txt
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x > $out/pytest_r2l.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2l.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc16 -s 3 -c 1 -o /tmp/prof_head python scripts/run_head.py 262144 256 1024 3 > $out/ncu_head.log 2>&1
ncu -i /tmp/prof_head.ncu-rep --page raw --csv > $out/prof_head_raw.csv 2>> $out/ncu_head.log
ncu -i /tmp/prof_head.ncu-rep --page source --csv > $out/prof_head_src.csv 2>> $out/ncu_head.log
tail -3 $out/ncu_head.log
