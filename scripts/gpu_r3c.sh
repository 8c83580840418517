#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 120 python scripts/run_chain.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_chain16 -s 3 -c 1 -o /tmp/prof_chain python scripts/run_chain.py 262144 3 > $out/ncu_chain.log 2>&1
ncu -i /tmp/prof_chain.ncu-rep --page raw --csv > $out/prof_chain_raw.csv 2>> $out/ncu_chain.log
ncu -i /tmp/prof_chain.ncu-rep --page source --csv > $out/prof_chain_src.csv 2>> $out/ncu_chain.log
tail -2 $out/ncu_chain.log
