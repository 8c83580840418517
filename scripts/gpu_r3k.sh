#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 120 python scripts/run_netvlad.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:netvlad_tail -s 3 -c 1 -o /tmp/prof_tail python scripts/run_netvlad.py 3 > $out/ncu_tail.log 2>&1
ncu -i /tmp/prof_tail.ncu-rep --page source --csv > $out/prof_tail_src.csv 2>> $out/ncu_tail.log
tail -2 $out/ncu_tail.log
