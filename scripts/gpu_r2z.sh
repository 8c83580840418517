#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:knn_query -o /tmp/prof_knn python scripts/profile_knn.py ncu > $out/ncu_knn.log 2>&1
ncu -i /tmp/prof_knn.ncu-rep --page raw --csv > $out/prof_knn_raw.csv 2>> $out/ncu_knn.log
ncu -i /tmp/prof_knn.ncu-rep --page source --csv > $out/prof_knn_src.csv 2>> $out/ncu_knn.log
tail -2 $out/ncu_knn.log
