#!/bin/bash
tag=${1:-knn}
out=gpurun_out
mkdir -p $out
python scripts/profile_knn.py > $out/knn_time_$tag.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"knn_query" \
    -o /tmp/prof_$tag python scripts/profile_knn.py ncu > $out/ncu_knn.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>> $out/ncu_knn.log
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv > $out/prof_${tag}_src.csv 2>> $out/ncu_knn.log
ncu -i /tmp/prof_$tag.ncu-rep --page details --csv > $out/prof_${tag}_details.csv 2>> $out/ncu_knn.log
cat $out/knn_time_$tag.txt
