#!/bin/bash
# Runs on the GPU box (under gpurun): per-launch device times of one eager forward + a full-metric capture of
# the kernels named in $2 (regex), raw + source pages exported as CSV (the .ncu-rep stays on the box).
#   bash scripts/ncu_fwd.sh <tag> [kernel-regex] [source-page-regex]
tag=${1:-r1}
kre=${2:-"gemm_tc16_kernel|netvlad_tc|knn_query|flexconv_ca|three_interp|fps_cluster|se_pool_excite|l2norm|conv_pointset"}
sre=${3:-"flexconv_ca|se_pool_excite"}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/launches_$tag.csv python scripts/profile_forward.py > $out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$kre" \
    -o /tmp/prof_$tag python scripts/profile_forward.py > $out/ncu_p.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>> $out/ncu_p.log
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv -k regex:"$sre" > $out/prof_${tag}_src.csv 2>> $out/ncu_p.log
ls -la /tmp/prof_$tag.ncu-rep | tail -2
