#!/bin/bash
# Runs on the GPU box (under gpurun): per-launch device times of one eager forward + a full-metric
# capture of the heavy kernels; exports CSV on the box (the .ncu-rep files are too big to bring back).
#   bash scripts/ncu_step.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/launches_$tag.csv python scripts/profile_forward.py > $out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_tc16_kernel|gemm_tc_kernel|netvlad_aggregate|netvlad_tc|knn_query|flexconv_tc|flexconv_ca|three_interp|conv_pointset|fps_reg|fps_cluster|flex_pool|netvlad_project" \
    -o /tmp/prof_$tag python scripts/profile_forward.py > $out/ncu_p.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>> $out/ncu_p.log
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv -k regex:"knn_query" > $out/prof_${tag}_src_knn.csv 2>> $out/ncu_p.log
# backward kernels + NMS
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/launches_${tag}_bwd.csv python scripts/profile_backward.py > $out/ncu_lb.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"flex_scatter|gemm_tn_partial|flex_moments_c0|sgemm_kernel|row_scatter|nms_select|knn_query" \
    -o /tmp/prof_${tag}_bwd python scripts/profile_backward.py > $out/ncu_pb.log 2>&1
ncu -i /tmp/prof_${tag}_bwd.ncu-rep --page raw --csv > $out/prof_${tag}_bwd_raw.csv 2>> $out/ncu_pb.log
ls -la /tmp/prof_$tag.ncu-rep $out | tail -20
