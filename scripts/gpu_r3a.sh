#!/bin/bash
TOOLS="memcheck synccheck racecheck" PARTS="knn three_nn flexconv gemm netvlad" bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitizer_summary_r3a.txt
