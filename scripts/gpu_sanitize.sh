#!/bin/bash
# compute-sanitizer over one small launch of every hand-synchronised kernel; logs -> gpurun_out/sanitizer_*.log
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 120 python scripts/sanitize_target.py > $out/sanitizer_plain.log 2>&1; echo "plain rc=$?"
for tool in ${TOOLS:-memcheck synccheck racecheck initcheck}; do
  for part in ${PARTS:-knn fps three_nn flexconv gather gemm netvlad}; do
    timeout 420 $CS --tool $tool --print-limit 20 python scripts/sanitize_target.py $part > $out/sanitizer_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitizer_${tool}_${part}.log | tail -1)"
  done
done
