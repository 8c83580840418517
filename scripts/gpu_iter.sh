#!/bin/bash
# GPU iteration helper (under gpurun): focused tests, FlexConv A/B microbench, then the full check.
tag=${1:-it}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "flex_conv_pm or fused_epilogue or three_interpolate or se_pool or strided_group" > $out/pytest_focus_$tag.log 2>&1; echo "focus rc=$?" | tee $out/summary_$tag.txt
tail -n 15 $out/pytest_focus_$tag.log
{
for shape in "32 8192 8 64 64" "32 8192 8 32 64" "32 1024 8 64 128" "32 1024 8 128 128" "32 1024 8 128 256" "8 8192 8 128 128"; do
  echo "== $shape  k8"; timeout 120 python scripts/run_flexconv.py $shape 20
  echo "== $shape  generic"; DH3D_FLEXCONV_K8=0 timeout 120 python scripts/run_flexconv.py $shape 20
done
echo "== 32 8192 8 32 64 ca(min_din=32) k8"; DH3D_FLEXCONV_CA_MIN_DIN=32 timeout 120 python scripts/run_flexconv.py 32 8192 8 32 64 20
} > $out/flexconv_ab_$tag.txt 2>&1
cat $out/flexconv_ab_$tag.txt
bash scripts/gpu_check.sh $tag
