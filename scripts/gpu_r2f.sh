#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_retrieval_gpu.py tests/test_ops_gpu.py tests/test_ref_cuda_gpu.py -m gpu -q -x > $out/pytest_r2f.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2f.log
timeout 300 python scripts/op_times.py far_outlier_padding > $out/op_times_outlier_r2f.log 2>&1; head -5 $out/op_times_outlier_r2f.log
# FlexConv staging A/B (north_star: "stages each point's K neighbour features via TMA"): cp.async ring (default) vs
# TMA tile::gather4 vs per-thread LDG, same shape, ncu metrics of the fused kernel
M="gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__m_xbar2l1tex_read_bytes.sum"
for mode in ca g4 regs; do
  for shape in "32 8192 8 64 64" "8 8192 8 128 128"; do
    tag=$(echo $shape | tr ' ' '_')
    DH3D_FLEXCONV=$mode timeout 300 ncu --metrics $M --clock-control none -k regex:flexconv -s 2 -c 3 --csv --log-file $out/ncu_flexconv_ab_${mode}_${tag}.csv python scripts/run_flexconv.py $shape 3 > $out/ncu_flexconv_ab_${mode}_${tag}.log 2>&1
    DH3D_FLEXCONV=$mode timeout 120 python scripts/run_flexconv.py $shape 20 | tee -a $out/flexconv_ab_r2f.txt
  done
done
bash scripts/gpu_sanitize.sh 2>&1 | tee $out/sanitizer_summary.txt
