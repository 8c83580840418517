#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 120 python scripts/run_head.py > $out/head_r2k.txt 2>&1; echo "head rc=$?"; cat $out/head_r2k.txt
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x > $out/pytest_r2k.log 2>&1; echo "tests rc=$?"; tail -5 $out/pytest_r2k.log
