#!/bin/bash
timeout 120 python scripts/run_chain.py 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -k "chain" 2>&1 | tail -2
