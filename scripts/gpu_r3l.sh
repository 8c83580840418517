#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 120 python scripts/run_netvlad.py 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q -x -k "netvlad or model or retriev" > $out/pytest_r3l.log 2>&1; echo "tests rc=$?"; tail -2 $out/pytest_r3l.log
timeout 300 python scripts/timeline.py $out/timeline_r3l.txt > $out/timeline_r3l.log 2>&1; tail -4 $out/timeline_r3l.txt | cut -c1-110
