#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python scripts/timeline.py $out/timeline_r2p.txt > $out/timeline_r2p.log 2>&1; echo "rc=$?"; tail -45 $out/timeline_r2p.log
