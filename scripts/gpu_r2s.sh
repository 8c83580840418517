#!/bin/bash
out=gpurun_out; mkdir -p $out
for ca in 0 1; do for g in 0 3 2; do echo "== DH3D_CA=$ca DH3D_CA_G=$g"; DH3D_CA=$ca DH3D_CA_G=$g timeout 200 python scripts/exp_sorted.py 2>&1 | grep -v Warn; done; done > $out/exp_sorted.txt 2>&1
cat $out/exp_sorted.txt
