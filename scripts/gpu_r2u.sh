#!/bin/bash
out=gpurun_out; mkdir -p $out
for s in "32 1024 8 64 128" "32 1024 8 128 128" "32 1024 8 128 256" "32 1024 8 64 128"; do timeout 100 python scripts/run_flexconv.py $s 20 2>&1 | tail -1; done | tee $out/flexconv_r2u.txt
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2u.json > $out/bench_r2u.json 2> $out/bench_r2u.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2u.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
timeout 300 python scripts/timeline.py $out/timeline_r2u.txt > $out/timeline_r2u.log 2>&1; grep -E "flexconv|replay" $out/timeline_r2u.log | cut -c1-100
