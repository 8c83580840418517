#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 200 python scripts/exp_sorted.py 2>&1 | grep -v Warn | grep random > $out/exp_r2w.txt; cat $out/exp_r2w.txt
for s in "8 8192 8 128 128" "8 8192 16 128 128" "8 8192 32 128 128"; do timeout 100 python scripts/run_flexconv.py $s 20 2>&1 | tail -1; done | tee $out/flexconv_r2w.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "flex or sweep" > $out/pytest_r2w.log 2>&1; echo "tests rc=$?"; tail -2 $out/pytest_r2w.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2w.json > $out/bench_r2w.json 2> $out/bench_r2w.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2w.json'))
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
