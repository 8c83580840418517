#!/bin/bash
out=gpurun_out; mkdir -p $out
for s in "32 1024 8 128 256" "32 1024 8 128 128" "32 1024 8 128 256"; do timeout 100 python scripts/run_flexconv.py $s 20 2>&1 | tail -1; done | tee $out/flexconv_r3m.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "flex or sweep or model or layers" > $out/pytest_r3m.log 2>&1; echo "tests rc=$?"; tail -2 $out/pytest_r3m.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes > $out/bench_r3m.json 2> $out/bench_r3m.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3m.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f' % (d['value'], d['ms_per_step']))
for r in d['op_roofline']:
    if 'Co256' in r['op']: print('  %-60s %8.4f ms' % (r['op'], r['ms']))
PY
