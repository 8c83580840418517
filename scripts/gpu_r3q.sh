#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -k "three_nn or knn or fps or model or sweep or nms or ref_cuda or retriev" > $out/pytest_r3q.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r3q.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes > $out/bench_r3q.json 2> $out/bench_r3q.err; echo "bench rc=$?"; tail -2 $out/bench_r3q.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3q.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f launches/step %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']/d['steps']))
PY
timeout 300 python scripts/timeline.py $out/timeline_r3q.txt > $out/timeline_r3q.log 2>&1; sed -n 6,18p $out/timeline_r3q.txt | cut -c1-100; tail -1 $out/timeline_r3q.txt
