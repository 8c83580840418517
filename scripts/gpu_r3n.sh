#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "fps or knn_sort" > $out/pytest_r3n.log 2>&1; echo "fps tests rc=$?"; tail -6 $out/pytest_r3n.log
timeout 120 python - <<PY
import torch, sys
sys.path.insert(0,'.')
from dh3d_b200 import ops
from bench import synth_clouds
pts = synth_clouds(32, 8192, 0).cuda()
ws = ops.knn_sort(pts)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
print('fps exhaustive %.4f ms  presorted %.4f ms  sort %.4f ms' % (t(lambda: ops.farthest_point_sample(1024, pts)), t(lambda: ops.farthest_point_sample(1024, pts, sorted_ws=ws)), t(lambda: ops.knn_sort(pts))))
PY
timeout 900 python -m pytest tests -m gpu -q -x -k "model or retriev or three_nn or layers" > $out/pytest_r3n2.log 2>&1; echo "model tests rc=$?"; tail -3 $out/pytest_r3n2.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes > $out/bench_r3n.json 2> $out/bench_r3n.err; echo "bench rc=$?"; tail -2 $out/bench_r3n.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3n.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
timeout 300 python scripts/timeline.py $out/timeline_r3n.txt > $out/timeline_r3n.log 2>&1; head -20 $out/timeline_r3n.txt | cut -c1-100; tail -1 $out/timeline_r3n.txt
