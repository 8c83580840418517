#!/bin/bash
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "fps" 2>&1 | tail -2
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
print('fps exhaustive %.4f ms  presorted %.4f ms' % (t(lambda: ops.farthest_point_sample(1024, pts)), t(lambda: ops.farthest_point_sample(1024, pts, sorted_ws=ws))))
PY
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes > gpurun_out/bench_r3o.json 2> gpurun_out/bench_r3o.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3o.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
