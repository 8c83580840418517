#!/bin/bash
# 2-GPU check of the two-lane timed region (+ all-gather at its end), launched the way the driver does
out=gpurun_out; mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_r3t_n2.json 2> $out/bench_r3t_n2.err; echo "bench n2 rc=$?"; tail -3 $out/bench_r3t_n2.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3t_n2.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f  sustained %.0f one-in-flight %.0f e2e %.0f' % (d['value'], d['ms_per_step'], (d.get('sustained') or {}).get('value', 0), d['one_step_in_flight']['value'], d['e2e']['value']))
PY
