#!/bin/bash
# r3t: InFlightForward as the public API behind bench.py's two-steps-in-flight loop
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "flight or graph_replay or overlap" > $out/pytest_r3t.log 2>&1; echo "tests rc=$?"; tail -2 $out/pytest_r3t.log
timeout 600 python bench.py --op-table $out/op_table_r3t.json > $out/bench_r3t.json 2> $out/bench_r3t.err; echo "bench rc=$?"; tail -2 $out/bench_r3t.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3t.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f one-in-flight %s sustained %.0f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('one_step_in_flight') or {}).get('value'), (d.get('sustained') or {}).get('value', 0), d['gpu_launches']))
PY
timeout 200 python bench.py --workload retrieval > $out/bench_r3t_retrieval_n1.json 2> $out/bench_r3t_retrieval_n1.err; echo "retrieval rc=$?"
python - <<PY
import json
for f in ('bench_r3t_retrieval_n1',):
    d=[json.loads(l) for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][0]
    print(f, 'value %.0f e2e %.0f launches %d self-first %.3f' % (d['value'], d['e2e']['value'], d['gpu_launches'], d['config']['retrieval_self_match_first']))
PY
