#!/bin/bash
# 2-GPU validation of the multi-rank bench paths (launched the way the driver does)
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/topo_n2.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > $out/bench_r3h_n2.json 2> $out/bench_r3h_n2.err; echo "bench n2 rc=$?"; tail -3 $out/bench_r3h_n2.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r3h_n2.json'))
print('value %.0f  ms/step %.4f  sustained %.0f' % (d['value'], d['ms_per_step'], (d.get('sustained') or {}).get('value', 0)))
print('e2e', json.dumps(d['e2e'])[:700])
print('modes', json.dumps(d['e2e_modes'])[:900])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload retrieval > $out/bench_r3h_retrieval_n2.json 2> $out/bench_r3h_retrieval_n2.err; echo "retrieval n2 rc=$?"; tail -3 $out/bench_r3h_retrieval_n2.err
cut -c1-1800 $out/bench_r3h_retrieval_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 0 | cut -c1-300
timeout 600 python bench.py --workload retrieval > $out/bench_r3h_retrieval_n1.json 2> $out/bench_r3h_retrieval_n1.err; echo "retrieval n1 rc=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_r3h_retrieval_n1.json','gpurun_out/bench_r3h_retrieval_n2.json'):
    d=json.load(open(f)); print(f, 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))
PY
