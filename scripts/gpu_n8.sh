#!/bin/bash
# 8-GPU run of the retrieval job (BASELINE configs[3]: 4096 clouds sharded over 8 GPUs), launched the way the driver does
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/topo_n8.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload retrieval > $out/bench_r3r_retrieval_n8.json 2> $out/bench_r3r_retrieval_n8.err; echo "retrieval n8 rc=$?"; tail -3 $out/bench_r3r_retrieval_n8.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r3r_retrieval_n8.json')); print('value %.0f e2e %.0f ms_total %.2f gather+retrieval %.3f ms' % (d['value'], d['e2e']['value'], d['ms_total'], d['gather_plus_retrieval_ms']))
PY
