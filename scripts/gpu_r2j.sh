#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q -x -k "netvlad or oracle or rowdot" > $out/pytest_r2j.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2j.log
timeout 900 bash scripts/ncu_fwd.sh r2j "gemm_tc16_kernel|gemm_join16|netvlad_tc2|netvlad_project|netvlad_head|netvlad_finalize|knn_query|knn_sort|flexconv_ca|three_interp|fps_cluster|se_pool_excite|conv_pointset|flex_pool" "flexconv_ca_kernel"
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-sensitivity --no-modes --op-table $out/op_table_r2j.json > $out/bench_r2j.json 2> $out/bench_r2j.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2j.json'))
print('value %.0f  ms/step %.4f' % (d['value'], d['ms_per_step']))
for r in d['op_roofline'][:6]: print('  %-60s %8.4f ms' % (r['op'], r['ms']))
PY
