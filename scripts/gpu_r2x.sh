#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_r2x.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r2x.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 200 > $out/smi_r2x.txt 2>&1 &
SMI=$!
timeout 900 python bench.py --op-table $out/op_table_r2x.json > $out/bench_r2x.json 2> $out/bench_r2x.err; echo "bench rc=$?"
kill $SMI
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2x.json'))
print('value %.0f  ms/step %.4f e2e %.0f roofline %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], {k: d['roofline'][k] for k in ('achieved','frac','kernel','traffic')}))
print('cpu', d['cpu_baseline']['value'], 'sustained', d.get('sustained',{}).get('value'))
PY
timeout 900 bash scripts/ncu_fwd.sh r2x "gemm_tc16_kernel|gemm_head16|gemm_join16|netvlad_tc2|netvlad_tail|knn_query|knn_sort|flexconv_ca|three_interp|fps_cluster|se_pool_excite|conv_pointset|flex_pool|group_point" "flexconv_ca_kernel|gemm_head16"
timeout 300 python scripts/timeline.py $out/timeline_r2x.txt > $out/timeline_r2x.log 2>&1; tail -3 $out/timeline_r2x.log | cut -c1-120
