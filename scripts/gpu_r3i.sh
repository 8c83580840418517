#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_r3i.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r3i.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 200 > $out/smi_r3i.txt 2>&1 &
SMI=$!
timeout 900 python bench.py --op-table $out/op_table_r3i.json > $out/bench_r3i.json 2> $out/bench_r3i.err; echo "bench rc=$?"
kill $SMI
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3i.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f roofline %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], {k: d['roofline'][k] for k in ('achieved','frac','kernel','traffic')}))
print('cpu', d['cpu_baseline']['value'], 'sustained', d.get('sustained',{}).get('value'), 'launches/step', d['gpu_launches']/d['steps'])
PY
timeout 900 bash scripts/ncu_fwd.sh r3i "gemm_tc16_kernel|gemm_head16|gemm_join16|gemm_chain16|netvlad_tc2|netvlad_tail|knn_query|knn_sort|flexconv_ca|three_interp|fps_cluster|se_pool_excite|conv_pointset|flex_pool|group_point" "flexconv_ca_kernel|gemm_head16|gemm_chain16"
timeout 300 python scripts/timeline.py $out/timeline_r3i.txt > $out/timeline_r3i.log 2>&1; tail -1 $out/timeline_r3i.txt | cut -c1-120
TOOLS="memcheck synccheck racecheck" PARTS="gemm three_nn" bash scripts/gpu_sanitize.sh 2>&1 | tee $out/sanitizer_summary_r3i.txt
