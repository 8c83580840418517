#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_r3p.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r3p.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 200 > $out/smi_r3p.txt 2>&1 &
SMI=$!
timeout 900 python bench.py --op-table $out/op_table_r3p.json > $out/bench_r3p.json 2> $out/bench_r3p.err; echo "bench rc=$?"
kill $SMI
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3p.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f roofline %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], {k: d['roofline'][k] for k in ('achieved','frac','kernel')}))
print('cpu', d['cpu_baseline']['value'], 'sustained', d.get('sustained',{}).get('value'), 'launches/step', d['gpu_launches']/d['steps'])
print('sens', {k: v.get('clouds_per_s') for k,v in d['data_sensitivity'].items() if isinstance(v, dict)})
PY
timeout 900 bash scripts/ncu_fwd.sh r3p "gemm_tc16_kernel|gemm_head16|gemm_join16|gemm_chain16|netvlad_tc2|netvlad_tail|knn_query|knn_sort|flexconv_ca|three_interp|fps_|se_pool_excite|conv_pointset|flex_pool|group_point" "fps_bucket|flexconv_ca_kernel"
timeout 300 python scripts/timeline.py $out/timeline_r3p.txt > $out/timeline_r3p.log 2>&1; tail -1 $out/timeline_r3p.txt | cut -c1-120
TOOLS="memcheck synccheck racecheck" PARTS="fps knn" bash scripts/gpu_sanitize.sh 2>&1 | tee $out/sanitizer_summary_r3p.txt
