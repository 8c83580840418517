#!/bin/bash
# Final validation of round 2 (r3s): ticket-ordered NetVLAD tail + two steps in flight in bench.py.
# Ordered by priority (the GPU budget may cut the call short); every leg writes its own file.
out=gpurun_out; mkdir -p $out
# 1. gate: the re-written tail (parity, B > 32, four concurrent streams).  If it fails, the rest runs on the previous tail
#    (libdh3d_b200_prevtail.so was an ad-hoc build of the previous commit's netvlad.cu for this one call; the gate passed).
timeout 200 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k netvlad > $out/pytest_r3s_netvlad.log 2>&1; rc=$?
echo "netvlad gate rc=$rc"; tail -2 $out/pytest_r3s_netvlad.log
if [ $rc -ne 0 ]; then
  echo "FALLBACK: previous tail build"; cp dh3d_b200/build/libdh3d_b200_prevtail.so dh3d_b200/libdh3d_b200.so; echo prevtail > $out/r3s_FALLBACK
else
  rm -f $out/r3s_FALLBACK
fi
# 2. the default bench line
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 200 > $out/smi_r3s.txt 2>&1 &
SMI=$!
timeout 600 python bench.py --op-table $out/op_table_r3s.json > $out/bench_r3s.json 2> $out/bench_r3s.err; echo "bench rc=$?"; tail -2 $out/bench_r3s.err
kill $SMI
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_r3s.json') if l.startswith('{')][0]
print('value %.0f  ms/step %.4f e2e %.0f one-in-flight %s sustained %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('one_step_in_flight') or {}).get('value'), (d.get('sustained') or {}).get('value', 0)))
print([ (r['op'][:40], r['ms']) for r in d['op_roofline'] if 'netvlad' in r['op']])
PY
# 3. the whole GPU suite + smoke
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_r3s.log 2>&1; echo "tests rc=$?"; tail -3 $out/pytest_r3s.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
# 4. retrieval job on one GPU, two lanes and one (pairs with profiles/bench_r3r_retrieval_n8.json, measured with one)
timeout 200 python bench.py --workload retrieval > $out/bench_r3s_retrieval_n1.json 2> $out/bench_r3s_retrieval_n1.err; echo "retrieval rc=$?"
timeout 200 python bench.py --workload retrieval --in-flight 1 > $out/bench_r3s_retrieval_n1_serial.json 2> $out/bench_r3s_retrieval_n1_serial.err; echo "retrieval serial rc=$?"
python - <<PY
import json
for f in ('bench_r3s_retrieval_n1','bench_r3s_retrieval_n1_serial'):
    d=[json.loads(l) for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][0]
    print(f, 'value %.0f e2e %.0f launches %d' % (d['value'], d['e2e']['value'], d['gpu_launches']))
PY
# 5. in-flight experiment (incl. the bit-identity check) and the timeline of one replay
timeout 200 python scripts/exp_inflight.py > $out/inflight_r3s.txt 2>&1; tail -5 $out/inflight_r3s.txt
timeout 200 python scripts/timeline.py $out/timeline_r3s.txt > $out/timeline_r3s.log 2>&1; tail -4 $out/timeline_r3s.txt | cut -c1-120
# 6. launch list + full-metric capture of the tail kernel; sanitizer over the NetVLAD launches
timeout 400 bash scripts/ncu_fwd.sh r3s "netvlad_tail|netvlad_tc2" "netvlad_tail"
TOOLS="memcheck racecheck synccheck" PARTS="netvlad" timeout 400 bash scripts/gpu_sanitize.sh 2>&1 | tee $out/sanitizer_summary_r3s.txt
