#!/bin/bash
# round 2, second session: final evidence call (parity suite, bench record, timelines, probes, ncu launch list + full captures)
set -u
out=gpurun_out/r2d3
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt gpurun_out/parity_*.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"; tail -c 300 "$out/bench_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers.csv" 2>/dev/null
python - <<'P' | tee -a "$out/summary.txt"
import json
d=json.loads(open('gpurun_out/r2d3/bench_full.json').read().strip().splitlines()[-1])
print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', d.get('latency_ms_one_frame_alone'), 'launches', d.get('gpu_launches_per_step'))
print('conv', d['roofline']['launch_us'], d['roofline']['frac'])
for k,v in d['roofline_grid_sample3d'].items():
    if isinstance(v,dict) and 'ms' in v: print(k, round(v['ms']*1e3,1),'us', round(v['frac'],3), 'dirty', round(v['ms_dirty_flush']*1e3,1), round(v['frac_dirty_flush'],3))
P
timeout 300 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1; tail -1 "$out/layers_auto.txt"
DBG=$PWD/emoportraits_b200/csrc/libemoport_dbg.so
for e in 0 1; do EMO_LIB=$DBG EMO_CONV_EPI=$e timeout 300 python tools/conv_layer_bench.py > "$out/layers_epi$e.txt" 2>&1; done
for n in 1 2 4; do timeout 300 python bench.py --steps 40 --warmup 5 --quick --inflight $n > "$out/bench_inflight$n.json" 2>/dev/null; python -c "
import json
d=json.loads(open('$out/bench_inflight$n.json').read().strip().splitlines()[-1]); print('inflight $n', round(d['value'],1), round(d['e2e']['value'],1))" | tee -a "$out/summary.txt"; done
for e in auto 0 1; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_$e.txt" 2>&1; done
timeout 200 tools/gs3_lab > "$out/gs3_lab.txt" 2>&1
timeout 200 tools/gs3_check > "$out/gs3_check.txt" 2>&1
timeout 300 python tools/apply_probe.py > "$out/apply_probe.txt" 2>&1
bash tools/profile.sh r2 > "$out/profile.log" 2>&1
du -sh gpurun_out
cat "$out/summary.txt"
