#!/bin/bash
# round 2, second session, call 3: bias-initialised accumulation (conv) + 4x4x4 bricks (grid_sample_3d): parity suite, bench, timelines, per-layer table
set -u
out=gpurun_out/r2d2b
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt gpurun_out/parity_*.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest full rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
for e in auto 0 1; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_$e.txt" 2>&1; done
timeout 300 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1; tail -1 "$out/layers_auto.txt" | tee -a "$out/summary.txt"
timeout 200 tools/gs3_check > "$out/gs3_check.txt" 2>&1; echo "check rc=$?" | tee -a "$out/summary.txt"
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
python - <<'P' | tee -a "$out/summary.txt"
import json
d=json.loads(open('gpurun_out/r2d2b/bench_full.json').read().strip().splitlines()[-1])
print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', d.get('latency_ms_one_frame_alone'), 'conv', d['roofline']['launch_us'], d['roofline']['frac'], 'all', d['roofline']['all_convs']['graph_timed']['frac'], 's2', d['stage2']['value'])
for k,v in d['roofline_grid_sample3d'].items():
    if isinstance(v,dict) and 'ms' in v: print(k, round(v['ms']*1e3,1),'us', round(v['frac'],3), 'dirty', round(v['ms_dirty_flush']*1e3,1), round(v['frac_dirty_flush'],3))
P
cat "$out/summary.txt"
