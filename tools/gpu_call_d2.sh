#!/bin/bash
# round 2, second session, call 2: brick-order adoption check (lab + C-ABI check + op tests), bench record, gs3 ncu capture
set -u
out=gpurun_out/r2d2
mkdir -p "$out"
timeout 200 tools/gs3_lab --quick > "$out/gs3_lab.txt" 2>&1; echo "lab rc=$?" | tee -a "$out/summary.txt"
timeout 200 tools/gs3_check > "$out/gs3_check.txt" 2>&1; echo "check rc=$?" | tee -a "$out/summary.txt"
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "grid_sample" -rA > "$out/pytest_gs.txt" 2>&1; echo "pytest gs rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_gs.txt" >> "$out/summary.txt"
timeout 600 python -m pytest tests/test_masks.py -q -m gpu -rA > "$out/pytest_masks.txt" 2>&1; echo "pytest masks rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_masks.txt" | tail -12 >> "$out/summary.txt"
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; echo "bench rc=$?" | tee -a "$out/summary.txt"
python - <<'P' | tee -a "$out/summary.txt"
import json
d=json.loads(open('gpurun_out/r2d2/bench_full.json').read().strip().splitlines()[-1])
print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'lat', d.get('latency_ms_one_frame_alone'))
for k,v in d['roofline_grid_sample3d'].items():
    if isinstance(v,dict) and 'ms' in v: print(k, round(v['ms']*1e3,1),'us', round(v['frac'],3), 'dirty', round(v['ms_dirty_flush']*1e3,1), round(v['frac_dirty_flush'],3))
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gs3_cl -c 8 -o /tmp/prof_gs3_r2b python tools/prof_kernels.py > "$out/prof_gs3.log" 2>&1
ncu -i /tmp/prof_gs3_r2b.ncu-rep --page raw --csv > "$out/prof_gs3_r2b.raw.csv" 2>/dev/null
for e in auto 1; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_$e.txt" 2>&1; done
rm -f gpurun_out/stage_parity.txt gpurun_out/parity_*.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest full rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
ls -la "$out"; cat "$out/summary.txt"
