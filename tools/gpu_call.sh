#!/bin/bash
# Round-2 call 11: fused finalize with parallel part loads, register statistics in the resamplers; per-layer table incl. ResNet tails.
set -u
out=gpurun_out/r2c11
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
timeout 300 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1; tail -6 "$out/layers_auto.txt"
for n in 3 1; do timeout 300 python bench.py --steps 40 --warmup 5 --quick --inflight $n > "$out/bench_if$n.json" 2> "$out/bench_if$n.err"; python -c "
import json
d=json.loads(open('$out/bench_if$n.json').read().strip().splitlines()[-1]); print('inflight $n', round(d['value'],1), round(d['e2e']['value'],1), round(d.get('latency_ms_one_frame_alone') or 0,3), d.get('gpu_launches_per_step'))" | tee -a "$out/summary.txt"; done
EMO_NO_LAYER_CSV=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --eager --quick > "$out/bench_under_ncu.log" 2>&1
cat "$out/summary.txt"
