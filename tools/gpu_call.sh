#!/bin/bash
set -u
out=gpurun_out/r2c7
mkdir -p "$out"
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
timeout 200 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1
EMO_CONV_EPI=0 timeout 200 python tools/conv_layer_bench.py > "$out/layers_epi0.txt" 2>&1
EMO_CONV_EPI=1 timeout 200 python tools/conv_layer_bench.py > "$out/layers_epi1.txt" 2>&1
for n in 3 1; do timeout 300 python bench.py --steps 30 --warmup 5 --quick --inflight $n > "$out/bench_if$n.json" 2> "$out/bench_if$n.err"; python -c "
import json,sys
d=json.loads(open('$out/bench_if$n.json').read().strip().splitlines()[-1]); print('inflight $n', d['value'], d['e2e']['value'], d.get('latency_ms_one_frame_alone'), d.get('gpu_launches_per_step'))" | tee -a "$out/summary.txt"; done
tail -8 "$out/summary.txt"; for f in "$out"/layers_*.txt; do tail -1 $f; done; cat "$out/layers_auto.txt"
