#!/bin/bash
# Round-2 evidence call: final A/B of the epilogue selection inside ONE box, bench record, ncu launch list + full captures.
set -u
out=gpurun_out/r2c8
mkdir -p "$out"
DBG=$PWD/emoportraits_b200/csrc/libemoport_dbg.so
b() { name=$1; shift; timeout 400 env "$@" python bench.py --steps 40 --warmup 5 --quick --inflight 3 > "$out/bench_$name.json" 2> "$out/bench_$name.err"; python -c "
import json
d=json.loads(open('$out/bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d.get('latency_ms_one_frame_alone') or 0,3), d.get('gpu_launches_per_step'))" | tee -a "$out/summary.txt"; }
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b product A=1
b dbg_auto EMO_LIB=$DBG
b dbg_epi0 EMO_LIB=$DBG EMO_CONV_EPI=0
b dbg_epi1 EMO_LIB=$DBG EMO_CONV_EPI=1
b dbg_yr0 EMO_LIB=$DBG EMO_CONV_YREUSE=0
b product_again A=1
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; tail -c 600 "$out/bench_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers.csv" 2>/dev/null
timeout 200 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1; tail -1 "$out/layers_auto.txt"
for c in 48 96 192 864; do EMO_ACC_CHUNK=$c timeout 200 python tools/conv_layer_bench.py > "$out/layers_chunk$c.txt" 2>&1; tail -1 "$out/layers_chunk$c.txt"; done
timeout 300 python tools/apply_probe.py > "$out/apply_probe.txt" 2>&1
for e in auto 0 1; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_$e.txt" 2>&1; done
bash tools/profile.sh r2 > "$out/profile.log" 2>&1
cat "$out/summary.txt"
