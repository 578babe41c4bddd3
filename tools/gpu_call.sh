#!/bin/bash
# Round-2 evidence call: parity suite, bench record, per-layer tables per kernel form, probes, ncu launch list + full captures.
set -u
out=gpurun_out/r2c12
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt
DBG=$PWD/emoportraits_b200/csrc/libemoport_dbg.so
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b() { name=$1; shift; timeout 400 env "$@" python bench.py --steps 40 --warmup 5 --quick > "$out/bench_$name.json" 2> "$out/bench_$name.err"; python -c "
import json
d=json.loads(open('$out/bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d.get('latency_ms_one_frame_alone') or 0,3), d.get('gpu_launches_per_step'))" | tee -a "$out/summary.txt"; }
b product A=1
b instrumented_auto EMO_LIB=$DBG
b instrumented_inwarp_final_phase EMO_LIB=$DBG EMO_CONV_EPI=0
b instrumented_tma_epilogue_everywhere EMO_LIB=$DBG EMO_CONV_EPI=1
b instrumented_per_tap_main_loop EMO_LIB=$DBG EMO_CONV_YREUSE=0
b product_again A=1
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; tail -c 300 "$out/bench_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers.csv" 2>/dev/null
timeout 300 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1; tail -1 "$out/layers_auto.txt"
EMO_CONV_EPI=0 timeout 300 python tools/conv_layer_bench.py > "$out/layers_epi0.txt" 2>&1
EMO_CONV_EPI=1 timeout 300 python tools/conv_layer_bench.py > "$out/layers_epi1.txt" 2>&1
EMO_CONV_YREUSE=0 timeout 300 python tools/conv_layer_bench.py > "$out/layers_pertap.txt" 2>&1
for c in 48 192; do EMO_ACC_CHUNK=$c timeout 300 python tools/conv_layer_bench.py > "$out/layers_chunk$c.txt" 2>&1; done
timeout 300 python tools/apply_probe.py > "$out/apply_probe.txt" 2>&1
timeout 200 tools/gs3_lab > "$out/gs3_lab.txt" 2>&1
timeout 120 tools/gs3_check > "$out/gs3_check.txt" 2>&1
for e in auto 0 1; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_$e.txt" 2>&1; done
bash tools/profile.sh r2 > "$out/profile.log" 2>&1
du -sh gpurun_out
cat "$out/summary.txt"
