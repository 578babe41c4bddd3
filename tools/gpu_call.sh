#!/bin/bash
# Round-2 evidence call: parity suite, accumulation-chunk A/B inside ONE box, bench record, per-layer tables, probes, ncu.
set -u
out=gpurun_out/r2c9
mkdir -p "$out"
DBG=$PWD/emoportraits_b200/csrc/libemoport_dbg.so
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b() { name=$1; chunk=$2; shift; shift; timeout 400 env "$@" python -c "
import sys, runpy
sys.path.insert(0, '.')
import emoportraits_b200.ops as o
o.ACC_CHUNK_BF16 = $chunk
sys.argv = ['bench.py', '--steps', '40', '--warmup', '5', '--quick', '--inflight', '3']
runpy.run_path('bench.py', run_name='__main__')" > "$out/bench_$name.json" 2> "$out/bench_$name.err"; python -c "
import json
d=json.loads(open('$out/bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d.get('latency_ms_one_frame_alone') or 0,3), d.get('gpu_launches_per_step'))" | tee -a "$out/summary.txt"; }
b chunk96 96 A=1
b chunk48 48 A=1
b chunk192 192 A=1
b chunk96_again 96 A=1
b chunk96_rowreuse_off 96 EMO_LIB=$DBG EMO_CONV_YREUSE=0
b chunk96_epi0 96 EMO_LIB=$DBG EMO_CONV_EPI=0
b chunk96_epi1 96 EMO_LIB=$DBG EMO_CONV_EPI=1
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; tail -c 400 "$out/bench_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers.csv" 2>/dev/null
timeout 200 python tools/conv_layer_bench.py > "$out/layers_auto.txt" 2>&1; tail -1 "$out/layers_auto.txt"
EMO_CONV_EPI=0 timeout 200 python tools/conv_layer_bench.py > "$out/layers_epi0.txt" 2>&1
EMO_CONV_EPI=1 timeout 200 python tools/conv_layer_bench.py > "$out/layers_epi1.txt" 2>&1
EMO_CONV_YREUSE=0 timeout 200 python tools/conv_layer_bench.py > "$out/layers_pertap.txt" 2>&1
for c in 48 192 864; do EMO_ACC_CHUNK=$c timeout 200 python tools/conv_layer_bench.py > "$out/layers_chunk$c.txt" 2>&1; tail -1 "$out/layers_chunk$c.txt"; done
timeout 300 python tools/apply_probe.py > "$out/apply_probe.txt" 2>&1
timeout 200 tools/gs3_lab > "$out/gs3_lab.txt" 2>&1
timeout 120 tools/gs3_check > "$out/gs3_check.txt" 2>&1
for e in auto 0 1; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_$e.txt" 2>&1; done
bash tools/profile.sh r2 > "$out/profile.log" 2>&1
rm -f gpurun_out/*.ncu-rep.tmp
du -sh gpurun_out
cat "$out/summary.txt"
