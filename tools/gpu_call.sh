#!/bin/bash
# Round-2 call 10: fp16-plane accumulation chunk 48 (parity + A/B in one box), bench record with graph-timed rooflines.
set -u
out=gpurun_out/r2c10
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b() { name=$1; c16=$2; shift; shift; timeout 400 python -c "
import sys, runpy
sys.path.insert(0, '.')
import emoportraits_b200.ops as o
o.ACC_CHUNK_F16 = $c16
sys.argv = ['bench.py', '--steps', '40', '--warmup', '5', '--quick', '--inflight', '3']
runpy.run_path('bench.py', run_name='__main__')" > "$out/bench_$name.json" 2> "$out/bench_$name.err"; python -c "
import json
d=json.loads(open('$out/bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d.get('latency_ms_one_frame_alone') or 0,3), d.get('gpu_launches_per_step'))" | tee -a "$out/summary.txt"; }
b f16chunk48 48
b f16chunk24 24
b f16chunk96 96
b f16chunk48_again 48
timeout 900 python bench.py > "$out/bench_full.json" 2> "$out/bench_full.err"; tail -c 300 "$out/bench_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers.csv" 2>/dev/null
cat "$out/summary.txt"
