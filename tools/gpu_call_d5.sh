#!/bin/bash
# co-residency experiment: can the HBM-bound apply passes of one frame run beside the tensor-bound convolutions of another?
set -u
out=gpurun_out/r2d5
mkdir -p "$out"
DBG=$PWD/emoportraits_b200/csrc/libemoport_dbg.so
b() { name=$1; shift; timeout 400 env EMO_LIB=$DBG "$@" python bench.py --steps 40 --warmup 5 --quick > "$out/bench_$name.json" 2> "$out/bench_$name.err"; python -c "
import json
d=json.loads(open('$out/bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d.get('latency_ms_one_frame_alone') or 0,3))" | tee -a "$out/summary.txt"; }
b base3 A=1
b apply128 EMO_APPLY_THREADS=128
b apply128_smem200 EMO_APPLY_THREADS=128 EMO_CONV_SMEM_KB=200
b apply64_smem200 EMO_APPLY_THREADS=64 EMO_CONV_SMEM_KB=200
b smem200 EMO_CONV_SMEM_KB=200
b base3_again A=1
cat "$out/summary.txt"
