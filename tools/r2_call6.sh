#!/bin/bash
set -u
out=gpurun_out/r2c6
mkdir -p "$out"
for yr in 1 0; do for e in 0 1 2; do EMO_CONV_YREUSE=$yr EMO_CONV_EPI=$e timeout 200 python tools/conv_layer_bench.py > "$out/layers_yr${yr}_epi$e.txt" 2>&1; done; done
tail -2 "$out"/layers_*.txt
timeout 200 tools/gs3_lab > "$out/gs3_lab.txt" 2>&1; echo "gs3_lab rc=$?"; cat "$out/gs3_lab.txt"
timeout 120 tools/gs3_check > "$out/gs3_check.txt" 2>&1; echo "gs3_check rc=$?"; grep -v ragged "$out/gs3_check.txt"
for n in 3 4; do timeout 300 python bench.py --steps 30 --warmup 5 --quick --inflight $n > "$out/bench_if$n.json" 2> "$out/bench_if$n.err"; python -c "
import json,sys
d=json.loads(open('$out/bench_if$n.json').read().strip().splitlines()[-1]); print('inflight $n', d['value'], d['e2e']['value'])"; done
