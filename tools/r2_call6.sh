#!/bin/bash
set -u
out=gpurun_out/r2c6
mkdir -p "$out"
for yr in 1 0; do for e in 0 1 2; do EMO_CONV_YREUSE=$yr EMO_CONV_EPI=$e timeout 200 python tools/conv_layer_bench.py > "$out/layers_yr${yr}_epi$e.txt" 2>&1; done; done
tail -2 "$out"/layers_*.txt
