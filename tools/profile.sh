#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU): tools/summarize_profiles.py <tag> turns the outputs into the committed
# summaries.  Numbers printed under ncu are never bench values.
set -x
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
# (1) every launch of one eager bench run with its device time (cold-cache, serialised: compare SHARES)
EMO_NO_LAYER_CSV=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --eager --quick > $out/bench_under_ncu.log 2>&1
# (2) full captures of the roofline kernels (tools/prof_kernels.py launches each three times; the third is captured)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|splitk_post' -c 18 -o $out/prof_conv_$tag \
    python tools/prof_kernels.py > $out/prof_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|gn_head' -c 6 -o $out/prof_elem_$tag \
    python tools/prof_kernels.py > $out/prof_elem.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs3_cl -c 8 -o $out/prof_gs3_$tag \
    python tools/prof_kernels.py > $out/prof_gs3.log 2>&1
ls -la $out/*.ncu-rep
