#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU).  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
# (1) every launch of one eager bench run with its device time (cold-cache, serialised: compare SHARES)
EMO_NO_LAYER_CSV=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# (2) full captures of the two roofline kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 9 -o gpurun_out/prof_conv_r1 \
    python tools/prof_kernels.py > gpurun_out/prof_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs3_cl -c 6 -o gpurun_out/prof_gs3_r1 \
    python tools/prof_kernels.py > gpurun_out/prof_gs3.log 2>&1
ls -la gpurun_out/
