#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU): tools/summarize_profiles.py <tag> turns the outputs into the committed
# summaries.  Numbers printed under ncu are never bench values.  The .ncu-rep files stay on the box (gpurun brings back at most
# 64 MiB): only their raw-page CSV exports and one single-kernel report (source page of the dominant conv) come home.
set -x
tag=${1:-r2}
out=gpurun_out
tmp=/tmp/emo_prof
mkdir -p $out $tmp
# (1) every launch of one eager bench run with its device time (cold-cache, serialised: compare SHARES)
EMO_NO_LAYER_CSV=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --eager --quick > $out/bench_under_ncu.log 2>&1
# (2) full captures of the roofline kernels (tools/prof_kernels.py launches each three times)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|splitk_post' -c 18 -o $tmp/prof_conv_$tag \
    python tools/prof_kernels.py > $out/prof_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|gn_head' -c 6 -o $tmp/prof_elem_$tag \
    python tools/prof_kernels.py > $out/prof_elem.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs3_cl -c 8 -o $tmp/prof_gs3_$tag \
    python tools/prof_kernels.py > $out/prof_gs3.log 2>&1
for n in prof_conv prof_elem prof_gs3; do ncu -i $tmp/${n}_$tag.ncu-rep --page raw --csv > $out/${n}_$tag.raw.csv 2>/dev/null; done
# the dominant conv layer alone (third launch of the first shape), small enough to travel: source page / SASS of the product
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 2 -c 1 -o $out/prof_conv_dominant_$tag \
    python tools/prof_kernels.py > /dev/null 2>&1
ls -la $tmp $out
du -sh $out
