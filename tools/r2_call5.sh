#!/bin/bash
# Round-2 GPU call 5: row-reuse conv main loop (halo A tile), post-op launch collapse, grid_sample laboratory.
set -u
out=gpurun_out/r2c5
mkdir -p "$out"
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
EMO_CONV_EPI=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_conv_modes_gpu.py -q -m gpu > "$out/pytest_epi2.txt" 2>&1; echo "pytest epi2 rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_epi2.txt" >> "$out/summary.txt"
for yr in 1 0; do for e in 0 1 2; do EMO_CONV_YREUSE=$yr timeout 120 python tools/conv_timeline.py $e > "$out/timeline_yr${yr}_epi$e.txt" 2>&1; done; done
b() { name=$1; shift; timeout 400 env "$@" > "$out/bench_$name.json" 2> "$out/bench_$name.err"; cp gpurun_out/conv_layers.csv "$out/conv_layers_$name.csv" 2>/dev/null; }
b yr1_epi1_if2 EMO_CONV_YREUSE=1 EMO_CONV_EPI=1 python bench.py --steps 30 --warmup 5 --quick
b yr1_epi2_if2 EMO_CONV_YREUSE=1 EMO_CONV_EPI=2 python bench.py --steps 30 --warmup 5 --quick
b yr0_epi1_if2 EMO_CONV_YREUSE=0 EMO_CONV_EPI=1 python bench.py --steps 30 --warmup 5 --quick
b yr1_epi0_if2 EMO_CONV_YREUSE=1 EMO_CONV_EPI=0 python bench.py --steps 30 --warmup 5 --quick
b yr1_epi1_if1 EMO_CONV_YREUSE=1 EMO_CONV_EPI=1 python bench.py --steps 30 --warmup 5 --quick --inflight 1
b yr1_epi1_if3 EMO_CONV_YREUSE=1 EMO_CONV_EPI=1 python bench.py --steps 30 --warmup 5 --quick --inflight 3
b yr1_epi2_if3 EMO_CONV_YREUSE=1 EMO_CONV_EPI=2 python bench.py --steps 30 --warmup 5 --quick --inflight 3
for f in "$out"/bench_*.json; do
  python - "$f" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rf = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 2), d["unit"], "e2e", round(d.get("e2e", {}).get("value", 0), 2),
          "lat1", round(d.get("latency_ms_one_frame_alone") or 0, 3), "launches/step", d.get("gpu_launches_per_step"),
          "top", rf.get("kernel", "")[:60], round(rf.get("frac", 0), 3), "allconv", round(rf.get("all_convs", {}).get("frac", 0), 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
timeout 200 tools/gs3_lab > "$out/gs3_lab.txt" 2>&1; echo "gs3_lab rc=$?" | tee -a "$out/summary.txt"
tail -30 "$out/summary.txt"; cat "$out/gs3_lab.txt"; head -12 "$out"/timeline_yr1_epi0.txt
