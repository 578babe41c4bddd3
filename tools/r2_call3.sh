#!/bin/bash
# Round-2 GPU call 3: TMA epilogue (conv EPI = 2) vs the round-1 forms, wrapper-boundary kernels, new bench keys.
set -u
out=gpurun_out/r2c3
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
grep -E "bicubic|default vs plain" "$out/pytest_gpu.txt" >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b() { name=$1; shift; timeout 400 env "$@" > "$out/bench_$name.json" 2> "$out/bench_$name.err"; cp gpurun_out/conv_layers.csv "$out/conv_layers_$name.csv" 2>/dev/null; }
b epi2_if2 EMO_CONV_EPI=2 python bench.py --steps 30 --warmup 5 --quick
b epi1_if2 EMO_CONV_EPI=1 python bench.py --steps 30 --warmup 5 --quick
b epi0_if2 EMO_CONV_EPI=0 python bench.py --steps 30 --warmup 5 --quick
b epi2_if1 EMO_CONV_EPI=2 python bench.py --steps 30 --warmup 5 --quick --inflight 1
b epi1_if1 EMO_CONV_EPI=1 python bench.py --steps 30 --warmup 5 --quick --inflight 1
b epi2_if3 EMO_CONV_EPI=2 python bench.py --steps 30 --warmup 5 --quick --inflight 3
timeout 900 python bench.py --steps 30 --warmup 5 > "$out/bench_full.json" 2> "$out/bench_full.err"
for f in "$out"/bench_*.json; do
  python - "$f" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rf = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 2), d["unit"], "e2e", round(d.get("e2e", {}).get("value", 0), 2),
          "lat1", d.get("latency_ms_one_frame_alone"), "wrapper", d.get("e2e_wrapper"), "stage2", (d.get("stage2") or {}).get("value"),
          "top", rf.get("kernel", "")[:60], round(rf.get("frac", 0), 3), "launch_us", round(rf.get("launch_us", 0), 1), "allconv", round(rf.get("all_convs", {}).get("frac", 0), 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
EMO_CONV_DBG=0 timeout 200 python tools/conv_bound_probe.py full > "$out/conv_probe_epi2.txt" 2>&1
EMO_CONV_EPI=1 timeout 200 python tools/conv_bound_probe.py full > "$out/conv_probe_epi1.txt" 2>&1
EMO_CONV_EPI=0 timeout 200 python tools/conv_bound_probe.py full > "$out/conv_probe_epi0.txt" 2>&1
tail -40 "$out/summary.txt"; cat "$out"/conv_probe_epi*.txt
