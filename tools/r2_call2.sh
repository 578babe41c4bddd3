#!/bin/bash
# Round-2 GPU call 2: the adopted defaults (sub-pixel up-convs, fp16 two-plane networks, folded down-convs) + programmatic
# dependent launch: full GPU suite, PDL on/off A/B at 1-3 frames in flight, elementwise-pass roofline probe, fresh launch list.
set -u
out=gpurun_out/r2c2
mkdir -p "$out"
rm -f gpurun_out/stage_parity.txt
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -20 >> "$out/summary.txt"
grep -E "white-noise frames|default vs plain" "$out/pytest_gpu.txt" >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b() { name=$1; shift; timeout 400 env "$@" > "$out/bench_$name.json" 2> "$out/bench_$name.err"; }
b pdl_if2 A=1 python bench.py --steps 30 --warmup 5 --quick
b pdl_if1 A=1 python bench.py --steps 30 --warmup 5 --quick --inflight 1
b pdl_if3 A=1 python bench.py --steps 30 --warmup 5 --quick --inflight 3
b nopdl_if2 EMO_PDL=0 python bench.py --steps 30 --warmup 5 --quick
b nopdl_if1 EMO_PDL=0 python bench.py --steps 30 --warmup 5 --quick --inflight 1
b nopdl_if3 EMO_PDL=0 python bench.py --steps 30 --warmup 5 --quick --inflight 3
timeout 600 python bench.py --steps 30 --warmup 5 > "$out/bench_full.json" 2> "$out/bench_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers.csv" 2>/dev/null
for f in "$out"/bench_*.json; do
  python - "$f" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rf = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 2), d["unit"], "e2e", round(d.get("e2e", {}).get("value", 0), 2),
          "launches/step", d.get("gpu_launches_per_step"), "top", rf.get("kernel", "")[:60], round(rf.get("frac", 0), 3), "allconv", round(rf.get("all_convs", {}).get("frac", 0), 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
timeout 300 python tools/apply_probe.py > "$out/apply_probe.txt" 2>&1; echo "apply_probe rc=$?" | tee -a "$out/summary.txt"
EMO_NO_LAYER_CSV=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file "$out/launches_r2.csv" \
    python bench.py --steps 2 --warmup 3 --eager --quick > "$out/bench_under_ncu.log" 2>&1; echo "ncu rc=$?" | tee -a "$out/summary.txt"
cat "$out/summary.txt" | tail -40; cat "$out/apply_probe.txt"
