#!/bin/bash
# Round-2 GPU call 4: conv post-op (fused split-K finalize), deterministic split-K, phase timeline of the conv kernel.
set -u
out=gpurun_out/r2c4
mkdir -p "$out"
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(FAILED|ERROR)|passed|failed" "$out/pytest_gpu.txt" | tail -30 >> "$out/summary.txt"
grep -E "conv post-op" "$out/pytest_gpu.txt" >> "$out/summary.txt"
EMO_CONV_EPI=2 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_conv_modes_gpu.py -q -m gpu > "$out/pytest_epi2.txt" 2>&1; echo "pytest epi2 rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_epi2.txt" >> "$out/summary.txt"
for e in 0 1 2; do timeout 120 python tools/conv_timeline.py $e > "$out/timeline_epi$e.txt" 2>&1; done
b() { name=$1; shift; timeout 400 env "$@" > "$out/bench_$name.json" 2> "$out/bench_$name.err"; }
b if1 A=1 python bench.py --steps 30 --warmup 5 --quick --inflight 1
b if2 A=1 python bench.py --steps 30 --warmup 5 --quick
b if3 A=1 python bench.py --steps 30 --warmup 5 --quick --inflight 3
for f in "$out"/bench_*.json; do
  python - "$f" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rf = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 2), d["unit"], "e2e", round(d.get("e2e", {}).get("value", 0), 2),
          "lat1", d.get("latency_ms_one_frame_alone"), "launches/step", d.get("gpu_launches_per_step"),
          "top", rf.get("kernel", "")[:60], round(rf.get("frac", 0), 3), "allconv", round(rf.get("all_convs", {}).get("frac", 0), 3))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -30 "$out/summary.txt"; cat "$out"/timeline_epi*.txt
