#!/bin/bash
# Round-2 GPU call 1: un-masked GPU suite, A/B of every opt-in variant written blind at the end of round 1, probes.
set -u
out=gpurun_out/r2c1
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/smi.txt" 2>&1
timeout 120 tools/gs3_check > "$out/gs3_check.txt" 2>&1; echo "gs3_check rc=$?" | tee -a "$out/summary.txt"
timeout 1500 python -m pytest tests -q -m gpu -rA -s > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" | tee -a "$out/summary.txt"
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed" "$out/pytest_gpu.txt" | tail -150 >> "$out/summary.txt"
cp gpurun_out/parity_*.txt gpurun_out/stage_parity.txt "$out/" 2>/dev/null
b() { name=$1; shift; timeout 400 env "$@" python bench.py --steps 30 --warmup 5 --quick > "$out/bench_$name.json" 2> "$out/bench_$name.err"; }
timeout 600 python bench.py --steps 30 --warmup 5 > "$out/bench_default_full.json" 2> "$out/bench_default_full.err"
cp gpurun_out/conv_layers.csv "$out/conv_layers_default.csv" 2>/dev/null
b default A=1
b upconv_ps EMO_UPCONV_PS=1
cp gpurun_out/conv_layers.csv "$out/conv_layers_ps.csv" 2>/dev/null
b h2 EMO_H2_NETS=warp,expression,head_pose
cp gpurun_out/conv_layers.csv "$out/conv_layers_h2.csv" 2>/dev/null
b h2_ps EMO_H2_NETS=warp,expression,head_pose EMO_UPCONV_PS=1
b apply_v1 EMO_APPLY_V1=1
b apply_occ EMO_APPLY_OCC=1
b apply_pf EMO_APPLY_PF=1
timeout 400 python bench.py --steps 30 --warmup 5 --quick --inflight 1 > "$out/bench_inflight1.json" 2> "$out/bench_inflight1.err"
timeout 400 python bench.py --steps 30 --warmup 5 --quick --inflight 3 > "$out/bench_inflight3.json" 2> "$out/bench_inflight3.err"
EMO_GS3_VEC2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$out/bench_gs3_vec2.json" 2> "$out/bench_gs3_vec2.err"
timeout 600 python bench.py --workload stage2 --steps 10 --warmup 3 > "$out/bench_stage2_default.json" 2> "$out/bench_stage2_default.err"
EMO_UPCONV_PS=1 EMO_POOLCONV_FOLD=1 timeout 600 python bench.py --workload stage2 --steps 10 --warmup 3 > "$out/bench_stage2_folds.json" 2> "$out/bench_stage2_folds.err"
for f in "$out"/bench_*.json; do
  python - "$f" <<'PY' | tee -a "$out/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline_grid_sample3d", {})
    rf = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 2), d["unit"], "e2e", round(d.get("e2e", {}).get("value", 0), 2),
          "top", rf.get("kernel", "")[:60], round(rf.get("frac", 0), 3), "allconv", round(rf.get("all_convs", {}).get("frac", 0), 3),
          "gs3", {k: round(v.get("frac", 0), 3) for k, v in r.items() if isinstance(v, dict) and "frac" in v}, d["config"].get("experimental"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
timeout 300 python tools/conv_bound_probe.py > "$out/conv_bound_probe.txt" 2>&1; echo "bound_probe rc=$?" | tee -a "$out/summary.txt"
if [ -x tools/mma_probe ]; then timeout 60 tools/mma_probe > "$out/mma_probe.txt" 2>&1; echo "mma_probe rc=$?" | tee -a "$out/summary.txt"; fi
cat "$out/summary.txt" | tail -60
