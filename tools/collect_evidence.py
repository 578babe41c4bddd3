"""Turn the scratch outputs of the round's evidence call (gpurun_out/<dir>) into the committed files under profiles/
(run on the CPU box):  python tools/collect_evidence.py r2c8 r2"""
import json, pathlib, re, shutil, sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
src = ROOT / "gpurun_out" / (sys.argv[1] if len(sys.argv) > 1 else "r2c8")
tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
PR = ROOT / "profiles"

# ---- parity tables (tests/test_model_gpu.py, tests/test_stage_parity_gpu.py write them)
out = [f"# Parity of the device path against the reference fixtures / the live CPU oracle ({tag}, B200)", "",
       "`err` = max-abs difference, `ref_max` = largest magnitude of the reference tensor.  End-to-end cases: `tests/test_model_gpu.py`",
       "(`refpose` = the reference's pose matrices injected; `device` = everything on the device); stage-isolated: every network fed the",
       "oracle's input (`tests/test_stage_parity_gpu.py`, 256^2 and 512^2).", ""]
for f in sorted(src.glob("parity_*.txt")):
    if "s2_" in f.name:
        continue
    out += [f"## {f.stem.replace('parity_', '')}", "", "| tap | err | ref_max |", "|---|---:|---:|"]
    for ln in f.read_text().splitlines():
        p = ln.split()
        if len(p) >= 5:
            out.append(f"| {p[0]} | {p[2]} | {p[4]} |")
    out.append("")
sp = src / "stage_parity.txt"
if sp.exists():
    out += ["## stage-isolated", "", "| size | stage | err | ref max | relative |", "|---|---|---:|---:|---:|"]
    for ln in sp.read_text().splitlines():
        m = re.match(r"(\d+) (\S+) err=(\S+) scale=(\S+) rel=(\S+)", ln)
        if m:
            out.append(f"| {m.group(1)} | {m.group(2)} | {m.group(3)} | {m.group(4)} | {m.group(5)} |")
    out.append("")
for f in sorted(src.glob("parity_s2_*.txt")):
    out += [f"## stage 2: {f.stem.replace('parity_', '')}", "", "```", f.read_text().strip(), "```", ""]
(PR / f"parity_{tag}.md").write_text("\n".join(out) + "\n")

# ---- per-layer conv table
rows = {}
names = []
for f in sorted(src.glob("layers_*.txt")):
    name = f.stem.replace("layers_", "")
    names.append(name)
    for ln in f.read_text().splitlines():
        m = re.match(r".*\| (.+?)\s+([\d.]+) us x(\d+)\s+([\d.]+) TFLOP/s", ln)
        if m:
            rows.setdefault(m.group(1).strip(), {})[name] = (float(m.group(2)), int(m.group(3)), float(m.group(4)))
if rows:
    out = [f"# Convolution shapes of one driver frame, device time per launch ({tag})", "",
           "`tools/conv_layer_bench.py`: CUDA-graph replays of 10 back-to-back launches (no host launch cost, no tensor-map encoding in the",
           "number).  `auto` = the product library's per-layer choice; `epi0` / `epi1` = instrumented build forced to the in-warp final phase /",
           "the TMA epilogue; `chunkN` = N MMAs per TMEM accumulation chunk for every layer (product default: 96 bf16 planes / 24 fp16 planes).  TFLOP/s = algorithmic (one product = one",
           "flop pair; three MMAs are issued per product).", "",
           "| layer | x per frame | " + " | ".join(f"{n} us" for n in names) + " | auto TFLOP/s |", "|---|---:|" + "---:|" * (len(names) + 1)]
    tot = {n: 0.0 for n in names}
    for k, v in rows.items():
        n_ = next(iter(v.values()))[1]
        out.append(f"| {k} | {n_} | " + " | ".join(f"{v[n][0]:.1f}" if n in v else "" for n in names) + f" | {v.get('auto', (0, 0, 0))[2]:.0f} |")
        for n in names:
            if n in v:
                tot[n] += v[n][0] * v[n][1]
    out.append("| **sum over the frame** | | " + " | ".join(f"**{tot[n] / 1000:.3f} ms**" for n in names) + " | |")
    (PR / f"conv_layers_{tag}.md").write_text("\n".join(out) + "\n")

# ---- plain copies
for a, b in [("gs3_lab.txt", f"gs3_lab_{tag}.txt"), ("apply_probe.txt", f"apply_probe_{tag}.txt"), ("gs3_check.txt", f"gs3_check_{tag}.txt"),
             ("conv_layers.csv", f"conv_layers_{tag}.csv"), ("bench_full.json", f"bench_{tag}_n1.json"), ("bench_n2.json", f"bench_{tag}_n2.json")]:
    if (src / a).exists():
        shutil.copy(src / a, PR / b)
tl = []
for f in sorted(src.glob("timeline_*.txt")):
    tl += [f"==== {f.stem} ====", f.read_text().strip(), ""]
if tl:
    (PR / f"conv_timeline_{tag}.txt").write_text("\n".join(tl) + "\n")
ab = src / "summary.txt"
if ab.exists():
    (PR / f"ab_{tag}.txt").write_text(ab.read_text())
print("written:", sorted(p.name for p in PR.glob(f"*{tag}*")))
