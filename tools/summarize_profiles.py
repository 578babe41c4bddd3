"""Turn the scratch ncu outputs under gpurun_out/ into the committed summaries under profiles/ (run on the CPU box)."""
import csv, io, pathlib, re, subprocess, sys, collections

ROOT = pathlib.Path(__file__).resolve().parents[1]
GO, PR = ROOT / "gpurun_out", ROOT / "profiles"
PR.mkdir(exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launches():
    f = GO / f"launches_{tag}.csv"
    if not f.exists():
        return
    lines = f.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = collections.OrderedDict()
    order = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        t = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        t_us = t / 1000.0 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1000.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t_us
        order.append((name, t_us))
    tot = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3 --eager`",
           "", "Times are cold-cache and serialised (ncu replays every launch in isolation): compare SHARES, not absolutes.",
           f"Captured launches: {len(order)}; total {tot/1000:.2f} ms (source pass + 5 driver frames + microbenches).", "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100*v[1]/tot:.1f}% |")
    # one driver frame = the launches between two consecutive pose_theta_kernel launches (taken near the end of the run)
    idx = [i for i, (n, _) in enumerate(order) if "pose_theta" in n]
    if len(idx) >= 3:
        a, b = idx[-3], idx[-2]
        fr = collections.OrderedDict()
        for n, t in order[a:b]:
            v = fr.setdefault(n, [0, 0.0]); v[0] += 1; v[1] += t
        ftot = sum(v[1] for v in fr.values())
        out += ["", f"## One driver frame ({b - a} launches, {ftot/1000:.2f} ms summed cold-cache durations)", "",
                "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
        for k, v in sorted(fr.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100*v[1]/ftot:.1f}% |")
        # phases by position in the launch order of the frame (see the sequence in Model.driver_pass)
        seq = order[a:b]
        gs = [i for i, (n, _) in enumerate(seq) if "gs3_cl" in n]
        rs = [i for i, (n, _) in enumerate(seq) if "resize_bilinear" in n]
        if gs and rs:
            lin = [i for i, (n, _) in enumerate(seq) if "linear_kernel" in n and i < gs[0]]
            marks = [("pose_theta + expression encoder (aligned 224^2 crop, ResNet)", 0, lin[0] if lin else gs[0]),
                     ("embedding MLPs + warp generator (3-D convs, fp16 two-plane operands)", lin[0] if lin else gs[0], gs[0]),
                     ("grid_sample_3d x2", gs[0], gs[-1] + 1),
                     ("decoder (two planes)", gs[-1] + 1, rs[0]),
                     ("head-pose regressor of the next frame (ResNet18)", rs[0], len(seq))]
            out += ["", "| phase | launches | us |", "|---|---:|---:|"]
            for name, lo, hi in marks:
                out.append(f"| {name} | {hi - lo} | {sum(t for _, t in seq[lo:hi]):.1f} |")
    (PR / f"launches_{tag}_summary.md").write_text("\n".join(out) + "\n")
    print("\n".join(out[:14]))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def rep(name):
    f, fcsv = GO / f"{name}_{tag}.ncu-rep", GO / f"{name}_{tag}.raw.csv"
    if fcsv.exists():       # exported on the GPU box (tools/profile.sh): the report itself is too large to travel
        text = fcsv.read_text()
    elif f.exists():
        text = subprocess.run(["ncu", "-i", str(f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader(io.StringIO(text)))
    if len(rows) < 3:
        print("could not read", f); return
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none: {name}_{tag}.ncu-rep", ""]
    idx = {h: i for i, h in enumerate(hdr)}
    for row in rows[2:]:
        out.append(f"## launch {row[idx['ID']]}: {row[idx['Kernel Name']][:90]}  grid {row[idx.get('Grid Size', 0)]} block {row[idx.get('Block Size', 0)]}")
        for k in KEYS:
            cands = [h for h in hdr if h == k or h.startswith(k)]
            for h in cands[:1]:
                out.append(f"  {h} = {row[idx[h]]} {units[idx[h]]}")
        out.append("")
    (PR / f"{name}_{tag}.txt").write_text("\n".join(out) + "\n")
    print(f"wrote profiles/{name}_{tag}.txt ({len(rows)-2} launches)")


launches()
rep("prof_conv")
rep("prof_elem")
rep("prof_gs3")
