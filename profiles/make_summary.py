#!/usr/bin/env python
"""Turns the ncu captures a gpurun call left in gpurun_out/ into the committed summaries under profiles/.

    python profiles/make_summary.py r01c          # reads gpurun_out/r01c_full_<kernel>.ncu-rep, gpurun_out/r01c_launches.csv

Writes profiles/<round>_kernels_ncu_full.{md,json}, profiles/<round>_launch_list.md and
profiles/<round>_launches_timed_region.csv. Needs `ncu` on PATH (only to READ the reports; no GPU).
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GOUT = os.path.join(ROOT, "gpurun_out")

# kernel -> (title, algorithmic bytes per element, elements per launch at the microbench shape)
W = 4096 * 4096 * 9
A = 256 * 256 * 56 * 56
KERNELS = OrderedDict([
    ("adam_kernel", ("fused Adam (ssq_adam_step), flat buffer of 151 M", 28, W)),
    ("ada_fwd_kernel", ("K1b fwd + regulariser (ssq_fq_adaround_fwd), weights [4096,4096,3,3]", 12, W)),
    ("ada_bwd_kernel", ("K1b bwd + regulariser gradient (ssq_fq_adaround_bwd)", 16, W)),
    ("recon_loss_kernel", ("K3 loss + dpred, p = 2 (ssq_recon_loss), activations [256,256,56,56]", 12, A)),
    ("fq_affine_fwd_vec", ("K1a fwd per-channel weights (ssq_fq_affine_fwd)", 8, W)),
    ("fq_affine_bwd_kernel", ("K1a bwd per-channel weights (ssq_fq_affine_bwd)", 12, W)),
    ("gather_rows_kernel", ("mini-batch gather (ssq_gather_rows), activations [256,256,56,56]", 8, A)),
    ("fq_shift_fwd_vec", ("K1c fwd, adaShift soft, S = 3 (ssq_fq_shift_fwd)", 12, W)),
    ("fq_shift_bwd_vec", ("K1c bwd, adaShift soft, S = 3 (ssq_fq_shift_bwd)", 16, W)),
    ("fq_shift_fwd_vec_deq", ("K1c fwd, dequantised mixture, S = 3 (ssq_fq_shift_fwd, -s 16)", 8, W)),
    ("fq_shift_bwd_vec_deq", ("K1c bwd, dequantised mixture, S = 3 (ssq_fq_shift_bwd, -s 16)", 8, W)),
    ("ada_bwd_adam_mt_kernel", ("launch 3: K1b bwd + regulariser gradient + Adam in one pass (ssq_fq_adaround_bwd_adam_mt)", 32, W)),
    ("ada_fwd_mt_kernel", ("K1b fwd multi-tensor (ssq_fq_adaround_fwd_mt)", 12, W)),
    ("inp_scale_sweep_kernel", ("K2b sweep: max t per column in one pass over the weights (inside ssq_inp_scale_search), level 16", 4, W)),
    ("export_vec_kernel", ("integer export, 2-bit + alpha (ssq_export_codes)", 8.25, W)),
    ("import_vec_kernel", ("integer import, 2-bit (ssq_import_codes)", 4.25, W)),
])
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def kernels(tag):
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6554.9) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6554.9
    md = [f"# {tag} — ncu --set full of each hot kernel on DRAM-resident inputs\n",
          "Command per kernel (after `python bench.py --micro-only` exited 0 without ncu): `ncu --set full --clock-control none "
          "--import-source on -k regex:<kernel> -s 3 -c 1 -o gpurun_out/" + tag + "_full_<kernel> python bench.py --micro-only` "
          "(profiles/capture_r01.sh, profiles/capture_r01e.sh).",
          "`traffic` = dram__bytes_read.sum + dram__bytes_write.sum of that one launch; algorithmic bytes = SURVEY §8d per-element "
          "figure x elements. ncu's DRAM % is against its own ~8.2 TB/s peak; the roofline fraction in bench.py is against the "
          f"measured {peak:.1f} GB/s copy peak (MEASURED_PEAKS.json). Durations under ncu are cold-cache single launches; the "
          "CUDA-event numbers are in DESIGN.md §4.\n"]
    js = {}
    for k, (title, bpe, elems) in KERNELS.items():
        path = os.path.join(GOUT, f"{tag}_full_{k}.ncu-rep")
        if not os.path.exists(path):
            continue
        r = raw_page(path)
        val = lambda m: float(r[m][1].replace(",", "")) if r.get(m, ("", ""))[1] not in ("", None) else float("nan")
        rd = val("dram__bytes_read.sum") * SCALE.get(r["dram__bytes_read.sum"][0], 1.0)
        wr = val("dram__bytes_write.sum") * SCALE.get(r["dram__bytes_write.sum"][0], 1.0)
        us = val("gpu__time_duration.sum") * SCALE.get(r["gpu__time_duration.sum"][0], 1.0)
        alg = bpe * elems
        inst = val("smsp__inst_executed.sum")
        md.append(f"## {k} — {title}\n")
        md.append(f"`{r['Kernel Name'][1][:110]}`\n")
        md.append("| metric | value |\n|---|---|")
        for m in METRICS:
            if m in r:
                md.append(f"| `{m}` | {r[m][1]} {r[m][0]} |")
        md.append(f"| algorithmic bytes | {alg / 1e6:.0f} MB ({bpe} B/elem x {elems / 1e6:.0f} M) |")
        md.append(f"| traffic / algorithmic | {(rd + wr) / 1e6:.0f} MB / {alg / 1e6:.0f} MB = {(rd + wr) / alg:.3f} |")
        md.append(f"| achieved under ncu | {alg / us / 1e3:.0f} GB/s algorithmic ({alg / us / 1e3 / peak:.3f} of measured copy peak), {(rd + wr) / us / 1e3:.0f} GB/s DRAM |")
        md.append(f"| thread instructions / element | {inst * 32 / elems:.1f} |\n")
        js[k] = {"title": title, "duration_us": us, "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6, "traffic_MB": (rd + wr) / 1e6,
                 "algorithmic_MB": alg / 1e6, "traffic_over_algorithmic": (rd + wr) / alg, "gbs_algorithmic": alg / us / 1e3,
                 "regs": val("launch__registers_per_thread"), "grid": val("launch__grid_size"),
                 "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "occupancy_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "thread_instr_per_elem": inst * 32 / elems}
    open(os.path.join(OUT, f"{tag}_kernels_ncu_full.md"), "w").write("\n".join(md) + "\n")
    json.dump(js, open(os.path.join(OUT, f"{tag}_kernels_ncu_full.json"), "w"), indent=1)
    print("kernels:", list(js))


def launch_list(tag, steps=3, units=9):
    path = os.path.join(GOUT, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    recs = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        recs.append((r["Kernel Name"], v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0}.get(unit, 1e-3)))
    # launch order of bench.py: 3 eager warm-up iterations per engine (3*units loop_advance launches), then `steps` untimed + `steps`
    # timed steps (one graph replay per unit each), then the per-unit timing loops and the in-step event pass. The timed region is
    # therefore loop_advance launches [3*units + steps*units, 3*units + 2*steps*units)
    adv = [i for i, (n, _) in enumerate(recs) if "loop_advance" in n or "iter_prologue" in n]
    first = 3 * units + steps * units
    start, stop = adv[first], adv[first + steps * units]
    region = recs[start:stop]
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in region:
        agg[n][0] += 1; agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for n, v in agg.items() if n.startswith("ssq::") or "ssq::" in n)
    with open(os.path.join(OUT, f"{tag}_launches_timed_region.csv"), "w") as f:
        f.write("index,kernel,duration_us\n")
        for i, (n, us) in enumerate(region):
            f.write(f"{i},\"{n[:120]}\",{us:.3f}\n")
    md = [f"# {tag} — ncu launch list of the bench step (gpu__time_duration.sum, --clock-control none)\n",
          "Command (after the same command exited 0 without ncu): `ncu --metrics gpu__time_duration.sum --clock-control none --csv "
          "python bench.py --steps 3 --warmup 3 --images 64 --skip-e2e --skip-act --skip-cpu --skip-micro --skip-tf32 --skip-shift --cudnn-benchmark 0`\n",
          f"Whole run: {len(recs)} launches. Timed region below = the {steps} timed steps x {units} units = {steps * units} captured iterations "
          f"({len(region)} kernel nodes, {len(region) / (steps * units):.0f} per iteration on average). Per-launch times are cold-cache and "
          "serialised: compare SHARES.\n",
          f"Total {total / 1e3:.3f} ms for {steps} steps = {total / 1e3 / steps:.3f} ms/step under ncu. ssq kernels: {100 * ours / total:.1f} % of "
          f"the step; cuDNN/ATen: {100 - 100 * ours / total:.1f} %.\n",
          "| share | us/step | launches/step | kernel |\n|---|---|---|---|"]
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| {100 * us / total:5.2f} % | {us / steps:8.1f} | {c / steps:5.1f} | `{n[:100]}` |")
    open(os.path.join(OUT, f"{tag}_launch_list.md"), "w").write("\n".join(md) + "\n")
    print(f"launch list: {len(region)} nodes, ssq share {100 * ours / total:.1f} %")


def k2_table(tag):
    """per-launch metric list of the scale-search kernels (ncu --metrics ..., `bench.py --k2-only`) -> profiles/<tag>_k2_ncu.md"""
    path = os.path.join(GOUT, f"{tag}_k2_ncu.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    by_id = OrderedDict()
    for r in rows:
        d = by_id.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        d[r["Metric Name"]] = v * SCALE.get(r.get("Metric Unit", ""), 1.0) if r["Metric Name"] in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum") else v
    # group identical (kernel, grid) launches: report the LAST one (the warmed microbench launches come last)
    groups = OrderedDict()
    for d in by_id.values():
        key = (d["name"].split("(")[0], d.get("launch__grid_size", d["grid"]))
        g = groups.setdefault(key, {"n": 0})
        g["n"] += 1
        g["last"] = d
    md = [f"# {tag} — scale-search kernels (K2a / K2b) under ncu\n",
          "Command (after `python bench.py --k2-only` exited 0 without ncu): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
          "dram__bytes_write.sum,smsp__issue_active...,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum --clock-control none -k regex:mse_search|mse_rank|"
          "mse_settle|inp_scale|row_minmax python bench.py --k2-only` (profiles/capture_r02_final.sh; the inp_scale rows were re-captured by profiles/capture_r02h.sh after "
          "the K2b call became four launches). One row per (kernel, grid) — the last "
          "launch of that shape; durations are cold-cache single launches (CUDA-event numbers: BENCH extra.scale_search).\n",
          "| kernel | grid | launches | us | DRAM MB (r+w) | DRAM % of ncu peak | issue-active % | warps-active % | regs | warp-instr (M) | XU-pipe instr (M) |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    js = []
    for (name, grid), g in groups.items():
        d = g["last"]
        f = lambda m, dflt=float("nan"): d.get(m, dflt)
        md.append(f"| `{name.replace('ssq::', '')[:60]}` | {int(f('launch__grid_size', 0))} | {g['n']} | {f('gpu__time_duration.sum'):.1f} | "
                  f"{(f('dram__bytes_read.sum', 0) + f('dram__bytes_write.sum', 0)) / 1e6:.1f} | {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                  f"{f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                  f"{int(f('launch__registers_per_thread', 0))} | {f('smsp__inst_executed.sum', 0) / 1e6:.2f} | {f('sm__inst_executed_pipe_xu.sum', 0) / 1e6:.2f} |")
        js.append({"kernel": name, "grid": f('launch__grid_size', 0), "launches": g["n"], "us": f('gpu__time_duration.sum'),
                   "dram_MB": (f('dram__bytes_read.sum', 0) + f('dram__bytes_write.sum', 0)) / 1e6,
                   "issue_active_pct": f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                   "dram_pct": f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')})
    open(os.path.join(OUT, f"{tag}_k2_ncu.md"), "w").write("\n".join(md) + "\n")
    json.dump(js, open(os.path.join(OUT, f"{tag}_k2_ncu.json"), "w"), indent=1)
    print("k2 table:", len(groups), "rows")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01c"
    kernels(tag)
    launch_list(tag)
    k2_table(tag)
