#!/usr/bin/env python
"""Collects the N-GPU measurements a gpurun call left in gpurun_out/ (profiles/capture_r02b.sh: N = 1, capture_r02c.sh: N = 2, 8)
into profiles/r02_scale/*.json and the table profiles/r02_scale_configs.md.     python profiles/make_scale_md.py"""
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles", "r02_scale")
os.makedirs(OUT, exist_ok=True)


def load(name):
    src = os.path.join(G, name)
    dst = os.path.join(OUT, name)
    if os.path.exists(src):
        shutil.copyfile(src, dst)
    if not os.path.exists(dst):
        return None
    for line in open(dst):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                continue
    return None


md = ["# r02 — BASELINE configs on 1 / 2 / 8 B200 of one box (weak scaling: every rank holds a shard of the calibration set and draws its own mini-batch of 32)\n",
      "Commands: `profiles/capture_r02b.sh` (N = 1), `profiles/capture_r02c.sh N` (N = 2, 8); raw JSON lines under `profiles/r02_scale/`. "
      "ms per step = max over ranks of the device time (CUDA events between barriers); it/s counts the batch-32 iterations all ranks complete; "
      "efficiency = (ms per step at N = 1) / (ms per step at N). fp32 convolutions (cuDNN), synthetic randn images, random-init weights.\n"]

# ---- configs[1] ResNet-18 (bench.py)
b1 = None
p = os.path.join(ROOT, "profiles", "r02_bench_full.json")
if os.path.exists(p):
    b1 = json.load(open(p))
rows = [(1, b1)] + [(n, load(f"r02_bench_n{n}.json")) for n in (2, 4, 8)]
md += ["## configs[1] — ResNet-18 W2A4 block reconstruction, 9 units (`bench.py --gpus N`)\n",
       "| N | ms/step | it/s (all ranks) | weak efficiency | e2e ms/step | e2e it/s | e2e efficiency |", "|---|---|---|---|---|---|---|"]
base = base_e = None
for n, d in rows:
    if not d:
        continue
    ms, e = d["ms_per_step"], d.get("e2e") or {}
    base = base or ms
    base_e = base_e or e.get("ms_per_step")
    md.append(f"| {n} | {ms:.3f} | {d['value']:.0f} | {base / ms:.3f} | {e.get('ms_per_step', float('nan')):.3f} | {e.get('value', float('nan')):.0f} | "
              f"{(base_e / e['ms_per_step']) if e.get('ms_per_step') and base_e else float('nan'):.3f} |")
d8 = load("r02_bench_n8_dma.json")
if d8 and d8.get("e2e"):
    md.append(f"\nN = 8 with `--host-stage dma` (per-row copy-engine transfers instead of the SM pull): e2e {d8['e2e']['ms_per_step']:.3f} ms/step, "
              f"{d8['e2e']['value']:.0f} it/s.")
md.append("\nRound 1 (`SCALE_r01.json`, NCCL all-reduce + Adam on every rank): 4.82 / 5.28 / 5.80 / 6.07 ms at N = 1 / 2 / 4 / 8 "
          "(efficiency 0.913 / 0.831 / 0.795); e2e 5.59 / 5.89 / 7.67 / 9.37 ms.\n")

# ---- configs[4] RegNetX-3200M
md += ["## configs[4] — RegNetX-3200M W2A4 block reconstruction, all 26 units (`examples/scale_configs.py --config regnet`)\n",
       "| N | ms/step (26 iterations) | it/s (all ranks) | weak efficiency | exchange | α elements | gradient bytes/step | peak HBM GiB per rank |", "|---|---|---|---|---|---|---|---|"]
base = None
for n in (1, 2, 4, 8):
    d = load(f"r02_regnet_n{n}.json")
    if not d:
        continue
    base = base or d["ms_per_step"]
    md.append(f"| {n} | {d['ms_per_step']:.2f} | {d['iters_per_s']:.0f} | {base / d['ms_per_step']:.3f} | {d['exchange']} | {d['alpha_elems'] / 1e6:.1f} M | "
              f"{d['grad_bytes_per_step'] / 1e6:.0f} MB | {d['peak_hbm_gib']} |")
md.append("")

# ---- configs[2] ResNet-50 shifted-scale layer reconstruction
md += ["## configs[2] — ResNet-50 W4A4 shifted-scale layer reconstruction (`layer_recon_shiftedScale`, `examples/scale_configs.py --config resnet50_shift`)\n"]
res = {n: load(f"r02_resnet50_shift_n{n}.json") for n in (1, 2, 4, 8)}
layers = list(res[1]["layers"]) if res.get(1) else []
md += ["| layer | " + " | ".join(f"N = {n}: µs/iter (it/s, efficiency)" for n in res if res[n]) + " |", "|---|" + "---|" * sum(1 for n in res if res[n])]
for L in layers:
    base = res[1]["layers"][L]["us_per_iter"]
    cells = []
    for n in res:
        if not res[n]:
            continue
        r = res[n]["layers"][L]
        cells.append(f"{r['us_per_iter']:.0f} ({r['iters_per_s']:.0f}, {base / r['us_per_iter']:.3f})")
    md.append(f"| `{L}` | " + " | ".join(cells) + " |")
md.append("\nThe exchange of this loop is the same peer-memory kernel (`ssq_grad_exchange_adam`; the loop's parameters are the group logits and the "
          "AdaRound α of one layer); on two GPUs it is bit-identical to the NCCL all-reduce + `ssq_adam_step` path (`tests/test_multi_gpu.py`).\n")

# ---- configs[3] MobileNetV2
m = load("r02_mobilenetv2_n1.json")
if m:
    md += ["## configs[3] — MobileNetV2 W3A3, one B200 (`examples/scale_configs.py --config mobilenetv2_mse`)\n",
           f"* {m['quant_modules']} quantised layers, {m['weight_rows']} weight rows (depthwise rows of 9 among them), {m['weight_elems'] / 1e6:.2f} M weights",
           f"* whole-model MSE weight-scale search (K2a, 80 candidates per row): **{m['k2a_weight_scale_search_all_layers_ms']:.2f} ms** of kernel time for all layers; "
           f"first quantised forward (search + per-layer host glue): {m['first_quantised_forward_s'] * 1e3:.0f} ms",
           f"* ChannelQuantMSE input-scale search (K2b) of every layer at level {m['k2b_level']}: **{m['k2b_inp_scale_search_all_layers_ms']:.2f} ms** for all layers "
           f"(6 launches per layer: launch-latency-sized at these shapes); the public flow `channelShift_wMSE_flow` "
           f"(builds and initialises {m['channelquantmse_layers_built']} ChannelQuantMSE quantisers): {m['channelShift_wMSE_flow_s'] * 1e3:.0f} ms",
           f"* AdaRound block reconstruction of all {m['block_recon']['units']} units: {m['block_recon']['ms_per_step']:.2f} ms per step = "
           f"**{m['block_recon']['iters_per_s']:.0f} it/s**\n"]

# ---- host link
md += ["## Host link (what the e2e / host-resident mode depends on; `scratch/host_link_probe.py`)\n",
       "| N | NUMA nodes | host cores | per-rank alone: copy engine / SM pull GB/s | all ranks at once: per rank (min–max) | aggregate copy engine / SM pull GB/s |", "|---|---|---|---|---|---|"]
for n in (1, 2, 4, 8):
    d = load(f"r02_host_link_n{n}.json")
    if not d:
        continue
    al = d["ranks"]
    a_d = [r["alone"]["dma_gbs"] for r in al]; a_p = [r["alone"]["pull_gbs"] for r in al]
    t_d = [r["together"]["dma_gbs"] for r in al]; t_p = [r["together"]["pull_gbs"] for r in al]
    md.append(f"| {n} | {len(d['numa_nodes'])} | {d['cpus_allowed']} | {min(a_d):.1f}–{max(a_d):.1f} / {min(a_p):.1f}–{max(a_p):.1f} | "
              f"{min(t_d):.1f}–{max(t_d):.1f} / {min(t_p):.1f}–{max(t_p):.1f} | {d['aggregate_together_gbs']['dma']:.0f} / {d['aggregate_together_gbs']['pull']:.0f} |")
md.append("")
open(os.path.join(ROOT, "profiles", "r02_scale_configs.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
