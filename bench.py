#!/usr/bin/env python
"""bench.py — headline benchmark of the PTQ-calibration hot path (BASELINE.json metric):
    recon iters/s (ResNet-18 W2A4, 1024 calib imgs); fake-quant HBM GB/s

One "step" = one reconstruction iteration (mini-batch 32) on EACH of the 9 reconstructed units of ResNet-18
(8 residual blocks + fc; the stem is ignore_reconstruction upstream), i.e. 9 iterations of the hot loop of
quant/block_recon.py:89-105. `value` = iterations/s with the cached features resident in HBM; `e2e` = the same
through the host-resident cache mode (pinned host features — post-ReLU tensors zero-packed —, per-step H2D of the mini-batch,
D2H of the loss; the mode block_reconstruction(..., host_resident=True) selects).
The roofline object reports the dominant kernel on DRAM-resident inputs (> L2), next to its in-step share.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
N > 1 is launched by torchrun (one rank per GPU); ranks hold disjoint shards of the calibration set, each draws
its own mini-batch of 32, and the flat AdaRound gradient is all-reduced (SUM) every iteration ("weak" scaling:
value counts batch-32 iterations completed by all ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WQ = {'n_bits': 2, 'channel_wise': True, 'scale_method': 'mse'}
AQ = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
RECON = dict(weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0)          # Brecq/main_imagenet.py:201-202
SCHED_ITERS = 20000
BATCH = 32
WORKLOAD = ("ResNet-18 W2A4 block reconstruction (configs[1]): 9 units (8 blocks + fc), weight-rounding phase, "
            "mini-batch 32, randn 224x224 calibration images, random-init weights")


# The contract is ONE JSON line on stdout. Libraries write to fd 1 behind Python's back (NCCL prints its version banner
# there at communicator creation), so fd 1 is pointed at stderr for the whole run and the result line goes to a
# private duplicate of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(text: str):
    _REAL_STDOUT.write(text + "\n")
    _REAL_STDOUT.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.idx = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- reference arm
def reference_loop(steps: int, warmup: int, sample_images: int = 64, units=None, device: str = "cpu"):
    """The reference's own op chain (oracle/ref_loop_torch.py: the ATen ops of quant/block_recon.py:89-105 in the same order,
    torch.optim.Adam — ONE optimizer per unit, built before the timed loop as block_recon.py:60 does —, pinned to the real
    reference loops by tests/golden/recon_loop.npz). device='cpu': all host cores = the CPU baseline / `--impl reference`
    arm. device='cuda': the same chain on ATen's CUDA kernels + cuDNN, i.e. what the unmodified reference does on this
    GPU (extra.gpu_reference; TF32 off like the headline). One step = one iteration on each of the 9 units at batch 32;
    features are synthetic relu(randn) of the real shapes, `sample_images` per unit."""
    from oracle import ref_loop_torch as R
    on_gpu = device != "cpu"
    if not on_gpu:
        torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    torch.manual_seed(1005)
    built = []
    specs = units or (R.RESNET18_UNITS + [("fc", 512, 1000, 1, 1)])
    total = steps + warmup
    for i, (name, cin, cout, stride, hw) in enumerate(specs):
        if name == "fc":
            unit = R.synthetic_fc_unit(cin, cout, 8, seed=i)
            x = torch.relu(torch.randn(sample_images, cin))
        else:
            unit = R.synthetic_resnet18_unit(name, cin, cout, stride, 2, seed=i)
            x = torch.relu(torch.randn(sample_images, cin, hw, hw))
        y = R.fp_unit_outputs(unit, x)
        tab = torch.stack([torch.randperm(sample_images)[:BATCH] for _ in range(total)])
        if on_gpu:
            unit, x, y, tab = R.unit_to(unit, device), x.to(device), y.to(device), tab.to(device)
        built.append((unit, x, y, tab, {}))
    alphas = [None] * len(built)
    sync = (lambda: torch.cuda.synchronize()) if on_gpu else (lambda: None)
    t_timed = 0.0
    for s in range(total):
        if s == warmup:
            sync(); t0 = time.perf_counter()
        for u, (unit, x, y, tab, state) in enumerate(built):
            # one iteration at the right point of the 20k schedule is what matters for cost: regulariser on
            alphas[u], _ = R.recon_weight_loop(unit, x, y, tab[s:s + 1], 1, weight=RECON['weight'], b_range=RECON['b_range'],
                                               warmup=RECON['warmup'], p=RECON['p'], alphas=alphas[u],
                                               start_count=SCHED_ITERS // 2 + s, t_max=SCHED_ITERS, state=state)
    sync(); t_timed = time.perf_counter() - t0
    ms = 1e3 * t_timed / max(steps, 1)
    return {"iters_per_s": len(built) / (ms / 1e3), "ms_per_step": ms, "cores": cores,
            "sample": f"{steps} steps x {len(built)} units, batch {BATCH}, {sample_images} synthetic feature images per unit "
                      f"(real ResNet-18 224x224 unit shapes), torch {torch.__version__} {'CUDA (ATen + cuDNN fp32)' if on_gpu else 'CPU'}, "
                      "one Adam per unit built outside the timed loop"}


def cpu_reference(steps: int, warmup: int, sample_images: int = 64, units=None):
    return reference_loop(steps, warmup, sample_images, units, "cpu")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "recon iters/s (ResNet-18 W2A4, 1024 calib imgs)", "value": r["iters_per_s"],
            "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, args.gpus),
            "reference_arm": "reference CPU path (oracle/ref_loop_torch.py: the ATen-CPU op chain of quant/block_recon.py:89-105, pinned to the "
                             "real reference loops by tests/golden/recon_loop.npz) on all host cores; rank 0 only; a bounded sample of the "
                             "workload in `config`: " + r["sample"],
            "cpu_baseline": {"value": r["iters_per_s"], "unit": "iters/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["iters_per_s"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
def build_model(dev, n_images, res=224, seed=1005):
    from shiftedscalequantization_b200 import quant as Q, zoo
    torch.manual_seed(seed)
    cnn = zoo.resnet18().to(dev).eval()
    qnn = Q.QuantModel(cnn, dict(WQ), dict(AQ)).to(dev).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(n_images, 3, res, res)
    return Q, qnn, cali


def recon_units(Q, qnn):
    units = []

    def walk(m):
        for _n, c in m.named_children():
            if isinstance(c, (Q.QuantModule, Q.BaseQuantBlock)):
                if not c.ignore_reconstruction:
                    units.append(c)
            else:
                walk(c)
    walk(qnn)
    return units


CAPTURE_S = []      # seconds spent in save_inp_oup_data per unit (feature capture, reported beside the loop numbers)


def make_engines(Q, qnn, cali, dev, act_quant, multi_gpu, host_resident=False, feats=None, scaling='weak', host_stage='pull'):
    """replicates the setup half of block_reconstruction/layer_reconstruction for every unit, keeping the engines"""
    from shiftedscalequantization_b200.engine import ReconEngine
    from shiftedscalequantization_b200.quant.adaptive_rounding import AdaRoundQuantizer
    from shiftedscalequantization_b200.quant.data_utils import save_inp_oup_data
    engines, feats_out = [], []
    for ui, unit in enumerate(recon_units(Q, qnn)):
        is_block = isinstance(unit, Q.BaseQuantBlock)
        mods = [m for m in unit.modules() if isinstance(m, Q.QuantModule)]
        qnn.set_quant_state(False, False)
        unit.set_quant_state(True, act_quant)
        aqs = []
        if not act_quant:
            for m in mods:
                if not isinstance(m.weight_quantizer, AdaRoundQuantizer):
                    m.weight_quantizer = AdaRoundQuantizer(uaq=m.weight_quantizer, round_mode='learned_hard_sigmoid',
                                                           weight_tensor=m.org_weight.data)
                m.weight_quantizer.soft_targets = True
        else:
            if is_block:
                aqs.append(unit.act_quantizer)
            aqs += [m.act_quantizer for m in mods if m.act_quantizer.delta is not None]
        if feats is None:
            torch.cuda.synchronize(dev); t_c = time.perf_counter()
            inps, outs = save_inp_oup_data(qnn, unit, cali, True, act_quant, BATCH)
            torch.cuda.synchronize(dev); CAPTURE_S.append(time.perf_counter() - t_c)
        else:
            inps, outs = feats[ui]
        feats_out.append((inps, outs))
        kw = dict(RECON)
        if act_quant:
            kw.update(p=2.4)
        eng = ReconEngine(unit, mods, inps, outs, None, act_quant=act_quant, iters=SCHED_ITERS, lr=4e-4, opt_mode='mse',
                          batch_size=BATCH, multi_gpu=multi_gpu, act_quantizers=aqs, use_graph=True, verbose=False,
                          host_resident=host_resident, device=dev, scaling=scaling, host_stage=host_stage, **kw)
        # time a representative point of the 20k schedule: past warm-up, so the regulariser path is live
        eng.step_dev.fill_(int(SCHED_ITERS * 0.5)); eng.host_step = int(SCHED_ITERS * 0.5)
        eng.capture()
        qnn.set_quant_state(False, False)
        unit.set_quant_state(True, act_quant)
        engines.append(eng)
        for m in mods:                                   # what the finished unit looks like to later units
            if not act_quant:
                m.weight_quantizer.soft_targets = False
    return engines, feats_out


def release(engines):
    for e in engines:
        e.close()


def timed_steps(engines, steps, warmup, dev, world, read_loss=False):
    import torch.distributed as td
    if read_loss:
        for e in engines:
            e.enable_loss_readback()
    for _ in range(warmup):
        for e in engines:
            e.unit.set_quant_state(True, e.act_quant)
            e.step()
            if read_loss:
                e.read_loss()
    torch.cuda.synchronize(dev)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    for _ in range(steps):
        for e in engines:
            e.step()
            if read_loss:
                e.read_loss()                            # D2H read of this unit's previous result (pinned ring, lag 1)
    if read_loss:
        for e in engines:
            e.read_loss(latest=True)                     # drain: the last step's results are read inside the timed region
    t1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize(dev)
    ms_dev = t0.elapsed_time(t1)
    ms_wall = 1e3 * (time.perf_counter() - w0)
    ms = max(ms_dev, ms_wall) if read_loss else ms_dev
    if world > 1:
        t = torch.tensor([ms], device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms = float(t)
    return ms / steps


def kernel_microbench(dev, peak_gbs):
    """fake-quant / loss kernels on DRAM-resident tensors (working set >> 126 MB L2), CUDA events, 3 warm-ups.
    Shapes from SURVEY.md §8d: weights [4096,4096,3,3] (604 MB), activations [256,256,56,56] (822 MB)."""
    from shiftedscalequantization_b200 import ops
    res = {}

    def timeit(fn, nbytes, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"ms": ms, "gbs": gbs, "frac": gbs / peak_gbs, "bytes": nbytes}

    torch.manual_seed(0)
    oc, k = 4096, 4096 * 9
    n = oc * k
    w = torch.randn(oc, 4096, 3, 3, device=dev) * 0.02
    d = (w.abs().amax(dim=(1, 2, 3), keepdim=True) / 3 * 1.2).contiguous()
    z = torch.full_like(d, 2.0)
    alpha = ops.adaround_init_alpha(w, d)
    g = torch.randn_like(w)
    bdev = ops.scalar_dev(11.0, dev)
    res["fq_affine_fwd(weights,per-channel)"] = timeit(lambda: ops.fq_affine_fwd(w, d, z, 0.0, 3.0), 8 * n)
    res["fq_affine_bwd(weights,per-channel)"] = timeit(lambda: ops.fq_affine_bwd(g, w, d, z, 0.0, 3.0), 12 * n)
    res["fq_adaround_fwd(+reg)"] = timeit(lambda: ops.adaround_fwd(w, alpha, d, z, 0.0, 3.0, True, bdev, 0.01, want_reg=True), 12 * n)
    ga = torch.empty_like(w)
    res["fq_adaround_bwd(+reg grad)"] = timeit(lambda: ops.adaround_bwd(g, w, alpha, d, z, 0.0, 3.0, bdev, 0.01, out=ga), 16 * n)
    m = torch.zeros_like(w); v = torch.zeros_like(w)
    step = torch.ones(1, dtype=torch.int64, device=dev); lr = ops.scalar_dev(1e-3, dev)
    res["adam_step"] = timeit(lambda: ops.adam_step(alpha, g, m, v, lr, step), 28 * n)
    # launch 3 of an iteration: alpha gradient + Adam in one pass (read g_wq, w, alpha, m, v; write alpha, m, v = 32 B/element)
    a2 = alpha.clone()
    tab = ops.AdaRoundTable([dict(w=w, alpha=a2, delta=d, zero_point=z, wq=ga, galpha=None, qmin=0.0, qmax=3.0)])
    lr0 = ops.scalar_dev(0.0, dev)                               # lr = 0: the timed launches leave alpha where it is
    res["fq_adaround_bwd_adam(+reg grad)"] = timeit(lambda: tab.backward_adam([g], bdev, 0.01, a2, m, v, lr0, step), 32 * n)
    del a2, tab
    # K1c: shifted-scale mixture, S = 3 shifts, one group per input channel (alpha [IC,3]); candidates recomputed from w
    shifts = [0.96875, 1.03125, 1.0]
    sd = torch.stack([(d * st).reshape(-1) for st in shifts]).contiguous()
    pr = ops.shift_probs_fwd(torch.randn(4096, 3, device=dev))
    dflat, zflat = d.reshape(-1).contiguous(), z.reshape(-1).contiguous()
    res["fq_shift_fwd(adaShift,soft)"] = timeit(lambda: ops.fq_shift_fwd(w, sd, dflat, zflat, pr, alpha, ops.SHIFT_ADASHIFT, False, False, 0.0, 3.0, False), 12 * n)
    res["fq_shift_bwd(adaShift,soft)"] = timeit(lambda: ops.fq_shift_bwd(g, w, sd, dflat, zflat, pr, alpha, ops.SHIFT_ADASHIFT, False, 0.0, 3.0, False, True), 16 * n)
    res["fq_shift_fwd(dequant mix)"] = timeit(lambda: ops.fq_shift_fwd(w, sd, dflat, zflat, pr, None, ops.SHIFT_DEQUANT, False, False, 0.0, 3.0, False), 8 * n)
    res["fq_shift_bwd(dequant mix)"] = timeit(lambda: ops.fq_shift_bwd(g, w, sd, dflat, zflat, pr, None, ops.SHIFT_DEQUANT, False, 0.0, 3.0, False, False), 8 * n)
    # the same bytes as a 1x1 convolution [8192,18432,1,1]: every element its own group (contiguous probabilities, float4 loads)
    w1 = w.view(8192, 18432, 1, 1); b1 = alpha.view_as(w1); g1 = g.view_as(w1)
    d1 = (w1.abs().amax(dim=(1, 2, 3)) / 3 * 1.2).contiguous(); z1 = torch.full_like(d1, 2.0)
    sd1 = torch.stack([d1 * st for st in shifts]).contiguous()
    pr1 = ops.shift_probs_fwd(torch.randn(18432, 3, device=dev))
    res["fq_shift_fwd(adaShift,soft,1x1)"] = timeit(lambda: ops.fq_shift_fwd(w1, sd1, d1, z1, pr1, b1, ops.SHIFT_ADASHIFT, False, False, 0.0, 3.0, False), 12 * n)
    res["fq_shift_bwd(adaShift,soft,1x1)"] = timeit(lambda: ops.fq_shift_bwd(g1, w1, sd1, d1, z1, pr1, b1, ops.SHIFT_ADASHIFT, False, 0.0, 3.0, False, True), 16 * n)
    del w1, b1, g1
    packed = ops.export_codes(w, d, z, 0.0, 3.0, 2, alpha=alpha)
    res["export_codes(2-bit,+alpha)"] = timeit(lambda: ops.export_codes(w, d, z, 0.0, 3.0, 2, alpha=alpha), 8 * n + n // 4)
    res["import_codes(2-bit)"] = timeit(lambda: ops.import_codes(packed, w.shape, d, z, 0.0, 2), 4 * n + n // 4)
    del w, alpha, g, ga, m, v, packed
    torch.cuda.empty_cache()
    x = torch.relu(torch.randn(256, 256, 56, 56, device=dev))
    t = torch.relu(torch.randn(256, 256, 56, 56, device=dev))
    na = x.numel()
    ds, zs = ops.scalar_dev(0.25, dev).reshape(()), ops.scalar_dev(0.0, dev).reshape(())
    res["fq_affine_fwd(acts,per-tensor)"] = timeit(lambda: ops.fq_affine_fwd(x, ds, zs, 0.0, 15.0), 8 * na)
    res["fq_affine_bwd(acts,per-tensor)"] = timeit(lambda: ops.fq_affine_bwd(t, x, ds, zs, 0.0, 15.0), 12 * na)
    res["recon_loss(fwd+dpred)"] = timeit(lambda: ops.recon_loss(x, t, 2.0), 12 * na)
    res["recon_loss(fwd+dpred,p=2.4)"] = timeit(lambda: ops.recon_loss(x, t, 2.4), 12 * na)
    idx = torch.randperm(256, device=dev)
    res["gather_rows"] = timeit(lambda: ops.gather_rows(x, idx, out=t), 8 * na)
    del x, t
    torch.cuda.empty_cache()
    return res


def scale_search_bench(Q, qnn, dev, peak_gbs):
    """K2a / K2b on the roofline: the 80-candidate MSE clip search over every weight tensor of the model
    (per-channel rows) and over one activation-sized tensor (grid-wide variant), row shapes k = 9 / 576 / 4608 on their own,
    and the ChannelQuantMSE input-scale search. CUDA events, after warm-up. K2a is instruction-bound by design (80 candidates
    x ~14 instructions + 2 MUFU per 4 bytes), so it reports candidate evaluations/s and the fraction of the SM's issue rate
    beside the (small) HBM fraction; K2b is one pass over the weights and reports the HBM fraction of its 4 B/element."""
    from shiftedscalequantization_b200 import ops
    mods = [m for m in qnn.modules() if isinstance(m, Q.QuantModule)]
    rows = [(m.org_weight.reshape(m.org_weight.shape[0], -1).contiguous(), m.weight_quantizer.n_levels) for m in mods]

    def timeit(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def run():
        for r, nl in rows:
            ops.mse_scale_search(r, nl, False)
    w_ms = timeit(run, reps=3, warm=1)
    # what QuantModel.forward does: every layer's search issued together, round-robin on 4 side streams
    w_many_ms = timeit(lambda: ops.mse_scale_search_many([(r, nl, False) for r, nl in rows]), reps=3, warm=1)
    act = torch.relu(torch.randn(64, 64, 112, 112, device=dev)).reshape(1, -1)
    a_ms = timeit(lambda: ops.mse_scale_search(act, 16, False), reps=3, warm=1)
    # issue-rate roof: 148 SMs x 4 schedulers x 1 warp-instruction/clk; ncu counts 25 warp-instructions per 32 (element,
    # candidate) pairs for the row kernel (profiles/r02_k2_ncu.md: smsp__inst_executed / pairs, 3.2 of them on the XU pipe: MUFU.LG2,
    # MUFU.EX2, FRND at 8 cycles each), so pairs/s <= SMs*4*32*clk/25 — the XU pipe gives the same bound within 5 %
    sm_clk = 1.965e9
    pair_roof = 148 * 4 * 32 * sm_clk / 25.0

    def k2a(x2d, nl):
        ms = timeit(lambda: ops.mse_scale_search(x2d, nl, False))
        n = x2d.numel()
        return {"ms": round(ms, 4), "rows": x2d.shape[0], "k": x2d.shape[1], "elems": n, "gbs": round(4 * n / ms / 1e6, 1),
                "hbm_frac": round(4 * n / ms / 1e6 / peak_gbs, 4), "cand_evals_per_s": round(80 * n / ms * 1e3, 0),
                "issue_frac_est": round(80 * n / (ms * 1e-3) / pair_roof, 3)}
    out = {"weights_all_layers_ms": w_many_ms, "weights_all_layers_one_stream_ms": w_ms, "channels": int(sum(r.shape[0] for r, _ in rows)),
           "weight_elems": int(sum(r.numel() for r, _ in rows)), "activation_tensor_ms": a_ms, "activation_elems": int(act.numel()),
           "round1_ms": {"weights_all_layers": 4.63, "activation_tensor": 17.07, "source": "BENCH_r01.json extra.scale_search"},
           "k2a": {}, "k2b": {}}
    torch.manual_seed(3)
    out["k2a"]["rows_k9(depthwise)"] = k2a(torch.randn(65536, 9, device=dev) * 0.05, 8)
    out["k2a"]["rows_k576"] = k2a(torch.randn(16384, 576, device=dev) * 0.05, 4)
    out["k2a"]["rows_k4608"] = k2a(torch.randn(4096, 4608, device=dev) * 0.05, 4)
    out["k2a"]["tensor_51M(activations)"] = {**k2a(act, 16)}
    del act

    def k2b(oc, k, level, bits=2, thr=1.5):
        w = torch.randn(oc, k, device=dev) * 0.02
        mn = w.min(1)[0].clamp(max=0); mx = w.max(1)[0].clamp(min=0)
        L = 2 ** bits
        d = ((mx - mn) / (L - 1)).contiguous(); raw = (-mn).contiguous()
        cand = torch.tensor([i / level for i in range(level, 0, -1)], dtype=torch.float32, device=dev)
        lo = float(torch.tensor(0.0 - 0.5 / (L - 1) * thr, dtype=torch.float32)); hi = float(torch.tensor(1.0 + 0.5 / (L - 1) * thr, dtype=torch.float32))
        s = torch.ones(k, device=dev)
        ms = timeit(lambda: ops.inp_scale_search(w, d, raw, cand, L - 1, lo, hi, s))
        ms_b = timeit(lambda: ops.inp_scale_search(w, d, raw, cand, L - 1, lo, hi, s, force_brute=True), reps=2, warm=1) \
            if oc * k * level <= 2 ** 31 else None
        n = oc * k
        return {"ms": round(ms, 4), "oc": oc, "k": k, "level": level, "gbs": round(4 * n / ms / 1e6, 1),
                "hbm_frac": round(4 * n / ms / 1e6 / peak_gbs, 3), "brute_force_ms": None if ms_b is None else round(ms_b, 4)}
    out["k2b"]["[512,256,3,3]_level16"] = k2b(512, 2304, 16)
    out["k2b"]["[512,256,3,3]_level1024"] = k2b(512, 2304, 1024)
    out["k2b"]["[4096,4096,3,3]_level16(604MB)"] = k2b(4096, 36864, 16)
    out["k2b"]["[4096,4096,3,3]_level1024(604MB)"] = k2b(4096, 36864, 1024)
    torch.cuda.empty_cache()
    out["note"] = ("K2a: MUFU-ranked candidates, libm-settled survivors (identical argmin); compute-bound (80 candidates x |d|^2.4 per 4 bytes): "
                   "issue_frac_est = candidate evaluations/s over the issue-rate roof at the 25 warp-instructions per 32 evaluations ncu counts "
                   "(1.965 GHz); measured smsp__issue_active 76-83 % and XU-pipe ~75 % at these shapes: profiles/r02_k2_ncu.md. K2b: one pass, 4 B/element. Reference: Python loops, 48.5 s for the ResNet-18 weights on 8 CPU cores (SURVEY probe)")
    return out


def shifted_loop_bench(dev, n_images=256):
    """a17/a18: the shifted-scale loops through their public functions (block_recon_fused_shiftedScale,
    block_recon_shiftedScale, layer_recon_shiftedScale), captured-graph iteration vs the eager autograd loop; device time
    of the iteration loop only (ChannelQuant init, capture and read-outs excluded), real 224x224 unit shapes"""
    import contextlib
    import io
    from shiftedscalequantization_b200 import quant as Q, zoo
    from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
    from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
    from shiftedscalequantization_b200.quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale
    out = {}
    cases = [("resnet18", 2, "model.layer2.0", "fused"), ("resnet18", 2, "model.layer4.1", "shift"),
             ("resnet50", 4, "model.layer3.2.conv2", "layer_shift")]
    for arch, bits, path, kind in cases:
        res = {}
        for captured, iters in ((True, 600), (False, 100)):
            torch.manual_seed(1005)
            cnn = zoo.build(arch).to(dev).eval()
            qnn = Q.QuantModel(cnn, {'n_bits': bits, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).to(dev).eval()
            qnn.set_first_last_layer_to_8bit()
            cali = torch.randn(n_images, 3, 224, 224)
            qnn.set_quant_state(True, False)
            with torch.no_grad():
                qnn(cali[:32].to(dev))
            unit = qnn
            for part in path.split('.'):
                unit = unit[int(part)] if part.isdigit() else getattr(unit, part)
            for m in ([unit] if isinstance(unit, Q.QuantModule) else [m for m in unit.modules() if isinstance(m, Q.QuantModule)]):
                m.weight_quantizer = ChannelQuant(1.0, uaq=m.weight_quantizer, weight_tensor=m.org_weight.data,
                                                  shiftTarget=[0.96875, 1.03125, 1.0], name=m.pathName)
            for mode, wq_on in (('if', True), ('of', False)):
                qnn.set_quant_state(wq_on, False)
                unit.cache_features = mode
                with torch.no_grad():
                    for i in range(0, n_images, BATCH):
                        qnn(cali[i:i + BATCH].to(dev))
                unit.cache_features = 'none'
            qnn.set_quant_state(False, False)
            unit.set_quant_state(True, False)
            LS.USE_CAPTURED_LOOP = captured
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                if kind == "fused":
                    block_recon_fused_shiftedScale(unit, iters=iters, lmda=[0.01, 0.01], model=qnn)
                elif kind == "shift":
                    LS.block_recon_shiftedScale(unit, iters=iters, lmda=0.01, model=qnn)
                else:
                    LS.layer_recon_shiftedScale(unit, iters=iters, lmda=0.01, model=qnn)
            st = dict(LS.LAST_LOOP_STATS)
            res["captured" if captured else "eager"] = {"iters_per_s": st["iters"] / (st["loop_ms"] * 1e-3), "iters": st["iters"],
                                                        "ssq_launches_per_iter": st["launches_per_iter"]}
            del qnn, cnn, unit
            torch.cuda.empty_cache()
        LS.USE_CAPTURED_LOOP = True
        res["speedup"] = res["captured"]["iters_per_s"] / res["eager"]["iters_per_s"]
        out[f"{arch}:{path}:{kind}"] = res
    return out


ENTRY_TO_MICRO = {"ssq_recon_loss": "recon_loss(fwd+dpred)", "ssq_gather_rows": "gather_rows", "ssq_adam_step": "adam_step",
                  "ssq_adam_step_end_iteration": "adam_step", "ssq_iter_prologue": "fq_adaround_fwd(+reg)",
                  "ssq_fq_adaround_bwd_adam_mt": "fq_adaround_bwd_adam(+reg grad)",
                  "ssq_fq_adaround_fwd_mt": "fq_adaround_fwd(+reg)", "ssq_fq_adaround_bwd_mt": "fq_adaround_bwd(+reg grad)",
                  "ssq_fq_affine_fwd": "fq_affine_fwd(acts,per-tensor)", "ssq_fq_affine_bwd": "fq_affine_bwd(acts,per-tensor)"}


def in_step_profile(engines, dev):
    """one eager (un-captured) pass over all units with a CUDA-event pair around each of our launches, plus the
    whole-pass time, to get each kernel's share of the step"""
    from shiftedscalequantization_b200 import ops
    graphs = [e.graph for e in engines]
    for e in engines:
        e.graph = None
    for e in engines:                    # warm the eager path
        e.unit.set_quant_state(True, e.act_quant); e.step()
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ops.profile_begin()
    e0.record()
    for e in engines:
        e.step()
    e1.record()
    prof = ops.profile_end()
    total = e0.elapsed_time(e1)
    for e, g in zip(engines, graphs):
        e.graph = g
    return prof, total


def public_api_bench(dev, n_images, iters=600):
    """the call a user makes: Q.block_reconstruction / Q.layer_reconstruction on fresh units (feature capture, engine set-up,
    graph capture, `iters` iterations, hard rounding) — wall time of the whole call and the loop-only rate, to be read beside
    the per-unit engine numbers (they must agree: it is the same engine)"""
    from shiftedscalequantization_b200.quant import block_recon as BR
    Q, qnn, cali = build_model(dev, n_images)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    kw = dict(cali_data=cali, iters=iters, weight=RECON['weight'], asym=True, b_range=RECON['b_range'], warmup=RECON['warmup'],
              act_quant=False, opt_mode='mse', batch_size=BATCH)
    out = {}
    for name, unit, fn in (("layer1.0", qnn.model.layer1[0], Q.block_reconstruction), ("layer1.1", qnn.model.layer1[1], Q.block_reconstruction),
                           ("layer2.0", qnn.model.layer2[0], Q.block_reconstruction)):
        torch.cuda.synchronize(dev); t0 = time.perf_counter()
        fn(qnn, unit, **kw)
        torch.cuda.synchronize(dev); wall = time.perf_counter() - t0
        st = dict(BR.LAST_RUN_STATS)
        out[name] = {"call_s": round(wall, 3), "loop_iters_per_s": round(st["iters"] / (st["loop_ms"] * 1e-3), 1), "iters": st["iters"],
                     "capture_s": round(st.get("capture_s", 0.0), 3), "capture_mode": st.get("capture_mode"),
                     "ssq_launches_per_iter": st["launches_per_iter"]}
    del qnn
    torch.cuda.empty_cache()
    return out


def readme_flags_bench(dev, n_images, iters=300):
    """configs[1] as written: `--bias_cal --bias_ch_quant` (README.md:20,33-34; no code upstream, semantics in DESIGN.md §6) on
    all nine units through the public functions: bias_cal alone = block/layer_reconstruction(bias_cal=True) (gamma^z, varphi^z
    learned with the AdaRound alphas); with bias_ch_quant the BasicBlock units go through ChannelQuant +
    block_recon_fused_shiftedScale(bias_cal=True) (input-channel group R). Loop-only iterations/s per unit."""
    from shiftedscalequantization_b200.quant import block_recon as BR
    from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
    from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
    from shiftedscalequantization_b200.quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale
    import contextlib
    import io
    out = {}
    for flags in ("bias_cal", "bias_cal+bias_ch_quant"):
        Q, qnn, cali = build_model(dev, n_images)
        qnn.set_quant_state(True, False)
        with torch.no_grad():
            qnn(cali[:64].to(dev))
        kw = dict(cali_data=cali, iters=iters, weight=RECON['weight'], asym=True, b_range=RECON['b_range'], warmup=RECON['warmup'],
                  act_quant=False, opt_mode='mse', batch_size=BATCH, bias_cal=True)
        rates = {}
        for unit in recon_units(Q, qnn):
            name = getattr(unit, "pathName", "") or "fc"
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                if isinstance(unit, Q.BaseQuantBlock) and "ch_quant" in flags:
                    for m in unit.modules():
                        if isinstance(m, Q.QuantModule):
                            m.weight_quantizer = ChannelQuant(1.0, uaq=m.weight_quantizer, weight_tensor=m.org_weight.data,
                                                              shiftTarget=[0.96875, 1.03125, 1.0], name=m.pathName)
                    unit.clear_cached_features()
                    for mode, wq_on in (('if', True), ('of', False)):
                        qnn.set_quant_state(wq_on, False)
                        unit.cache_features = mode
                        with torch.no_grad():
                            for i in range(0, cali.shape[0], BATCH):
                                qnn(cali[i:i + BATCH].to(dev))
                        unit.cache_features = 'none'
                    qnn.set_quant_state(False, False)
                    unit.set_quant_state(True, False)
                    block_recon_fused_shiftedScale(unit, iters=iters, lmda=[RECON['weight'], RECON['weight']], model=qnn, bias_cal=True)
                    unit.clear_cached_features()
                    st = dict(LS.LAST_LOOP_STATS)
                elif isinstance(unit, Q.BaseQuantBlock):
                    Q.block_reconstruction(qnn, unit, **kw)
                    st = dict(BR.LAST_RUN_STATS)
                else:
                    Q.layer_reconstruction(qnn, unit, **kw)
                    st = dict(BR.LAST_RUN_STATS)
            rates[name] = {"iters_per_s": round(st["iters"] / (st["loop_ms"] * 1e-3), 1), "ssq_launches_per_iter": st["launches_per_iter"]}
        hm = len(rates) / sum(1.0 / v["iters_per_s"] for v in rates.values())
        out[flags] = {"per_unit": rates, "harmonic_mean_iters_per_s": round(hm, 1), "iters_per_unit": iters}
        del qnn
        torch.cuda.empty_cache()
    return out


def code_agreement_bench(dev):
    """north_star: hard integer codes vs the REAL reference after a long loop. tests/golden/long_horizon.npz holds the codes the
    reference produced on the CPU (block_reconstruction on layer1.0, layer_reconstruction on fc of a seeded ResNet-18, 64 randn
    32x32 images, 2 000 iterations; tests/golden/make_golden_round2.py); the same flow runs here through the public API and
    the fraction of identical codes is reported next to the reference's agreement with ITSELF under another CPU convolution
    backend (the noise floor: cuDNN and CPU convolutions differ in the last bits, Adam amplifies that)."""
    import numpy as np
    from shiftedscalequantization_b200 import ops, quant as Q, zoo
    path = os.path.join(ROOT, "tests", "golden", "long_horizon.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    out = {}
    for iters in (2000,):
        torch.manual_seed(1005)
        cnn = zoo.resnet18(num_classes=10).to(dev).eval()
        qnn = Q.QuantModel(cnn, {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).to(dev).eval()
        qnn.set_first_last_layer_to_8bit()
        cali = torch.randn(64, 3, 32, 32)
        qnn.set_quant_state(True, False)
        with torch.no_grad():
            qnn(cali[:32].to(dev))
        kw = dict(cali_data=cali, iters=iters, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=False, opt_mode='mse', batch_size=32)
        block = qnn.model.layer1[0]
        torch.manual_seed(377)
        Q.block_reconstruction(qnn, block, **kw)
        torch.manual_seed(378)
        Q.layer_reconstruction(qnn, qnn.model.fc, **kw)
        rep = {}
        for name, m in (("block.conv1", block.conv1), ("block.conv2", block.conv2), ("fc", qnn.model.fc)):
            q = m.weight_quantizer
            _, codes = ops.adaround_fwd(m.org_weight.detach(), q.alpha.detach(), q.delta.detach(), q.zero_point.detach(), 0.0,
                                        float(q.n_levels - 1), soft=False, want_codes=True)
            same = codes.cpu().numpy().astype(np.uint8) == g[f"i{iters}.{name}.codes"]
            key = f"i{iters}.{name}.self_agreement"
            rep[name] = {"identical_codes": float(same.mean()), "codes": int(same.size),
                         "reference_vs_itself": float(g[key]) if key in g.files else None}
        out[f"{iters}_iterations"] = rep
    out["note"] = ("fraction of hard integer codes identical to the real reference's (CPU) after the loop; 20 000 iterations: "
                   "tests/test_round2_gpu.py::test_long_horizon_code_agreement, numbers in profiles/r02_parity_report.json")
    return out


def run_ours(args):
    from shiftedscalequantization_b200 import dist as D
    from shiftedscalequantization_b200 import ops
    rank, local, world = D.init_from_env()
    if world > 1:
        log(f"[rank {int(os.environ.get('RANK', 0))}] bound to {D.bind_to_gpu_cpus(local)} GPU-local cores (0 = unchanged)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)   # pass 0 under a profiler: autotuning mis-times there
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)       # default: true fp32 convolutions, like the CPU path
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    if args.micro_only:
        micro = kernel_microbench(dev, peak_gbs)
        for k, v in micro.items():
            emit(f"{k:40s} {v['ms']:8.4f} ms  {v['gbs']:8.1f} GB/s  {v['frac']:.3f} of {peak_gbs:.0f}")
        return
    n_total = args.images
    lo, hi = D.shard_range(n_total, rank, world)
    t_setup = time.perf_counter()
    Q, qnn, cali = build_model(dev, n_total)
    strong = world > 1 and args.scaling == "strong"
    if not strong:
        cali = cali[lo:hi]            # weak: each rank owns a shard; strong: every rank holds the whole cache
    # weight-scale init: the first quantised forward runs the MSE search of all 21 layers (K2a)
    # an FP forward first: CUDA module loading, cuDNN autotuning of the 21 convolutions and the allocator's first blocks are paid
    # here, so the next number is the scale search itself (21 K2a launches + the host glue of init_quantization_scale)
    qnn.set_quant_state(False, False)
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    torch.cuda.synchronize(dev); fp_first_s = time.perf_counter() - t0
    qnn.set_quant_state(True, False)
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    with torch.no_grad():
        qnn(cali[:64].to(dev))
    torch.cuda.synchronize(dev); scale_search_s = time.perf_counter() - t0
    log(f"[rank {rank}] first FP forward (CUDA/cuDNN warm-up) {fp_first_s * 1e3:.1f} ms; first quantised forward = weight scale search "
        f"(5800 channels x 80 candidates) {scale_search_s * 1e3:.1f} ms")
    search = scale_search_bench(Q, qnn, dev, peak_gbs) if world == 1 else None
    if args.k2_only:
        emit(json.dumps({"first_fp_forward_ms": fp_first_s * 1e3, "first_quantised_forward_ms": scale_search_s * 1e3, "scale_search": search}))
        return
    engines, feats = make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=world > 1, scaling=args.scaling)
    setup_s = time.perf_counter() - t_setup
    launches_per_step = sum(e.launches_per_iter for e in engines)
    capture_s = list(CAPTURE_S)
    log(f"[rank {rank}] feature capture per unit (s): {[round(c, 2) for c in capture_s]}")
    log(f"[rank {rank}] setup {setup_s:.1f}s, {len(engines)} units, {launches_per_step} ssq launches per step")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step = timed_steps(engines, args.steps, args.warmup, dev, world)
    clocks = sampler.stop() if rank == 0 else None
    n_units = len(engines)
    work = 1 if strong else world      # strong: the ranks split ONE batch-32 iteration; weak: each rank completes its own
    value = work * n_units / (ms_step * 1e-3)

    extra = {}
    per_unit = None
    if world == 1:
        # SURVEY §8d protocol: per-unit iterations/s (each unit replayed on its own, CUDA events) + harmonic mean
        per_unit = {}
        for ui, e in enumerate(engines):
            for _ in range(5):
                e.step()
            torch.cuda.synchronize(dev)
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            n_unit_iters = max(args.steps, args.unit_iters)      # >= 500 timed iterations per unit
            for _ in range(n_unit_iters):
                e.step()
            t1.record(); torch.cuda.synchronize(dev)
            us = 1e3 * t0.elapsed_time(t1) / n_unit_iters
            name = getattr(e.unit, "pathName", "") or f"unit{ui}"
            per_unit[name] = {"us_per_iter": round(us, 1), "iters_per_s": round(1e6 / us, 1),
                              "alpha_elems": int(e.flat.numel()), "ssq_launches": e.launches_per_iter}
        hm = len(per_unit) / sum(1.0 / v["iters_per_s"] for v in per_unit.values())
        per_unit["harmonic_mean_iters_per_s"] = round(hm, 1)
        per_unit["timed_iterations_per_unit"] = max(args.steps, args.unit_iters)
        prof, eager_ms = in_step_profile(engines, dev)
    release(engines)
    e2e = None
    # ---- end to end: features in pinned host memory, per-step H2D of the mini-batch, D2H of the loss
    if not args.skip_e2e:
        # the pinned cache is placed by first touch: allocate it (and run the step) from the GPU's own NUMA node, as the multi-rank
        # launch does for the whole process; the affinity is restored afterwards so that the CPU baseline keeps every host core
        saved_affinity = os.sched_getaffinity(0)
        bound = D.bind_to_gpu_cpus(local) if world == 1 else 0
        log(f"[rank {rank}] e2e: bound to {bound} GPU-local cores (0 = unchanged) of {len(saved_affinity)}")
        eng_h, _ = make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=world > 1, host_resident=True, feats=feats,
                                scaling=args.scaling, host_stage=args.host_stage)
        ms_e2e = timed_steps(eng_h, max(args.steps // 2, 3), max(args.warmup // 2, 3), dev, world, read_loss=True)
        if world == 1:
            os.sched_setaffinity(0, saved_affinity)
        e2e = {"value": work * n_units / (ms_e2e * 1e-3), "unit": "iters/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": sum(e.h2d_bytes_per_step() for e in eng_h), "d2h_bytes_per_step": 4 * n_units,
               "h2d_dense_bytes_per_step": sum(4 * (e.cur_inp.numel() + e.cur_out.numel()) for e in eng_h),
               "host_stage": args.host_stage,
               "path": "ReconEngine(host_resident=True, host_stage='pull'): pinned host feature cache, post-ReLU tensors kept zero-packed "
                       "(non-zero values on the host; bit mask + chunk offsets, 3 % of the dense size, on the device); inside each captured "
                       "iteration a 16-24-CTA kernel (ssq_pull_rows_host[_packed]) reads the NEXT mini-batch's input and target rows out "
                       "of mapped host memory over PCIe beside the current iteration's kernels and expands them to dense rows "
                       "(h2d_bytes_per_step = the bytes that cross PCIe, h2d_dense_bytes_per_step = the dense rows they become); every "
                       "iteration's loss is copied to a pinned 2-deep ring and read on the host one launch later (all losses read inside "
                       "the timed region)"}
        log(f"[rank {rank}] e2e per-unit PCIe bytes / dense bytes: "
            + ", ".join(f"{e.h2d_bytes_per_step() / max(4 * (e.cur_inp.numel() + e.cur_out.numel()), 1):.2f}" for e in eng_h))
        release(eng_h)
    torch.cuda.empty_cache()

    if world > 1:
        if rank == 0:
            line = base_line(args, value, ms_step, world, clocks, launches_per_step, e2e, peak_src)
            emit(json.dumps(line))
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
        return

    # ---- activation phase (W2A4's second half): LSQ step-size learning, 5000 iters/unit upstream
    act = None
    if not args.skip_act:
        qnn.set_quant_state(True, True)
        with torch.no_grad():
            qnn(cali[:64].to(dev))
        qnn.disable_network_output_quantization()
        eng_a, _ = make_engines(Q, qnn, cali, dev, act_quant=True, multi_gpu=False)
        ms_a = timed_steps(eng_a, max(args.steps // 2, 3), 3, dev, 1)
        act = {"iters_per_s": n_units / (ms_a * 1e-3), "ms_per_step": ms_a, "launches_per_step": sum(e.launches_per_iter for e in eng_a)}
        release(eng_a)
        del eng_a
        torch.cuda.empty_cache()
    tf32_extra = None
    if not args.skip_tf32 and not args.tf32:
        # PyTorch's own default (cudnn.allow_tf32=True) for the convolutions the path does not own; reported beside the
        # fp32 headline, never instead of it
        torch.backends.cudnn.allow_tf32 = True
        qnn.set_quant_state(True, False)
        eng_t, _ = make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=False, feats=feats)
        ms_t = timed_steps(eng_t, max(args.steps // 2, 3), 3, dev, 1)
        tf32_extra = {"iters_per_s": n_units / (ms_t * 1e-3), "ms_per_step": ms_t, "conv_math": "tf32 (torch default)"}
        release(eng_t)
        del eng_t
        torch.backends.cudnn.allow_tf32 = False
    del feats
    del engines, qnn
    torch.cuda.empty_cache()
    shifted = shifted_loop_bench(dev) if not args.skip_shift else None
    more = {}
    if not args.skip_extras:
        more["public_api"] = public_api_bench(dev, min(n_total, 256))
        more["readme_flags"] = readme_flags_bench(dev, min(n_total, 256))
        more["code_agreement"] = code_agreement_bench(dev)
        g = reference_loop(steps=5, warmup=2, device=str(dev))
        more["gpu_reference"] = {"iters_per_s": g["iters_per_s"], "ms_per_step": g["ms_per_step"], "sample": g["sample"],
                                 "note": "the reference's own op chain (~530 ATen launches per iteration) on this B200, fp32 cuDNN: "
                                         "separates 'B200 vs host CPU' from '3 launches vs 530'"}

    # ---- roofline: DRAM-resident microbench of every kernel + in-step shares
    if args.skip_micro:
        line = base_line(args, value, ms_step, world, clocks, launches_per_step, e2e, peak_src)
        line["extra"] = {"first_fp_forward_ms": fp_first_s * 1e3, "first_quantised_forward_ms": scale_search_s * 1e3, "scale_search": search, "setup_s": setup_s, "act_phase": act,
                         "shifted_loops": shifted, "per_unit": per_unit, **more}
        emit(json.dumps(line))
        return
    micro = kernel_microbench(dev, peak_gbs)
    ours_ms = sum(ms for _n, ms in prof.values())
    shares = {k: {"launches": n, "ms": ms, "share_of_step": ms / eager_ms} for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    dominant = next(iter(shares))
    mk = ENTRY_TO_MICRO.get(dominant, "recon_loss(fwd+dpred)")
    traffic = None
    try:    # dram__bytes_read+write of one launch at this microbench shape, from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_kernels_ncu_full.json")))
        try:
            prof.update(json.load(open(os.path.join(ROOT, "profiles", "r02_kernels_ncu_full.json"))))
        except Exception:
            pass
        key = {"fq_adaround_fwd(+reg)": "ada_fwd_kernel", "fq_adaround_bwd(+reg grad)": "ada_bwd_kernel", "adam_step": "adam_kernel",
               "fq_adaround_bwd_adam(+reg grad)": "ada_bwd_adam_mt_kernel",
               "recon_loss(fwd+dpred)": "recon_loss_kernel", "gather_rows": "gather_rows_kernel",
               "fq_affine_fwd(weights,per-channel)": "fq_affine_fwd_vec",
               "fq_affine_bwd(weights,per-channel)": "fq_affine_bwd_kernel"}.get(mk)    # captures taken at exactly these shapes
        if key in prof:
            traffic = prof[key]["traffic_MB"] * 1e6
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dominant, "achieved": micro[mk]["gbs"], "peak": peak_gbs, "unit": "GB/s",
            "frac": micro[mk]["frac"], "traffic": traffic, "algorithmic_bytes": micro[mk]["bytes"], "peak_source": peak_src,
            "measured_on": f"{mk}: DRAM-resident microbench ({micro[mk]['bytes'] / 1e6:.0f} MB algorithmic per launch, inputs > L2), "
                           "CUDA events, 3 warm-ups + 10 timed launches",
            "in_step": {"eager_step_ms": eager_ms, "ssq_kernels_ms": ours_ms, "shares": shares},
            "all_kernels": {k: {"gbs": round(v["gbs"], 1), "frac": round(v["frac"], 3), "ms": round(v["ms"], 4)} for k, v in micro.items()}}

    cpu = cpu_reference(steps=2, warmup=1) if not args.skip_cpu else None
    line = base_line(args, value, ms_step, world, clocks, launches_per_step, e2e, peak_src)
    line["roofline"] = roof
    if cpu:
        line["cpu_baseline"] = {"value": cpu["iters_per_s"], "unit": "iters/s", "cores": cpu["cores"], "kind": "port", "sample": cpu["sample"]}
    fq = {k: round(v["gbs"], 1) for k, v in micro.items() if k.startswith("fq_")}
    line["extra"] = {"first_fp_forward_ms": fp_first_s * 1e3, "first_quantised_forward_ms": scale_search_s * 1e3, "scale_search": search, "setup_s": setup_s, "act_phase": act,
                     "fake_quant_hbm_gbs": fq, "fake_quant_hbm_frac_min": round(min(v["frac"] for k, v in micro.items() if k.startswith("fq_")), 3),
                     "tf32": tf32_extra, "shifted_loops": shifted, "per_unit": per_unit,
                     "feature_capture_s": {"per_unit": [round(c, 3) for c in capture_s], "total": round(sum(capture_s), 3),
                                           "note": "save_inp_oup_data (quant/data_utils.py:8-37), 1024 images, batch 32, asym=True: excluded from iters/s"},
                     "projected_full_run_s": (9 * 20000) / value + ((9 * 5000) / act["iters_per_s"] if act else 0), **more}
    emit(json.dumps(line))


def bench_config(args, world):
    """the workload description; identical keys and values in both arms (`--impl ours` / `--impl reference`)"""
    strong = world > 1 and args.scaling == "strong"
    return {"workload": WORKLOAD,
            "calib_images": args.images, "per_rank_batch": BATCH // world if strong else BATCH,
            "global_batch": BATCH if strong else BATCH * world,
            "step": "one iteration on each of the 9 units",
            "conv_math": "tf32" if args.tf32 else "fp32", "cudnn_benchmark": bool(args.cudnn_benchmark),
            "l2": "inputs larger than L2: one step streams ~215 MB of mini-batch features (random rows of the 6.4 GB cache) plus the "
                  "weights / alpha / Adam state of all 9 units (~230 MB), so nothing a unit touches survives in the 126 MB L2 until its "
                  "next iteration; the roofline kernels are timed on 0.6-0.8 GB tensors",
            "multi_gpu": ("single GPU" if world == 1 else
                          "strong: every rank holds the cache, the ranks split one global mini-batch of 32, SUM of the flat "
                          "alpha gradient (= the 1-GPU gradient) every iteration" if strong else
                          "weak: calibration images sharded by rank, each rank draws its own mini-batch of 32, SUM of the "
                          "flat alpha gradient every iteration")}


def base_line(args, value, ms_step, world, clocks, launches_per_step, e2e, peak_src):
    return {"metric": "recon iters/s (ResNet-18 W2A4, 1024 calib imgs)", "value": value, "unit": "iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * (args.steps + args.warmup)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="multi-GPU partitioning (DESIGN.md §7)")
    ap.add_argument("--tf32", type=int, default=0)
    ap.add_argument("--cudnn-benchmark", type=int, default=1)
    ap.add_argument("--unit-iters", type=int, default=500, help="timed iterations per unit of the per-unit table (lower it only under a profiler)")
    ap.add_argument("--host-stage", default="pull", choices=["pull", "dma"], help="e2e: how the mini-batch crosses PCIe (SM pull inside the graph | per-row copy-engine transfers)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-act", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-tf32", action="store_true")
    ap.add_argument("--skip-shift", action="store_true", help="no shifted-scale loop timing")
    ap.add_argument("--micro-only", action="store_true", help="only the DRAM-resident kernel microbench")
    ap.add_argument("--skip-micro", action="store_true", help="no roofline microbench (short profiler runs)")
    ap.add_argument("--k2-only", action="store_true", help="only the scale-search (K2a / K2b) roofline numbers")
    ap.add_argument("--skip-extras", action="store_true", help="no public-API / README-flags / code-agreement / GPU-reference extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
