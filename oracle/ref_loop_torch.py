"""CPU restatement of the reference's reconstruction hot loop — TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference's loop (quant/block_recon.py:89-105, quant/layer_recon.py:79-96) is a chain of stock ATen CPU ops
plus torch.optim.Adam; this file restates that chain with the same ATen ops in the same order, on a functional
description of a unit (so it carries no product code):

    unit = {"kind": "layer" | "basic" | "bottleneck" | "invres",
            "layers": {name: {"weight", "bias", "conv": {stride,padding,dilation,groups} | None (linear),
                              "act": "relu" | "relu6" | None, "delta", "zero_point", "n_levels"}},
            "tail_act": "relu" | None, "use_res_connect": bool}

  UniformAffine / AdaRound forward   quant/quant_layer.py:92-97, quant/adaptive_rounding.py:49-64
  block forward                      quant/quant_block.py:99-117 (basic), :153-166 (bottleneck/regnet), :228-239
  LossFunction                       quant/block_recon.py:142-182
  LinearTempDecay                    quant/block_recon.py:185-202
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Pinned against the real reference run on the CPU (tests/test_host_cpu.py, tests/test_oracle_golden.py):
  recon_loop.npz   block loop body + the real layer_reconstruction (12 iterations): losses and alphas to 1e-6
  families.npz     the real block_reconstruction on a ResNet-50 bottleneck and a RegNetX-3200M block (16 iterations): alphas to 1e-5
  bias_cal.npz     alpha_out / beta_out in the optimiser (train_affine=True): loss trace, alphas, gamma, varphi to 1e-5
  act_phase.npz    the real block_reconstruction(act_quant=True): learned activation step sizes to 1e-5
  loss.npz         rec_loss (mse / fisher_diag / fisher_full): values and gradients
  long_horizon.npz 2 000 iterations of the real block_reconstruction: hard codes 100 % identical
                   (tests/golden/check_oracle_long_horizon.py; 5 minutes, run by hand)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GAMMA, ZETA = -0.1, 1.1


def round_ste(x):
    return (x.round() - x).detach() + x


def uaq_forward(x, delta, zero_point, n_levels, sym=False):
    x_int = round_ste(x / delta) + zero_point
    if sym:
        x_quant = torch.clamp(x_int, -n_levels // 2, n_levels // 2 - 1)
    else:
        x_quant = torch.clamp(x_int, 0, n_levels - 1)
    return (x_quant - zero_point) * delta


def soft_targets(alpha):
    return torch.clamp(torch.sigmoid(alpha) * (ZETA - GAMMA) + GAMMA, 0, 1)


def adaround_forward(w, alpha, delta, zero_point, n_levels, soft):
    x_floor = torch.floor(w / delta)
    x_int = x_floor + (soft_targets(alpha) if soft else (alpha >= 0).float())
    x_quant = torch.clamp(x_int + zero_point, 0, n_levels - 1)
    return (x_quant - zero_point) * delta


def init_alpha(w, delta):
    x_floor = torch.floor(w / delta)
    rest = (w / delta) - x_floor
    return -torch.log((ZETA - GAMMA) / (rest - GAMMA) - 1)


def lp_loss(pred, tgt, p=2.0):
    return (pred - tgt).abs().pow(p).sum(1).mean()


def temperature(t, t_max, rel_start_decay, start_b, end_b):
    start_decay = rel_start_decay * t_max
    if t < start_decay:
        return start_b
    rel_t = (t - start_decay) / (t_max - start_decay)
    return end_b + (start_b - end_b) * max(0.0, (1 - rel_t))


_ACT = {"relu": torch.relu, "relu6": F.relu6, None: (lambda t: t)}


def _layer(spec, x, wq, act_state=None, name=None):
    out = F.conv2d(x, wq, spec["bias"], **spec["conv"]) if spec["conv"] is not None else F.linear(x, wq, spec["bias"])
    out = out * spec["alpha_out"] + spec["beta_out"] if "alpha_out" in spec else out     # quant_layer.py:258-259
    out = _ACT[spec["act"]](out)
    if act_state is not None and name in act_state:                                      # quant_layer.py:269-271
        d, z, nl = act_state[name]
        out = uaq_forward(out, d, z, nl)
    return out


def unit_forward(unit, x, wqs, act_state=None):
    L, kind = unit["layers"], unit["kind"]
    f = lambda n, t: _layer(L[n], t, wqs[n], act_state, n)
    if kind == "layer":
        (n,) = L.keys()
        return f(n, x)
    if kind == "invres":
        out = x
        for n in L:
            out = f(n, out)
        out = x + out if unit["use_res_connect"] else out
    else:
        residual = f("downsample", x) if "downsample" in L else x
        out = f("conv1", x)
        out = f("conv2", out)
        if kind == "bottleneck":
            out = f("conv3", out)
        out = out + residual
    out = _ACT[unit.get("tail_act")](out)
    if act_state is not None and "__block__" in act_state:
        d, z, nl = act_state["__block__"]
        out = uaq_forward(out, d, z, nl)
    return out


def rec_loss(pred, tgt, grad=None, mode="mse", p=2.0):
    """LossFunction's reconstruction term (quant/block_recon.py:154-162)"""
    if mode == "mse":
        return lp_loss(pred, tgt, p=p)
    if mode == "fisher_diag":
        return ((pred - tgt).pow(2) * grad.pow(2)).sum(1).mean()
    if mode == "fisher_full":
        a = (pred - tgt).abs()
        g = grad.abs()
        batch_dotprod = torch.sum(a * g, (1, 2, 3)).view(-1, 1, 1, 1)
        return (batch_dotprod * a * g).mean() / 100
    raise ValueError(mode)


def recon_weight_loop(unit, cached_inps, cached_outs, idx_table, iters, weight=0.01, b_range=(20, 2), warmup=0.2,
                      p=2.0, alphas=None, start_count=0, t_max=None, state=None, opt_mode="mse", cached_grads=None,
                      train_affine=False):
    """the weight-rounding loop: returns (alphas, losses). `idx_table[i]` is the mini-batch of iteration i.
    t_max/start_count let a caller run a slice of a longer schedule (the CPU baseline times iterations from the
    middle of the 20 000-iteration schedule, where the regulariser is live and b is non-integer); `state` (a dict) keeps
    the optimizer between such slices, as the reference keeps ONE torch.optim.Adam for the whole loop (block_recon.py:60)."""
    L = unit["layers"]
    if alphas is None:
        alphas = {n: init_alpha(s["weight"].detach(), s["delta"].detach()).requires_grad_(True) for n, s in L.items()}
    opt = state.get("opt") if state is not None else None
    if opt is None:
        params = list(alphas.values())
        if train_affine:
            # README --bias_cal: the output-channel affine of every layer (alpha_out / beta_out, quant/quant_layer.py:231-238, applied
            # at :258-259 by _layer above) joins the same optimiser (upstream's commented lines, layer_recon_fused_shiftedScale.py:67-68)
            for s_ in L.values():
                s_["alpha_out"].requires_grad_(True); s_["beta_out"].requires_grad_(True)
                params += [s_["alpha_out"], s_["beta_out"]]
        opt = torch.optim.Adam(params)
        if state is not None:
            state["opt"] = opt
    t_max = iters if t_max is None else t_max
    loss_start = t_max * warmup
    losses, count = [], start_count
    for i in range(iters):
        idx = idx_table[i]
        cur_inp, cur_out = cached_inps[idx], cached_outs[idx]
        opt.zero_grad()
        wqs = {n: adaround_forward(s["weight"], alphas[n], s["delta"], s["zero_point"], s["n_levels"], True) for n, s in L.items()}
        out_quant = unit_forward(unit, cur_inp, wqs)
        count += 1
        rec = rec_loss(out_quant, cur_out, None if cached_grads is None else cached_grads[idx], opt_mode, p)
        b = temperature(count, t_max, warmup, b_range[0], b_range[1])
        if count < loss_start:
            b = rnd = 0
        else:
            rnd = 0
            for n in L:
                rnd += weight * (1 - ((soft_targets(alphas[n]) - .5).abs() * 2).pow(b)).sum()
        total = rec + rnd
        total.backward()
        opt.step()
        losses.append(float(total.detach()) if torch.is_tensor(total) else float(total))
    return alphas, losses


def recon_act_loop(unit, cached_inps, cached_outs, idx_table, iters, act_state, alphas, lr=4e-4, p=2.4):
    """the activation step-size loop (LSQ): act_state[name] = (delta 0-dim tensor requiring grad, zero_point, n_levels)"""
    L = unit["layers"]
    params = [v[0] for v in act_state.values()]
    opt = torch.optim.Adam(params, lr=lr)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=iters, eta_min=0.)
    losses = []
    with torch.no_grad():
        wqs = {n: adaround_forward(s["weight"], alphas[n], s["delta"], s["zero_point"], s["n_levels"], False) for n, s in L.items()}
    for i in range(iters):
        idx = idx_table[i]
        opt.zero_grad()
        out_quant = unit_forward(unit, cached_inps[idx], wqs, act_state)
        err = lp_loss(out_quant, cached_outs[idx], p=p)
        err.backward()
        opt.step()
        sch.step()
        losses.append(float(err.detach()))
    return losses


# ---- synthetic ResNet-18 units for the CPU baseline (shapes of SURVEY.md App. C) --------------------------------
RESNET18_UNITS = [  # name, cin, cout, stride, spatial_in (224 input)
    ("layer1.0", 64, 64, 1, 56), ("layer1.1", 64, 64, 1, 56), ("layer2.0", 64, 128, 2, 56), ("layer2.1", 128, 128, 1, 28),
    ("layer3.0", 128, 256, 2, 28), ("layer3.1", 256, 256, 1, 14), ("layer4.0", 256, 512, 2, 14), ("layer4.1", 512, 512, 1, 7),
]


def _max_init(w, n_levels):
    """per-channel 'max' scale for the synthetic baseline units (cost of the loop does not depend on delta)"""
    flat = w.reshape(w.shape[0], -1)
    mn = flat.min(1)[0].clamp(max=0); mx = flat.max(1)[0].clamp(min=0)
    delta = ((mx - mn) / (n_levels - 1)).clamp(min=1e-8)
    zp = (-mn / delta).round()
    shape = (-1,) + (1,) * (w.dim() - 1)
    return delta.view(shape), zp.view(shape)


def _as_reference_parameters(spec):
    """upstream keeps weight, bias, delta, zero_point, alpha_out, beta_out as nn.Parameters on the graph
    (quant_layer.py:87-88,203,233-238), so autograd also produces their (unused) gradients every iteration"""
    for k in ("weight", "bias", "delta", "zero_point", "alpha_out", "beta_out"):
        if k in spec and spec[k] is not None:
            spec[k].requires_grad_(True)
    return spec


def synthetic_resnet18_unit(name, cin, cout, stride, n_bits=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    nl = 2 ** n_bits

    def conv(ci, co, k, s):
        w = torch.randn(co, ci, k, k, generator=g) * (2.0 / (k * k * co)) ** 0.5
        d, z = _max_init(w, nl)
        return _as_reference_parameters(dict(
            weight=w, bias=torch.zeros(co), conv=dict(stride=s, padding=k // 2, dilation=1, groups=1), act=None,
            delta=d, zero_point=z, n_levels=nl, alpha_out=torch.ones(1, co, 1, 1), beta_out=torch.zeros(1, co, 1, 1)))

    layers = {"conv1": conv(cin, cout, 3, stride), "conv2": conv(cout, cout, 3, 1)}
    layers["conv1"]["act"] = "relu"
    if stride != 1 or cin != cout:
        layers["downsample"] = conv(cin, cout, 1, stride)
    return {"kind": "basic", "layers": layers, "tail_act": "relu"}


def synthetic_fc_unit(cin=512, cout=1000, n_bits=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(cout, cin, generator=g) * 0.01
    d, z = _max_init(w, 2 ** n_bits)
    return {"kind": "layer", "layers": {"fc": _as_reference_parameters(dict(
        weight=w, bias=torch.zeros(cout), conv=None, act=None, delta=d, zero_point=z, n_levels=2 ** n_bits,
        alpha_out=torch.ones(1, cout), beta_out=torch.zeros(1, cout)))}}


def unit_to(unit, device):
    """the same functional unit with its tensors on `device` (leaf-ness and requires_grad preserved): the GPU-reference leg of
    bench.py runs this very op chain with ATen's CUDA kernels, as the reference does on a GPU"""
    def mv(v):
        if torch.is_tensor(v):
            t = v.detach().to(device)
            return t.requires_grad_(True) if v.requires_grad else t
        return v
    out = {k: v for k, v in unit.items() if k != "layers"}
    out["layers"] = {n: {k: mv(v) for k, v in s.items()} for n, s in unit["layers"].items()}
    return out


def fp_unit_outputs(unit, x):
    """FP output of the unit (no output affine: quant_layer.py:258 applies it only when weight quant is on)"""
    with torch.no_grad():
        fp = {"kind": unit["kind"], "tail_act": unit.get("tail_act"), "use_res_connect": unit.get("use_res_connect"),
              "layers": {n: {k: v for k, v in s.items() if k not in ("alpha_out", "beta_out")} for n, s in unit["layers"].items()}}
        return unit_forward(fp, x, {n: s["weight"] for n, s in unit["layers"].items()})
