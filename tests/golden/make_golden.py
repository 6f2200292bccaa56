"""Generates tests/golden/*.npz by running the REAL reference (/root/reference, read-only) on seeded
inputs on the CPU. Run in the build container only:  python tests/golden/make_golden.py
The vectors are committed; nothing at test/bench time reads /root/reference.

Shims needed to import the reference: a stub `icecream` and a stub
`pretrained.PyTorch_CIFAR10.cifar10_models.resnet.BasicBlockCIFAR`.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("SSQ_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    ic = types.ModuleType("icecream")

    class _IC:
        def __call__(self, *a, **k):
            return a[0] if a else None

        def configureOutput(self, **k):
            pass

        def disable(self):
            pass

    ic.ic = _IC()
    sys.modules["icecream"] = ic
    for name in ["pretrained", "pretrained.PyTorch_CIFAR10", "pretrained.PyTorch_CIFAR10.cifar10_models",
                 "pretrained.PyTorch_CIFAR10.cifar10_models.resnet"]:
        sys.modules[name] = types.ModuleType(name)

    class BasicBlockCIFAR(nn.Module):
        pass

    sys.modules["pretrained.PyTorch_CIFAR10.cifar10_models.resnet"].BasicBlockCIFAR = BasicBlockCIFAR
    import quant  # noqa: F401


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def gen_weights(shape, seed, scale=0.05, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale + offset


def main():
    import_reference()
    from quant.quant_layer import UniformAffineQuantizer, lp_loss
    from quant.adaptive_rounding import AdaRoundQuantizer
    from quant.channelQuant import ChannelQuant
    from quant.channelQuantMSE import ChannelQuantMSE
    from quant.block_recon import LinearTempDecay, LossFunction

    torch.manual_seed(1005)

    # ---------------------------------------------------------------- UAQ: mse/max init + forward + backward
    cases = {}
    specs = [
        ("w_conv_b2", (8, 4, 3, 3), 2, False, True, "mse"),
        ("w_conv_b4", (8, 4, 3, 3), 4, False, True, "mse"),
        ("w_dw_b3", (12, 1, 3, 3), 3, False, True, "mse"),       # depthwise rows of 9 (ragged, non-vector path)
        ("w_fc_b8", (10, 16), 8, False, True, "mse"),
        ("w_conv_b4_sym", (8, 4, 3, 3), 4, True, True, "mse"),
        ("w_conv_b4_max", (8, 4, 3, 3), 4, False, True, "max"),
        ("a_tensor_b4", (4, 6, 5, 5), 4, False, False, "mse"),    # activation, per-tensor
        # NOTE per-tensor 'max' (and per-tensor symmetric 'mse') cannot run in the reference: zero_point is a
        # python int there and nn.Parameter(int) raises (quant_layer.py:88,139) — no golden vector exists.
    ]
    for i, (name, shape, bits, sym, cw, method) in enumerate(specs):
        x = gen_weights(shape, 100 + i, 0.05 if cw else 1.0)
        if not cw:
            x = torch.relu(x)
        q = UniformAffineQuantizer(n_bits=bits, symmetric=sym, channel_wise=cw, scale_method=method)
        xr = x.clone().requires_grad_(True)
        y = q(xr)
        gy = gen_weights(shape, 200 + i, 1.0)
        y.backward(gy)
        # codes as the reference computes them
        with torch.no_grad():
            x_int = torch.round(x / q.delta) + q.zero_point
            lo, hi = (-q.n_levels // 2, q.n_levels // 2 - 1) if sym else (0, q.n_levels - 1)
            codes = torch.clamp(x_int, lo, hi)
        raw = q.raw_zero_point
        raw = npy(raw) if torch.is_tensor(raw) else np.asarray(raw, dtype=np.float32)
        cases[name] = dict(x=npy(x), delta=npy(q.delta), zp=npy(q.zero_point), raw=raw, y=npy(y), codes=npy(codes),
                           gy=npy(gy), gx=npy(xr.grad), gdelta=npy(q.delta.grad), gzp=npy(q.zero_point.grad),
                           meta=np.array([bits, int(sym), int(cw), int(method == "mse")]))
    save("uaq", **{f"{k}.{f}": v for k, d in cases.items() for f, v in d.items()})

    # ---------------------------------------------------------------- AdaRound
    ada = {}
    for i, (name, shape, bits) in enumerate([("conv_b2", (8, 4, 3, 3), 2), ("dw_b4", (12, 1, 3, 3), 4), ("fc_b8", (10, 16), 8)]):
        w = gen_weights(shape, 300 + i)
        uaq = UniformAffineQuantizer(n_bits=bits, channel_wise=True, scale_method="mse")
        uaq(w)
        aq = AdaRoundQuantizer(uaq=uaq, round_mode="learned_hard_sigmoid", weight_tensor=w)
        alpha0 = npy(aq.alpha)
        # move alpha away from init so soft targets exercise the clamp on both sides
        with torch.no_grad():
            aq.alpha.add_(gen_weights(shape, 310 + i, 3.0))
        aq.soft_targets = True
        wq_soft = aq(w)
        gw = gen_weights(shape, 320 + i, 1.0)
        wq_soft.backward(gw)
        galpha = npy(aq.alpha.grad)
        aq.alpha.grad = None
        aq.soft_targets = False
        wq_hard = aq(w)
        with torch.no_grad():
            codes_hard = torch.clamp(torch.floor(w / aq.delta) + (aq.alpha >= 0).float() + aq.zero_point, 0, aq.n_levels - 1)
        # regulariser through the reference's LossFunction expression (block_recon.py:173-174)
        regs = {}
        for b in (20, 11.3, 2.0):
            aq.alpha.grad = None
            rv = aq.get_soft_targets()
            r = 0.01 * (1 - ((rv - .5).abs() * 2).pow(b)).sum()
            r.backward()
            regs[f"reg_b{b}"] = npy(r)
            regs[f"greg_b{b}"] = npy(aq.alpha.grad)
        ada[name] = dict(w=npy(w), delta=npy(aq.delta), zp=npy(aq.zero_point), alpha0=alpha0, alpha=npy(aq.alpha),
                         wq_soft=npy(wq_soft), gw=npy(gw), galpha=galpha, wq_hard=npy(wq_hard), codes_hard=npy(codes_hard),
                         h=npy(aq.get_soft_targets()), bits=np.array(bits), **regs)
    save("adaround", **{f"{k}.{f}": v for k, d in ada.items() for f, v in d.items()})

    # ---------------------------------------------------------------- losses + temperature schedule
    loss = {}
    for i, (name, shape) in enumerate([("conv", (4, 6, 5, 5)), ("fc", (8, 10))]):
        pred = gen_weights(shape, 400 + i, 1.0).requires_grad_(True)
        tgt = gen_weights(shape, 410 + i, 1.0)
        for p in (2.0, 2.4):
            pred.grad = None
            l = lp_loss(pred, tgt, p=p)
            l.backward()
            loss[f"{name}.lp{p}"] = npy(l); loss[f"{name}.dlp{p}"] = npy(pred.grad)
        loss[f"{name}.pred"] = npy(pred); loss[f"{name}.tgt"] = npy(tgt)
        if len(shape) == 4:
            g = gen_weights(shape, 420 + i, 1.0).abs() + 1.0
            pred.grad = None
            l = ((pred - tgt).pow(2) * g.pow(2)).sum(1).mean(); l.backward()
            loss[f"{name}.fdiag"] = npy(l); loss[f"{name}.dfdiag"] = npy(pred.grad)
            pred.grad = None
            a = (pred - tgt).abs(); ga = g.abs()
            bd = torch.sum(a * ga, (1, 2, 3)).view(-1, 1, 1, 1)
            l = (bd * a * ga).mean() / 100; l.backward()
            loss[f"{name}.ffull"] = npy(l); loss[f"{name}.dffull"] = npy(pred.grad)
            loss[f"{name}.fisher"] = npy(g)
    td = LinearTempDecay(200, rel_start_decay=0.2, start_b=20, end_b=2)
    loss["temp.t"] = np.arange(0, 202)
    loss["temp.b"] = np.array([float(td(t)) for t in range(0, 202)], dtype=np.float64)
    save("loss", **loss)

    # ---------------------------------------------------------------- ChannelQuant (shifted scale)
    cq = {}
    shifts = [0.96875, 1.03125, 1.0]
    for i, (name, shape, bits) in enumerate([("conv_b2", (8, 6, 3, 3), 2), ("fc_b4", (6, 10), 4), ("dw_b3", (8, 1, 3, 3), 3)]):
        w = gen_weights(shape, 500 + i)
        uaq = UniformAffineQuantizer(n_bits=bits, channel_wise=True, scale_method="mse")
        uaq(w)
        # -- shift-only path: init_v -> 'learned_hard_sigmoid'
        q = ChannelQuant(1.0, uaq, w, shiftTarget=list(shifts))
        y_none = q(w)                                   # 'none' mode
        q.device = "cpu"
        q.init_v(w.clone())
        alpha_init = npy(q.alpha)
        with torch.no_grad():
            q.alpha.add_(gen_weights(tuple(q.alpha.shape), 510 + i, 1.5))
        y_soft = q(w)
        gy = gen_weights(shape, 520 + i, 1.0)
        y_soft.backward(gy)
        galpha_soft = npy(q.alpha.grad)
        q.alpha.grad = None
        ent = 0.7 * (-torch.sum(q.get_sig_soft_targets() * torch.log(q.get_sig_soft_targets() + 1e-10)))
        ent.backward()
        gent = npy(q.alpha.grad)
        q.hard_targets = True
        y_hard = q(w)
        q.hard_targets = False
        d = dict(w=npy(w), delta=npy(q.delta), zp=npy(q.zero_point), shifts=np.array(shifts), y_none=npy(y_none),
                 alpha_init=alpha_init, alpha=npy(q.alpha), p=npy(q.get_sig_soft_targets()), y_soft=npy(y_soft), gy=npy(gy),
                 galpha_soft=galpha_soft, ent=npy(ent), gent=gent, y_hard=npy(y_hard), bits=np.array(bits),
                 xq=np.stack([npy(t) for t in q.x_q], -1))
        # -- adaround mode after shift (update_delta + init_beta)
        q.update_delta(); q.init_beta(w.clone()); q.opt_mode = "adaround"
        with torch.no_grad():
            q.beta.add_(gen_weights(shape, 530 + i, 2.0))
        y_ar = q(w); y_ar.backward(gy)
        d.update(ar_delta=npy(q.delta), ar_beta=npy(q.beta), ar_y=npy(y_ar), ar_gbeta=npy(q.beta.grad))
        q.hard_round = True
        d["ar_y_hard"] = npy(q(w))
        # -- fused path: init_v_beta -> 'adaShift'
        q2 = ChannelQuant(1.0, uaq, w, shiftTarget=list(shifts))
        q2.device = "cpu"
        q2.init_v_beta(w.clone())
        q2.opt_mode = "adaShift"
        d["as_alpha_init"] = npy(q2.alpha); d["as_beta_init"] = npy(q2.beta)
        with torch.no_grad():
            q2.alpha.add_(gen_weights(tuple(q2.alpha.shape), 540 + i, 1.5))
            q2.beta.add_(gen_weights(shape, 550 + i, 2.0))
        y_as = q2(w); y_as.backward(gy)
        d.update(as_alpha=npy(q2.alpha), as_beta=npy(q2.beta), as_y=npy(y_as), as_galpha=npy(q2.alpha.grad),
                 as_gbeta=npy(q2.beta.grad), as_xq=np.stack([npy(t) for t in q2.x_q], -1))
        q2.alpha.grad = None
        for b2 in (20, 7.7):
            rv = q2.get_sig_soft_targets()
            r = 0.3 * (1 - ((rv - .5).abs() * 2).pow(b2)).sum(); r.backward()
            d[f"as_regS_b{b2}"] = npy(r); d[f"as_gregS_b{b2}"] = npy(q2.alpha.grad); q2.alpha.grad = None
        q2.hard_round = True; q2.hard_targets = True
        d["as_y_hard"] = npy(q2(w))
        cq[name] = d
    save("channelquant", **{f"{k}.{f}": v for k, d in cq.items() for f, v in d.items()})

    # ---------------------------------------------------------------- ChannelQuantMSE
    cm = {}
    for i, (name, shape, bits) in enumerate([("conv_b2", (16, 6, 3, 3), 2), ("conv_b4", (8, 4, 3, 3), 4)]):
        w = gen_weights(shape, 600 + i)
        uaq = UniformAffineQuantizer(n_bits=bits, channel_wise=True, scale_method="max")
        uaq(w)
        d = dict(w=npy(w), delta=npy(uaq.delta), raw=npy(uaq.raw_zero_point), bits=np.array(bits))
        for level, thr in [(1, 1.0), (4, 1.5), (16, 1.5), (64, 2.0)]:
            q = ChannelQuantMSE(1.0, uaq, w, level=level, threshold=thr, opt_mode="max")
            q.init_scale(w)
            d[f"inp_scale_l{level}"] = npy(q.inp_scale); d[f"y_l{level}"] = npy(q(w)); d[f"codes_l{level}"] = npy(q.quant(w))
            d[f"thr_l{level}"] = np.array(thr)
        cm[name] = d
    save("channelquantmse", **{f"{k}.{f}": v for k, d in cm.items() for f, v in d.items()})

    # ---------------------------------------------------------------- Adam trajectory (torch.optim.Adam defaults)
    p = nn.Parameter(gen_weights((37,), 700, 1.0))
    opt = torch.optim.Adam([p])
    traj = {"p0": npy(p)}
    for s in range(1, 6):
        g = gen_weights((37,), 700 + s, 0.3)
        p.grad = g.clone(); opt.step()
        traj[f"g{s}"] = npy(g); traj[f"p{s}"] = npy(p)
    sched_p = nn.Parameter(torch.zeros(1))
    o2 = torch.optim.Adam([sched_p], lr=4e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(o2, T_max=50, eta_min=0.)
    lrs = []
    for s in range(50):
        lrs.append(o2.param_groups[0]["lr"]); o2.step(); sch.step()
    traj["cosine_lr"] = np.array(lrs, dtype=np.float64)
    save("adam", **traj)

    # ---------------------------------------------------------------- reconstruction loops (tiny ResNet-18, CPU)
    from quant import QuantModel, QuantModule, layer_reconstruction
    from quant.data_utils import save_inp_oup_data
    from models.resnet import resnet18 as ref_resnet18
    torch.manual_seed(1005)
    cnn = ref_resnet18(num_classes=10).eval()
    wq = {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}     # 'max': the mse init of 5.8k channels takes ~50 s here
    aq = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
    qnn = QuantModel(model=cnn, weight_quant_params=wq, act_quant_params=aq).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(32, 3, 16, 16)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali)
    # the network itself is not stored: zoo.resnet18(num_classes=10) under torch.manual_seed(1005) is bit-identical
    # to the reference constructor (checked in tests/test_host_cpu.py against the probe weights below)
    rl = {"cali": npy(cali), "probe.conv1_w": npy(qnn.model.conv1.org_weight[:4]),
          "probe.l4_w": npy(qnn.model.layer4[1].conv2.org_weight[:2])}
    # -- block loop: the body of quant/block_recon.py:89-105 driven here because :88-93 hard-code 'cuda'
    block = qnn.model.layer1[0]
    iters, bs = 12, 16
    qnn.set_quant_state(False, False); block.set_quant_state(True, False)
    mods = [m for _n, m in block.named_modules() if isinstance(m, QuantModule)]
    for m in mods:
        m.weight_quantizer = AdaRoundQuantizer(uaq=m.weight_quantizer, round_mode='learned_hard_sigmoid', weight_tensor=m.org_weight.data)
        m.weight_quantizer.soft_targets = True
    opt = torch.optim.Adam([m.weight_quantizer.alpha for m in mods])
    lf = LossFunction(block, round_loss='relaxation', weight=0.01, max_count=iters, rec_loss='mse', b_range=(20, 2),
                      decay_start=0, warmup=0.2, p=2.0)
    inps, outs = save_inp_oup_data(qnn, block, cali, True, False, bs)
    torch.manual_seed(77)
    idx_tab, losses = [], []
    for i in range(iters):
        idx = torch.randperm(inps.size(0))[:bs]
        idx_tab.append(npy(idx))
        opt.zero_grad()
        err = lf(block(inps[idx]), outs[idx])
        err.backward(retain_graph=True)
        opt.step()
        losses.append(float(err))
    rl.update({"block.inps": npy(inps), "block.outs": npy(outs), "block.idx": np.stack(idx_tab), "block.losses": np.array(losses)})
    for n, m in (("conv1", block.conv1), ("conv2", block.conv2)):
        rl.update({f"block.{n}.weight": npy(m.org_weight), f"block.{n}.bias": npy(m.org_bias), f"block.{n}.delta": npy(m.weight_quantizer.delta),
                   f"block.{n}.zp": npy(m.weight_quantizer.zero_point), f"block.{n}.alpha": npy(m.weight_quantizer.alpha)})
    for m in mods:
        m.weight_quantizer.soft_targets = False
    # -- layer loop: the real layer_reconstruction (runs on CPU unmodified); the RNG stream starts at seed 78
    torch.manual_seed(78)
    layer_reconstruction(qnn, qnn.model.fc, cali_data=cali, iters=iters, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                         act_quant=False, opt_mode='mse', batch_size=bs)
    fc = qnn.model.fc
    rl.update({"fc.weight": npy(fc.org_weight), "fc.bias": npy(fc.org_bias), "fc.delta": npy(fc.weight_quantizer.delta),
               "fc.zp": npy(fc.weight_quantizer.zero_point), "fc.alpha": npy(fc.weight_quantizer.alpha)})
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        rl["final_logits"] = npy(qnn(cali[:8]))
    save("recon_loop", **rl)
    # checkpoint compatibility: parameter/buffer names and shapes of the reference's state_dict at this point
    import json
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in qnn.state_dict().items()}, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
