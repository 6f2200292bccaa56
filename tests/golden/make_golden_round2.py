"""Generates the round-2 golden vectors by running the REAL reference (/root/reference, read-only) on the CPU:

  layer_shift.npz   quant/layer_recon_shiftedScale.py:262-338 `layer_recon_shiftedScale` on a ResNet-50 conv2 layer
                    (BASELINE configs[2]): shift first, AdaRound on top, and the act=True (regulariser off) flavour.
                    The function hard-codes torch.device('cuda') at :268; its source is loaded with that one
                    expression replaced by the model's device — nothing else is touched.
  families.npz      the real `block_reconstruction` (quant/block_recon.py:10-116, `device = 'cuda'` at :88 replaced the
                    same way) for 16 iterations on a ResNet-50 bottleneck with downsample and on a RegNetX-3200M block.
  long_horizon.npz  the real `block_reconstruction` on ResNet-18 layer1.0 and the real, unmodified `layer_reconstruction`
                    on fc for 2 000 and 20 000 iterations (the north_star horizon): final AdaRound alphas and hard codes.

  channelquantact.npz  quant/channelQuantAct.py:45-54, the 'adaround' / 'none' forward modes + autograd of beta.

Run in the build container only:  python tests/golden/make_golden_round2.py [layer_shift] [families] [long_horizon] [channelquantact]
The vectors are committed; nothing at test/bench time reads /root/reference.
"""
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, import_reference, npy, save  # noqa: E402

SHIFTS = [0.96875, 1.03125, 1.0]      # ShiftedScaleQuant.py:388


def load_patched(modname, relpath, old, new):
    """the reference module at `relpath`, executed from its own source with ONE device expression replaced"""
    src = open(os.path.join(REF, relpath)).read()
    assert src.count(old) == 1, (relpath, old, src.count(old))
    mod = types.ModuleType(modname)
    mod.__file__ = os.path.join(REF, relpath)
    exec(compile(src.replace(old, new), mod.__file__, "exec"), mod.__dict__)
    return mod


def add_path_name_shim():
    """QuantBottleneck / QuantResBottleneckBlock / QuantInvertedResidual lack setPathName upstream"""
    from quant.quant_block import BaseQuantBlock, QuantBasicBlock
    from quant import QuantModule
    if "setPathName" not in BaseQuantBlock.__dict__:
        def set_path_name(self, name):
            self.pathName = name
            for n, m in self.named_modules():
                if isinstance(m, QuantModule):
                    m.pathName = name + '.' + n
        BaseQuantBlock.setPathName = set_path_name
    return QuantBasicBlock


def build(arch, bits, res, n, seed=1005, num_classes=10, scale='max'):
    from quant import QuantModel
    import models.resnet as R
    import models.regnet as G
    torch.manual_seed(seed)
    ctor = getattr(R, arch, None)
    cnn = (ctor(num_classes=num_classes) if ctor is not None else getattr(G, arch)()).eval()   # RegNet: 1000 classes, fixed
    wq = {'n_bits': bits, 'channel_wise': True, 'scale_method': scale}
    aq = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
    qnn = QuantModel(model=cnn, weight_quant_params=wq, act_quant_params=aq).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(n, 3, res, res)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:32])
    return qnn, cali


def swap_and_cache_layer(qnn, layer, cali, bs=16):
    """ChannelQuant swap + the 'if'/'of' cache protocol for ONE QuantModule (ShiftedScaleQuant.py:244-255, 384-392)"""
    from quant.channelQuant import ChannelQuant
    layer.weight_quantizer = ChannelQuant(1.0, uaq=layer.weight_quantizer, weight_tensor=layer.org_weight.data,
                                          shiftTarget=list(SHIFTS), name=layer.pathName)
    for mode, wq_on in (('if', True), ('of', False)):
        qnn.set_quant_state(wq_on, False)
        layer.cache_features = mode
        with torch.no_grad():
            for i in range(0, cali.shape[0], bs):
                qnn(cali[i:i + bs])
        layer.cache_features = 'none'
    qnn.set_quant_state(False, False)
    layer.set_quant_state(True, False)


def gen_layer_shift():
    add_path_name_shim()
    LS = load_patched("ref_layer_recon_shiftedScale_cpu", "quant/layer_recon_shiftedScale.py",
                      "device = torch.device('cuda')", "device = next(model.parameters()).device")
    iters = 24
    out = {"shifts": np.array(SHIFTS), "iters": np.array(iters)}
    for tag, act in (("A", False), ("B", True)):
        qnn, cali = build("resnet50", 4, 16, 48)
        layer = qnn.model.layer1[0].conv2                 # 3x3, 64 -> 64: alpha [IC, S]
        swap_and_cache_layer(qnn, layer, cali)
        if tag == "A":
            out["cali"] = npy(cali)
            out["probe.conv2_w"] = npy(layer.org_weight[:4])
            out["A.inp"] = npy(torch.cat(layer.cached_inp_features)); out["A.out"] = npy(torch.cat(layer.cached_out_features))
        torch.manual_seed(191)
        soft, hard = LS.layer_recon_shiftedScale(layer, iters=iters, lmda=0.01, model=qnn, act=act)
        out[f"{tag}.shift.losses"] = np.array([soft, hard], dtype=np.float64)
        out[f"{tag}.shift.alpha"] = npy(layer.weight_quantizer.alpha)
        with torch.no_grad():
            out[f"{tag}.shift.hard_out"] = npy(layer(torch.cat(layer.cached_inp_features)[:8]))
        if tag == "A":
            torch.manual_seed(192)
            soft, hard = LS.layer_recon_shiftedScale(layer, iters=iters, lmda=0.01, model=qnn, adaround=True)
            out["A.ada.losses"] = np.array([soft, hard], dtype=np.float64)
            out["A.ada.beta"] = npy(layer.weight_quantizer.beta)
            out["A.ada.delta"] = npy(layer.weight_quantizer.delta)
            out["A.ada.layer_hard_round"] = np.array(bool(getattr(layer, "hard_round", False)))
            out["A.ada.quantizer_hard_round"] = np.array(bool(layer.weight_quantizer.hard_round))
            with torch.no_grad():
                out["A.ada.out_after"] = npy(layer(torch.cat(layer.cached_inp_features)[:8]))
    save("layer_shift", **out)


def _patched_block_recon():
    return load_patched("ref_block_recon_cpu", "quant/block_recon.py", "device = 'cuda'",
                        "device = next(model.parameters()).device")


def gen_families():
    add_path_name_shim()
    BR = _patched_block_recon()
    from quant import QuantModule
    out = {}
    for tag, arch, bits, pick in (("r50", "resnet50", 4, lambda q: q.model.layer1[0]),
                                  ("rx32", "regnetx_3200m", 2, lambda q: q.model.s2.b1)):
        qnn, cali = build(arch, bits, 32, 32)
        block = pick(qnn)
        iters, bs = 16, 16
        torch.manual_seed(277)
        BR.block_reconstruction(qnn, block, cali_data=cali, iters=iters, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                                act_quant=False, opt_mode='mse', batch_size=bs)
        out[f"{tag}.cali"] = npy(cali)
        for n, m in block.named_modules():
            if isinstance(m, QuantModule):
                out[f"{tag}.{n}.probe_w"] = npy(m.org_weight.reshape(-1)[:16])
                out[f"{tag}.{n}.delta"] = npy(m.weight_quantizer.delta)
                out[f"{tag}.{n}.zp"] = npy(m.weight_quantizer.zero_point)
                out[f"{tag}.{n}.alpha"] = npy(m.weight_quantizer.alpha)
        qnn.set_quant_state(False, False); block.set_quant_state(True, False)
        from quant.data_utils import save_inp_oup_data
        inps, _outs = save_inp_oup_data(qnn, block, cali[:16], True, False, bs)
        qnn.set_quant_state(False, False); block.set_quant_state(True, False)
        with torch.no_grad():
            out[f"{tag}.hard_out"] = npy(block(inps[:8]))
    save("families", **out)


def hard_codes(m):
    """integer codes of the hard forward (quant/adaptive_rounding.py:50-59): floor(w/delta) + (alpha >= 0) + zp, clamped"""
    q = m.weight_quantizer
    w = m.org_weight.data
    x = torch.floor(w / q.delta) + (q.alpha >= 0).float()
    return torch.clamp(x + q.zero_point, 0, q.n_levels - 1).to(torch.uint8)


def _long_horizon_run(BR, layer_reconstruction, iters):
    qnn, cali = build("resnet18", 2, 32, 64)
    kw = dict(cali_data=cali, iters=iters, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=False,
              opt_mode='mse', batch_size=32)
    block = qnn.model.layer1[0]
    t0 = time.time()
    torch.manual_seed(377)
    BR.block_reconstruction(qnn, block, **kw)
    t_block = time.time() - t0
    t0 = time.time()
    torch.manual_seed(378)
    layer_reconstruction(qnn, qnn.model.fc, **kw)
    t_fc = time.time() - t0
    return qnn, cali, {"block.conv1": block.conv1, "block.conv2": block.conv2, "fc": qnn.model.fc}, (t_block, t_fc)


def gen_long_horizon():
    BR = _patched_block_recon()
    from quant import layer_reconstruction
    out = {}
    for iters in (2000, 20000):
        qnn, cali, mods, secs = _long_horizon_run(BR, layer_reconstruction, iters)
        if "cali_probe" not in out:
            out["cali_probe"] = npy(cali.reshape(-1)[:64])
            out["probe.conv1_w"] = npy(qnn.model.conv1.org_weight[:4])
        for n, m in mods.items():
            out[f"i{iters}.{n}.alpha16"] = npy(m.weight_quantizer.alpha).astype(np.float16)
            out[f"i{iters}.{n}.codes"] = npy(hard_codes(m))
        out[f"i{iters}.cpu_seconds"] = np.array(secs)
        print(f"iters={iters}: block {secs[0]:.1f}s fc {secs[1]:.1f}s", flush=True)
        # the reference against ITSELF with another convolution backend (oneDNN off -> ATen's native CPU kernels): the same
        # last-bit differences a cuDNN run has, i.e. the noise floor of any code-agreement figure at this horizon
        with torch.backends.mkldnn.flags(enabled=False):
            _q2, _c2, mods2, _ = _long_horizon_run(BR, layer_reconstruction, iters)
        for n, m in mods2.items():
            same = npy(hard_codes(m)) == out[f"i{iters}.{n}.codes"]
            out[f"i{iters}.{n}.self_agreement"] = np.array(float(same.mean()))
            print(f"  reference vs reference (native conv backend) {n}: {same.mean():.6f}", flush=True)
    save("long_horizon", **out)


def gen_channelquantact():
    """quant/channelQuantAct.py:45-54: the 'adaround' forward of the activation twin (hard rounding only: the soft branch calls
    an undefined get_soft_round upstream; the caller supplies beta and hard_round, the class never initialises them)"""
    from quant.quant_layer import UniformAffineQuantizer
    from quant.channelQuantAct import ChannelQuantAct
    g = torch.Generator().manual_seed(4242)
    x = torch.relu(torch.randn(4, 8, 6, 6, generator=g)) * 1.7
    uaq = UniformAffineQuantizer(n_bits=4, channel_wise=False, scale_method='mse', leaf_param=True)
    uaq(x)
    q = ChannelQuantAct(uaq)
    q.opt_mode = 'adaround'
    q.beta = torch.nn.Parameter(torch.randn(x.shape, generator=g) * 2.0)
    out = {"x": npy(x), "delta": npy(uaq.delta), "zp": npy(uaq.zero_point), "beta": npy(q.beta)}
    q.hard_round = False
    try:                                   # upstream: the soft branch calls an undefined get_soft_round (:48)
        q(x)
        out["soft_raises"] = np.array(False)
    except AttributeError:
        out["soft_raises"] = np.array(True)
    q.hard_round = True
    with torch.no_grad():
        out["y_hard"] = npy(q(x))
    q.opt_mode = 'none'
    with torch.no_grad():
        out["y_none"] = npy(q(x))
    save("channelquantact", **out)


if __name__ == "__main__":
    import_reference()
    todo = sys.argv[1:] or ["layer_shift", "families", "long_horizon", "channelquantact"]
    for name in todo:
        {"layer_shift": gen_layer_shift, "families": gen_families, "long_horizon": gen_long_horizon,
         "channelquantact": gen_channelquantact}[name]()
