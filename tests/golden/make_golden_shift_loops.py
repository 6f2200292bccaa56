"""Generates tests/golden/shift_loops.npz by running the REAL reference's shifted-scale reconstruction loops
(/root/reference, read-only; quant/layer_recon_shiftedScale.py:12-124, quant/layer_recon_fused_shiftedScale.py:23-141)
on the CPU with seeded inputs. Run in the build container only:  python tests/golden/make_golden_shift_loops.py
The vectors are committed; nothing at test/bench time reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, npy, save  # noqa: E402

SHIFTS = [0.96875, 1.03125, 1.0]      # ShiftedScaleQuant.py:388
ITERS = 20


def build(seed=1005):
    from quant import QuantModel
    from models.resnet import resnet18 as ref_resnet18
    torch.manual_seed(seed)
    cnn = ref_resnet18(num_classes=10).eval()
    wq = {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}
    aq = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
    qnn = QuantModel(model=cnn, weight_quant_params=wq, act_quant_params=aq).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(48, 3, 16, 16)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:32])
    return qnn, cali


def swap_and_cache(qnn, block, cali, bs=16):
    """ChannelQuant swap + the 'if'/'of' cache protocol (ShiftedScaleQuant.py:244-255, 384-392)"""
    from quant import QuantModule
    from quant.channelQuant import ChannelQuant
    for _n, m in block.named_modules():
        if isinstance(m, QuantModule):
            m.weight_quantizer = ChannelQuant(1.0, uaq=m.weight_quantizer, weight_tensor=m.org_weight.data,
                                              shiftTarget=list(SHIFTS), name=m.pathName)
    qnn.set_quant_state(True, False)
    block.cache_features = 'if'
    with torch.no_grad():
        for i in range(0, cali.shape[0], bs):
            qnn(cali[i:i + bs])
    block.cache_features = 'none'
    qnn.set_quant_state(False, False)
    block.cache_features = 'of'
    with torch.no_grad():
        for i in range(0, cali.shape[0], bs):
            qnn(cali[i:i + bs])
    block.cache_features = 'none'
    block.set_quant_state(True, False)


def mods_of(block):
    from quant import QuantModule
    return [(n, m) for n, m in block.named_modules() if isinstance(m, QuantModule)]


def main():
    import_reference()
    from quant.layer_recon_shiftedScale import block_recon_shiftedScale
    from quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale
    out = {"shifts": np.array(SHIFTS), "iters": np.array(ITERS)}
    # ---- A: shift first (entropy regulariser), then AdaRound on top (pow regulariser)
    qnn, cali = build()
    block = qnn.model.layer2[0]                    # conv1, conv2 and a 1x1 downsample: three QuantModules
    swap_and_cache(qnn, block, cali)
    out["cali"] = npy(cali)
    out["A.inp"] = npy(torch.cat(block.cached_inp_features)); out["A.out"] = npy(torch.cat(block.cached_out_features))
    torch.manual_seed(91)
    soft, hard = block_recon_shiftedScale(block, iters=ITERS, lmda=0.01, model=qnn)
    out["A.shift.losses"] = np.array([soft, hard], dtype=np.float64)
    for n, m in mods_of(block):
        out[f"A.shift.{n}.alpha"] = npy(m.weight_quantizer.alpha)
    with torch.no_grad():
        out["A.shift.hard_out"] = npy(block(torch.cat(block.cached_inp_features)[:8]))
    torch.manual_seed(92)
    soft, hard = block_recon_shiftedScale(block, iters=ITERS, lmda=0.01, model=qnn, adaround=True)
    out["A.ada.losses"] = np.array([soft, hard], dtype=np.float64)
    for n, m in mods_of(block):
        out[f"A.ada.{n}.beta"] = npy(m.weight_quantizer.beta)
        out[f"A.ada.{n}.delta"] = npy(m.weight_quantizer.delta)
    with torch.no_grad():
        out["A.ada.hard_out"] = npy(block(torch.cat(block.cached_inp_features)[:8]))
    # ---- B: fused shift + rounding (adaShift)
    qnn, cali = build()
    block = qnn.model.layer2[0]
    swap_and_cache(qnn, block, cali)
    torch.manual_seed(93)
    soft, hard = block_recon_fused_shiftedScale(block, iters=ITERS, lmda=[0.01, 0.02], model=qnn)
    out["B.fused.losses"] = np.array([soft, hard], dtype=np.float64)
    for n, m in mods_of(block):
        out[f"B.fused.{n}.alpha"] = npy(m.weight_quantizer.alpha)
        out[f"B.fused.{n}.beta"] = npy(m.weight_quantizer.beta)
    with torch.no_grad():
        out["B.fused.hard_out"] = npy(block(torch.cat(block.cached_inp_features)[:8]))
    # ---- C: activation step sizes through the shifted loop (act=True), after hard weights from A-style shift
    qnn, cali = build()
    block = qnn.model.layer2[0]
    swap_and_cache(qnn, block, cali)
    torch.manual_seed(94)
    block_recon_shiftedScale(block, iters=ITERS, lmda=0.01, model=qnn)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        qnn(cali[:32])                               # initialises the activation step sizes
    block.set_quant_state(True, True)
    out["C.delta0"] = np.array([float(block.act_quantizer.delta)] + [float(m.act_quantizer.delta) for _n, m in mods_of(block)
                                                                    if not m.act_quantizer.disable_act_quant and m.act_quantizer.delta is not None])
    torch.manual_seed(95)
    soft, hard = block_recon_shiftedScale(block, iters=ITERS, lmda=0.01, model=qnn, act=True)
    out["C.act.losses"] = np.array([soft, hard], dtype=np.float64)
    out["C.delta1"] = np.array([float(block.act_quantizer.delta)] + [float(m.act_quantizer.delta) for _n, m in mods_of(block)
                                                                    if not m.act_quantizer.disable_act_quant and m.act_quantizer.delta is not None])
    save("shift_loops", **out)


if __name__ == "__main__":
    main()
