"""Generates tests/golden/bias_cal.npz: the REAL reference's block-reconstruction loop body (quant/block_recon.py:89-105)
with the output-channel affine gamma^z / varphi^z of every QuantModule (alpha_out / beta_out, quant/quant_layer.py:231-238,
applied at :258-259) handed to the optimiser next to the AdaRound alphas — README `--bias_cal`; upstream keeps those two
`opt_params +=` lines commented (quant/layer_recon_fused_shiftedScale.py:67-68). Everything evaluated here is the
reference's own forward/autograd/LossFunction/torch.optim.Adam on the CPU; only the parameter list is extended.
Run in the build container only:  python tests/golden/make_golden_bias_cal.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, npy, save  # noqa: E402


def main():
    import_reference()
    from quant import QuantModel, QuantModule
    from quant.adaptive_rounding import AdaRoundQuantizer
    from quant.block_recon import LossFunction
    from quant.data_utils import save_inp_oup_data
    from models.resnet import resnet18 as ref_resnet18
    torch.manual_seed(1005)
    cnn = ref_resnet18(num_classes=10).eval()
    wq = {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}
    aq = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
    qnn = QuantModel(model=cnn, weight_quant_params=wq, act_quant_params=aq).eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(32, 3, 16, 16)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali)
    block = qnn.model.layer2[0]                    # conv1, conv2, downsample
    iters, bs = 16, 16
    qnn.set_quant_state(False, False); block.set_quant_state(True, False)
    mods = [(n, m) for n, m in block.named_modules() if isinstance(m, QuantModule)]
    for _n, m in mods:
        m.weight_quantizer = AdaRoundQuantizer(uaq=m.weight_quantizer, round_mode='learned_hard_sigmoid', weight_tensor=m.org_weight.data)
        m.weight_quantizer.soft_targets = True
    params = []
    for _n, m in mods:
        params += [m.weight_quantizer.alpha, m.alpha_out, m.beta_out]
    opt = torch.optim.Adam(params)
    lf = LossFunction(block, round_loss='relaxation', weight=0.01, max_count=iters, rec_loss='mse', b_range=(20, 2),
                      decay_start=0, warmup=0.2, p=2.0)
    inps, outs = save_inp_oup_data(qnn, block, cali, True, False, bs)
    out = {"cali": npy(cali), "inps": npy(inps), "outs": npy(outs), "iters": np.array(iters), "bs": np.array(bs)}
    torch.manual_seed(177)
    idx_tab, losses = [], []
    for i in range(iters):
        idx = torch.randperm(inps.size(0))[:bs]
        idx_tab.append(npy(idx))
        opt.zero_grad()
        err = lf(block(inps[idx]), outs[idx])
        err.backward(retain_graph=True)
        if i == 0:
            for n, m in mods:
                out[f"{n}.g_alpha_out0"] = npy(m.alpha_out.grad); out[f"{n}.g_beta_out0"] = npy(m.beta_out.grad)
        opt.step()
        losses.append(float(err))
    out["idx"] = np.stack(idx_tab); out["losses"] = np.array(losses)
    for n, m in mods:
        out[f"{n}.alpha"] = npy(m.weight_quantizer.alpha)
        out[f"{n}.alpha_out"] = npy(m.alpha_out); out[f"{n}.beta_out"] = npy(m.beta_out)
    for _n, m in mods:
        m.weight_quantizer.soft_targets = False
    with torch.no_grad():
        out["hard_out"] = npy(block(inps[:8]))
    save("bias_cal", **out)


if __name__ == "__main__":
    main()
