#!/usr/bin/env python
"""Golden vectors for the ACTIVATION phase of block reconstruction (LSQ step-size learning, quant/block_recon.py:38-48,66-105 with
act_quant=True: Adam lr 4e-4 + CosineAnnealingLR, lp_loss p = 2.4): the REAL reference on the CPU (`device = 'cuda'` at :88 replaced
by the model's device, as for the other loop goldens). ResNet-18 (10 classes, 16x16 images) layer1.0: a short weight phase installs
the AdaRound quantisers, the activation quantisers are initialised by one forward, then 16 iterations learn the two step sizes
(conv1's output and the block's output; conv2's own quantiser is disabled inside a BasicBlock, quant_block.py).

    python tests/golden/make_golden_act_phase.py        # needs /root/reference; writes tests/golden/act_phase.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_round2 as G2                                      # noqa: E402
from make_golden import import_reference, npy, save                  # noqa: E402


def main():
    import_reference()
    BR = G2._patched_block_recon()
    from quant import QuantModule
    from quant.data_utils import save_inp_oup_data
    qnn, cali = G2.build("resnet18", 2, 16, 32)
    block = qnn.model.layer1[0]
    bs, iters = 16, 16
    torch.manual_seed(41)
    BR.block_reconstruction(qnn, block, cali_data=cali, iters=6, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                            act_quant=False, opt_mode='mse', batch_size=bs)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        qnn(cali)                                                     # activation-quantiser scale init (per-tensor mse search)
    qnn.disable_network_output_quantization()
    out = {"cali": npy(cali), "iters": np.array(iters), "bs": np.array(bs)}
    for n, m in block.named_modules():
        if isinstance(m, QuantModule):
            q = m.weight_quantizer
            out[f"{n}.weight"] = npy(m.org_weight); out[f"{n}.bias"] = npy(m.org_bias)
            out[f"{n}.delta"] = npy(q.delta); out[f"{n}.zp"] = npy(q.zero_point); out[f"{n}.alpha"] = npy(q.alpha)
    aq = {"conv1": block.conv1.act_quantizer, "__block__": block.act_quantizer}
    for k, q in aq.items():
        out[f"{k}.act_delta0"] = npy(q.delta); out[f"{k}.act_zp"] = npy(q.zero_point); out[f"{k}.act_levels"] = np.array(q.n_levels)
    # the features the loop will cache (block_recon.py:52): same call, same state
    qnn.set_quant_state(False, False); block.set_quant_state(True, True)
    inps, outs = save_inp_oup_data(qnn, block, cali, True, True, bs)
    out["inps"] = npy(inps); out["outs"] = npy(outs)
    torch.manual_seed(43)
    BR.block_reconstruction(qnn, block, cali_data=cali, iters=iters, asym=True, act_quant=True, opt_mode='mse', lr=4e-4, p=2.4, batch_size=bs)
    for k, q in aq.items():
        out[f"{k}.act_delta1"] = npy(q.delta)
    torch.manual_seed(43)
    out["idx"] = np.stack([npy(torch.randperm(inps.size(0))[:bs]) for _ in range(iters)])
    save("act_phase", **out)
    for k in aq:
        print(k, float(out[f"{k}.act_delta0"]), "->", float(out[f"{k}.act_delta1"]))


if __name__ == "__main__":
    main()
