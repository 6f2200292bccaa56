#!/usr/bin/env python
"""Golden vectors for ChannelQuant.init_shift_candidates (quant/channelQuant.py:240-277; its only call site upstream is commented
out at :281, the method itself runs): the REAL reference class on the CPU, seeded weights, scale from its own mse init.

    python tests/golden/make_golden_shift_candidates.py        # needs /root/reference; writes tests/golden/shift_candidates.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                                             # noqa: E402  (import_reference, gen_weights, npy, save)


def main():
    MG.import_reference()
    from quant.quant_layer import UniformAffineQuantizer
    from quant.channelQuant import ChannelQuant
    out = {}
    for i, (name, shape, bits, scale) in enumerate([("conv_b2", (16, 6, 3, 3), 2, 0.05), ("conv_b4", (8, 12, 3, 3), 4, 0.05),
                                                    ("fc_b3", (12, 20), 3, 0.1), ("dw_b3", (8, 1, 3, 3), 3, 0.05)]):
        w = MG.gen_weights(shape, 900 + i, scale)
        uaq = UniformAffineQuantizer(n_bits=bits, channel_wise=True, scale_method="mse")
        uaq(w)
        q = ChannelQuant(1.0, uaq, w, shiftTarget=[0.96875, 1.03125, 1.0])
        q.init_shift_candidates(w.clone())
        out[f"{name}.w"] = MG.npy(w); out[f"{name}.delta"] = MG.npy(q.delta); out[f"{name}.zp"] = MG.npy(q.zero_point)
        out[f"{name}.bits"] = np.array(bits)
        out[f"{name}.shiftTarget"] = np.array([float(s) for s in q.shiftTarget], dtype=np.float64)
    MG.save("shift_candidates", **out)


if __name__ == "__main__":
    main()
