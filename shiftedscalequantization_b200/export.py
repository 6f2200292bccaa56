"""True-integer export of a calibrated QuantModel.

The reference keeps dequantised fp32 weights only; its integer codes exist as the intermediate `x_quant` of the hard
forwards (quant/quant_layer.py:92-96, quant/adaptive_rounding.py:50-58, quant/channelQuantMSE.py:134-141,
quant/channelQuant.py:49-94). `export_int_weights` stores that intermediate bit-packed (2-bit weights: 16x smaller than
fp32) with the parameters needed to dequantise; `import_int_weights` rebuilds, bit for bit, the weights the quantised
model's forward uses. Both directions are `ssq_export_codes` / `ssq_import_codes` kernels (no host arithmetic).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .quant.adaptive_rounding import AdaRoundQuantizer
from .quant.channelQuant import ChannelQuant
from .quant.channelQuantMSE import ChannelQuantMSE
from .quant.quant_layer import QuantModule, UniformAffineQuantizer

FORMAT = "ssq-int-v1"


def _describe(q, w):
    """(code_delta, dequant_delta, zero_point, qmin, qmax, alpha, in_scale) of quantiser q in its HARD forward"""
    n_levels = q.n_levels
    unsigned = (0.0, float(n_levels - 1))
    if isinstance(q, AdaRoundQuantizer):
        if q.round_mode == 'learned_hard_sigmoid':
            if q.soft_targets:
                raise ops._lib.SsqError("export needs hard rounding: set soft_targets = False first")
            return q.delta, q.delta, q.zero_point, *unsigned, q.alpha, None
        if q.round_mode in ('nearest', 'nearest_ste'):
            return q.delta, q.delta, q.zero_point, *unsigned, None, None
        raise ops._lib.SsqError(f"round_mode {q.round_mode!r} has no deterministic integer code")
    if isinstance(q, ChannelQuantMSE):
        return q.delta, q.delta, q._zero(), *unsigned, None, q.inp_scale.reshape(-1)
    if isinstance(q, ChannelQuant):
        qmin, qmax = q._bounds()
        if q.opt_mode == 'none':
            d = q.delta * q.shiftedScale
            return d, d, q.zero_point, qmin, qmax, None, None
        if q.opt_mode == 'adaround':
            if not q.hard_round:
                raise ops._lib.SsqError("export needs hard rounding: set hard_round = True first")
            d = q.delta * q.shiftedScale
            return d, d, q.zero_point, qmin, qmax, q.beta, None
        if q.opt_mode == 'adaShift':
            # hard: floor(x / (delta * s_selected)) + (beta >= 0), dequantised with delta * 1.0 (channelQuant.py:51-64)
            if not (q.hard_round and q.hard_targets):
                raise ops._lib.SsqError("export needs hard targets and hard rounding")
            return q.get_delta(), q.delta * q.shiftedScale, q.zero_point, qmin, qmax, q.beta, None
        raise ops._lib.SsqError(f"ChannelQuant mode {q.opt_mode!r}: the mixture of dequantised candidates is not one integer grid")
    if isinstance(q, UniformAffineQuantizer):
        if q.sym:
            return q.delta, q.delta, q.zero_point, float(-n_levels // 2), float(n_levels // 2 - 1), None, None
        return q.delta, q.delta, q.zero_point, *unsigned, None, None
    raise ops._lib.SsqError(f"no integer export for {type(q).__name__}")


@torch.no_grad()
def export_int_weights(model: torch.nn.Module) -> Dict[str, dict]:
    """{module path: {codes uint8 [OC, row_bytes], delta, zero_point, qmin, n_bits, shape[, in_scale]}} for every
    QuantModule whose weight quantiser is initialised. Tensors stay on the device; use torch.save on the result."""
    out = {"__format__": FORMAT}
    for name, m in model.named_modules():
        if not isinstance(m, QuantModule):
            continue
        q = m.weight_quantizer
        if getattr(q, 'delta', None) is None:
            continue
        w = m.weight.detach()
        code_d, deq_d, zp, qmin, qmax, alpha, in_scale = _describe(q, w)
        code_d = code_d.detach().contiguous()
        packed = ops.export_codes(w, code_d, zp.detach(), qmin, qmax, q.n_bits,
                                  alpha=None if alpha is None else alpha.detach(), in_scale=in_scale)
        entry = {"codes": packed, "delta": deq_d.detach().clone().contiguous(), "zero_point": ops.match_param(zp.detach(), deq_d.detach()).clone(),
                 "qmin": float(qmin), "n_bits": int(q.n_bits), "shape": tuple(w.shape)}
        if in_scale is not None:
            entry["in_scale"] = in_scale.detach().clone()
        # the layer's output affine gamma^z / varphi^z (quant_layer.py:258-259) is part of what the layer computes once
        # bias_cal has trained it; identity (1, 0) is left out
        if not m._output_affine_is_identity():
            entry["out_scale"] = m.alpha_out.detach().reshape(-1).clone()
            entry["out_offset"] = m.beta_out.detach().reshape(-1).clone()
        out[name] = entry
    return out


@torch.no_grad()
def dequantize(entry: dict) -> torch.Tensor:
    """fp32 weight encoded by one export entry"""
    return ops.import_codes(entry["codes"], entry["shape"], entry["delta"], entry["zero_point"], entry["qmin"], entry["n_bits"],
                            in_scale=entry.get("in_scale"))


@torch.no_grad()
def import_int_weights(model: torch.nn.Module, blob: Dict[str, dict], strict: bool = True) -> int:
    """write the dequantised weights into the matching Conv2d / Linear / QuantModule `.weight` of `model`
    (a float model then computes what the quantised model computed; for a QuantModule that is its quantiser-off path). A
    trained output affine (bias_cal) is folded into weight and bias (W <- gamma_oc W, b <- gamma b + varphi: equal to the
    quantised forward up to fp32 rounding; bit-exact when the affine is the identity). Returns the number of layers written."""
    if blob.get("__format__") != FORMAT:
        raise ops._lib.SsqError("not an ssq integer export")
    mods = dict(model.named_modules())
    n = 0
    for name, entry in blob.items():
        if name == "__format__":
            continue
        m = mods.get(name)
        if m is None or not hasattr(m, "weight"):
            if strict:
                raise KeyError(f"no module {name!r} with a weight in the target model")
            continue
        wq = dequantize(entry)
        if tuple(m.weight.shape) != tuple(wq.shape):
            raise ops._lib.SsqError(f"{name}: shape {tuple(m.weight.shape)} vs exported {tuple(wq.shape)}")
        a, b = entry.get("out_scale"), entry.get("out_offset")
        bias = m.bias.data if getattr(m, "bias", None) is not None else None
        if a is not None:                   # fold gamma^z / varphi^z: (conv(x, W) + bias) * a + b == conv(x, a W) + (a bias + b)
            wq = wq * a.view((-1,) + (1,) * (wq.dim() - 1))
            bias = b.clone() if bias is None else bias * a + b
            if m.bias is None:
                m.bias = torch.nn.Parameter(bias.clone())
            else:
                m.bias.data.copy_(bias)
        m.weight.data.copy_(wq)
        if isinstance(m, QuantModule):      # the FP path of a QuantModule reads org_weight / org_bias
            m.org_weight.copy_(wq)
            if a is not None:
                if m.org_bias is None:
                    m.org_bias = bias.clone()
                else:
                    m.org_bias.copy_(bias)
        n += 1
    return n


def packed_bytes(blob: Dict[str, dict]) -> int:
    return sum(e["codes"].numel() for k, e in blob.items() if k != "__format__")
