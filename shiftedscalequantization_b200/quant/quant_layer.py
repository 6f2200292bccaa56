"""Host-side mirror of the reference's quant/quant_layer.py: same class names, constructor signatures,
attributes and error behaviour; every quantiser computation is a libssq_b200 kernel.

Reference surface reproduced here (upstream path:line):
  StraightThrough / round_ste / lp_loss      quant/quant_layer.py:10-32
  UniformAffineQuantizer                     quant/quant_layer.py:35-185
  QuantModule (forward path only)            quant/quant_layer.py:188-311
The greedy / dist search methods of QuantModule (quant_layer.py:313-528) target an older
ChannelQuant API that no longer exists upstream and are out of scope.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class StraightThrough(nn.Module):
    def __init__(self, channel_num: int = 1):
        super().__init__()

    def forward(self, input):
        return input


def round_ste(x: torch.Tensor):
    """rint with a straight-through gradient (quant_layer.py:18-22). Helper kept for API parity;
    the fused kernels apply the same rule internally."""
    return (x.round() - x).detach() + x


def lp_loss(pred, tgt, p=2.0, reduction='none'):
    """L_p reconstruction loss (quant_layer.py:25-32) as one fused reduction kernel."""
    return ops.ReconLoss.apply(pred, tgt, float(p), 'mse' if reduction == 'none' else 'mse_all', None)


def _bounds(n_levels: int, sym: bool):
    # quant_layer.py:93-96 — python ints: -L//2 and L//2-1 when symmetric, else 0 and L-1
    return (float(-n_levels // 2), float(n_levels // 2 - 1)) if sym else (0.0, float(n_levels - 1))


class UniformAffineQuantizer(nn.Module):
    """Uniform affine fake-quantiser with STE backward; scale init by 'max' or 'mse' clip search.

    :param n_bits: bit width (1..8)
    :param symmetric: signed clamp range, zero_point fixed at 0 by the 'mse' init
    :param channel_wise: one (delta, zero_point) per slice along dim 0
    :param scale_method: 'max' ('..scale..' variants shrink the range) or 'mse'
    """

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False, scale_method: str = 'max',
                 leaf_param: bool = False, tune_delta_zero: bool = False, ch: int = 64, disable_act_quant: bool = False):
        super().__init__()
        assert 1 <= n_bits <= 8, 'bitwidth not supported'
        self.sym = symmetric
        self.n_bits = n_bits
        self.n_levels = 2 ** self.n_bits
        self.delta = None
        self.zero_point = None
        self.raw_zero_point = None
        self.inited = False
        self.leaf_param = leaf_param
        self.channel_wise = channel_wise
        self.scale_method = scale_method
        self.disable_act_quant = disable_act_quant
        if tune_delta_zero or disable_act_quant:
            return
        # placeholder parameters so a state_dict saved after calibration loads into a fresh model
        if leaf_param:
            shape = ()
        elif type(ch) is int:
            shape = (ch, 1)
        elif len(ch) == 2:
            shape = (ch[0], 1)
        else:
            shape = (ch[0], 1, 1, 1)
        self.delta = nn.Parameter(torch.zeros(shape))
        self.zero_point = nn.Parameter(torch.zeros(shape))

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor):
        if self.inited is False:
            delta, zero_point, self.raw_zero_point = self.init_quantization_scale(x, self.channel_wise)
            self.delta = nn.Parameter(delta)
            self.zero_point = nn.Parameter(zero_point)
            self.inited = True
        qmin, qmax = _bounds(self.n_levels, self.sym)
        return ops.FakeQuantAffine.apply(x, self.delta, self.zero_point, qmin, qmax)

    # ------------------------------------------------------------------ scale initialisation
    def init_quantization_scale(self, x: torch.Tensor, channel_wise: bool = False):
        """Returns (delta, zero_point, raw_zero_point) shaped [C,1,..] (channel_wise) or 0-dim."""
        x = x.detach()
        rows = x.reshape(x.shape[0], -1) if channel_wise else x.reshape(1, -1)
        if 'max' in self.scale_method:
            delta, zp, raw = self._init_max(rows)
        elif self.scale_method == 'mse':
            delta, zp, raw = self._init_mse(rows)
        else:
            raise NotImplementedError
        if channel_wise:
            shape = (-1, 1, 1, 1) if x.dim() == 4 else (-1, 1)
            return delta.view(shape), zp.view(shape), raw.view(shape)
        return delta.reshape(()), zp.reshape(()), raw.reshape(())

    def _init_max(self, rows: torch.Tensor):
        # quant_layer.py:124-142 is Python-double arithmetic on .item() values: the min/max come from one
        # reduction kernel, the doubles are evaluated here exactly as upstream.
        mn, mx = ops.row_minmax(rows)
        mn, mx = mn.tolist(), mx.tolist()
        deltas, zps, raws = [], [], []
        for x_min, x_max in zip(mn, mx):
            x_min = min(x_min, 0)
            x_max = max(x_max, 0)
            if 'scale' in self.scale_method:
                x_min = x_min * (self.n_bits + 2) / 8
                x_max = x_max * (self.n_bits + 2) / 8
            if self.sym:
                x_absmax = max(abs(x_min), x_max)
                x_min, x_max = -x_absmax if x_min < 0 else 0, x_absmax
            delta = float(x_max - x_min) / (self.n_levels - 1)
            if delta < 1e-8:
                delta = 1e-8
            deltas.append(delta)
            zps.append(round(-x_min / delta))
            raws.append(-x_min)
        mk = lambda v: torch.tensor(v, dtype=torch.float64).to(rows.dtype).to(rows.device)
        return mk(deltas), mk(zps), mk(raws)

    def _init_mse(self, rows: torch.Tensor):
        # multi-GPU: the per-channel search is sharded by output channel (rows) and all-gathered; replicas hold the
        # same weights, the kernel is deterministic, so every rank ends with bit-identical parameters
        from .. import dist as ssq_dist
        world = ssq_dist.world_size()
        if world > 1 and self.channel_wise and rows.shape[0] >= world:
            lo, hi = ssq_dist.shard_range(rows.shape[0])
            parts = ops.mse_scale_search(rows[lo:hi].contiguous(), self.n_levels, self.sym, 2.4)
            packed = torch.stack([parts[0], parts[1], parts[2], parts[4].to(torch.float32)], dim=1)
            full = ssq_dist.all_gather_rows(packed, rows.shape[0])
            delta, zp, raw, idx = full[:, 0].contiguous(), full[:, 1].contiguous(), full[:, 2].contiguous(), full[:, 3]
        else:
            delta, zp, raw, _score, idx = ops.mse_scale_search(rows, self.n_levels, self.sym, 2.4)
        if bool((idx < 0).any()):
            # upstream: every score is NaN (all-zero channel) so delta stays None and the assignment
            # `delta[c] = None` raises TypeError (quant_layer.py:114)
            raise TypeError("can't assign a NoneType to a torch.FloatTensor")
        return delta, zp, raw

    def quantize(self, x, max, min):
        """candidate fake-quant used by the search (quant_layer.py:168-175; always the unsigned clamp)"""
        dev = x.device
        mx = torch.as_tensor(max, dtype=torch.float32, device=dev)
        mn = torch.as_tensor(min, dtype=torch.float32, device=dev)
        delta = ((mx - mn) / (2 ** self.n_bits - 1)).reshape(1)
        zero_point = (-mn.reshape(1) / delta).round()
        return ops.fq_affine_fwd(x.detach(), delta, zero_point, 0.0, float(self.n_levels - 1))

    def bitwidth_refactor(self, refactored_bit: int):
        assert 2 <= refactored_bit <= 8, 'bitwidth not supported'
        self.n_bits = refactored_bit
        self.n_levels = 2 ** self.n_bits

    def extra_repr(self):
        return (f'bit={self.n_bits}, scale_method={self.scale_method}, symmetric={self.sym}, '
                f'channel_wise={self.channel_wise}, leaf_param={self.leaf_param}')


class QuantModule(nn.Module):
    """Conv2d / Linear whose weight (and optionally output activation) is fake-quantised.
    The contraction itself stays on torch (cuDNN / cuBLAS): it is a dense op the kernels do not own."""

    def __init__(self, org_module: Union[nn.Conv2d, nn.Linear], weight_quant_params: dict = {},
                 act_quant_params: dict = {}, disable_act_quant: bool = False, se_module=None):
        super().__init__()
        if isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = F.conv2d
        else:
            self.fwd_kwargs = dict()
            self.fwd_func = F.linear
        self.weight = org_module.weight
        self.org_weight = org_module.weight.data.clone()
        self.bias = org_module.bias
        self.org_bias = None if org_module.bias is None else org_module.bias.data.clone()
        self.use_weight_quant = False
        self.use_act_quant = False
        self.disable_act_quant = disable_act_quant
        # the shared dicts are mutated in place, as upstream does (quant_layer.py:216-218)
        weight_quant_params['ch'] = self.weight.shape
        self.weight_quantizer = UniformAffineQuantizer(**weight_quant_params)
        act_quant_params['disable_act_quant'] = disable_act_quant
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False
        self.se_module = se_module
        self.extra_repr = org_module.extra_repr
        self.cache_features = 'none'
        self.cached_inp_features = []
        self.cached_out_features = []
        n_ch = self.weight.shape[0]
        affine_shape = (1, n_ch, 1, 1) if self.weight.dim() == 4 else (1, n_ch)
        self.alpha_out = nn.Parameter(torch.ones(affine_shape))     # README's gamma^z
        self.beta_out = nn.Parameter(torch.zeros(affine_shape))     # README's varphi^z
        self._affine_key = None
        self._affine_identity = True
        self.train_output_affine = False    # True while gamma^z / varphi^z are being learned: always apply the affine
        self._engine_weight = None      # set by ReconEngine: weight already produced by a multi-tensor launch
        self._engine_bias = None        # set by ReconEngine when the output affine is folded: gamma*bias + varphi
        self.selection = None
        self.selectionInited = False
        self.pathName = ''
        self.dump_cnt = 0

    def _output_affine_is_identity(self) -> bool:
        """alpha_out == 1 and beta_out == 0 make `out*alpha_out+beta_out` a bit-exact no-op, so the
        pass over the activation is skipped. Re-checked (one device read) only when either parameter's
        storage or version changed."""
        a, b = self.alpha_out, self.beta_out
        key = (a.data_ptr(), a._version, b.data_ptr(), b._version, a.device)
        if key != self._affine_key:
            self._affine_identity = bool((a.detach() == 1).all()) and bool((b.detach() == 0).all())
            self._affine_key = key
        return self._affine_identity

    def forward(self, input: torch.Tensor):
        if self.cache_features == 'if':
            self.cached_inp_features += [input.cpu().clone().detach()]
        quantized = self.use_weight_quant and self.cache_features == 'none'
        folded = False
        if quantized:
            weight = self._engine_weight if self._engine_weight is not None else self.weight_quantizer(self.weight)
            folded = self._engine_weight is not None and self._engine_bias is not None
            bias = self._engine_bias if folded else self.bias      # folded: W_eff = gamma W_q, b_eff = gamma b + varphi
        else:
            weight, bias = self.org_weight, self.org_bias
        out = self.fwd_func(input, weight, bias, **self.fwd_kwargs)
        if quantized and not folded and (self.train_output_affine or not self._output_affine_is_identity()):
            out = ops.ChanAffine.apply(out, self.alpha_out, self.beta_out)
        if self.se_module is not None:
            out = self.se_module(out)
        out = self.activation_function(out)
        if not self.disable_act_quant and self.use_act_quant:
            out = self.act_quantizer(out)
        if self.cache_features == 'debug':
            torch.save(out, f'fc_{self.dump_cnt}.pt')
            self.dump_cnt += 1
        if self.cache_features == 'of':
            self.cached_out_features += [out.cpu().clone().detach()]
        return out

    def set_quant_init_state(self):
        self.weight_quantizer.inited = True
        self.act_quantizer.inited = True

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant

    def disable_cache_features(self):
        self.cache_features = 'none'

    def clear_cached_features(self):
        self.cached_inp_features = []
        self.cached_out_features = []

    def getLoss(self, A, B, p=2.0):
        """sum over cached batches of the L_p loss (quant_layer.py:309-311 region)"""
        loss = 0.0
        for a, b in zip(A, B):
            loss += float(lp_loss(a, b, p=p, reduction='none'))
        return loss
