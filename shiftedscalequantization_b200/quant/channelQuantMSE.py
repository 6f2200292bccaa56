"""ChannelQuantMSE — per-(input-channel x kh x kw) input scale on top of the per-output-channel step size;
mirror of the live part of the reference's quant/channelQuantMSE.py (:7-143; the rest of that file is commented out).
"""
import torch
from torch import nn

from .. import ops
from .quant_layer import UniformAffineQuantizer


class ChannelQuantMSE(nn.Module):
    @torch.no_grad()
    def __init__(self, delta, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor, shiftTarget: int = 2, act=False,
                 opt_mode='max', level=1, threshold=1.0, name='--'):
        super().__init__()
        self.RUN_CHANNEL_WISE = True
        self.act = act
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta * delta
        self.zero_point = uaq.zero_point
        self.n_levels = uaq.n_levels
        self.raw_zero_point = uaq.raw_zero_point
        self.device = weight_tensor.device
        self.isFC = len(self.delta.shape) != 4
        self.nchannel = (weight_tensor.shape[0], weight_tensor.shape[1])
        self.shiftTarget = shiftTarget
        self.x_q = []
        self.opt_mode = opt_mode
        self.hard_targets = False
        self.hard_round = False
        self.gamma, self.zeta = -0.1, 1.1
        self.alpha = None
        self.beta = None
        self.deltaQuant = None
        self.shiftedDone = False
        shape = (1, weight_tensor.shape[1]) if self.isFC else (1,) + tuple(weight_tensor.shape[1:])
        self.inp_scale = torch.ones(shape, device=self.device)
        self.scale_threshold = threshold
        self.scale_level = level
        self.name = name

    def _zero(self):
        # round(raw_zero_point / delta): the zero point is re-derived from the raw offset (channelQuantMSE.py:128,136)
        return torch.round(self.raw_zero_point / self.delta)

    def mse_calc(self, x, x_quant, ignore_inp_scale=False):
        """mean squared error of the dequantised codes (channelQuantMSE.py:53-68); host scalar"""
        zero = self._zero()
        x_float = (x_quant - zero) * self.delta if ignore_inp_scale else (x_quant - zero) * self.delta * self.inp_scale
        return torch.mean(torch.square(x_float - x)).item()

    def init_scale(self, x):
        """for each column keep the smallest candidate scale c in {level/level .. 1/level} whose rescaled codes stay
        inside the (threshold-widened) code range for every output channel (channelQuantMSE.py:70-110)"""
        mode = self.opt_mode
        if mode != 'max':
            raise NotImplementedError
        x_range = self.n_levels - 1
        min_lim = 0.0 - 0.5 / x_range * self.scale_threshold
        max_lim = 1.0 + 0.5 / x_range * self.scale_threshold
        cand = torch.tensor([i / self.scale_level for i in range(self.scale_level, 0, -1)], dtype=torch.float32, device=x.device)
        oc = x.shape[0]
        scale = self.inp_scale.clone().reshape(-1).contiguous()
        lo = float(torch.tensor(min_lim, dtype=torch.float32))      # python double -> fp32, as ATen compares
        hi = float(torch.tensor(max_lim, dtype=torch.float32))
        ops.inp_scale_search(x.detach().reshape(oc, -1), self.delta.detach().reshape(-1),
                             self.raw_zero_point.detach().reshape(-1).contiguous(), cand, x_range, lo, hi, scale)
        self.inp_scale = scale.view_as(self.inp_scale)

    def quant(self, x):
        _y, codes = ops.fq_affine_fwd(x.detach(), self.delta.detach(), self._zero().detach(), 0.0, float(self.n_levels - 1),
                                      in_scale=self.inp_scale.reshape(-1), want_codes=True)
        return codes

    def forward(self, x):
        return ops.fq_affine_fwd(x.detach(), self.delta.detach(), self._zero().detach(), 0.0, float(self.n_levels - 1),
                                 in_scale=self.inp_scale.reshape(-1))
