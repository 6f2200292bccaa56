"""AdaRoundQuantizer — mirror of the reference's quant/adaptive_rounding.py:6-74 over the K1b kernels."""
import torch
from torch import nn

from .. import ops
from .quant_layer import UniformAffineQuantizer, round_ste


class AdaRoundQuantizer(nn.Module):
    """Adaptive rounding (https://arxiv.org/abs/2004.10568): floor(w/delta) + h(alpha), h the rectified sigmoid.

    :param uaq: UniformAffineQuantizer providing n_bits / delta / zero_point
    :param weight_tensor: weights used to initialise alpha so that h(alpha) = frac(w/delta)
    :param round_mode: 'learned_hard_sigmoid' (trainable), 'nearest', 'nearest_ste', 'stochastic'
    """

    def __init__(self, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor, round_mode='learned_round_sigmoid'):
        super().__init__()
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta
        self.zero_point = uaq.zero_point
        self.n_levels = uaq.n_levels
        self.round_mode = round_mode
        self.alpha = None
        self.soft_targets = False
        self.gamma, self.zeta = -0.1, 1.1
        self.beta = 2 / 3
        self.init_alpha(x=weight_tensor.clone())

    def forward(self, x):
        hi = float(self.n_levels - 1)       # upstream always clamps unsigned here, even if sym (adaptive_rounding.py:58)
        if self.round_mode == 'learned_hard_sigmoid':
            if self.soft_targets:
                return ops.AdaRoundSoft.apply(x, self.alpha, self.delta, self.zero_point, 0.0, hi)
            return ops.adaround_fwd(x.detach(), self.alpha.detach(), self.delta.detach(), self.zero_point.detach(),
                                    0.0, hi, soft=False)
        if self.round_mode == 'nearest':
            return ops.fq_affine_fwd(x.detach(), self.delta.detach(), self.zero_point.detach(), 0.0, hi)
        if self.round_mode == 'nearest_ste':
            return ops.FakeQuantAffine.apply(x, self.delta, self.zero_point, 0.0, hi)
        if self.round_mode == 'stochastic':
            # RNG-driven debug mode (never used by the calibration loops): torch's bernoulli stream is the contract
            x_floor = torch.floor(x / self.delta)
            x_int = x_floor + torch.bernoulli((x / self.delta) - x_floor)
            print('Draw stochastic sample')
            x_quant = torch.clamp(x_int + self.zero_point, 0, self.n_levels - 1)
            return (x_quant - self.zero_point) * self.delta
        raise ValueError('Wrong rounding mode')

    def get_soft_targets(self):
        """h(alpha); differentiable, for callers that build the regulariser themselves"""
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def init_alpha(self, x: torch.Tensor):
        if self.round_mode != 'learned_hard_sigmoid':
            raise NotImplementedError
        self.alpha = nn.Parameter(ops.adaround_init_alpha(x.detach(), self.delta.detach()))
