"""ChannelQuantAct — activation twin of ChannelQuant; mirror of the reference's quant/channelQuantAct.py.

Upstream's init_v references undefined names (`x`, `self.x_q`, `self.init_alpha`, `self.isFC`; channelQuantAct.py:126-134)
and the shifted forward paths need that state, so only the 'none' and 'adaround' forward modes can execute there.
Those two are implemented on the kernels; the shifted modes raise the same NameError/AttributeError family a caller
would hit upstream, with a message saying why (parity for them is unpinned: SURVEY.md §8c)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .quant_layer import UniformAffineQuantizer


class ChannelQuantAct(nn.Module):
    @torch.no_grad()
    def __init__(self, uaq: UniformAffineQuantizer, shiftTarget: list = [2 / 2, 2 / 2]):
        super().__init__()
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta
        self.zero_point = uaq.zero_point
        self.n_levels = uaq.n_levels
        self.device = 'cuda:0'
        self.shiftedScale = 1.0
        self.shiftTarget = shiftTarget
        self.opt_mode = 'none'
        self.hard_targets = False
        self.hard_round = False
        self.gamma, self.zeta = -0.1, 1.1
        self.alpha = None
        self.beta = None
        self.shiftedDone = False

    def forward(self, x):
        hi = float(self.n_levels - 1)          # activations always use the unsigned clamp (channelQuantAct.py:36,45,55)
        d = (self.delta * self.shiftedScale).detach()
        if self.opt_mode == 'adaround':
            if self.hard_round:
                return ops.adaround_fwd(x.detach(), self.beta.detach(), d, self.zero_point.detach(), 0.0, hi, soft=False)
            return ops.AdaRoundSoft.apply(x, self.beta, d, self.zero_point.detach(), 0.0, hi)
        if self.opt_mode == 'none':
            return ops.fq_affine_fwd(x.detach(), d, self.zero_point.detach(), 0.0, hi)
        if self.opt_mode == 'adaShift' or self.opt_mode in 'learned_hard_sigmoid':
            return self.shifted_x_quant()
        raise ValueError('opt_mode is not defined')

    def shifted_x_quant(self):
        raise AttributeError("ChannelQuantAct has no candidate cache 'x_q': upstream's init_v (channelQuantAct.py:126-134) "
                             "cannot run, so the shifted activation modes are undefined")

    def get_sig_soft_targets(self):
        return torch.clamp(F.softmax(self.alpha, dim=-1) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def get_soft_targets(self):
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def get_soft_round(self):
        return torch.clamp(torch.sigmoid(self.beta) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def inverse_softmax(self, x):
        x = (x - self.gamma) / (self.zeta - self.gamma)
        logits = torch.log(x)
        return logits - torch.mean(logits, dim=-1, keepdim=True)

    @torch.no_grad()
    def init_v(self):
        raise NameError("name 'x' is not defined (upstream channelQuantAct.py:130 reads an undefined variable)")
