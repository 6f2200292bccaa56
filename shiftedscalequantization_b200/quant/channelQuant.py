"""ChannelQuant — the "shifted scale" weight quantiser; mirror of the reference's quant/channelQuant.py:7-307.

Each input channel (conv) or each element (FC) chooses, through a soft-max over `alpha`, one of S shifted step
sizes delta*s_i. Forward modes (opt_mode): 'none' (plain rounding), 'learned_hard_sigmoid' (soft/hard mixture of
the S dequantised candidates), 'adaround' (floor + h(beta)), 'adaShift' (mixture of the S integer floors + h(beta)).
The reference materialises the S candidates as weight-sized tensors `x_q`; the kernels recompute them from the
weight (same arithmetic, 12 B/elem instead of (S+2)*4), while `x_q` is still populated for API parity.
One-off initialisers (init_alpha, get_delta, init_shift_candidates) are host-side parameter bookkeeping on
[IC,S]-sized tensors and stay in torch.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .quant_layer import UniformAffineQuantizer, round_ste  # noqa: F401


class ChannelQuant(nn.Module):
    @torch.no_grad()
    def __init__(self, delta, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor, shiftTarget: list = [2 / 2, 2 / 2],
                 act=False, name='--'):
        super().__init__()
        self.RUN_CHANNEL_WISE = True
        self.act = act
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta * delta
        self.zero_point = uaq.zero_point
        self.n_levels = uaq.n_levels
        self.device = weight_tensor.device
        self.isFC = len(self.delta.shape) != 4
        self.nchannel = (weight_tensor.shape[0], weight_tensor.shape[1])
        self.shiftedScale = 1.0
        self.shiftTarget = shiftTarget
        self.x_q = []
        self.opt_mode = 'none'
        self.hard_targets = False
        self.hard_round = False
        self.gamma, self.zeta = -0.1, 1.1
        self.alpha = None
        self.beta = None
        self.deltaQuant = None
        self.shiftedDone = False
        self.name = name
        self._w_src = None          # weight the S candidates are derived from (argument of init_v / init_v_beta)
        self._shift_delta = None    # [S, OC] = delta * s_i (fp32 tensor * python float, like upstream)

    # ------------------------------------------------------------------ helpers
    def _bounds(self):
        if self.sym:
            return float(-self.n_levels // 2), float(self.n_levels // 2 - 1)
        return 0.0, float(self.n_levels - 1)

    def _remember_source(self, x):
        self._w_src = x.detach().contiguous()
        d = self.delta.detach()
        self._shift_delta = torch.stack([(d * st).reshape(-1) for st in self.shiftTarget]).contiguous()

    def _mixture(self, mode):
        """shifted_x_quant() (channelQuant.py:96-118) on recomputed candidates"""
        qmin, qmax = self._bounds()
        d, z = self.delta.detach(), ops.match_param(self.zero_point.detach(), self.delta.detach())
        beta = self.beta if mode == ops.SHIFT_ADASHIFT else None
        if self.hard_targets:
            p = ops.shift_probs_fwd(self.alpha.detach())
            b = None if beta is None else beta.detach()
            if mode == ops.SHIFT_ADASHIFT and not self.hard_round and beta.requires_grad and torch.is_grad_enabled():
                raise NotImplementedError('hard_targets with a trainable soft round is not used by any upstream loop')
            return ops.fq_shift_fwd(self._w_src, self._shift_delta, d, z, p, b, mode, True, self.hard_round, qmin, qmax, self.isFC)
        p = ops.ShiftProbs.apply(self.alpha)
        return ops.ShiftMix.apply(p, beta, self._w_src, self._shift_delta, d, z, mode, self.hard_round, qmin, qmax, self.isFC)

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        qmin, qmax = self._bounds()
        if self.opt_mode == 'adaShift':
            return self._mixture(ops.SHIFT_ADASHIFT)
        if self.opt_mode == 'adaround':
            d = (self.delta * self.shiftedScale).detach()
            if self.hard_round:
                return ops.adaround_fwd(x.detach(), self.beta.detach(), d, self.zero_point.detach(), qmin, qmax, soft=False)
            return ops.AdaRoundSoft.apply(x, self.beta, d, self.zero_point.detach(), qmin, qmax)
        if self.opt_mode == 'none':
            d = (self.delta * self.shiftedScale).detach()
            return ops.fq_affine_fwd(x.detach(), d, self.zero_point.detach(), qmin, qmax)
        if self.opt_mode in 'learned_hard_sigmoid':      # substring test, as upstream (channelQuant.py:81)
            return self._mixture(ops.SHIFT_DEQUANT)
        raise ValueError('opt_mode is not defined')

    def shifted_x_quant(self):
        return self._mixture(ops.SHIFT_ADASHIFT if self.opt_mode == 'adaShift' else ops.SHIFT_DEQUANT)

    def get_sig_soft_targets(self):
        return torch.clamp(F.softmax(self.alpha, dim=-1) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def get_soft_targets(self):
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def get_soft_round(self):
        return torch.clamp(torch.sigmoid(self.beta) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    # ------------------------------------------------------------------ initialisers (one-off, host-side bookkeeping)
    def init_alpha(self, x: torch.Tensor, clip=0.80, device='cuda'):
        """one-hot of the min-error shift per group, smoothed to (clip, rest...) with clip forced to 0.33
        (channelQuant.py:158-191), mapped through the inverse of the stretched soft-max"""
        clip = 0.33
        S = len(self.shiftTarget)
        errs = []
        for i in range(S):
            sq = (x - self.x_q[i]) ** 2
            errs.append(sq if self.isFC else torch.sum(sq, dim=(0, 2, 3) if self.RUN_CHANNEL_WISE else (2, 3)))
        _, min_index = torch.min(torch.stack(errs, dim=0), dim=0)
        if S == 1:
            rest, clip = 0, 1.0
        else:
            rest = (1.0 - clip) / (S - 1)
        alpha = torch.full((*min_index.shape, S), rest, dtype=torch.float, device=device)
        alpha[F.one_hot(min_index, S).bool().to(alpha.device)] = clip
        return self.inverse_softmax(alpha)

    def inverse_softmax(self, x):
        x = (x - self.gamma) / (self.zeta - self.gamma)
        logits = torch.log(x)
        return logits - torch.mean(logits, dim=-1, keepdim=True)

    @torch.no_grad()
    def init_v(self, x: torch.Tensor):
        """dequantised candidate per shift, alpha init, switch to the mixture forward (channelQuant.py:201-213)"""
        for st in self.shiftTarget:
            self.shiftedScale = st
            self.x_q.append(self(x))
        self.shiftedScale = 1.0
        self._remember_source(x)
        alpha = self.init_alpha(x, clip=(0.90 - 0.05 * len(self.shiftTarget)), device=self.device)
        self.alpha = nn.Parameter(alpha)
        self.opt_mode = 'learned_hard_sigmoid'

    def get_delta(self):
        """step size selected by argmax of the group probabilities, per (oc, ic) (channelQuant.py:221-237)"""
        p = self.get_sig_soft_targets()
        if p.dim() == 2:
            p = p.unsqueeze(0)
        max_index = torch.argmax(p, dim=-1)
        delta = self.delta * self.shiftTarget[0]
        for i in range(1, len(self.shiftTarget)):
            mask = max_index == i
            if not self.isFC:
                mask = mask.unsqueeze(-1).unsqueeze(-1)
            delta = torch.where(mask, self.delta * self.shiftTarget[i], delta)
        return delta

    @torch.no_grad()
    def init_shift_candidates(self, x):
        """rank-vote the 14 scales i/8 by per-group L2.4 error and keep the two best plus 1.0
        (channelQuant.py:239-277; upstream's call site is commented out at :281)"""
        num_of_candi = 3
        candidates = [i / 8 for i in range(1, 16) if i != 8]
        qmin, qmax = self._bounds()
        per_cand = []
        for st in candidates:
            x_float = ops.fq_affine_fwd(x, (self.delta * st).detach(), self.zero_point.detach(), qmin, qmax)
            err = (x_float - x).abs().pow(2.4)
            if self.RUN_CHANNEL_WISE:
                err = err.sum(dim=0) if self.isFC else err.sum(dim=(0, -1, -2))
            else:
                err = (err.sum(dim=(-1, -2)) if self.isFC else err).flatten()
            per_cand.append(err)
        table = torch.stack(per_cand, dim=0)
        scores = {i: 0 for i in range(table.shape[0])}
        order = torch.argsort(table, dim=0)[:num_of_candi].cpu()
        for col in range(table.shape[1]):
            for j in range(num_of_candi):
                scores[int(order[j, col])] += num_of_candi - j
        top = [k for k, _v in sorted(scores.items(), key=lambda kv: kv[1], reverse=True)][:num_of_candi - 1]
        self.shiftTarget = [candidates[i] for i in top] + [1.0]

    @torch.no_grad()
    def init_v_beta(self, x: torch.Tensor):
        """integer-floor candidate per shift, alpha init, beta init on the argmax-selected step size
        (channelQuant.py:279-294)"""
        print(f"{self.name}, Optimal shift candidates: ", self.shiftTarget)
        for st in self.shiftTarget:
            self.shiftedScale = st
            self.x_q.append(torch.floor(x / (self.delta * self.shiftedScale)))
        self.shiftedScale = 1.0
        self._remember_source(x)
        self.alpha = self.init_alpha(x, clip=(0.90 - 0.05 * len(self.shiftTarget)), device=self.device)
        delta = self.get_delta()
        beta = ops.adaround_init_alpha(x.detach(), delta.detach().contiguous())
        self.alpha = nn.Parameter(self.alpha)
        self.beta = nn.Parameter(beta)

    @torch.no_grad()
    def update_delta(self):
        self.delta = self.get_delta()

    @torch.no_grad()
    def init_beta(self, x: torch.Tensor):
        self.beta = nn.Parameter(ops.adaround_init_alpha(x.detach(), self.delta.detach().contiguous()))
