"""Fused shift + rounding reconstruction — mirror of the reference's quant/layer_recon_fused_shiftedScale.py
(print_ratio :13-21, block_recon_fused_shiftedScale :23-141, layer_recon_fused_shiftedScale :144-221,
FusedScaleLossFunction :223-309, FusedLinearTempDecayShift :382-399).

Every QuantModule runs ChannelQuant in 'adaShift' mode: integer floors of the S shifted scales mixed by the group
probabilities, plus the AdaRound soft round h(beta). Adam steps `alpha` only (beta is initialised but, as upstream,
not handed to the optimiser). Two regularisers: lambda_R on h(beta) with temperature b, lambda_S on the group
probabilities with a temperature on a 3/4-length schedule.
"""
import numpy as np
import torch

from .. import ops
from .channelQuantAct import ChannelQuantAct
from . import layer_recon_shiftedScale as _ls
from .layer_recon_shiftedScale import _LazyScalars, _probe, _run_loop, _run_captured, shifted_b_table
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule, UniformAffineQuantizer, lp_loss


def print_ratio(quantizers):
    """share of groups that picked each shift"""
    for qt in quantizers:
        soft_target = qt.get_sig_soft_targets().detach().cpu().numpy()
        values, counts = np.unique(np.argmax(soft_target, axis=-1), return_counts=True)
        total = np.sum(counts)
        dump = ' '.join(f'{k}:{v:.3f}' for k, v in zip(values, counts / total))
        print(f'{qt.name}[{total}] : {dump}')


class FusedLinearTempDecayShift:
    def __init__(self, t_max: int, rel_start_decay: float = 0.2, start_b: int = 10, end_b: int = 2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b = start_b
        self.end_b = end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay) if self.t_max != 0 else 1
        return self.end_b + (self.start_b - self.end_b) * max(0.0, (1 - rel_t))


class FusedScaleLossFunction(_LazyScalars):
    def __init__(self, block, quantizer, round_loss: str = 'relaxation', lmda: list = [1., 1.], max_count: int = 2000,
                 b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0, p: float = 2.0,
                 adaround: bool = False):
        self.block = block
        self.quantizer = quantizer
        self.round_loss = round_loss
        self.lmdaR, self.lmdaS = lmda[0], lmda[1]
        self.loss_start = max_count * warmup
        self.itr = max_count
        self.p = p
        self.b = 0
        rel = warmup + (1 - warmup) * decay_start
        self.temp_decay = FusedLinearTempDecayShift(max_count, rel_start_decay=rel, start_b=b_range[0], end_b=b_range[1])
        self.temp_decay_shift = FusedLinearTempDecayShift(max_count * 3 / 4, rel_start_decay=rel, start_b=b_range[0], end_b=b_range[1])
        self.count = 0
        self._r = self._s = 0

    @property
    def round_loss_val(self):
        return f'R:{float(self._r):.3f} S:{float(self._s):.3f}'

    def __call__(self, pred, tgt, grad=None):
        rec_loss = lp_loss(pred, tgt, p=self.p)
        round_lossR = round_lossS = 0
        b = self.temp_decay(self.count)
        b2 = self.temp_decay_shift(self.count)
        if self.count < self.loss_start or self.round_loss == 'none':
            b = b2 = 0
        elif self.round_loss == 'relaxation':
            b_dev, b2_dev = ops.scalar_dev(b, pred.device), ops.scalar_dev(b2, pred.device)
            for qt in self.quantizer:
                round_lossR = round_lossR + ops.RoundReg.apply(qt.beta, b_dev, self.lmdaR)
                round_lossS = round_lossS + ops.ShiftProbsReg.apply(qt.alpha, 1, b2_dev, self.lmdaS)
        else:
            raise NotImplementedError
        total_loss = rec_loss + round_lossR + round_lossS
        self._total, self._rec = total_loss.detach(), rec_loss.detach()
        self._r = round_lossR.detach() if torch.is_tensor(round_lossR) else round_lossR
        self._s = round_lossS.detach() if torch.is_tensor(round_lossS) else round_lossS
        self.b = b
        self.count += 1
        return total_loss

    def report(self):
        return 'Total loss:\t{:.6f} (rec:{:.6f}, round:{})\tb={:.2f}'.format(
            float(self.total_loss), float(self.rec_loss), self.round_loss_val, self.b)


def _fused_captured(unit, loss_func, quantizers, lr, cached_inp, cached_out, iters, batch_size, describe, extra_slots=()):
    """captured-graph version of the loop at upstream :84-110: temperature b for h(beta), b2 (3/4-length schedule) for the
    group probabilities, both gated off during warm-up through their device tables"""
    lr_table = torch.full((max(iters, 1),), float(lr))
    b_tables = [shifted_b_table(loss_func, iters, loss_func.temp_decay), shifted_b_table(loss_func, iters, loss_func.temp_decay_shift)]
    lmdaR, lmdaS = loss_func.lmdaR, loss_func.lmdaS

    def reg_fn(live):
        if loss_func.round_loss != 'relaxation':
            return []
        return [sum(ops.RoundReg.apply(q.beta, live[0], lmdaR) for q in quantizers),
                sum(ops.ShiftProbsReg.apply(q.alpha, 1, live[1], lmdaS) for q in quantizers)]

    def on_regs(vals):
        if vals:
            loss_func._r, loss_func._s = vals[0].reshape(()), vals[1].reshape(())

    slots = [(q, 'alpha') for q in quantizers] + list(extra_slots)
    start_loss, _eng = _run_captured(unit, loss_func, slots, lr_table, b_tables, reg_fn, cached_inp, cached_out, iters,
                                     batch_size, describe, on_regs)
    return start_loss


def block_recon_fused_shiftedScale(block: BaseQuantBlock, iters: int = 20000, lmda: list = [1., 1.], model=None,
                                   test_loader=None, act=False, adaround=False, useShiftedScale=True, bias_cal=False):
    """bias_cal (README flag; upstream's commented `opt_params += [module.alpha_out] / [module.beta_out]`, :67-68): the
    output-channel scale/offset of every layer joins the optimiser."""
    block.train()
    warmup, p, b_range, lr, batch_size = 0.2, 2.0, (20, 2), 0.001, 32
    device = next(model.parameters()).device
    quantizers, opt_params = [], []
    modules = [m for _n, m in block.named_modules() if isinstance(m, QuantModule)]
    if act:
        # upstream swaps in ChannelQuantAct and calls its init_v, which cannot run (channelQuantAct.py:126-134)
        for m in modules:
            if m.act_quantizer.disable_act_quant:
                continue
            m.act_quantizer = ChannelQuantAct(uaq=m.act_quantizer, shiftTarget=[2 / 2, 1 / 2])
            m.act_quantizer.init_v()
    else:
        for m in modules:
            q = m.weight_quantizer
            q.init_v_beta(x=m.org_weight.data.clone().detach())
            opt_params.append(q.alpha)
            quantizers.append(q)
            q.opt_mode = 'adaShift'
    extra_slots = []
    if bias_cal and not act:
        for m in modules:
            m.train_output_affine = True
            extra_slots += [(m, 'alpha_out'), (m, 'beta_out')]
            opt_params += [m.alpha_out, m.beta_out]
    print("number of elements in opt_params: {}".format(sum(q.numel() for q in opt_params)))
    loss_func = FusedScaleLossFunction(block, quantizers, round_loss='none' if act else 'relaxation', lmda=lmda,
                                       max_count=iters, b_range=b_range, decay_start=0, warmup=warmup, p=p)
    cached_inp = torch.cat(block.cached_inp_features).to(device)
    cached_out = torch.cat(block.cached_out_features).to(device)
    describe = lambda s0, lf: f"{s0:.6f} -> {lf.rec_loss:.6f} {lf.round_loss_val} "
    if _ls.USE_CAPTURED_LOOP and iters >= 8 and opt_params:
        start_loss = _fused_captured(block, loss_func, quantizers, lr, cached_inp, cached_out, iters, batch_size, describe,
                                     extra_slots)
        optimizer = torch.optim.Adam([q.alpha for q in quantizers] + [getattr(o, a) for o, a in extra_slots], lr=lr)
    else:
        optimizer = torch.optim.Adam(opt_params, lr=lr)
        start_loss = _run_loop(block, loss_func, optimizer, None, cached_inp, cached_out, iters, batch_size, describe)
    out = [_probe(block, loss_func, optimizer, cached_inp, cached_out, batch_size)]
    print(f"Soft Round : {start_loss:.6f} -> {loss_func.rec_loss:.6f} {loss_func.round_loss_val}")
    if not act:
        for m in modules:
            m.weight_quantizer.hard_round = True
            m.weight_quantizer.hard_targets = True
            m.weight_quantizer.shiftedDone = True
    out.append(_probe(block, loss_func, optimizer, cached_inp, cached_out, batch_size))
    print(f"Hard Round : {start_loss:.6f} -> {loss_func.rec_loss:.6f} {loss_func.round_loss_val}")
    print_ratio(quantizers)
    for m in modules:
        if getattr(m, 'train_output_affine', False):
            m.train_output_affine = False
            m._affine_key = None
    torch.cuda.empty_cache()
    model.eval()
    return out


def layer_recon_fused_shiftedScale(layer: QuantModule, iters: int = 20000, lmda: list = [1., 1.], model=None,
                                   test_loader=None, act=False, adaround=False, useShiftedScale=True):
    """Upstream's version raises UnboundLocalError on its first statement that touches `opt_params`
    (layer_recon_fused_shiftedScale.py:156) and so has no defined behaviour; this is the single-layer analogue of the
    block function with the settings that function body spells out (Adam default lr, p = 1.0)."""
    model.train()
    warmup, b_range, batch_size, p = 0.2, (20, 2), 32, 1.0
    device = next(model.parameters()).device
    q = layer.weight_quantizer
    q.init_v_beta(x=layer.org_weight.data.clone().detach())
    opt_params = [q.alpha]
    q.opt_mode = 'adaShift'
    loss_func = FusedScaleLossFunction(layer, [q], round_loss='none' if act else 'relaxation', lmda=lmda, max_count=iters,
                                       b_range=b_range, decay_start=0, warmup=warmup, p=p, adaround=adaround)
    cached_inp = torch.cat(layer.cached_inp_features).to(device)
    cached_out = torch.cat(layer.cached_out_features).to(device)
    print("number of elements in opt_params: {}".format(sum(t.numel() for t in opt_params)))
    describe = lambda s0, lf: f"{s0:.6f} -> {lf.rec_loss:.6f} {lf.round_loss_val} "
    if _ls.USE_CAPTURED_LOOP and iters >= 8:
        start_loss = _fused_captured(layer, loss_func, [q], 1e-3, cached_inp, cached_out, iters, batch_size, describe)
        optimizer = torch.optim.Adam([q.alpha])
    else:
        optimizer = torch.optim.Adam(opt_params)
        start_loss = _run_loop(layer, loss_func, optimizer, None, cached_inp, cached_out, iters, batch_size, describe)
    out = [_probe(layer, loss_func, optimizer, cached_inp, cached_out, batch_size)]
    print(f"Soft Round : {start_loss:.6f} -> {loss_func.rec_loss:.6f} {loss_func.round_loss_val}")
    if adaround:
        layer.hard_round = True
    else:
        q.hard_targets = True
        q.shiftedDone = True
    out.append(_probe(layer, loss_func, optimizer, cached_inp, cached_out, batch_size))
    print(f"Hard Round : {start_loss:.6f} -> {loss_func.rec_loss:.6f} {loss_func.round_loss_val}")
    print_ratio([q])
    torch.cuda.empty_cache()
    model.eval()
    return out
