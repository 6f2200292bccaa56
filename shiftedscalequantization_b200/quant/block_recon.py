"""block_reconstruction / LossFunction / LinearTempDecay — mirror of the reference's quant/block_recon.py.

The public function keeps the upstream signature and side effects (quantiser swap, quant-state toggles,
hard rounding at the end); the 20 000-iteration loop itself runs in ..engine.ReconEngine.
LossFunction stays usable on its own (callers such as notebooks build it directly); it evaluates the same
three fused kernels — reconstruction loss, rounding regulariser — under autograd.
"""
import torch

from .. import ops
from ..engine import AutogradReconEngine, ReconEngine, brecq_b_table, temperature
from .adaptive_rounding import AdaRoundQuantizer
from .data_utils import save_grad_data, save_inp_oup_data
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule, StraightThrough, lp_loss
from .quant_model import QuantModel


LAST_RUN_STATS = {}      # {'iters', 'loop_ms', 'launches_per_iter', 'capture_s', 'capture_mode'} of the most recent call (bench.py reads it)


def _quant_modules(unit):
    return [m for _n, m in unit.named_modules() if isinstance(m, QuantModule)]


def _run_with_output_affine(unit, modules, cached_inps, cached_outs, *, iters, weight, b_range, warmup, p, batch_size, multi_gpu):
    """README `--bias_cal`: the output-channel scale gamma^z (alpha_out) and offset varphi^z (beta_out) of every layer in
    the unit are learned together with the AdaRound alphas (upstream's commented `opt_params +=` lines,
    quant/layer_recon_fused_shiftedScale.py:67-68; forward at quant/quant_layer.py:258-259), same Adam (lr 1e-3), same loss.
    The forward runs through the quantisers' autograd functions + the fused per-channel affine kernel (gradients of
    gamma/varphi = per-channel reductions over N,H,W), captured as one graph per iteration."""
    slots = []
    for m in modules:
        slots += [(m.weight_quantizer, 'alpha'), (m, 'alpha_out'), (m, 'beta_out')]
        m.train_output_affine = True
    quantizers = [m.weight_quantizer for m in modules]
    reg_fn = lambda live: [sum(ops.RoundReg.apply(q.alpha, live[0], weight) for q in quantizers)]
    eng = AutogradReconEngine(unit, slots, cached_inps, cached_outs, iters=iters, batch_size=batch_size, p=p,
                              lr_table=torch.full((max(iters, 1),), 1e-3),
                              b_tables=[brecq_b_table(iters, warmup, b_range, round_loss=True)], reg_fn=reg_fn,
                              multi_gpu=multi_gpu)

    def report(i):
        count = i + 1
        rec, rnd = float(eng.loss_dev), float(eng.reg_vals[0])
        print('Total loss:\t{:.3f} (rec:{:.3f}, round:{:.3f})\tb={:.2f}\tcount={}'.format(rec + rnd, rec, rnd, float(eng.live[0]), count))

    try:
        eng.run(every=500, on_report=report, report_offset=1)      # count % 500 == 0, as upstream's LossFunction (:176)
        LAST_RUN_STATS.update(iters=iters, loop_ms=eng.loop_ms(), launches_per_iter=eng.launches_per_iter)
    finally:
        eng.close()
        for m in modules:
            m.train_output_affine = False
            m._affine_key = None            # the parameters were stepped through the flat buffer: re-derive "is identity"


def reconstruct_unit(model, unit, cali_data, *, is_block, batch_size, iters, weight, opt_mode, asym, include_act_func,
                     b_range, warmup, act_quant, lr, p, multi_gpu, eval, bias_cal=False, scaling='weak', host_resident=False):
    """shared body of block_reconstruction (block_recon.py:10-116) and layer_reconstruction (layer_recon.py:10-104)"""
    if eval:
        iters = 0          # structure-only call: swap the quantisers, learn nothing (block_recon.py:36-37)
    model.set_quant_state(False, False)
    unit.set_quant_state(True, act_quant)
    round_mode = 'learned_hard_sigmoid'
    if not include_act_func:
        org_act_func = unit.activation_function
        unit.activation_function = StraightThrough()
    modules = _quant_modules(unit)
    act_quantizers = []
    if not act_quant:
        for m in modules:
            m.weight_quantizer = AdaRoundQuantizer(uaq=m.weight_quantizer, round_mode=round_mode,
                                                   weight_tensor=m.org_weight.data)
            m.weight_quantizer.soft_targets = True
    else:
        if is_block and hasattr(unit.act_quantizer, 'delta'):
            act_quantizers.append(unit.act_quantizer)
        act_quantizers += [m.act_quantizer for m in modules if m.act_quantizer.delta is not None]

    if not eval:
        # host_resident = upstream's keep_gpu=False: the caches never exist on the device as a whole (data_utils.py:34-36)
        keep_gpu = not host_resident
        device = next(model.parameters()).device
        import time
        from .data_utils import capture_stats
        torch.cuda.synchronize(device); t_cap = time.perf_counter()
        cached_inps, cached_outs = save_inp_oup_data(model, unit, cali_data, asym, act_quant, batch_size, keep_gpu=keep_gpu)
        torch.cuda.synchronize(device)
        LAST_RUN_STATS.update(capture_s=time.perf_counter() - t_cap, capture_mode=capture_stats(model).last_mode)
        cached_grads = save_grad_data(model, unit, cali_data, act_quant, batch_size=batch_size, keep_gpu=keep_gpu) \
            if opt_mode != 'mse' else None
    if iters > 0:
        foldable = all(m.weight_quantizer.delta.numel() == m.weight.shape[0] and m.se_module is None for m in modules) \
            if (bias_cal and not act_quant) else False
        if bias_cal and not act_quant and (bias_cal == 'exact' or not foldable):
            # the reference expression itself, out*alpha_out + beta_out on the activation (two roundings), under autograd
            if opt_mode != 'mse':
                raise NotImplementedError('bias_cal is defined for the mse reconstruction loss')
            _run_with_output_affine(unit, modules, cached_inps.to(device), cached_outs.to(device), iters=iters, weight=weight, b_range=b_range,
                                    warmup=warmup, p=p, batch_size=batch_size, multi_gpu=multi_gpu)
        else:
            engine = ReconEngine(unit, modules, cached_inps, cached_outs, cached_grads, act_quant=act_quant, iters=iters,
                                 weight=weight, b_range=b_range, warmup=warmup, p=p, lr=lr, opt_mode=opt_mode,
                                 batch_size=batch_size, multi_gpu=multi_gpu, act_quantizers=act_quantizers, scaling=scaling,
                                 host_resident=host_resident, device=device, fold_output_affine=bool(bias_cal) and not act_quant)
            try:
                engine.run()
                LAST_RUN_STATS.update(iters=iters, loop_ms=engine.loop_ms(), launches_per_iter=engine.launches_per_iter)
            finally:
                engine.close()
    if not eval:
        del cached_inps, cached_outs, cached_grads
        torch.cuda.empty_cache()

    for m in modules:                       # finish: hard rounding (block_recon.py:110-112, layer_recon.py:100)
        m.weight_quantizer.soft_targets = False
    if not include_act_func:
        unit.activation_function = org_act_func


def block_reconstruction(model: QuantModel, block: BaseQuantBlock, cali_data: torch.Tensor,
                         batch_size: int = 32, iters: int = 20000, weight: float = 0.01, opt_mode: str = 'mse',
                         asym: bool = False, include_act_func: bool = True, b_range: tuple = (20, 2),
                         warmup: float = 0.0, act_quant: bool = False, lr: float = 4e-5, p: float = 2.0,
                         multi_gpu: bool = False, eval: bool = False, bias_cal: bool = False, scaling: str = 'weak',
                         host_resident: bool = False):
    """Optimise the rounding (or, with act_quant, the activation step sizes) of every layer in `block` so the
    block output matches the FP block output on the calibration data. Arguments as upstream; `bias_cal` (README flag,
    keyword-only in practice) additionally learns every layer's output-channel scale/offset in the weight phase — True: folded
    into the weight launch of the 3-launch engine (W_eff = gamma W_q, b_eff = gamma b + varphi; equal to the reference
    expression up to fp32 rounding), 'exact': the reference's out*alpha_out+beta_out on the activation under autograd;
    `scaling` ('weak' | 'strong') picks the multi-GPU partitioning (engine.ReconEngine); `host_resident=True` keeps the cached
    features in pinned host memory (upstream's keep_gpu=False, quant/data_utils.py:34-36) and pulls every mini-batch over PCIe."""
    reconstruct_unit(model, block, cali_data, is_block=True, batch_size=batch_size, iters=iters, weight=weight,
                     opt_mode=opt_mode, asym=asym, include_act_func=include_act_func, b_range=b_range, warmup=warmup,
                     act_quant=act_quant, lr=lr, p=p, multi_gpu=multi_gpu, eval=eval, bias_cal=bias_cal, scaling=scaling,
                     host_resident=host_resident)


class LossFunction:
    """rec_loss + rounding regulariser (block_recon.py:119-182). `count` is incremented before use."""

    def __init__(self, block, round_loss: str = 'relaxation', weight: float = 1., rec_loss: str = 'mse',
                 max_count: int = 2000, b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0,
                 p: float = 2.):
        self.block = block
        self.layer = block
        self.round_loss = round_loss
        self.weight = weight
        self.rec_loss = rec_loss
        self.loss_start = max_count * warmup
        self.p = p
        self.temp_decay = LinearTempDecay(max_count, rel_start_decay=warmup + (1 - warmup) * decay_start,
                                          start_b=b_range[0], end_b=b_range[1])
        self.count = 0

    def __call__(self, pred, tgt, grad=None):
        self.count += 1
        if self.rec_loss == 'mse':
            rec_loss = lp_loss(pred, tgt, p=self.p)
        elif self.rec_loss in ('fisher_diag', 'fisher_full'):
            rec_loss = ops.ReconLoss.apply(pred, tgt, 2.0, self.rec_loss, grad)
        else:
            raise ValueError('Not supported reconstruction loss function: {}'.format(self.rec_loss))
        b = self.temp_decay(self.count)
        if self.count < self.loss_start or self.round_loss == 'none':
            b = round_loss = 0
        elif self.round_loss == 'relaxation':
            round_loss = 0
            b_dev = ops.scalar_dev(b, pred.device)
            for m in _quant_modules(self.block):
                round_loss = round_loss + ops.RoundReg.apply(m.weight_quantizer.alpha, b_dev, self.weight)
        else:
            raise NotImplementedError
        total_loss = rec_loss + round_loss
        if self.count % 500 == 0:
            print('Total loss:\t{:.3f} (rec:{:.3f}, round:{:.3f})\tb={:.2f}\tcount={}'.format(
                float(total_loss), float(rec_loss), float(round_loss), b, self.count))
        return total_loss


class LinearTempDecay:
    """temperature b(t): start_b until rel_start_decay*t_max, then linear to end_b (block_recon.py:185-202)"""

    def __init__(self, t_max: int, rel_start_decay: float = 0.2, start_b: int = 10, end_b: int = 2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b = start_b
        self.end_b = end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, (1 - rel_t))
