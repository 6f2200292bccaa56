"""Shifted-scale reconstruction loops — mirror of the reference's quant/layer_recon_shiftedScale.py
(block_recon_shiftedScale :12-124, layer_recon_shiftedScale :262-338, ScaleLossBlockFunction :340-411,
ScaleLossFunction :413-486, LinearTempDecayShift :488-505).

Two ways to use them, as upstream: shift first (learn which shifted step size each input channel takes, entropy
regulariser) and/or AdaRound on top (`adaround=True`, pow regulariser on h(beta)). They read the unit's
`cached_inp_features` / `cached_out_features` lists filled by the 'if'/'of' cache modes.
Differences that change no value: the loss objects keep their scalars on the device and convert on access
(upstream calls .item() twice per iteration), and the layer variant runs on the model's device instead of a
hard-coded 'cuda'.
"""
import torch
from tqdm import tqdm

from .. import ops
from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule, UniformAffineQuantizer, lp_loss


class LinearTempDecayShift:
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


class _LazyScalars:
    """loss bookkeeping without a host sync per iteration"""
    _total = _rec = _round = 0.0

    @staticmethod
    def _val(v):
        return float(v) if torch.is_tensor(v) else v

    total_loss = property(lambda s: s._val(s._total))
    rec_loss = property(lambda s: s._val(s._rec))
    round_loss_val = property(lambda s: s._val(s._round))


class ScaleLossBlockFunction(_LazyScalars):
    """lp reconstruction loss + (entropy of the group probabilities | pow regulariser on h(beta)) summed over the
    unit's QuantModules. `count` is incremented AFTER use here (upstream :407), unlike block_recon.LossFunction."""

    def __init__(self, block, round_loss: str = 'relaxation', lmda: float = 1., max_count: int = 2000,
                 b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0, p: float = 2.0,
                 adaround: bool = False):
        self.block = block
        self.round_loss = round_loss
        self.lmda = lmda
        self.loss_start = max_count * warmup
        self.itr = max_count
        self.p = p
        self.b = 0
        self.temp_decay = LinearTempDecayShift(max_count, rel_start_decay=warmup + (1 - warmup) * decay_start,
                                               start_b=b_range[0], end_b=b_range[1])
        self.adaround = adaround
        self.count = 0

    def _quantizers(self):
        if isinstance(self.block, QuantModule):
            return [self.block.weight_quantizer]
        return [m.weight_quantizer for _n, m in self.block.named_modules() if isinstance(m, QuantModule)]

    def __call__(self, pred, tgt, grad=None):
        rec_loss = lp_loss(pred, tgt, p=self.p)
        b = self.temp_decay(self.count)
        if self.count < self.loss_start or self.round_loss == 'none':
            b = round_loss = 0
        elif self.round_loss == 'relaxation':
            round_loss = 0
            b_dev = ops.scalar_dev(b, pred.device) if self.adaround else None
            for q in self._quantizers():
                if self.adaround:
                    round_loss = round_loss + ops.RoundReg.apply(q.beta, b_dev, self.lmda)
                else:
                    round_loss = round_loss + ops.ShiftProbsReg.apply(q.alpha, 0, None, self.lmda)
        else:
            raise NotImplementedError
        total_loss = rec_loss + round_loss
        self._total, self._rec, self._round = total_loss.detach(), rec_loss.detach(), \
            (round_loss.detach() if torch.is_tensor(round_loss) else round_loss)
        self.b = b
        self.count += 1
        return total_loss

    def report(self):
        return 'Total loss:\t{:.6f} (rec:{:.6f}, round:{:.6f})\tb={:.2f}'.format(
            float(self.total_loss), float(self.rec_loss), float(self.round_loss_val), self.b)


class ScaleLossFunction(ScaleLossBlockFunction):
    """single-layer flavour (upstream duplicates the class; here `layer` is simply the unit)"""

    def __init__(self, layer, round_loss: str = 'relaxation', lmda: float = 1., max_count: int = 2000,
                 b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0, p: float = 2.0,
                 adaround: bool = False):
        super().__init__(layer, round_loss, lmda, max_count, b_range, decay_start, warmup, p, adaround)
        self.layer = layer


def _prepare_weight_params(modules, adaround):
    """quantiser state + the (owner, attribute) slots of the tensors Adam steps (upstream :37-50 / :270-279)"""
    slots = []
    for m in modules:
        q = m.weight_quantizer
        w = m.org_weight.data
        if adaround:
            if q.opt_mode == 'learned_hard_sigmoid':        # shift was learned first: freeze its choice into delta
                q.update_delta()
            q.init_beta(x=w.clone().detach())
            q.opt_mode = 'adaround'
            slots.append((q, 'beta'))
        else:
            q.init_v(x=w.clone().detach())
            slots.append((q, 'alpha'))
    return slots


def _act_delta_params(unit):
    """upstream :24-33. named_modules() yields every QuantModule and then its act_quantizer again as a
    UniformAffineQuantizer, so those step sizes are listed — and stepped by Adam — twice per iteration; kept."""
    slots = []
    for _n, m in unit.named_modules():
        if isinstance(m, QuantModule):
            if not m.act_quantizer.disable_act_quant:
                slots.append((m.act_quantizer, 'delta'))
        elif isinstance(m, UniformAffineQuantizer):
            if not m.disable_act_quant:
                slots.append((m, 'delta'))
    return slots


USE_CAPTURED_LOOP = True      # False: the eager autograd loop below (kept as the behavioural cross-check)
MULTI_GPU = False             # True under torchrun: SUM all-reduce of the flat gradient per iteration (upstream's loops are single-GPU)
LAST_LOOP_STATS = {}          # {'iters', 'loop_ms', 'captured', 'launches_per_iter'} of the most recent loop (bench.py reads it)


def _run_captured(unit, loss_func, slots, lr_table, b_tables, reg_fn, cached_inp, cached_out, iters, batch_size, describe,
                  on_regs):
    """the loop body of upstream :52-95 as one replayed CUDA graph per iteration (engine.AutogradReconEngine): same
    index stream (torch.randperm(N)[:B] per iteration), same temperature/lr per iteration, same kernels as the eager
    path; the host reads device scalars only at the i % 500 == 0 read-outs."""
    from ..engine import AutogradReconEngine
    from .. import dist as ssq_dist
    eng = AutogradReconEngine(unit, slots, cached_inp, cached_out, iters=iters, batch_size=batch_size, p=loss_func.p,
                              lr_table=lr_table, b_tables=b_tables, reg_fn=reg_fn, multi_gpu=MULTI_GPU and ssq_dist.world_size() > 1)
    state = {'start': 0.0}
    bar = tqdm(total=iters, desc='', dynamic_ncols=True)

    def sync_loss_object():
        loss_func._rec = eng.loss_dev.detach().reshape(()).clone()      # clones: the originals live in the graph's pool
        vals = [v.clone() for v in eng.reg_vals]
        on_regs(vals)
        reg_sum = sum(vals) if vals else 0.0
        loss_func._total = loss_func._rec + reg_sum
        loss_func.b = float(eng.live[0]) if eng.live else 0

    def report(i):
        sync_loss_object()
        state['start'] = max(state['start'], loss_func.rec_loss)
        bar.update(min(500, iters - i))
        bar.set_description(describe(state['start'], loss_func))

    eng.run(every=500, on_report=report)
    if iters > 0:
        sync_loss_object()
    loss_func.count += iters
    bar.close()
    if iters > 0:
        LAST_LOOP_STATS.update(iters=iters, loop_ms=eng.loop_ms(), captured=True, launches_per_iter=eng.launches_per_iter)
    eng.close()
    return state['start'], eng


def _run_loop(unit, loss_func, optimizer, scheduler, cached_inp, cached_out, iters, batch_size, describe):
    start_loss = 0.0
    bar = tqdm(range(iters), desc='', dynamic_ncols=True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in bar:
        perm = torch.randperm(cached_inp.size(0))[:batch_size]
        cur_inp, cur_out = cached_inp[perm], cached_out[perm]
        optimizer.zero_grad()
        err = loss_func(unit(cur_inp), cur_out)
        err.backward()
        optimizer.step()
        if scheduler is not None:
            scheduler.step()
        if i % 500 == 0:
            start_loss = max(start_loss, loss_func.rec_loss)
            bar.set_description(describe(start_loss, loss_func))
    t1.record()
    if iters > 0:
        t1.synchronize()
        LAST_LOOP_STATS.update(iters=iters, loop_ms=t0.elapsed_time(t1), captured=False, launches_per_iter=None)
    return start_loss


def _probe(unit, loss_func, optimizer, cached_inp, cached_out, batch_size):
    """reconstruction loss on the first batch (the 'Soft Round' / 'Hard Round' read-outs, upstream :97-121)"""
    optimizer.zero_grad()
    with torch.no_grad():                   # a read-out: upstream builds (and drops) an autograd graph here
        loss_func(unit(cached_inp[:batch_size]), cached_out[:batch_size])
    return loss_func.rec_loss


def shifted_b_table(loss_func, iters, temp_decay=None, gate_only=False):
    """temperature seen by iteration i of the shifted losses: count is used BEFORE its increment (upstream :378,:407);
    0 while count < loss_start or round_loss == 'none' = regulariser off. gate_only: 1.0 where on (entropy has no b)."""
    tab = torch.zeros(max(iters, 1), dtype=torch.float32)
    if loss_func.round_loss == 'none':
        return tab
    decay = temp_decay or loss_func.temp_decay
    for i in range(iters):
        if i >= loss_func.loss_start:
            tab[i] = 1.0 if gate_only else float(decay(i))
    return tab


def _shifted_recon(unit, modules, iters, lmda, model, act, adaround, loss_cls, train_target, device, loss_off=None):
    """act: optimise the activation step sizes (block variant, upstream :21-35); loss_off: build the loss with
    round_loss='none' (defaults to `act`; the layer variant passes act=False, loss_off=<its act flag>: upstream :283
    switches the regulariser off for act=True but still optimises the weight parameters, :270-279)"""
    from ..engine import cosine_lr_table
    loss_off = act if loss_off is None else bool(loss_off)
    warmup, p, b_range, lr, batch_size = 0.2, 2.0, (20, 2), 4e-4, 32
    scheduler = None
    if act:
        slots = _act_delta_params(unit)
    else:
        slots = _prepare_weight_params(modules, adaround)
        print("number of elements in opt_params: {}".format(sum(getattr(o, a).numel() for o, a in slots)))
    loss_func = loss_cls(unit, round_loss='none' if loss_off else 'relaxation', lmda=lmda, max_count=iters, b_range=b_range,
                         decay_start=0, warmup=warmup, p=p, adaround=adaround)
    cached_inp = torch.cat(unit.cached_inp_features).to(device)
    cached_out = torch.cat(unit.cached_out_features).to(device)
    describe = lambda s0, lf: f"{s0:.6f} -> {lf.rec_loss:.6f} {lf.round_loss_val:.3f} "
    if USE_CAPTURED_LOOP and iters >= 8:
        lr_table = cosine_lr_table(lr, iters) if act else torch.full((max(iters, 1),), 1e-3)
        b_tables = [shifted_b_table(loss_func, iters, gate_only=not adaround)]
        quantizers = loss_func._quantizers() if not loss_off else []

        def reg_fn(live):
            if loss_off or not quantizers:
                return []
            if adaround:
                return [sum(ops.RoundReg.apply(q.beta, live[0], lmda) for q in quantizers)]
            return [sum(ops.ShiftProbsReg.apply(q.alpha, 0, live[0], lmda) for q in quantizers)]

        def on_regs(vals):
            loss_func._round = vals[0].reshape(()) if vals else 0

        start_loss, _eng = _run_captured(unit, loss_func, slots, lr_table, b_tables, reg_fn, cached_inp, cached_out,
                                         iters, batch_size, describe, on_regs)
        optimizer = torch.optim.Adam([getattr(o, a) for o, a in slots], lr=lr if act else 1e-3)
    else:
        opt_params = [getattr(o, a) for o, a in slots]
        # foreach=False: the per-tensor loop, which is how the reference's verified torch 1.11 (and the CPU path) steps a
        # duplicated entry — two sequential updates
        optimizer = torch.optim.Adam(opt_params, lr=lr, foreach=False) if act else torch.optim.Adam(opt_params)
        if act:
            scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=iters, eta_min=0.)
        start_loss = _run_loop(unit, loss_func, optimizer, scheduler, cached_inp, cached_out, iters, batch_size, describe)
    out = [_probe(unit, loss_func, optimizer, cached_inp, cached_out, batch_size)]
    print(f"Soft Round : {start_loss:.6f} -> {loss_func.rec_loss:.6f} {loss_func.round_loss_val:.3f}")
    return out, loss_func, optimizer, cached_inp, cached_out, start_loss, batch_size


def block_recon_shiftedScale(block: BaseQuantBlock, iters: int = 20000, lmda: float = 1., model=None, test_loader=None,
                             act=False, adaround=False, useShiftedScale=True):
    """Learn the shifted-scale choice (or, with adaround, the rounding) of every layer in `block`; returns
    [soft reconstruction loss, hard reconstruction loss] on the first cached batch."""
    block.train()
    device = next(model.parameters()).device
    modules = [m for _n, m in block.named_modules() if isinstance(m, QuantModule)]
    out, lf, opt, ci, co, s0, bs = _shifted_recon(block, modules, iters, lmda, model, act, adaround,
                                                  ScaleLossBlockFunction, None, device)
    if not act:
        for m in modules:
            if adaround:
                m.weight_quantizer.hard_round = True
                m.weight_quantizer(m.weight)
            else:
                m.weight_quantizer.hard_targets = True
                m.weight_quantizer.shiftedDone = True
    out.append(_probe(block, lf, opt, ci, co, bs))
    print(f"Hard Round : {s0:.6f} -> {lf.rec_loss:.6f} {lf.round_loss_val:.3f}")
    torch.cuda.empty_cache()
    model.eval()
    return out


def layer_recon_shiftedScale(layer: QuantModule, iters: int = 20000, lmda: float = 1., model=None, test_loader=None,
                             act=False, adaround=False, useShiftedScale=True):
    """Single-layer variant. As upstream, the weight parameters are optimised even when act=True is passed (act only
    switches the regulariser off for the whole loop, :283), and with adaround the hard switch is written to
    `layer.hard_round` (not the quantiser's) — quirks kept (:325-326). Pinned by tests/golden/layer_shift.npz."""
    model.train()
    device = next(model.parameters()).device
    out, lf, opt, ci, co, s0, bs = _shifted_recon(layer, [layer], iters, lmda, model, False, adaround,
                                                  ScaleLossFunction, None, device, loss_off=bool(act))
    if adaround:
        layer.hard_round = True
    else:
        layer.weight_quantizer.hard_targets = True
        layer.weight_quantizer.shiftedDone = True
    out.append(_probe(layer, lf, opt, ci, co, bs))
    print(f"Hard Round : {s0:.6f} -> {lf.rec_loss:.6f} {lf.round_loss_val:.3f}")
    torch.cuda.empty_cache()
    model.eval()
    return out
