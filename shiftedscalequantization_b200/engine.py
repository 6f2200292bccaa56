"""ReconEngine — the reconstruction hot loop (reference: quant/block_recon.py:89-105, quant/layer_recon.py:79-96)
re-designed for B200:

  * all trainable state of a unit lives in ONE flat device buffer (every AdaRound alpha of the block, or every
    activation step size), with the nn.Parameters the API exposes as views into it — one multi-tensor
    fake-quant launch, one fused Adam launch, one all-reduce bucket per iteration;
  * the mini-batch indices (the reference's CPU `torch.randperm(N)[:B]` stream), the temperature b and the
    learning rate are precomputed tables read through a device-side step counter, so the whole iteration —
    gather, fake-quant, cuDNN convs, loss, backward, Adam — is captured once in a CUDA graph and replayed;
  * the reconstruction loss reads the cached FP outputs in place through the index (no gathered copy), writes
    d(loss)/d(pred) in the same pass, and the regulariser and its gradient ride inside the fake-quant kernels.

Per-iteration launches of ours (weight phase, one GPU): iter_prologue (schedules + mini-batch gather + all layers' soft
weights + regulariser), recon_loss (+ dpred), adaround_bwd_adam_mt (alpha gradients + Adam + end of iteration) = 3, against
~530 ATen dispatches upstream. With several GPUs the last one splits into adaround_bwd_mt, the gradient
exchange and adam_step_end_iteration.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import dist as ssq_dist
from . import ops

PULL_CTAS = int(os.environ.get("SSQ_PULL_CTAS", "16"))     # CTAs of the host-pull kernels (scratch/pull_probe.py: 16 reach the link's rate)


def temperature(t: int, t_max: int, rel_start_decay: float, start_b, end_b):
    """LinearTempDecay.__call__ (quant/block_recon.py:193-202), host doubles"""
    start_decay = rel_start_decay * t_max
    if t < start_decay:
        return start_b
    rel_t = (t - start_decay) / (t_max - start_decay)
    return end_b + (start_b - end_b) * max(0.0, (1 - rel_t))


def brecq_b_table(iters: int, warmup: float, b_range, round_loss: bool) -> torch.Tensor:
    """b used at iteration i (0-based) by LossFunction: count = i+1 is incremented BEFORE use
    (block_recon.py:153); b = 0 (regulariser off) while count < iters*warmup (:167)."""
    out = torch.zeros(max(iters, 1), dtype=torch.float32)
    if not round_loss:
        return out
    loss_start = iters * warmup
    for i in range(iters):
        count = i + 1
        if count >= loss_start:
            out[i] = float(temperature(count, iters, warmup, b_range[0], b_range[1]))
    return out


def index_table(n_samples: int, batch: int, iters: int) -> torch.Tensor:
    """the exact CPU RNG call sequence of the reference loop (block_recon.py:90), consumed up front"""
    tab = torch.empty((max(iters, 1), batch), dtype=torch.int64)
    for i in range(iters):
        tab[i] = torch.randperm(n_samples)[:batch]
    return tab


def cosine_lr_table(lr: float, iters: int) -> torch.Tensor:
    """learning rates Adam sees under CosineAnnealingLR(T_max=iters, eta_min=0) (block_recon.py:72-73),
    produced by the real scheduler so the recursive form's rounding is reproduced"""
    p = nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=lr)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=max(iters, 1), eta_min=0.)
    out = torch.empty(max(iters, 1), dtype=torch.float32)
    for i in range(iters):
        out[i] = opt.param_groups[0]['lr']
        opt.step()
        sch.step()
    return out


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class ReconEngine:
    """Runs `iters` reconstruction iterations of one unit (QuantModule or BaseQuantBlock).

    weight phase (act_quant=False): optimises the AdaRound alpha of every QuantModule in the unit.
    activation phase (act_quant=True): optimises the activation step sizes (LSQ) with frozen hard-rounded weights.
    """

    def __init__(self, unit: nn.Module, modules: Sequence[nn.Module], cached_inps: torch.Tensor,
                 cached_outs: torch.Tensor, cached_grads: Optional[torch.Tensor] = None, *,
                 act_quant: bool, iters: int, weight: float, b_range=(20, 2), warmup: float = 0.0, p: float = 2.0,
                 lr: float = 4e-5, opt_mode: str = 'mse', batch_size: int = 32, multi_gpu: bool = False,
                 act_quantizers: Sequence[nn.Module] = (), use_graph: Optional[bool] = None,
                 idx_table: Optional[torch.Tensor] = None, verbose: bool = True,
                 host_resident: bool = False, device: Optional[torch.device] = None, host_stage: str = 'pull',
                 scaling: str = 'weak', host_pack: Optional[bool] = None, fold_output_affine: bool = False):
        """host_resident=True keeps the cached features in (pinned) host memory, as the reference does with
        keep_gpu=False (quant/data_utils.py:34-36, `cached_inps[idx].to(device)` at block_recon.py:91-92): every
        step moves its mini-batch rows host->device. host_stage='pull' (default): a kernel inside the captured
        iteration reads the NEXT mini-batch's rows out of the mapped pinned cache while the current iteration
        computes (no host work per step); 'dma': one cudaMemcpyAsync per row on a copy stream, issued by the host.
        host_pack (pull mode; default on): cached tensors whose rows are a multiple of 1024 elements and at most 80 % non-zero —
        post-ReLU features — are kept zero-packed on the host (non-zero values only; bit mask and chunk offsets on the device, 3 %
        of the dense size) and expanded by the pulling kernel, so fewer bytes cross PCIe; lossless, same trajectory.
        scaling (multi_gpu only): 'weak' = every rank draws its own mini-batch of `batch_size` from its shard and the
        gradients are SUMMED (the reference's link.allreduce semantics, block_recon.py:100-102; global batch = R x batch);
        'strong' = the ranks split ONE global mini-batch of `batch_size` (every rank holds the whole cache and the same
        index table, rank r takes columns [r*b, (r+1)*b)); the loss gradient is scaled by 1/R and the regulariser counted
        once, so the summed gradient equals the single-GPU gradient up to summation order and the trajectory is the
        1-GPU trajectory.
        fold_output_affine (weight phase; README --bias_cal): every layer's output-channel scale gamma^z (alpha_out) and
        offset varphi^z (beta_out) is learned with the alphas, FOLDED into the weight launch: W_eff = gamma_oc W_q,
        b_eff = gamma b + varphi (quant_layer.py:258-259 up to fp32 rounding; no activation-sized pass), their gradients
        from the weight / bias gradients of the folded layer (ssq_affine_grad_mt)."""
        self.unit, self.modules = unit, list(modules)
        self.fold_affine = bool(fold_output_affine) and not act_quant
        self.sym = None
        self.host_resident = bool(host_resident)
        if scaling not in ('weak', 'strong'):
            raise ValueError('scaling must be "weak" or "strong"')
        if host_stage not in ('pull', 'dma'):
            raise ValueError('host_stage must be "pull" or "dma"')
        self.host_pull = self.host_resident and host_stage == 'pull'
        self.host_dma = self.host_resident and host_stage == 'dma'
        self.host_pack = self.host_pull and (True if host_pack is None else bool(host_pack))
        self.dev = torch.device(device) if device is not None else cached_inps.device
        if self.dev.type != 'cuda':
            raise ops._lib.SsqError('reconstruction runs on CUDA only (no CPU fallback)')
        if self.host_resident:
            pin = lambda t: t if (t is None or t.is_pinned()) else t.contiguous().pin_memory()
            cached_inps, cached_outs, cached_grads = pin(cached_inps.cpu()), pin(cached_outs.cpu()), \
                pin(None if cached_grads is None else cached_grads.cpu())
        self.cached_inps = cached_inps.contiguous()
        self.cached_outs = cached_outs.contiguous()
        self.cached_grads = None if cached_grads is None else cached_grads.contiguous()
        self.act_quant, self.iters, self.weight, self.p, self.opt_mode = act_quant, int(iters), float(weight), float(p), opt_mode
        self.batch = min(int(batch_size), self.cached_inps.shape[0])
        self.multi_gpu = bool(multi_gpu) and ssq_dist.world_size() > 1 and ssq_dist.EXCHANGE != 'none'   # 'none': measurement aid
        self.strong = self.multi_gpu and scaling == 'strong'
        self.use_graph = (self.iters >= 8) if use_graph is None else bool(use_graph)
        self.verbose = verbose
        self.launches_per_iter = 0
        self.graph = None
        self.keep_grad = False          # True: the fused backward also writes the alpha gradients to gflat (inspection)
        n = self.cached_inps.shape[0]
        # ---- device-side schedules -------------------------------------------------------------------
        tab = idx_table if idx_table is not None else index_table(n, self.batch, self.iters)
        self.grad_scale, self.reg_share = None, 1.0
        if self.strong:
            world, rk = ssq_dist.world_size(), ssq_dist.rank()
            if self.batch % world:
                raise ValueError(f'strong scaling needs batch_size ({self.batch}) divisible by the world size ({world})')
            b = self.batch // world
            tab = tab[:, rk * b:(rk + 1) * b].contiguous()          # this rank's slice of the global mini-batch
            self.batch = b
            self.reg_share = 1.0 / world
            self.grad_scale = torch.full((1,), 1.0 / world, device=self.dev)
        self.idx_table = tab.to(self.dev)
        self.idx_table_host = tab.cpu()
        self.host_step = 0
        self._copy_stream = None
        self._loss_ring = None
        self.b_table = brecq_b_table(self.iters, warmup, b_range, round_loss=not act_quant).to(self.dev)
        lr_tab = cosine_lr_table(lr, self.iters) if act_quant else torch.full((max(self.iters, 1),), 1e-3)
        self.lr_table = lr_tab.to(self.dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.idx_live = torch.zeros(self.batch, dtype=torch.int64, device=self.dev)
        self.b_live = torch.zeros(1, device=self.dev)
        self.lr_live = torch.zeros(1, device=self.dev)
        self.state = ops.IterationState(self.step_dev, self.idx_table, self.idx_live, self.b_table, self.b_live,
                                        self.lr_table, self.lr_live, max(self.iters, 1))
        self.cur_inp = torch.empty((self.batch,) + tuple(self.cached_inps.shape[1:]), device=self.dev)
        self.cur_out = torch.empty((self.batch,) + tuple(self.cached_outs.shape[1:]), device=self.dev) if self.host_resident else None
        self.cur_grad = torch.empty((self.batch,) + tuple(self.cached_grads.shape[1:]), device=self.dev) \
            if (self.host_resident and self.cached_grads is not None) else None
        self.loss_dev = torch.zeros(1, device=self.dev)
        self.reg_dev = torch.zeros(1, device=self.dev)
        if self.host_pull:
            self._pull_stream = torch.cuda.Stream(self.dev)
            self._pull_bufs = [(self._maybe_pack(src), torch.empty_like(cur), cur) for src, cur in
                               ((self.cached_inps, self.cur_inp), (self.cached_outs, self.cur_out), (self.cached_grads, self.cur_grad))
                               if src is not None]
            self._pull_primed = False
        # ---- freeze everything that is not optimised (no wasted wgrad / bias-grad kernels) ---------------
        self._frozen = [(q, q.requires_grad) for q in unit.parameters()]
        for q, _ in self._frozen:
            q.requires_grad_(False)
        try:
            if act_quant:
                self._setup_act_phase(list(act_quantizers))
            else:
                self._setup_weight_phase()
            # the peer-memory exchange keeps Adam's moments for this rank's shard only
            n_state = self.sym.shard if getattr(self, 'sym', None) is not None else self.flat.numel()
            self.exp_avg = torch.zeros(n_state, device=self.dev)
            self.exp_avg_sq = torch.zeros(n_state, device=self.dev)
        except BaseException:
            self.close()                 # a failed set-up must not leave the unit frozen or holding engine weights
            raise

    # ------------------------------------------------------------------------------------------ setup
    def _setup_weight_phase(self):
        qs = [m.weight_quantizer for m in self.modules]
        sizes = [_pad4(q.alpha.numel()) for q in qs]
        n_alpha = sum(sizes)
        ocs = [_pad4(m.weight.shape[0]) for m in self.modules] if self.fold_affine else []
        total = n_alpha + 2 * sum(ocs)                  # [alphas | gammas | varphis]
        # several GPUs: parameter and gradient buffers in symmetric memory, exchanged by one peer-memory kernel
        self.sym = ssq_dist.symmetric_unit_or_none(total, self.dev) if self.multi_gpu else None
        if self.sym is not None:
            self.flat, self.gflat = self.sym.flat, self.sym.gflat
        else:
            self.flat = torch.zeros(total, device=self.dev)
            self.gflat = torch.zeros_like(self.flat)
        self.n_alpha = n_alpha
        entries, self.wq_leaves, self.bias_leaves, off = [], [], [], 0
        g_off, p_off = n_alpha, n_alpha + sum(ocs)
        for i, (m, q, sz) in enumerate(zip(self.modules, qs, sizes)):
            n = q.alpha.numel()
            view = self.flat[off:off + n].view(q.alpha.shape)
            view.copy_(q.alpha.detach())
            q.alpha = nn.Parameter(view, requires_grad=False)      # API-visible parameter = view of the flat buffer
            wq = torch.empty_like(m.weight, memory_format=torch.contiguous_format).requires_grad_(True)
            self.wq_leaves.append(wq)
            m._engine_weight = wq
            e = dict(w=m.weight.detach(), alpha=view, delta=q.delta.detach(),
                     zero_point=ops.match_param(q.zero_point.detach(), q.delta.detach()), wq=wq.detach(),
                     galpha=self.gflat[off:off + n].view(q.alpha.shape),
                     qmin=0.0, qmax=float(q.n_levels - 1))
            if self.fold_affine:
                oc = m.weight.shape[0]
                gam, phi = self.flat[g_off:g_off + oc], self.flat[p_off:p_off + oc]
                gam.copy_(m.alpha_out.detach().reshape(-1)); phi.copy_(m.beta_out.detach().reshape(-1))
                m.alpha_out = nn.Parameter(gam.view(m.alpha_out.shape), requires_grad=False)
                m.beta_out = nn.Parameter(phi.view(m.beta_out.shape), requires_grad=False)
                beff = torch.empty(oc, device=self.dev).requires_grad_(True)
                self.bias_leaves.append(beff)
                m._engine_bias = beff
                e.update(gamma=gam, phi=phi, bias=None if m.bias is None else m.bias.detach().contiguous(), beff=beff.detach(),
                         ggamma=self.gflat[g_off:g_off + oc], gphi=self.gflat[p_off:p_off + oc])
                g_off += ocs[i]; p_off += ocs[i]
            entries.append(e)
            off += sz
        self.table = ops.AdaRoundTable(entries)

    def _setup_act_phase(self, act_quantizers: List[nn.Module]):
        self.act_quantizers = [q for q in act_quantizers if q.delta is not None]
        k = max(len(self.act_quantizers), 1)
        self.flat = torch.zeros(_pad4(k), device=self.dev)
        self.gflat = torch.zeros_like(self.flat)
        self.delta_params = []
        for i, q in enumerate(self.act_quantizers):
            view = self.flat[i:i + 1].view(q.delta.shape)
            view.copy_(q.delta.detach())
            q.delta = nn.Parameter(view, requires_grad=True)
            self.delta_params.append(q.delta)
        with torch.no_grad():                                        # weights are constants in this phase
            for m in self.modules:
                m._engine_weight = m.weight_quantizer(m.weight).detach()

    # ------------------------------------------------------------------------------------------ one iteration
    def _iteration(self):
        """one iteration; *step_dev (iterations completed) is read by every kernel and incremented by the last one"""
        table = None if self.act_quant else self.table
        reg = None if self.act_quant else self.reg_dev
        if not self.host_resident:
            ops.iter_prologue(self.state, self.cached_inps, self.cur_inp, table, self.weight, reg)
        else:
            ops.iter_prologue(self.state, None, None, table, self.weight, reg)
            if self.host_pull:
                # the rows pulled during the previous iteration become current (HBM->HBM); then fork: the SMs pull the
                # NEXT mini-batch (row *step_dev + 1 of the table) over PCIe beside this iteration's kernels
                main = torch.cuda.current_stream(self.dev)
                for _src, stage, cur in self._pull_bufs:
                    cur.copy_(stage, non_blocking=True)
                self._pull_stream.wait_stream(main)
                self._pull(1)
        with torch.enable_grad():
            out = self.unit(self.cur_inp)
        if self.host_resident:
            loss, dpred = ops.recon_loss(out.detach(), self.cur_out, self.p, self.opt_mode, fisher=self.cur_grad,
                                         gscale=self.grad_scale)
        else:
            loss, dpred = ops.recon_loss(out.detach(), self.cached_outs, self.p, self.opt_mode, fisher=self.cached_grads,
                                         tgt_index=self.idx_live, gscale=self.grad_scale)
        self.loss_dev = loss
        if self.act_quant:
            # a unit whose output does not depend on any step size (e.g. the head after
            # disable_network_output_quantization) yields zero gradients: Adam then leaves the deltas untouched
            grads = torch.autograd.grad([out], self.delta_params, [dpred.view_as(out)], allow_unused=True) \
                if out.requires_grad else [None] * len(self.delta_params)
            if self.delta_params:
                packed = torch.stack([torch.zeros((), device=self.dev) if g is None else g.reshape(()) for g in grads])
                self.gflat[:packed.numel()].copy_(packed)
        else:
            grads = torch.autograd.grad([out], self.wq_leaves + self.bias_leaves, [dpred.view_as(out)])
            gwqs = grads[:len(self.wq_leaves)]
            if self.fold_affine:                         # d/d gamma, d/d varphi: needs the alphas the forward used
                self.table.affine_grad(gwqs, grads[len(self.wq_leaves):])
            if not self.multi_gpu:
                na = self.n_alpha
                self.table.backward_adam(gwqs, self.b_live, self.weight * self.reg_share, self.flat, self.exp_avg,
                                         self.exp_avg_sq, self.lr_live, self.step_dev, store_grad=self.keep_grad)
                if self.fold_affine:                     # second parameter group; the counter already reads t (1-based)
                    ops.adam_step(self.flat[na:], self.gflat[na:], self.exp_avg[na:], self.exp_avg_sq[na:], self.lr_live, self.step_dev)
            elif self.sym is not None or self.fold_affine:
                self.table.backward_adam(gwqs, self.b_live, self.weight * self.reg_share, self.flat, None, None,
                                         self.lr_live, self.step_dev, apply_adam=False)
            else:
                self.table.backward(gwqs, self.b_live, self.weight * self.reg_share)
        if self.multi_gpu and getattr(self, 'sym', None) is not None:
            # SUM over the ranks (link.allreduce, block_recon.py:100-102) + Adam + hand-out of the new alphas: one kernel
            ops.grad_exchange_adam(self.sym, self.exp_avg, self.exp_avg_sq, self.lr_live, self.step_dev)
        elif self.act_quant or self.multi_gpu:
            if self.multi_gpu:
                ssq_dist.all_reduce_sum_(self.gflat)                 # NCCL path (activation step sizes; SSQ_EXCHANGE=nccl)
            ops.adam_step_end_iteration(self.flat, self.gflat, self.exp_avg, self.exp_avg_sq, self.lr_live, self.step_dev)
        if self.host_pull:
            torch.cuda.current_stream(self.dev).wait_stream(self._pull_stream)      # join the prefetch branch

    def _maybe_pack(self, src: torch.Tensor):
        """zero-packed form of a pinned cache tensor when that saves PCIe bytes (post-ReLU features), else the tensor itself"""
        if not (self.host_pack and ops.packable(src)):
            return src
        packed = ops.pack_rows_sparse(src, self.dev)
        return packed if packed.density <= 0.8 else src      # the packed stream reaches 0.87 of the dense pull's PCIe rate

    def _pull(self, lookahead: int):
        for src, stage, _cur in self._pull_bufs:
            if isinstance(src, ops.PackedRows):
                ops.pull_rows_host_packed(src, self.idx_table, self.step_dev, lookahead, self.idx_table.shape[0], stage,
                                          max_ctas=PULL_CTAS, stream=self._pull_stream)
            else:
                ops.pull_rows_host(src, self.idx_table, self.step_dev, lookahead, self.idx_table.shape[0], stage,
                                   max_ctas=PULL_CTAS, stream=self._pull_stream)

    def _prime_pull(self):
        """first mini-batch of a run: pull row *step_dev (not yet advanced) before the first iteration"""
        main = torch.cuda.current_stream(self.dev)
        self._pull_stream.wait_stream(main)
        self._pull(0)
        main.wait_stream(self._pull_stream)
        self._pull_primed = True

    # ------------------------------------------------------------------------------------------ graph capture
    def _snapshot(self):
        return (self.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone())

    def _restore(self, snap):
        self.flat.copy_(snap[0]); self.exp_avg.copy_(snap[1]); self.exp_avg_sq.copy_(snap[2]); self.step_dev.copy_(snap[3])
        if self.host_pull:
            self._pull_primed = False                            # the staged rows belong to another step now

    def capture(self, warm: int = 3):
        """warm up eagerly on a side stream (allocations, cuDNN plans, workspaces), roll the state back,
        then capture one iteration"""
        snap = self._snapshot()
        if self.host_dma:
            hs = self.host_step
            self._stage_batch_from_host()
            torch.cuda.synchronize(self.dev)
            self.host_step = hs; self._copy_stream = None        # restart the prefetch pipeline at the right step
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            if self.host_pull:
                self._prime_pull()
            for _ in range(warm):
                self._iteration()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        self._restore(snap)
        before = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        self.launches_per_iter = ops.launch_count() - before
        self._restore(snap)

    # ------------------------------------------------------------------------------------------ driver
    def _issue_stage(self, step_index: int, slot: int):
        """H2D of mini-batch `step_index` into staging slot `slot` on the copy stream: one cudaMemcpyAsync per row,
        straight from the pinned cache (no host gather; the batch-memcpy driver entry points are deliberately unused)"""
        rows = self.idx_table_host[min(step_index, self.idx_table_host.shape[0] - 1)]
        cs = self._copy_stream
        cs.wait_event(self._slot_free[slot])                 # the D2D out of this slot must have finished
        ops.stage_rows_h2d(self.cached_inps, rows, self._stage_inp[slot], cs)
        ops.stage_rows_h2d(self.cached_outs, rows, self._stage_out[slot], cs)
        if self.cur_grad is not None:
            ops.stage_rows_h2d(self.cached_grads, rows, self._stage_grad[slot], cs)
        self._slot_ready[slot].record(cs)

    def _stage_batch_from_host(self, prefetch: bool = True):
        """make the staged mini-batch current (device-to-device); with prefetch=True also start the transfer of the
        next one. step() defers the prefetch until the captured iteration has been launched, so the host time spent
        issuing the row copies and the PCIe transfer of step i+1 both overlap the iteration of step i."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.dev)
            mk = lambda t: [torch.empty_like(t) for _ in range(2)]
            self._stage_inp, self._stage_out = mk(self.cur_inp), mk(self.cur_out)
            self._stage_grad = mk(self.cur_grad) if self.cur_grad is not None else None
            self._slot_ready = [torch.cuda.Event() for _ in range(2)]
            self._slot_free = [torch.cuda.Event() for _ in range(2)]
            for e in self._slot_free:
                e.record(torch.cuda.current_stream(self.dev))
            self._issue_stage(self.host_step, self.host_step % 2)
        slot = self.host_step % 2
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(self._slot_ready[slot])
        self.cur_inp.copy_(self._stage_inp[slot], non_blocking=True)
        self.cur_out.copy_(self._stage_out[slot], non_blocking=True)
        if self.cur_grad is not None:
            self.cur_grad.copy_(self._stage_grad[slot], non_blocking=True)
        self._slot_free[slot].record(main)
        self.host_step += 1
        if prefetch:
            self._prefetch_next()

    def _prefetch_next(self):
        self._issue_stage(self.host_step, self.host_step % 2)

    def h2d_bytes_per_step(self) -> int:
        if not self.host_resident:
            return 0
        if self.host_pull:           # packed tensors move their non-zero values only (average over the cache's rows)
            return int(sum(round(self.batch * src.host_bytes_per_row()) if isinstance(src, ops.PackedRows) else 4 * cur.numel()
                           for src, _stage, cur in self._pull_bufs))
        n = self.cur_inp.numel() + self.cur_out.numel() + (0 if self.cur_grad is None else self.cur_grad.numel())
        return 4 * n

    def step(self):
        if self.host_dma:
            self._stage_batch_from_host(prefetch=False)
        elif self.host_pull and not self._pull_primed:
            self._prime_pull()
        if self.graph is not None:
            self.graph.replay()
        else:
            before = ops.launch_count()
            self._iteration()
            self.launches_per_iter = ops.launch_count() - before
        if self.host_dma:
            self._prefetch_next()            # issued after the iteration's launch: host + PCIe time hide behind it
        if self._loss_ring is not None:
            k = self._loss_k
            self._loss_ring[k].copy_(self.loss_dev, non_blocking=True)
            self._loss_evt[k].record(torch.cuda.current_stream(self.dev))
            self._loss_k = k ^ 1

    def enable_loss_readback(self):
        """device->host read of every iteration's loss through a 2-deep pinned ring: read_loss() returns the loss of
        the iteration before the one just launched, so the read never drains the device queue"""
        self._loss_ring = [torch.zeros(1).pin_memory() for _ in range(2)]
        self._loss_evt = [torch.cuda.Event() for _ in range(2)]
        self._loss_k, self._loss_seen = 0, [False, False]

    def read_loss(self, latest: bool = False):
        """loss of the previous step() (latest=False, lagged by one launch) or of the last one (latest=True, waits)"""
        k = (self._loss_k ^ 1) if latest else self._loss_k
        if not self._loss_evt[k].query():
            self._loss_evt[k].synchronize()
        return float(self._loss_ring[k])

    def run(self):
        if self.iters <= 0:
            return
        if self.use_graph and self.graph is None:
            self.capture()
        self._t0, self._t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._t0.record()
        for i in range(self.iters):
            self.step()
            count = i + 1
            if self.verbose and count % 500 == 0:
                rec, rnd = float(self.loss_dev), float(self.reg_dev)
                print('Total loss:\t{:.3f} (rec:{:.3f}, round:{:.3f})\tb={:.2f}\tcount={}'.format(
                    rec + rnd, rec, rnd, float(self.b_live), count))
        self._t1.record()

    def loop_ms(self) -> float:
        """device time of the last run()'s iteration loop (capture excluded); synchronises"""
        self._t1.synchronize()
        return self._t0.elapsed_time(self._t1)

    def close(self):
        if getattr(self, 'sym', None) is not None and self.graph is not None:
            self.sym.check()
        for m in self.modules:
            m._engine_weight = None
            m._engine_bias = None
            if self.fold_affine:
                m._affine_key = None                     # gamma / varphi moved: re-derive "is identity"
                for name in ('alpha_out', 'beta_out'):
                    getattr(m, name).requires_grad_(True)
        for q, flag in self._frozen:
            q.requires_grad_(flag)
        if self.act_quant:
            for q in getattr(self, 'act_quantizers', ()):
                q.delta.requires_grad_(True)
        else:
            for m in self.modules:
                a = getattr(m.weight_quantizer, 'alpha', None)
                if a is not None:
                    a.requires_grad_(True)
        self.graph = None


class AutogradReconEngine:
    """The same captured-iteration design for the loops whose forward is NOT the plain AdaRound multi-tensor launch:
    the shifted-scale loops (reference: quant/layer_recon_shiftedScale.py:52-95,282-318,
    quant/layer_recon_fused_shiftedScale.py:84-110), their activation phases, and reconstruction with the output-channel
    affine gamma^z / varphi^z trainable (README `--bias_cal`).

    The unit's forward runs through the quantisers' own autograd functions (K1a/K1b/K1c kernels); what changes against
    the eager loop is everything around it: every trainable tensor is a view of ONE flat buffer (`slots` = the
    (owner, attribute) pairs holding the nn.Parameters), the index / temperature / lr schedules are device tables read
    through device step counters, the regulariser gates are device scalars (b <= 0 = off), the loss is the fused
    K3 kernel reading its targets in place, Adam is one fused launch, and the whole iteration replays as one CUDA graph —
    no host synchronisation per iteration (upstream: two .item() calls, a CPU randperm + index H2D, ~20 optimizer launches).

    reg_fn(live) -> sequence of 0-dim tensors built from `live` (the device temperature scalars, one per entry of
    `b_tables`); their sum is the rounding / group regulariser of the loss.
    """

    def __init__(self, unit: nn.Module, slots, cached_inps: torch.Tensor, cached_outs: torch.Tensor, *, iters: int,
                 batch_size: int = 32, p: float = 2.0, lr_table: torch.Tensor, b_tables: Sequence[torch.Tensor] = (),
                 reg_fn=None, multi_gpu: bool = False, idx_table: Optional[torch.Tensor] = None,
                 use_graph: Optional[bool] = None, adam_betas=(0.9, 0.999), adam_eps: float = 1e-8):
        self.unit, self.iters, self.p = unit, int(iters), float(p)
        self.dev = cached_inps.device
        if self.dev.type != 'cuda':
            raise ops._lib.SsqError('reconstruction runs on CUDA only (no CPU fallback)')
        self.cached_inps, self.cached_outs = cached_inps.contiguous(), cached_outs.contiguous()
        n = self.cached_inps.shape[0]
        self.batch = min(int(batch_size), n)
        self.multi_gpu = bool(multi_gpu) and ssq_dist.world_size() > 1
        self.use_graph = (self.iters >= 8) if use_graph is None else bool(use_graph)
        self.reg_fn, self.betas, self.eps = reg_fn, adam_betas, adam_eps
        self.n_steps = max(self.iters, 1)
        tab = idx_table if idx_table is not None else index_table(n, self.batch, self.iters)
        self.idx_table = tab.to(self.dev)
        self.lr_table = lr_table.to(self.dev, torch.float32).contiguous()
        self.b_tables = [t.to(self.dev, torch.float32).contiguous() for t in b_tables]
        for t in [self.lr_table] + self.b_tables:
            if t.numel() < self.n_steps:
                raise ValueError('schedule table shorter than the loop')
        # one device step counter per temperature table: ssq_loop_advance copies row `step` and increments
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.extra_steps = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in self.b_tables[1:]]
        self.idx_live = torch.zeros(self.batch, dtype=torch.int64, device=self.dev)
        self.lr_live = torch.zeros(1, device=self.dev)
        self.live = [torch.zeros(1, device=self.dev) for _ in self.b_tables]
        self.cur_inp = torch.empty((self.batch,) + tuple(self.cached_inps.shape[1:]), device=self.dev)
        self.loss_dev = torch.zeros(1, device=self.dev)
        self.reg_vals: List[torch.Tensor] = []
        # ---- flat parameter buffer; the API-visible nn.Parameters become views of it.
        # A tensor listed k times is stepped k times per iteration with its own step count, which is what
        # torch.optim.Adam's per-tensor loop does with a duplicated entry (upstream lists every QuantModule's activation
        # step size twice in the act phase: layer_recon_shiftedScale.py:24-33 walks named_modules() and meets the
        # module's act_quantizer again as a UniformAffineQuantizer).
        self.slots = list(slots)
        uniq, mult = [], {}
        for owner, attr in self.slots:
            old = getattr(owner, attr)
            if id(old) in mult:
                mult[id(old)] += 1
            else:
                mult[id(old)] = 1
                uniq.append((owner, attr, old))
        uniq.sort(key=lambda e: mult[id(e[2])] > 1)              # singles first (stable), repeated ones behind them
        sizes = [_pad4(old.numel()) for _o, _a, old in uniq]
        # several GPUs, no tensor listed twice (the weight phases): parameters and gradients in symmetric memory, SUM over the
        # ranks + Adam + hand-out of the new parameters as ONE peer-memory kernel, as in ReconEngine; otherwise NCCL all-reduce
        self.sym = ssq_dist.symmetric_unit_or_none(sum(sizes), self.dev) \
            if (self.multi_gpu and sum(sizes) > 0 and all(v == 1 for v in mult.values())) else None
        if self.sym is not None:
            self.flat, self.gflat = self.sym.flat, self.sym.gflat
            self.xchg_step = torch.zeros(1, dtype=torch.int64, device=self.dev)      # iterations completed (the kernel steps with t = this + 1)
        else:
            self.flat = torch.zeros(sum(sizes), device=self.dev)
            self.gflat = torch.zeros_like(self.flat)
        self.params, self.gviews, self.repeats, off = [], [], [], 0
        self.n_single = 0
        for (owner, attr, old), sz in zip(uniq, sizes):
            k = old.numel()
            view = self.flat[off:off + k].view(old.shape)
            view.copy_(old.detach())
            par = nn.Parameter(view, requires_grad=True)
            for o2, a2 in self.slots:                            # every slot that held this tensor now holds the view
                if getattr(o2, a2) is old:
                    setattr(o2, a2, par)
            self.params.append(par)
            self.gviews.append(self.gflat[off:off + k].view(old.shape))
            if mult[id(old)] > 1:
                self.repeats.append((off, sz, mult[id(old)], torch.zeros(1, dtype=torch.int64, device=self.dev)))
            else:
                self.n_single = off + sz
            off += sz
        n_state = self.sym.shard if self.sym is not None else self.flat.numel()       # the exchange kernel steps this rank's shard only
        self.exp_avg = torch.zeros(n_state, device=self.dev)
        self.exp_avg_sq = torch.zeros(n_state, device=self.dev)
        mine = {id(q) for q in self.params}
        self._frozen = [(q, q.requires_grad) for q in unit.parameters() if id(q) not in mine]
        for q, _ in self._frozen:
            q.requires_grad_(False)
        self.graph = None
        self.launches_per_iter = 0

    # ------------------------------------------------------------------------------------------ one iteration
    def _iteration(self):
        ops.loop_advance(self.step_dev, self.idx_table, self.idx_live, self.b_tables[0] if self.b_tables else None,
                         self.live[0] if self.live else None, self.lr_table, self.lr_live, self.n_steps)
        for st, tab, lv in zip(self.extra_steps, self.b_tables[1:], self.live[1:]):
            ops.loop_advance(st, None, None, tab, lv, None, None, self.n_steps)
        ops.gather_rows(self.cached_inps, self.idx_live, out=self.cur_inp)
        with torch.enable_grad():
            out = self.unit(self.cur_inp)
            regs = list(self.reg_fn(self.live)) if self.reg_fn is not None else []
        loss, dpred = ops.recon_loss(out.detach(), self.cached_outs, self.p, 'mse', tgt_index=self.idx_live)
        self.loss_dev, self.reg_vals = loss, [r.detach() for r in regs]
        heads, seeds = [], []
        if out.requires_grad:
            heads.append(out); seeds.append(dpred.view_as(out))
        for r in regs:
            if r.requires_grad:
                heads.append(r); seeds.append(torch.ones_like(r))
        grads = torch.autograd.grad(heads, self.params, seeds, allow_unused=True) if heads else [None] * len(self.params)
        for view, g in zip(self.gviews, grads):
            if g is None:
                view.zero_()
            else:
                view.copy_(g.view_as(view))
        if self.sym is not None:
            ops.grad_exchange_adam(self.sym, self.exp_avg, self.exp_avg_sq, self.lr_live, self.xchg_step, self.betas, self.eps)
            return
        if self.multi_gpu:
            ssq_dist.all_reduce_sum_(self.gflat)
        n1 = self.n_single
        ops.adam_step(self.flat[:n1], self.gflat[:n1], self.exp_avg[:n1], self.exp_avg_sq[:n1], self.lr_live, self.step_dev,
                      self.betas, self.eps)
        for off, sz, k, ctr in self.repeats:
            for _ in range(k):
                ops.loop_advance(ctr, None, None, None, None, None, None, 1 << 62)
                ops.adam_step(self.flat[off:off + sz], self.gflat[off:off + sz], self.exp_avg[off:off + sz],
                              self.exp_avg_sq[off:off + sz], self.lr_live, ctr, self.betas, self.eps)

    def _state(self):
        return [self.flat, self.exp_avg, self.exp_avg_sq, self.step_dev] + self.extra_steps + [r[3] for r in self.repeats] + \
            ([self.xchg_step] if self.sym is not None else [])

    def capture(self, warm: int = 3):
        snap = [t.clone() for t in self._state()]
        restore = lambda: [t.copy_(s) for t, s in zip(self._state(), snap)]
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(warm):
                self._iteration()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        restore()
        before = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        self.launches_per_iter = ops.launch_count() - before
        restore()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            before = ops.launch_count()
            self._iteration()
            self.launches_per_iter = ops.launch_count() - before

    def run(self, every: int = 500, on_report=None, report_offset: int = 0):
        """runs the loop; on_report(i) is called after iteration i whenever (i + report_offset) % every == 0 (the
        reference's read-out cadences: i % 500 == 0 in the shifted loops, count = i + 1 in LossFunction) — the only
        points where the host looks at device values"""
        if self.iters <= 0:
            return
        if self.use_graph and self.graph is None:
            self.capture()
        self._t0, self._t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._t0.record()
        for i in range(self.iters):
            self.step()
            if on_report is not None and (i + report_offset) % every == 0:
                on_report(i)
        self._t1.record()

    def loop_ms(self) -> float:
        """device time of the last run()'s iteration loop (capture excluded); synchronises"""
        self._t1.synchronize()
        return self._t0.elapsed_time(self._t1)

    def close(self):
        for q, flag in self._frozen:
            q.requires_grad_(flag)
        had_graph, self.graph = self.graph is not None, None
        if getattr(self, 'sym', None) is not None and had_graph:
            self.sym.check()                  # a rendezvous of the exchange kernel that timed out invalidates the run
