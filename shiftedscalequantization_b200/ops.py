"""Tensor-level entry points over the C-ABI, and the autograd Functions the quantiser modules use.

PyTorch is plumbing here: it owns device memory and the stream; every arithmetic result comes from
libssq_b200.so. All tensors must be CUDA fp32; anything else raises (there is no CPU path).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import ctypes as C

import torch

from . import _lib
from ._lib import AdaRoundDesc, AffineDesc, IterState, MT_MAX, MT_TILE, SHIFT_ADASHIFT, SHIFT_DEQUANT

# --------------------------------------------------------------------------------------- plumbing
_launch_count = 0          # kernels launched through this module (bench.py reports it)
_ws_cache: dict = {}


def launch_count() -> int:
    return _launch_count


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _req(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor, got {type(t)}")
    if not t.is_cuda:
        raise _lib.SsqError(f"{name}: tensor is on {t.device}; the ssq kernels are CUDA-only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise _lib.SsqError(f"{name}: dtype {t.dtype} not supported (fp32 only, as the reference)")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_ws_retired = []


def workspace(device: torch.device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Zero-initialised scratch for the deterministic reductions (tickets reset themselves).
    One buffer per (device, tag); callers on the same stream may share it."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)      # a captured graph may still point at it: never hand it back to the allocator
        buf = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _ws(t: torch.Tensor, nchan: int = 1, tag: str = "default") -> torch.Tensor:
    return workspace(t.device, _lib.load().ssq_ws_bytes(int(nchan)), tag)


_prof = None               # when a list: (name, start_event, end_event) per launch, for bench.py's kernel shares


def profile_begin() -> None:
    global _prof
    _prof = []


def profile_end() -> dict:
    """{entry point: (launches, total ms)} measured with CUDA events on the launching stream"""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in rec:
        n, ms = out.get(name, (0, 0.0))
        out[name] = (n + 1, ms + e0.elapsed_time(e1))
    return out


def _call(name: str, *args) -> None:
    global _launch_count
    if _prof is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(getattr(_lib.load(), name)(*args), name)
        e1.record()
        _prof.append((name, e0, e1))
    else:
        _lib.check(getattr(_lib.load(), name)(*args), name)
    _launch_count += 1


def channel_layout(x: torch.Tensor, delta: torch.Tensor) -> Tuple[int, int]:
    """(inner, nchan) of the header's channel layout. delta is a scalar (per-tensor), [C,1,...] along
    dim 0 (quant_layer.py:67-77), or — after ChannelQuant.update_delta (channelQuant.py:221-237,296-298) —
    [OC,IC,1,1] / [OC,IC]: any shape whose non-unit dims are a prefix of x's."""
    n = x.numel()
    c = delta.numel()
    if c == 1:
        return max(n, 1), 1
    lead = list(delta.shape)
    while lead and lead[-1] == 1:
        lead.pop()
    if tuple(lead) != tuple(x.shape[:len(lead)]):
        raise _lib.SsqError(f"delta shape {tuple(delta.shape)} does not prefix tensor shape {tuple(x.shape)}")
    return max(n // c, 1), c


def match_param(param: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    """zero_point broadcast to delta's shape (they differ after update_delta)"""
    if param.numel() == delta.numel():
        return param
    return param.expand_as(delta).contiguous()


def scalar_dev(value: float, device) -> torch.Tensor:
    return torch.tensor([float(value)], dtype=torch.float32, device=device)


# --------------------------------------------------------------------------------------- K1a
def fq_affine_fwd(x, delta, zero_point, qmin: float, qmax: float, in_scale=None, want_codes=False):
    x = _req(x, "x"); delta = _req(delta, "delta"); zero_point = _req(match_param(zero_point, delta), "zero_point")
    inner, nchan = channel_layout(x, delta)
    if in_scale is not None:
        in_scale = _req(in_scale, "in_scale")
        if in_scale.numel() != inner:
            raise _lib.SsqError("in_scale must have IC*kh*kw elements")
    y = torch.empty_like(x)
    codes = torch.empty_like(x) if want_codes else None
    _call("ssq_fq_affine_fwd", x.data_ptr(), delta.data_ptr(), zero_point.data_ptr(), _ptr(in_scale),
          y.data_ptr(), _ptr(codes), x.numel(), inner, nchan, qmin, qmax, _stream(x))
    return (y, codes) if want_codes else y


def fq_affine_bwd(gy, x, delta, zero_point, qmin, qmax, need_gx=True, need_gparams=True):
    gy = _req(gy, "gy"); x = _req(x, "x"); delta = _req(delta, "delta"); zero_point = _req(zero_point, "zero_point")
    inner, nchan = channel_layout(x, delta)
    gx = torch.empty_like(x) if need_gx else None
    gd = torch.empty_like(delta) if need_gparams else None
    gz = torch.empty_like(zero_point) if need_gparams else None
    ws = _ws(x, nchan)
    _call("ssq_fq_affine_bwd", gy.data_ptr(), x.data_ptr(), delta.data_ptr(), zero_point.data_ptr(),
          _ptr(gx), _ptr(gd), _ptr(gz), x.numel(), inner, nchan, qmin, qmax, ws.data_ptr(), ws.numel(), _stream(x))
    return gx, gd, gz


class FakeQuantAffine(torch.autograd.Function):
    """UniformAffineQuantizer.forward (quant/quant_layer.py:92-97) with its autograd."""

    @staticmethod
    def forward(ctx, x, delta, zero_point, qmin, qmax):
        ctx.save_for_backward(x, delta, zero_point)
        ctx.bounds = (qmin, qmax)
        return fq_affine_fwd(x, delta, zero_point, qmin, qmax)

    @staticmethod
    def backward(ctx, gy):
        x, delta, zero_point = ctx.saved_tensors
        qmin, qmax = ctx.bounds
        need_x, need_d, need_z = ctx.needs_input_grad[:3]
        gx, gd, gz = fq_affine_bwd(gy, x, delta, zero_point, qmin, qmax, need_gx=need_x, need_gparams=need_d or need_z)
        if gx is not None and gx.shape != x.shape:
            gx = gx.view_as(x)
        return gx, (gd if need_d else None), (gz if need_z else None), None, None


# --------------------------------------------------------------------------------------- K1b
def adaround_fwd(w, alpha, delta, zero_point, qmin, qmax, soft: bool, b_dev=None, lam: float = 0.0,
                 want_codes=False, want_reg=False):
    w = _req(w, "w"); alpha = _req(alpha, "alpha"); delta = _req(delta, "delta")
    zero_point = _req(match_param(zero_point, delta), "zero_point")
    if alpha.shape != w.shape:
        raise _lib.SsqError("alpha must have the weight's shape")
    inner, nchan = channel_layout(w, delta)
    wq = torch.empty_like(w)
    codes = torch.empty_like(w) if want_codes else None
    reg = torch.empty(1, dtype=torch.float32, device=w.device) if want_reg else None
    ws = _ws(w, 1) if want_reg else None
    _call("ssq_fq_adaround_fwd", w.data_ptr(), alpha.data_ptr(), delta.data_ptr(), zero_point.data_ptr(),
          wq.data_ptr(), _ptr(codes), w.numel(), inner, nchan, qmin, qmax, int(bool(soft)),
          _ptr(b_dev), float(lam), _ptr(reg), _ptr(ws), 0 if ws is None else ws.numel(), _stream(w))
    out = [wq]
    if want_codes:
        out.append(codes)
    if want_reg:
        out.append(reg)
    return out[0] if len(out) == 1 else tuple(out)


def adaround_bwd(gwq, w, alpha, delta, zero_point, qmin, qmax, b_dev=None, lam: float = 0.0, greg=None,
                 out=None, accumulate=False):
    w = _req(w, "w"); alpha = _req(alpha, "alpha"); delta = _req(delta, "delta")
    zero_point = _req(match_param(zero_point, delta), "zero_point")
    if gwq is not None:
        gwq = _req(gwq, "gwq")
    inner, nchan = channel_layout(w, delta)
    galpha = out if out is not None else torch.empty_like(alpha)
    _call("ssq_fq_adaround_bwd", _ptr(gwq), w.data_ptr(), alpha.data_ptr(), delta.data_ptr(), zero_point.data_ptr(),
          galpha.data_ptr(), w.numel(), inner, nchan, qmin, qmax, _ptr(b_dev), float(lam), _ptr(greg),
          int(bool(accumulate)), _stream(w))
    return galpha


def adaround_init_alpha(w, delta):
    w = _req(w, "w"); delta = _req(delta, "delta")
    inner, nchan = channel_layout(w, delta)
    alpha = torch.empty_like(w)
    _call("ssq_adaround_init_alpha", w.data_ptr(), delta.data_ptr(), alpha.data_ptr(), w.numel(), inner, nchan, _stream(w))
    return alpha


class AdaRoundSoft(torch.autograd.Function):
    """AdaRoundQuantizer soft forward (quant/adaptive_rounding.py:49-59); gradient flows to alpha only
    (weight sees floor() => zero gradient in the reference; delta/zero_point are never optimised here)."""

    @staticmethod
    def forward(ctx, w, alpha, delta, zero_point, qmin, qmax):
        ctx.save_for_backward(w, alpha, delta, zero_point)
        ctx.bounds = (qmin, qmax)
        return adaround_fwd(w, alpha, delta, zero_point, qmin, qmax, soft=True)

    @staticmethod
    def backward(ctx, gwq):
        w, alpha, delta, zero_point = ctx.saved_tensors
        qmin, qmax = ctx.bounds
        galpha = adaround_bwd(gwq, w, alpha, delta, zero_point, qmin, qmax) if ctx.needs_input_grad[1] else None
        return None, galpha, None, None, None, None


def round_reg_fwd(v, b_dev, lam):
    v = _req(v, "v")
    reg = torch.empty(1, dtype=torch.float32, device=v.device)
    ws = _ws(v, 1)
    _call("ssq_round_reg_fwd", v.data_ptr(), v.numel(), b_dev.data_ptr(), float(lam), reg.data_ptr(),
          ws.data_ptr(), ws.numel(), _stream(v))
    return reg


def round_reg_bwd(v, b_dev, lam, greg=None, out=None, accumulate=False):
    v = _req(v, "v")
    gv = out if out is not None else torch.empty_like(v)
    _call("ssq_round_reg_bwd", v.data_ptr(), v.numel(), b_dev.data_ptr(), float(lam), _ptr(greg), gv.data_ptr(),
          int(bool(accumulate)), _stream(v))
    return gv


class RoundReg(torch.autograd.Function):
    """lambda * sum(1 - |2h(v)-1|^b)  (quant/block_recon.py:173-174); returns a 0-dim tensor."""

    @staticmethod
    def forward(ctx, v, b_dev, lam):
        ctx.save_for_backward(v, b_dev)
        ctx.lam = lam
        return round_reg_fwd(v, b_dev, lam).reshape(())

    @staticmethod
    def backward(ctx, g):
        v, b_dev = ctx.saved_tensors
        greg = g.reshape(1).contiguous()
        return round_reg_bwd(v, b_dev, ctx.lam, greg=greg), None, None


# multi-tensor AdaRound -------------------------------------------------------------------------
class AdaRoundTable:
    """Host-side descriptor table for the *_mt launches (one reconstruction unit)."""

    def __init__(self, entries: Sequence[dict]):
        if not 0 < len(entries) <= MT_MAX:
            raise _lib.SsqError(f"a unit may hold 1..{MT_MAX} quantised layers, got {len(entries)}")
        self.count = len(entries)
        self.table = (AdaRoundDesc * self.count)()
        self.keep = entries          # keeps the tensors alive
        tile = 0
        for i, e in enumerate(entries):
            w = _req(e["w"], "w"); alpha = _req(e["alpha"], "alpha")
            inner, nchan = channel_layout(w, e["delta"])
            d = self.table[i]
            d.w, d.alpha = w.data_ptr(), alpha.data_ptr()
            d.delta, d.zero_point = _req(e["delta"], "delta").data_ptr(), _req(e["zero_point"], "zp").data_ptr()
            d.wq = e["wq"].data_ptr()
            d.gwq = 0
            d.galpha = e["galpha"].data_ptr() if e.get("galpha") is not None else 0
            d.n, d.inner, d.nchan = w.numel(), inner, nchan
            d.tile_begin = tile
            d.qmin, d.qmax = e["qmin"], e["qmax"]
            tile += (w.numel() + MT_TILE - 1) // MT_TILE
        self.total_tiles = tile
        self.device = entries[0]["w"].device
        # folded output affine (gamma^z / varphi^z): every entry then carries gamma, phi, bias (or None), beff, ggamma, gphi
        self.aff = None
        if entries[0].get("gamma") is not None:
            self.aff = (AffineDesc * self.count)()
            rows = 0
            for i, e in enumerate(entries):
                f = self.aff[i]
                f.gamma, f.phi = _req(e["gamma"], "gamma").data_ptr(), _req(e["phi"], "phi").data_ptr()
                f.bias = _ptr(e.get("bias"))
                f.beff = e["beff"].data_ptr()
                f.gbeff = 0
                f.ggamma, f.gphi = e["ggamma"].data_ptr(), e["gphi"].data_ptr()
                f.row_begin = rows
                rows += self.table[i].nchan
                if self.table[i].nchan != e["gamma"].numel():
                    raise _lib.SsqError("the folded output affine needs per-output-channel weight quantisers")

    def forward(self, soft: bool, b_dev=None, lam: float = 0.0, reg_out=None):
        ws = workspace(self.device, _lib.load().ssq_ws_bytes(1), "mt") if reg_out is not None else None
        _call("ssq_fq_adaround_fwd_mt", self.table, self.count, self.total_tiles, int(bool(soft)), _ptr(b_dev),
              float(lam), _ptr(reg_out), _ptr(ws), 0 if ws is None else ws.numel(),
              torch.cuda.current_stream(self.device).cuda_stream)

    def backward(self, gwqs: Sequence[torch.Tensor], b_dev=None, lam: float = 0.0):
        for i, g in enumerate(gwqs):
            self.table[i].gwq = _req(g, "gwq").data_ptr()
        _call("ssq_fq_adaround_bwd_mt", self.table, self.count, self.total_tiles, _ptr(b_dev), float(lam),
              torch.cuda.current_stream(self.device).cuda_stream)

    def backward_adam(self, gwqs: Sequence[torch.Tensor], b_dev, lam: float, flat, exp_avg, exp_avg_sq, lr_dev, step_dev,
                      betas=(0.9, 0.999), eps=1e-8, store_grad=False, apply_adam=True):
        """gradient of every alpha + its Adam step + end of the iteration (*step_dev += 1) in one launch;
        apply_adam=False: gradients only (written to galpha), for the multi-GPU exchange"""
        for i, g in enumerate(gwqs):
            self.table[i].gwq = _req(g, "gwq").data_ptr()
        ws = workspace(self.device, _lib.load().ssq_ws_bytes(1), "mt")
        _call("ssq_fq_adaround_bwd_adam_mt", self.table, self.aff, self.count, self.total_tiles, _ptr(b_dev), float(lam),
              flat.data_ptr(), _ptr(exp_avg), _ptr(exp_avg_sq), lr_dev.data_ptr(), step_dev.data_ptr(),
              float(betas[0]), float(betas[1]), float(eps), int(bool(store_grad) or not apply_adam), int(bool(apply_adam)),
              ws.data_ptr(), ws.numel(), torch.cuda.current_stream(self.device).cuda_stream)

    def affine_grad(self, gwqs: Sequence[torch.Tensor], gbeffs: Sequence[torch.Tensor]):
        """gradients of the folded gamma / varphi from the folded layer's weight and bias gradients (before alpha moves)"""
        for i, (g, gb) in enumerate(zip(gwqs, gbeffs)):
            self.table[i].gwq = _req(g, "gwq").data_ptr()
            self.aff[i].gbeff = _req(gb, "gbeff").data_ptr()
        _call("ssq_affine_grad_mt", self.table, self.aff, self.count, torch.cuda.current_stream(self.device).cuda_stream)


class IterationState:
    """device-side iteration state of one reconstruction loop (include/ssq_b200.h, ssq_iter_state)"""

    def __init__(self, step_dev, idx_table, idx_live, b_table, b_live, lr_table, lr_live, n_steps: int):
        self.keep = (step_dev, idx_table, idx_live, b_table, b_live, lr_table, lr_live)
        self.c = IterState(step_dev.data_ptr(), idx_table.data_ptr(), _ptr(idx_live), _ptr(b_table), _ptr(b_live),
                           _ptr(lr_table), _ptr(lr_live), int(n_steps), int(idx_table.shape[-1]))
        self.device = step_dev.device


def iter_prologue(state: IterationState, cache, cur_inp, table: Optional[AdaRoundTable], lam: float = 0.0, reg_out=None):
    """launch 1 of an iteration: bookkeeping + mini-batch gather (cache None: skipped) + multi-tensor soft forward with the
    regulariser (table None: skipped)"""
    per_sample = 0 if cache is None else cache.numel() // cache.shape[0]
    ws = workspace(state.device, _lib.load().ssq_ws_bytes(1), "mt") if reg_out is not None else None
    _call("ssq_iter_prologue", C.byref(state.c), _ptr(cache), _ptr(cur_inp) if cache is not None else None, per_sample,
          table.table if table is not None else None, table.aff if table is not None else None,
          table.count if table is not None else 0,
          table.total_tiles if table is not None else 0, float(lam), _ptr(reg_out), _ptr(ws), 0 if ws is None else ws.numel(),
          torch.cuda.current_stream(state.device).cuda_stream)


# --------------------------------------------------------------------------------------- K1c
def shift_probs_fwd(alpha, reg_mode: int = -1, b_dev=None, lam: float = 0.0, want_reg=False):
    alpha = _req(alpha, "alpha")
    S = alpha.shape[-1]
    groups = alpha.numel() // S
    p = torch.empty_like(alpha)
    reg = torch.empty(1, dtype=torch.float32, device=alpha.device) if want_reg else None
    ws = _ws(alpha, 1) if want_reg else None
    _call("ssq_shift_probs_fwd", alpha.data_ptr(), p.data_ptr(), groups, S, int(reg_mode), _ptr(b_dev), float(lam),
          _ptr(reg), _ptr(ws), 0 if ws is None else ws.numel(), _stream(alpha))
    return (p, reg) if want_reg else p


def shift_probs_bwd(alpha, gp, reg_mode: int = -1, b_dev=None, lam: float = 0.0, greg=None):
    alpha = _req(alpha, "alpha")
    S = alpha.shape[-1]
    groups = alpha.numel() // S
    if gp is not None:
        gp = _req(gp, "gp")
    galpha = torch.empty_like(alpha)
    _call("ssq_shift_probs_bwd", alpha.data_ptr(), _ptr(gp), galpha.data_ptr(), groups, S, int(reg_mode), _ptr(b_dev),
          float(lam), _ptr(greg), _stream(alpha))
    return galpha


class ShiftProbs(torch.autograd.Function):
    """p = clamp(softmax(alpha,-1)*1.2-0.1, 0, 1)  (quant/channelQuant.py:120-121)."""

    @staticmethod
    def forward(ctx, alpha):
        ctx.save_for_backward(alpha)
        return shift_probs_fwd(alpha)

    @staticmethod
    def backward(ctx, gp):
        (alpha,) = ctx.saved_tensors
        return shift_probs_bwd(alpha, gp)


class ShiftProbsReg(torch.autograd.Function):
    """regulariser on the group probabilities: mode 0 entropy, mode 1 pow (see include/ssq_b200.h)."""

    @staticmethod
    def forward(ctx, alpha, reg_mode, b_dev, lam):
        ctx.save_for_backward(alpha, b_dev) if b_dev is not None else ctx.save_for_backward(alpha)
        ctx.cfg = (reg_mode, lam, b_dev is not None)
        _, reg = shift_probs_fwd(alpha, reg_mode, b_dev, lam, want_reg=True)
        return reg.reshape(())

    @staticmethod
    def backward(ctx, g):
        reg_mode, lam, has_b = ctx.cfg
        alpha = ctx.saved_tensors[0]
        b_dev = ctx.saved_tensors[1] if has_b else None
        return shift_probs_bwd(alpha, None, reg_mode, b_dev, lam, greg=g.reshape(1).contiguous()), None, None, None


def _shift_dims(w: torch.Tensor, per_element: bool):
    oc = w.shape[0]
    ic = w.shape[1] if w.dim() > 1 else 1
    kk = w.numel() // (oc * ic)
    if per_element:
        return oc, ic * kk, 1
    return oc, ic, kk


def fq_shift_fwd(w, shift_delta, delta, zero_point, p, beta, mode, hard_targets, hard_round, qmin, qmax, per_element):
    w = _req(w, "w"); shift_delta = _req(shift_delta, "shift_delta"); p = _req(p, "p")
    delta = _req(delta, "delta"); zero_point = _req(zero_point, "zero_point")
    if beta is not None:
        beta = _req(beta, "beta")
    oc, ic, kk = _shift_dims(w, per_element)
    S = p.shape[-1]
    y = torch.empty_like(w)
    _call("ssq_fq_shift_fwd", w.data_ptr(), shift_delta.data_ptr(), delta.data_ptr(), zero_point.data_ptr(), p.data_ptr(),
          _ptr(beta), y.data_ptr(), oc, ic, kk, S, int(per_element), int(mode), int(bool(hard_targets)),
          int(bool(hard_round)), qmin, qmax, _stream(w))
    return y


def fq_shift_bwd(gy, w, shift_delta, delta, zero_point, p, beta, mode, hard_round, qmin, qmax, per_element, need_gbeta):
    gy = _req(gy, "gy"); w = _req(w, "w"); p = _req(p, "p")
    oc, ic, kk = _shift_dims(w, per_element)
    S = p.shape[-1]
    gp = torch.empty_like(p)
    gbeta = torch.empty_like(w) if (need_gbeta and beta is not None) else None
    nbytes = _lib.load().ssq_shift_bwd_ws_bytes(oc, ic, kk, S, int(per_element))
    ws = workspace(w.device, nbytes, "shift")
    _call("ssq_fq_shift_bwd", gy.data_ptr(), w.data_ptr(), shift_delta.data_ptr(), delta.data_ptr(), zero_point.data_ptr(),
          p.data_ptr(), _ptr(beta), gp.data_ptr(), _ptr(gbeta), oc, ic, kk, S, int(per_element), int(mode),
          int(bool(hard_round)), qmin, qmax, ws.data_ptr(), ws.numel(), _stream(w))
    return gp, gbeta


class ShiftMix(torch.autograd.Function):
    """ChannelQuant 'learned_hard_sigmoid' / 'adaShift' soft forward (quant/channelQuant.py:49-64,96-118)."""

    @staticmethod
    def forward(ctx, p, beta, w, shift_delta, delta, zero_point, mode, hard_round, qmin, qmax, per_element):
        ctx.save_for_backward(p, beta if beta is not None else p.new_empty(0), w, shift_delta, delta, zero_point)
        ctx.cfg = (mode, hard_round, qmin, qmax, per_element, beta is not None)
        return fq_shift_fwd(w, shift_delta, delta, zero_point, p, beta, mode, False, hard_round, qmin, qmax, per_element)

    @staticmethod
    def backward(ctx, gy):
        p, beta, w, shift_delta, delta, zero_point = ctx.saved_tensors
        mode, hard_round, qmin, qmax, per_element, has_beta = ctx.cfg
        beta = beta if has_beta else None
        need_gbeta = has_beta and ctx.needs_input_grad[1] and not hard_round
        gp, gbeta = fq_shift_bwd(gy, w, shift_delta, delta, zero_point, p, beta, mode, hard_round, qmin, qmax,
                                 per_element, need_gbeta)
        return gp, (gbeta if need_gbeta else None), None, None, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------- K2
def mse_scale_search(x2d, n_levels: int, symmetric: bool, p_norm: float = 2.4, ws_tag: str = "search", out=None):
    """x2d: [rows, k]. Returns delta, zero_point, raw_zero_point, best_score (fp32 [rows]) and index (int32)."""
    x2d = _req(x2d, "x")
    rows, k = x2d.shape
    dev = x2d.device
    if out is None:
        delta = torch.empty(rows, dtype=torch.float32, device=dev)
        zp = torch.empty_like(delta); raw = torch.empty_like(delta); score = torch.empty_like(delta)
        idx = torch.empty(rows, dtype=torch.int32, device=dev)
    else:
        delta, zp, raw, score, idx = out
    ws = workspace(dev, _lib.load().ssq_mse_scale_search_ws_bytes(rows, k), ws_tag)
    _call("ssq_mse_scale_search", x2d.data_ptr(), rows, k, int(n_levels), int(bool(symmetric)), float(p_norm),
          delta.data_ptr(), zp.data_ptr(), raw.data_ptr(), score.data_ptr(), idx.data_ptr(),
          ws.data_ptr(), ws.numel(), _stream(x2d))
    return delta, zp, raw, score, idx


_search_streams = {}
_MANY_MIN_ELEMS = 256 * 1024          # mse_scale_search_many: smaller tensors are not worth a side stream


def mse_scale_search_many(jobs, n_streams: int = 4):
    """the searches of several independent tensors (jobs: (x2d, n_levels, symmetric[, p_norm]) each) issued round-robin on
    `n_streams` side streams forked from the current stream and joined back into it: a model's layers have 64-512 rows each,
    i.e. one launch fills a fraction of the 148 SMs x 8 resident CTAs, and their searches do not depend on each other
    (quant_layer.py:100-166 runs them one after the other only because each quantiser initialises itself lazily).
    Results are bit-identical to calling mse_scale_search per tensor; outputs are allocated on the calling stream."""
    if not jobs:
        return []
    dev = jobs[0][0].device
    main = torch.cuda.current_stream(dev)
    key = (dev.index, n_streams)
    if key not in _search_streams:
        _search_streams[key] = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    streams = _search_streams[key]
    outs = []
    for x2d, *_rest in jobs:                                   # outputs belong to the calling stream's allocator pool
        rows = x2d.shape[0]
        f = lambda dt=torch.float32: torch.empty(rows, dtype=dt, device=dev)
        outs.append((f(), f(), f(), f(), f(torch.int32)))
    # only tensors with enough work to be worth a fork go to the side streams (ResNet-18: the 256- and 512-channel layers, 1.69 ->
    # 1.16 ms for the model); small ones (depthwise 3x3, 64-channel layers) stay on the calling stream — sent through the side
    # streams too, MobileNetV2's 53 mostly tiny layers took 1.7 ms instead of 1.0
    big = [i for i, job in enumerate(jobs) if job[0].numel() >= _MANY_MIN_ELEMS]
    if len(big) >= 2:
        for s in streams:
            s.wait_stream(main)
    for n, i in enumerate(big if len(big) >= 2 else []):
        x2d, n_levels, symmetric = jobs[i][:3]
        p_norm = jobs[i][3] if len(jobs[i]) > 3 else 2.4
        with torch.cuda.stream(streams[n % n_streams]):
            mse_scale_search(x2d.contiguous(), n_levels, symmetric, p_norm, ws_tag=f"search{n % n_streams}", out=outs[i])
    for i, job in enumerate(jobs):
        if len(big) >= 2 and i in big:
            continue
        x2d, n_levels, symmetric = job[:3]
        mse_scale_search(x2d.contiguous(), n_levels, symmetric, job[3] if len(job) > 3 else 2.4, out=outs[i])
    if len(big) >= 2:
        for s in streams:
            main.wait_stream(s)
    return outs


def row_minmax(x2d):
    x2d = _req(x2d, "x")
    rows, k = x2d.shape
    mn = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    mx = torch.empty_like(mn)
    done = 0
    while done < rows:   # gridDim.y limit
        r = min(rows - done, 65535)
        ws = _ws(x2d, r, "search")
        _call("ssq_row_minmax", x2d[done:].data_ptr(), r, k, mn[done:].data_ptr(), mx[done:].data_ptr(),
              ws.data_ptr(), ws.numel(), _stream(x2d))
        done += r
    return mn, mx


def inp_scale_search(w2d, delta, raw_zero_point, cand, x_range: float, lo: float, hi: float, inp_scale, force_brute=False):
    """w2d [oc,k]; updates inp_scale ([k]) in place and returns it. force_brute: evaluate every (column, candidate) with
    the reference expression instead of the one-pass monotone search (cross-check; same result)."""
    w2d = _req(w2d, "w"); delta = _req(delta, "delta"); raw_zero_point = _req(raw_zero_point, "raw_zero_point")
    cand = _req(cand, "cand"); inp_scale = _req(inp_scale, "inp_scale")
    oc, k = w2d.shape
    ws = workspace(w2d.device, _lib.load().ssq_inp_scale_search_ws_bytes2(oc, k), "inpscale")
    _call("ssq_inp_scale_search_ex", w2d.data_ptr(), delta.data_ptr(), raw_zero_point.data_ptr(), cand.data_ptr(),
          cand.numel(), float(x_range), float(lo), float(hi), inp_scale.data_ptr(), oc, k, int(bool(force_brute)),
          ws.data_ptr(), ws.numel(), _stream(w2d))
    return inp_scale


# --------------------------------------------------------------------------------------- K3
LOSS_MODES = {"mse": 0, "mse_all": 0, "fisher_diag": 1, "fisher_full": 2}   # mse_all: lp_loss(reduction != 'none')


def _loss_dims(pred: torch.Tensor, mode: str = "mse"):
    batch = pred.shape[0] if pred.dim() > 0 else 1
    per_sample = pred.numel() // max(batch, 1)
    chan = pred.shape[1] if (pred.dim() > 1 and mode != "mse_all") else 1
    denom = pred.numel() / chan          # mean over everything but dim 1 (quant_layer.py:30); global mean for :32
    return batch, per_sample, float(denom)


def recon_loss(pred, tgt, p: float = 2.0, mode: str = "mse", fisher=None, tgt_index=None, want_grad=True,
               gscale=None):
    """Returns (loss[1], dpred or None). With tgt_index, tgt/fisher are the full cached tensors."""
    pred = _req(pred, "pred"); tgt = _req(tgt, "tgt")
    if fisher is not None:
        fisher = _req(fisher, "fisher")
    batch, per_sample, denom = _loss_dims(pred, mode)
    if tgt_index is None and tgt.shape != pred.shape:
        raise _lib.SsqError("pred/tgt shape mismatch")
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    dpred = torch.empty_like(pred) if want_grad else None
    ws = _ws(pred, batch if mode == "fisher_full" else 1, "loss")
    _call("ssq_recon_loss", pred.data_ptr(), tgt.data_ptr(), _ptr(fisher), _ptr(tgt_index), loss.data_ptr(),
          _ptr(dpred), batch, per_sample, denom, LOSS_MODES[mode], float(p), _ptr(gscale),
          ws.data_ptr(), ws.numel(), _stream(pred))
    return loss, dpred


def recon_loss_bwd(pred, tgt, gloss, p: float = 2.0, mode: str = "mse", fisher=None, tgt_index=None):
    pred = _req(pred, "pred"); tgt = _req(tgt, "tgt")
    batch, per_sample, denom = _loss_dims(pred, mode)
    dpred = torch.empty_like(pred)
    ws = _ws(pred, batch if mode == "fisher_full" else 1, "loss")
    _call("ssq_recon_loss_bwd", pred.data_ptr(), tgt.data_ptr(), _ptr(fisher), _ptr(tgt_index), gloss.data_ptr(),
          dpred.data_ptr(), batch, per_sample, denom, LOSS_MODES[mode], float(p), ws.data_ptr(), ws.numel(),
          _stream(pred))
    return dpred


class ReconLoss(torch.autograd.Function):
    """lp_loss(pred, tgt, p) / Fisher losses as one reduction (quant/quant_layer.py:25-32,
    quant/block_recon.py:154-162). Gradient flows to pred only."""

    @staticmethod
    def forward(ctx, pred, tgt, p, mode, fisher):
        loss, _ = recon_loss(pred, tgt, p, mode, fisher, want_grad=False)
        ctx.save_for_backward(pred, tgt, fisher if fisher is not None else pred.new_empty(0))
        ctx.cfg = (p, mode, fisher is not None)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        pred, tgt, fisher = ctx.saved_tensors
        p, mode, has_f = ctx.cfg
        dpred = recon_loss_bwd(pred, tgt, g.reshape(1).contiguous(), p, mode, fisher if has_f else None)
        return dpred.view_as(pred), None, None, None, None


# --------------------------------------------------------------------------------------- integer export
def packed_row_bytes(k: int, n_bits: int) -> int:
    return int(_lib.load().ssq_packed_row_bytes(int(k), int(n_bits)))


def export_codes(w, delta, zero_point, qmin: float, qmax: float, n_bits: int, alpha=None, in_scale=None):
    """bit-packed integer codes [rows, packed_row_bytes] (uint8, device) of the hard forward of w; see include/ssq_b200.h"""
    w = _req(w, "w"); delta = _req(delta, "delta"); zero_point = _req(match_param(zero_point, delta), "zero_point")
    rows = w.shape[0]
    k = w.numel() // max(rows, 1)
    inner, nchan = channel_layout(w, delta)
    if alpha is not None:
        alpha = _req(alpha, "alpha")
        if alpha.shape != w.shape:
            raise _lib.SsqError("alpha must have the weight's shape")
    if in_scale is not None:
        in_scale = _req(in_scale, "in_scale")
        if in_scale.numel() != k:
            raise _lib.SsqError("in_scale must have IC*kh*kw elements")
    packed = torch.zeros((rows, packed_row_bytes(k, n_bits)), dtype=torch.uint8, device=w.device)
    _call("ssq_export_codes", w.data_ptr(), _ptr(alpha), _ptr(in_scale), delta.data_ptr(), zero_point.data_ptr(),
          packed.data_ptr(), rows, k, inner, nchan, float(qmin), float(qmax), int(n_bits), _stream(w))
    return packed


def import_codes(packed, shape, delta, zero_point, qmin: float, n_bits: int, in_scale=None):
    """dequantised fp32 weight of `shape` from codes packed by export_codes"""
    delta = _req(delta, "delta"); zero_point = _req(match_param(zero_point, delta), "zero_point")
    if not packed.is_cuda or packed.dtype != torch.uint8:
        raise _lib.SsqError("packed must be a uint8 device tensor")
    packed = packed.contiguous()
    wq = torch.empty(tuple(shape), dtype=torch.float32, device=packed.device)
    rows = wq.shape[0]
    k = wq.numel() // max(rows, 1)
    if tuple(packed.shape) != (rows, packed_row_bytes(k, n_bits)):
        raise _lib.SsqError("packed tensor does not match shape / n_bits")
    inner, nchan = channel_layout(wq, delta)
    if in_scale is not None:
        in_scale = _req(in_scale, "in_scale")
    _call("ssq_import_codes", packed.data_ptr(), _ptr(in_scale), delta.data_ptr(), zero_point.data_ptr(), wq.data_ptr(),
          rows, k, inner, nchan, float(qmin), int(n_bits), _stream(wq))
    return wq


# --------------------------------------------------------------------------------------- affine / adam / loop
def chan_affine_fwd(x, a, b):
    x = _req(x, "x"); a = _req(a, "a"); b = _req(b, "b")
    nchan = a.numel()
    inner = x.numel() // (x.shape[0] * nchan) if x.numel() else 1
    y = torch.empty_like(x)
    _call("ssq_chan_affine_fwd", x.data_ptr(), a.data_ptr(), b.data_ptr(), y.data_ptr(), x.numel(), max(inner, 1), nchan, _stream(x))
    return y


def chan_affine_bwd(gy, x, a, need_gx=True):
    gy = _req(gy, "gy"); x = _req(x, "x"); a = _req(a, "a")
    nchan = a.numel()
    inner = x.numel() // (x.shape[0] * nchan) if x.numel() else 1
    gx = torch.empty_like(x) if need_gx else None
    ga = torch.empty_like(a); gb = torch.empty_like(a)
    ws = _ws(x, nchan)
    _call("ssq_chan_affine_bwd", gy.data_ptr(), x.data_ptr(), a.data_ptr(), _ptr(gx), ga.data_ptr(), gb.data_ptr(),
          x.numel(), max(inner, 1), nchan, ws.data_ptr(), ws.numel(), _stream(x))
    return gx, ga, gb


class ChanAffine(torch.autograd.Function):
    """out*alpha_out + beta_out (quant/quant_layer.py:258-259)."""

    @staticmethod
    def forward(ctx, x, a, b):
        ctx.save_for_backward(x, a)
        return chan_affine_fwd(x, a, b)

    @staticmethod
    def backward(ctx, gy):
        x, a = ctx.saved_tensors
        gx, ga, gb = chan_affine_bwd(gy, x, a, need_gx=ctx.needs_input_grad[0])
        return gx, ga.view_as(a), gb.view_as(a)


def adam_step(param, grad, exp_avg, exp_avg_sq, lr_dev, step_dev, betas=(0.9, 0.999), eps=1e-8):
    if param.numel() == 0:
        return
    _call("ssq_adam_step", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
          lr_dev.data_ptr(), float(betas[0]), float(betas[1]), float(eps), step_dev.data_ptr(), _stream(param))


def adam_step_pending(param, grad, exp_avg, exp_avg_sq, lr_dev, step_dev, betas=(0.9, 0.999), eps=1e-8):
    """adam_step with t = *step_dev + 1 that leaves the iteration open (a parameter group stepped before the launch that ends it)"""
    if param.numel() == 0:
        return
    _call("ssq_adam_step_pending", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
          lr_dev.data_ptr(), float(betas[0]), float(betas[1]), float(eps), step_dev.data_ptr(), _stream(param))


def adam_step_end_iteration(param, grad, exp_avg, exp_avg_sq, lr_dev, step_dev, betas=(0.9, 0.999), eps=1e-8):
    """adam_step with t = *step_dev + 1 that also ends the iteration (*step_dev += 1 once every CTA has retired)"""
    ws = workspace(param.device, _lib.load().ssq_ws_bytes(1), "mt")
    _call("ssq_adam_step_end_iteration", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
          lr_dev.data_ptr(), float(betas[0]), float(betas[1]), float(eps), step_dev.data_ptr(), ws.data_ptr(), ws.numel(), _stream(param))


def grad_exchange_adam(sym, exp_avg_shard, exp_avg_sq_shard, lr_dev, step_dev, betas=(0.9, 0.999), eps=1e-8, reduced_out=None):
    """the multi-GPU exchange step as ONE kernel over peer memory: reduce-scatter of the symmetric gradient buffers, Adam on
    this rank's shard, all-gather of the new parameters into every rank's buffer, end of iteration (include/ssq_b200.h)"""
    dev = sym.flat.device
    ws = workspace(dev, _lib.load().ssq_ws_bytes(1), "mt")
    _call("ssq_grad_exchange_adam", sym.flat_ptrs, sym.grad_ptrs, sym.pad_ptrs, sym.rank, sym.world, sym.n,
          exp_avg_shard.data_ptr(), exp_avg_sq_shard.data_ptr(), lr_dev.data_ptr(), step_dev.data_ptr(),
          float(betas[0]), float(betas[1]), float(eps), _ptr(reduced_out), sym.timeouts.data_ptr(), sym.epochs.data_ptr(),
          ws.data_ptr(), ws.numel(),
          torch.cuda.current_stream(dev).cuda_stream)


def gather_rows(src, index, out=None):
    src = _req(src, "src")
    batch = index.numel()
    per_sample = src.numel() // src.shape[0]
    if out is None:
        out = torch.empty((batch,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    _call("ssq_gather_rows", src.data_ptr(), index.data_ptr(), out.data_ptr(), batch, per_sample, _stream(src))
    return out


def stage_rows_h2d(host_src: torch.Tensor, rows: torch.Tensor, dev_dst: torch.Tensor, stream=None):
    """dev_dst[n] = host_src[rows[n]]; host_src pinned fp32, rows a CPU int64 tensor; async on `stream`"""
    if host_src.is_cuda or not host_src.is_pinned() or host_src.dtype != torch.float32:
        raise _lib.SsqError("host_src must be a pinned fp32 host tensor")
    rows = rows.contiguous()
    per_sample = host_src.numel() // host_src.shape[0]
    st = (stream or torch.cuda.current_stream(dev_dst.device)).cuda_stream
    _lib.check(_lib.load().ssq_stage_rows_h2d(host_src.data_ptr(), rows.data_ptr(), dev_dst.data_ptr(), rows.numel(),
                                               per_sample, st), "ssq_stage_rows_h2d")


def pull_rows_host(host_src: torch.Tensor, idx_table: torch.Tensor, step_dev: torch.Tensor, lookahead: int, n_steps: int,
                   dev_dst: torch.Tensor, max_ctas: int = 32, stream=None):
    """dev_dst[n] = host_src[idx_table[min(step_dev + lookahead, n_steps-1), n]], read by the SMs out of mapped pinned
    host memory; idx_table / step_dev live on the device, so the call is capturable and needs no host work per step"""
    if host_src.is_cuda or not host_src.is_pinned() or host_src.dtype != torch.float32:
        raise _lib.SsqError("host_src must be a pinned fp32 host tensor")
    _req(dev_dst, "dev_dst")
    if not (idx_table.is_cuda and step_dev.is_cuda and idx_table.dtype == torch.int64 and step_dev.dtype == torch.int64):
        raise _lib.SsqError("idx_table / step_dev must be int64 device tensors")
    per_sample = host_src.numel() // host_src.shape[0]
    batch = idx_table.shape[-1]
    if dev_dst.numel() != batch * per_sample:
        raise _lib.SsqError("dev_dst does not hold one mini-batch of rows")
    st = (stream or torch.cuda.current_stream(dev_dst.device)).cuda_stream
    _call("ssq_pull_rows_host", host_src.data_ptr(), idx_table.data_ptr(), step_dev.data_ptr(), int(lookahead), int(n_steps),
          dev_dst.data_ptr(), batch, per_sample, int(max_ctas), st)


PACK_CHUNK = 1024


class PackedRows:
    """zero-packed host-resident rows (include/ssq_b200.h, ssq_pull_rows_host_packed): the non-zero values of every row in pinned
    host memory, the bit mask and the per-chunk value offsets on the device"""

    def __init__(self, vals_host, mask_dev, chunk_off_dev, shape):
        self.vals, self.mask, self.chunk_off, self.shape = vals_host, mask_dev, chunk_off_dev, tuple(shape)
        self.per_sample = int(mask_dev.shape[1]) * 32
        self.nnz = int(vals_host.numel()) - 8

    @property
    def density(self) -> float:
        return self.nnz / float(self.shape[0] * self.per_sample)

    def host_bytes_per_row(self) -> float:
        return 4.0 * self.nnz / self.shape[0]


def packable(t: torch.Tensor) -> bool:
    return t is not None and t.dtype == torch.float32 and t.dim() >= 2 and t.shape[0] > 0 and (t[0].numel() % PACK_CHUNK) == 0


def pack_rows_sparse(t: torch.Tensor, dev, rows_per_slice: int = 64) -> PackedRows:
    """build the zero-packed form of a [N, ...] fp32 tensor (host or device) with device-side torch ops, `rows_per_slice` rows
    at a time (set-up cost, outside every timed region). An element is dropped iff its 32 bits are all zero."""
    if not packable(t):
        raise _lib.SsqError("pack_rows_sparse needs an fp32 [N, ...] tensor whose rows are a multiple of 1024 elements")
    n, per = t.shape[0], t[0].numel()
    w, c = per // 32, per // PACK_CHUNK
    weights = torch.ones(32, dtype=torch.int64, device=dev) << torch.arange(32, dtype=torch.int64, device=dev)
    masks, vals, counts = [], [], []
    for i in range(0, n, rows_per_slice):
        x = t[i:i + rows_per_slice].to(dev).reshape(-1, per).contiguous()
        nz = x.view(torch.int32) != 0
        counts.append(nz.view(-1, c, PACK_CHUNK).sum(-1, dtype=torch.int64))
        m = (nz.view(-1, w, 32).to(torch.int64) * weights).sum(-1)            # [rows, w] in [0, 2^32)
        masks.append(torch.where(m >= 2 ** 31, m - 2 ** 32, m).to(torch.int32))
        vals.append(x[nz].cpu())
    vals.append(torch.zeros(8))                                                # slack for the aligned over-read of the last chunk
    vals_host = torch.cat(vals).pin_memory()
    chunk_off = torch.zeros(n * c + 1, dtype=torch.int64, device=dev)
    chunk_off[1:] = torch.cat(counts).reshape(-1).cumsum(0)
    return PackedRows(vals_host, torch.cat(masks).contiguous(), chunk_off, t.shape)


def pull_rows_host_packed(packed: PackedRows, idx_table: torch.Tensor, step_dev: torch.Tensor, lookahead: int, n_steps: int,
                          dev_dst: torch.Tensor, max_ctas: int = 24, stream=None):
    """pull_rows_host from a zero-packed cache: only the non-zero values cross PCIe, the SMs expand them into dense rows"""
    if packed.vals.is_cuda or not packed.vals.is_pinned() or not packed.mask.is_cuda or not packed.chunk_off.is_cuda:
        raise _lib.SsqError("packed rows: values must be pinned host memory, mask / chunk offsets device tensors")
    _req(dev_dst, "dev_dst")
    if not (idx_table.is_cuda and step_dev.is_cuda and idx_table.dtype == torch.int64 and step_dev.dtype == torch.int64):
        raise _lib.SsqError("idx_table / step_dev must be int64 device tensors")
    batch = idx_table.shape[-1]
    if dev_dst.numel() != batch * packed.per_sample:
        raise _lib.SsqError("dev_dst does not hold one mini-batch of rows")
    st = (stream or torch.cuda.current_stream(dev_dst.device)).cuda_stream
    _call("ssq_pull_rows_host_packed", packed.mask.data_ptr(), packed.vals.data_ptr(), packed.chunk_off.data_ptr(), idx_table.data_ptr(),
          step_dev.data_ptr(), int(lookahead), int(n_steps), dev_dst.data_ptr(), batch, packed.per_sample, int(max_ctas), st)


def loop_advance(step_dev, idx_table, idx_live, b_table, b_live, lr_table, lr_live, n_steps: int):
    batch = 0 if idx_live is None else idx_live.numel()
    _call("ssq_loop_advance", step_dev.data_ptr(), _ptr(idx_table), _ptr(idx_live), batch, _ptr(b_table), _ptr(b_live),
          _ptr(lr_table), _ptr(lr_live), int(n_steps), _stream(step_dev))
