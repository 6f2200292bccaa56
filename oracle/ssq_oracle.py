"""CPU oracle for the PTQ-calibration hot path — TEST INFRASTRUCTURE ONLY.

A numpy (float32-disciplined) restatement of the reference's arithmetic, function by function, each
citing the upstream file:line it follows. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(shiftedscalequantization_b200/) never does and fails loudly without its CUDA library.

Pinning: every function here is checked against golden vectors produced by importing the real
reference from /root/reference (tests/golden/make_golden.py; run in the build container, vectors
committed under tests/golden/). Float conventions that matter for bit-exact codes:
  * tensor (op) python-scalar casts the scalar to fp32 first (ATen wrapped-number rule);
  * tensor / tensor and tensor / python-scalar are IEEE fp32 divisions on the CPU path
    (the reference's CUDA path multiplies by a reciprocal for python scalars — we follow the CPU path);
  * torch.round is round-half-even (np.rint); clamp bounds are inclusive for the gradient.
"""
from __future__ import annotations

import numpy as np

F = np.float32
ZETA, GAMMA = 1.1, -0.1
STRETCH = F(ZETA - GAMMA)      # python double 1.2000000000000002 -> fp32
GAMMA32 = F(GAMMA)


def _f(x):
    return np.asarray(x, dtype=F)


# ------------------------------------------------------------------------------------------------ helpers
def sigmoid(a):
    a = _f(a)
    return (F(1) / (F(1) + np.exp(-a, dtype=F))).astype(F)


def rect_sigmoid(a):
    """h(a), quant/adaptive_rounding.py:63-64"""
    return np.clip(sigmoid(a) * STRETCH + GAMMA32, F(0), F(1)).astype(F)


def rect_sigmoid_grad(a):
    s = sigmoid(a)
    v = s * STRETCH + GAMMA32
    return np.where((v >= 0) & (v <= 1), STRETCH * s * (F(1) - s), F(0)).astype(F)


def pow_scalar(x, e):
    """ATen pow(tensor, python scalar): exponent cast to fp32; fast paths for 2, 3, .5, 1, 0."""
    x = _f(x); e = F(e)
    if e == 2: return x * x
    if e == 1: return x.copy()
    if e == 3: return x * x * x
    if e == 0.5: return np.sqrt(x)
    if e == 0: return np.ones_like(x)
    return np.power(x, e, dtype=F)


def bounds(n_levels: int, sym: bool):
    """quant/quant_layer.py:93-96"""
    return (F(-(n_levels // 2)), F(n_levels // 2 - 1)) if sym else (F(0), F(n_levels - 1))


def _bcast(param, x):
    """delta/zero_point of shape [C,1,..] or scalar against x"""
    param = _f(param)
    if param.size == 1:
        return param.reshape(())
    return param.reshape((param.size,) + (1,) * (x.ndim - 1))


# ------------------------------------------------------------------------------------------------ K1a
def uaq_forward(x, delta, zero_point, qmin, qmax, in_scale=None):
    """quant/quant_layer.py:92-97; with in_scale quant/channelQuantMSE.py:134-143. Returns (y, codes)."""
    x = _f(x); d = _bcast(delta, x); z = _bcast(zero_point, x)
    if in_scale is None:
        u = x / d
    else:
        s = _f(in_scale).reshape((1,) + x.shape[1:])
        u = (x / s) / d
    q = np.clip(np.rint(u) + z, F(qmin), F(qmax)).astype(F)
    y = (q - z) * d
    if in_scale is not None:
        y = y * s
    return y.astype(F), q


def uaq_backward(gy, x, delta, zero_point, qmin, qmax):
    """autograd of quant/quant_layer.py:92-97 (round_ste :18-22): returns gx, gdelta, gzp (param shapes)."""
    gy = _f(gy); x = _f(x); d = _bcast(delta, x); z = _bcast(zero_point, x)
    u = x / d
    r = np.rint(u)
    xi = r + z
    inside = (xi >= F(qmin)) & (xi <= F(qmax))
    q = np.clip(xi, F(qmin), F(qmax))
    gx = np.where(inside, gy, F(0)).astype(F)
    gd_e = gy.astype(np.float64) * np.where(inside, (r - u), (q - z)).astype(np.float64)
    gz_e = np.where(inside, 0.0, -(gy * d).astype(np.float64))
    dshape = np.shape(delta)
    if np.size(delta) == 1:
        return gx, F(gd_e.sum()).reshape(dshape), F(gz_e.sum()).reshape(dshape)
    axes = tuple(range(1, x.ndim))
    return gx, gd_e.sum(axis=axes).astype(F).reshape(dshape), gz_e.sum(axis=axes).astype(F).reshape(dshape)


# ------------------------------------------------------------------------------------------------ K2a
def _quantize_candidate(x, new_max, new_min, n_bits):
    """quant/quant_layer.py:168-175 (always the unsigned clamp)"""
    lm1 = F(2 ** n_bits - 1)
    delta = (new_max - new_min) / lm1
    zp = np.rint(-new_min / delta)
    q = np.clip(np.rint(x / delta) + zp, F(0), lm1)
    return ((q - zp) * delta).astype(F)


def mse_search_row(x, n_bits: int, sym: bool = False, p: float = 2.4, return_scores=False):
    """quant/quant_layer.py:145-162 for one row / one tensor. Returns (delta, zero_point, raw_zero_point)
    as fp32 scalars; None when no candidate scores below 1e10 (the reference's delta-stays-None case)."""
    x = _f(x).ravel()
    x_max, x_min = x.max(), x.min()
    if sym:
        am = max(abs(x_min), x_max)
        x_min, x_max = (F(-am) if x_min < 0 else F(0)), F(am)
    best, out, scores = F(1e10), None, []
    lm1 = F(2 ** n_bits - 1)
    for i in range(80):
        f = F(1.0 - (i * 0.01))
        new_max, new_min = F(x_max * f), F(x_min * f)
        with np.errstate(all="ignore"):
            xq = _quantize_candidate(x, new_max, new_min, n_bits)
            # fp32 |d|^p, mean accumulated in fp64 (ATen's pairwise fp32 sum differs in the last bits)
            score = F(np.mean(pow_scalar(np.abs(x - xq), p).astype(np.float64)))
        scores.append(score)
        if score < best:
            best = score
            delta = F((new_max - new_min) / lm1)
            zp = F(np.rint(-new_min / delta)) if not sym else F(0)
            raw = F(-new_min) if not sym else F(0)
            out = (delta, zp, raw, i)
    if return_scores:
        return out, np.asarray(scores, dtype=F)
    return out


def mse_search(x, n_bits: int, sym: bool = False, channel_wise: bool = True):
    """quant/quant_layer.py:100-122: per-channel loop over dim 0. Returns arrays (delta, zp, raw, idx)."""
    x = _f(x)
    rows = x.reshape(x.shape[0], -1) if channel_wise else x.reshape(1, -1)
    res = [mse_search_row(r, n_bits, sym) for r in rows]
    if any(r is None for r in res):
        raise TypeError("delta stayed None for a channel (all-NaN scores), as in the reference")
    d, z, raw, idx = (np.asarray([r[j] for r in res]) for j in range(4))
    return d.astype(F), z.astype(F), raw.astype(F), idx.astype(np.int32)


def max_init(x, n_bits: int, sym: bool = False, scale_method: str = "max"):
    """quant/quant_layer.py:124-142 — Python-double arithmetic on the host."""
    x = _f(x)
    x_min = min(float(x.min()), 0)
    x_max = max(float(x.max()), 0)
    if "scale" in scale_method:
        x_min = x_min * (n_bits + 2) / 8
        x_max = x_max * (n_bits + 2) / 8
    if sym:
        am = max(abs(x_min), x_max)
        x_min, x_max = (-am if x_min < 0 else 0), am
    delta = float(x_max - x_min) / (2 ** n_bits - 1)
    if delta < 1e-8:
        delta = 1e-8
    zero_point = round(-x_min / delta)
    return F(delta), F(zero_point), F(-x_min)


# ------------------------------------------------------------------------------------------------ K1b
def adaround_init_alpha(w, delta):
    """quant/adaptive_rounding.py:66-72"""
    w = _f(w); d = _bcast(delta, w)
    u = w / d
    rest = u - np.floor(u)
    return (-np.log(STRETCH / (rest - GAMMA32) - F(1), dtype=F)).astype(F)


def adaround_forward(w, alpha, delta, zero_point, qmin, qmax, soft: bool):
    """quant/adaptive_rounding.py:49-59 (qmin,qmax = 0,L-1 there; ChannelQuant 'adaround' passes sym bounds,
    quant/channelQuant.py:66-78). Returns (wq, codes)."""
    w = _f(w); d = _bcast(delta, w); z = _bcast(zero_point, w)
    fl = np.floor(w / d)
    r = rect_sigmoid(alpha) if soft else (_f(alpha) >= 0).astype(F)
    q = np.clip((fl + r) + z, F(qmin), F(qmax)).astype(F)
    return ((q - z) * d).astype(F), q


def adaround_backward(gwq, w, alpha, delta, zero_point, qmin, qmax):
    """d wq / d alpha through the soft path (autograd of adaptive_rounding.py:50-59)."""
    gwq = _f(gwq); w = _f(w); d = _bcast(delta, w); z = _bcast(zero_point, w)
    xi = (np.floor(w / d) + rect_sigmoid(alpha)) + z
    inside = (xi >= F(qmin)) & (xi <= F(qmax))
    return np.where(inside, (gwq * d) * rect_sigmoid_grad(alpha), F(0)).astype(F)


def round_reg(alpha_or_h, b: float, lam: float, is_h: bool = False):
    """quant/block_recon.py:173-174: lam * sum(1 - |2(h-.5)|^b). b <= 0 => 0 (warm-up, :167)."""
    if b <= 0:
        return F(0)
    h = _f(alpha_or_h) if is_h else rect_sigmoid(alpha_or_h)
    t = np.abs(h - F(0.5)) * F(2)
    return F(lam * np.sum((F(1) - pow_scalar(t, b)).astype(np.float64)))


def round_reg_grad(alpha, b: float, lam: float):
    if b <= 0:
        return np.zeros_like(_f(alpha))
    h = rect_sigmoid(alpha)
    d = h - F(0.5)
    t = np.abs(d) * F(2)
    dr = -(F(b) * pow_scalar(t, b - 1.0)) * F(2) * np.sign(d)
    return (F(lam) * dr * rect_sigmoid_grad(alpha)).astype(F)


def linear_temp_decay(t, t_max, rel_start_decay, start_b, end_b):
    """quant/block_recon.py:185-202 (also layer_recon_shiftedScale.py:488-505)"""
    start_decay = rel_start_decay * t_max
    if t < start_decay:
        return start_b
    rel_t = (t - start_decay) / (t_max - start_decay) if t_max != start_decay else 1
    return end_b + (start_b - end_b) * max(0.0, (1 - rel_t))


# ------------------------------------------------------------------------------------------------ K3
def lp_loss(pred, tgt, p=2.0):
    """quant/quant_layer.py:30: (pred-tgt).abs().pow(p).sum(1).mean(); returns (loss, dpred)"""
    pred = _f(pred); tgt = _f(tgt)
    d = pred - tgt
    ad = np.abs(d)
    denom = pred.size / (pred.shape[1] if pred.ndim > 1 else 1)
    loss = F(np.sum(pow_scalar(ad, p).astype(np.float64)) / denom)
    inv = F(1) / F(denom)
    dpred = (inv * (F(p) * pow_scalar(ad, p - 1.0))) * np.sign(d).astype(F)
    return loss, dpred.astype(F)


def fisher_diag_loss(pred, tgt, grad):
    """quant/block_recon.py:156-157"""
    pred = _f(pred); tgt = _f(tgt); g = _f(grad)
    d = pred - tgt
    denom = pred.size / pred.shape[1]
    loss = F(np.sum(((d * d) * (g * g)).astype(np.float64)) / denom)
    return loss, ((F(1) / F(denom)) * (g * g) * (F(2) * d)).astype(F)


def fisher_full_loss(pred, tgt, grad):
    """quant/block_recon.py:158-162"""
    pred = _f(pred); tgt = _f(tgt)
    d = pred - tgt
    a = np.abs(d); g = np.abs(_f(grad))
    dots = np.sum((a * g).astype(np.float64), axis=tuple(range(1, pred.ndim)))
    scale = 1.0 / (pred.size * 100.0)
    loss = F(np.sum(dots * dots) * scale)
    dp = (2.0 * dots * scale).reshape((-1,) + (1,) * (pred.ndim - 1)) * g * np.sign(d)
    return loss, dp.astype(F)


# ------------------------------------------------------------------------------------------------ K1c
def softmax_last(a):
    a = _f(a)
    e = np.exp(a - a.max(axis=-1, keepdims=True), dtype=F)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F)).astype(F)


def shift_probs(alpha):
    """quant/channelQuant.py:120-121"""
    return np.clip(softmax_last(alpha) * STRETCH + GAMMA32, F(0), F(1)).astype(F)


def shift_probs_backward(alpha, gp):
    sm = softmax_last(alpha)
    v = sm * STRETCH + GAMMA32
    gs = np.where((v >= 0) & (v <= 1), _f(gp), F(0)) * STRETCH
    dot = np.sum(gs * sm, axis=-1, keepdims=True, dtype=F)
    return (sm * (gs - dot)).astype(F)


def entropy_reg(alpha, lam):
    """quant/layer_recon_shiftedScale.py:393: lam * (-sum p*log(p+1e-10)); returns (reg, dreg/dp)"""
    p = shift_probs(alpha)
    reg = F(lam * -np.sum((p * np.log(p + F(1e-10), dtype=F)).astype(np.float64)))
    dp = -(np.log(p + F(1e-10), dtype=F) + p / (p + F(1e-10)))
    return reg, (F(lam) * dp).astype(F)


def pow_reg_on_probs(alpha, b, lam):
    """quant/layer_recon_fused_shiftedScale.py:281-282"""
    p = shift_probs(alpha)
    if b <= 0:
        return F(0), np.zeros_like(p)
    d = p - F(0.5)
    t = np.abs(d) * F(2)
    reg = F(lam * np.sum((F(1) - pow_scalar(t, b)).astype(np.float64)))
    dp = -(F(b) * pow_scalar(t, b - 1.0)) * F(2) * np.sign(d)
    return reg, (F(lam) * dp).astype(F)


def _group_view(p, w):
    """p [IC,S] (conv) or [OC,IC,S] (FC, per element) -> broadcastable to w + trailing S"""
    if w.ndim == 4:
        return p.reshape(1, w.shape[1], 1, 1, p.shape[-1])
    return p.reshape(w.shape + (p.shape[-1],))


def shift_terms(w, delta, zero_point, shifts, qmin, qmax, mode):
    """candidates per shift: mode 'dequant' = init_v (quant/channelQuant.py:201-213, 'none' forward :79-94);
    mode 'floor' = init_v_beta (:279-288). Returns [..., S]."""
    w = _f(w); d = _bcast(delta, w); z = _bcast(zero_point, w)
    out = []
    for s in shifts:
        ds = (d * F(s)).astype(F)
        if mode == "dequant":
            q = np.clip(np.rint(w / ds) + z, F(qmin), F(qmax))
            out.append(((q - z) * ds).astype(F))
        else:
            out.append(np.floor(w / ds).astype(F))
    return np.stack(out, axis=-1)


def shift_mix(terms, p, w, hard: bool):
    """quant/channelQuant.py:96-118"""
    pv = _group_view(p, w)
    if hard:
        idx = np.argmax(np.broadcast_to(pv, terms.shape), axis=-1)   # first maximum
        return np.take_along_axis(terms, idx[..., None], axis=-1)[..., 0].astype(F)
    out = terms[..., 0] * pv[..., 0]
    for i in range(1, terms.shape[-1]):
        out = out + terms[..., i] * pv[..., i]
    return out.astype(F)


def shift_forward(w, delta, zero_point, shifts, p, qmin, qmax, mode, hard_targets=False, beta=None, hard_round=False):
    """mode 'dequant': ChannelQuant 'learned_hard_sigmoid' forward; mode 'adashift': quant/channelQuant.py:50-64"""
    w = _f(w)
    if mode == "dequant":
        return shift_mix(shift_terms(w, delta, zero_point, shifts, qmin, qmax, "dequant"), p, w, hard_targets)
    d = _bcast(delta, w); z = _bcast(zero_point, w)
    f = shift_mix(shift_terms(w, delta, zero_point, shifts, qmin, qmax, "floor"), p, w, hard_targets)
    r = (_f(beta) >= 0).astype(F) if hard_round else rect_sigmoid(beta)
    q = np.clip((f + r) + z, F(qmin), F(qmax))
    return ((q - z) * (d * F(1.0))).astype(F)


def shift_backward(gy, w, delta, zero_point, shifts, p, qmin, qmax, mode, beta=None, hard_round=False):
    """gradients wrt p (group shape of p) and beta for the soft-target forward"""
    gy = _f(gy); w = _f(w)
    if mode == "dequant":
        terms = shift_terms(w, delta, zero_point, shifts, qmin, qmax, "dequant")
        gm = gy; gbeta = None
    else:
        d = _bcast(delta, w); z = _bcast(zero_point, w)
        terms = shift_terms(w, delta, zero_point, shifts, qmin, qmax, "floor")
        f = shift_mix(terms, p, w, False)
        r = (_f(beta) >= 0).astype(F) if hard_round else rect_sigmoid(beta)
        xi = (f + r) + z
        inside = (xi >= F(qmin)) & (xi <= F(qmax))
        gm = np.where(inside, gy * d, F(0)).astype(F)
        gbeta = np.zeros_like(w) if hard_round else (gm * rect_sigmoid_grad(beta)).astype(F)
    ge = gm[..., None].astype(np.float64) * terms.astype(np.float64)
    if w.ndim == 4:
        gp = ge.sum(axis=(0, 2, 3))
    else:
        gp = ge
    return gp.astype(F).reshape(np.shape(p)), gbeta


# ------------------------------------------------------------------------------------------------ K2b
def init_shift_candidates(w, delta, zero_point, n_levels, sym=False, channel_wise_groups=True):
    """ChannelQuant.init_shift_candidates, quant/channelQuant.py:240-277 (RUN_CHANNEL_WISE = True): every scale i/8, i = 1..15
    without 8, is scored per input-channel group by sum |x_q(delta * st) - x|^2.4 over output channels and kernel positions; each
    group votes 3 / 2 / 1 for its three best scales; the two scales with the most votes (first-come on ties: Python's stable sort)
    plus 1.0 are returned."""
    w = _f(w)
    cands = [i / 8 for i in range(1, 16) if i != 8]
    qmin, qmax = bounds(n_levels, sym)
    is_fc = w.ndim == 2
    table = []
    for st in cands:
        d = (_f(delta) * F(st)).astype(F)
        y, _ = uaq_forward(w, d, zero_point, qmin, qmax)
        err = pow_scalar(np.abs(y - w), 2.4)
        table.append(err.sum(axis=0, dtype=F) if is_fc else err.sum(axis=(0, 2, 3), dtype=F))
    table = np.stack(table, 0)                                       # [14, groups]
    order = np.argsort(table, axis=0, kind="stable")[:3]
    scores = {i: 0 for i in range(len(cands))}
    for col in range(table.shape[1]):
        for j in range(3):
            scores[int(order[j, col])] += 3 - j
    top = [k for k, _v in sorted(scores.items(), key=lambda kv: kv[1], reverse=True)][:2]
    return [cands[i] for i in top] + [1.0]


def inp_scale_search(w, delta, raw_zero_point, n_levels, level, threshold, inp_scale=None):
    """quant/channelQuantMSE.py:70-110 ('max' mode). w [OC,IC,kh,kw] or [OC,IC]; returns inp_scale [1,IC,kh,kw]."""
    w = _f(w); d = _bcast(delta, w)
    x_range = n_levels - 1
    min_lim = F(0.0 - 0.5 / x_range * threshold)
    max_lim = F(1.0 + 0.5 / x_range * threshold)
    zero = np.rint(_bcast(raw_zero_point, w) / d)
    shape = (1,) + w.shape[1:]
    out = np.ones(shape, dtype=F) if inp_scale is None else _f(inp_scale).reshape(shape).copy()
    for c in [i / level for i in range(level, 0, -1)]:
        cs = F(c)
        xq = ((w / cs) / d + zero) / F(x_range)
        mn = xq.min(axis=0, keepdims=True); mx = xq.max(axis=0, keepdims=True)
        out = np.where((mn > min_lim) & (mx < max_lim), cs, out).astype(F)
    return out


def channelquantmse_forward(w, delta, raw_zero_point, inp_scale, n_levels):
    """quant/channelQuantMSE.py:134-143"""
    w = _f(w); d = _bcast(delta, w)
    zp = np.rint(_bcast(raw_zero_point, w) / d)
    s = _f(inp_scale).reshape((1,) + w.shape[1:])
    q = np.clip(np.rint((w / s) / d) + zp, F(0), F(n_levels - 1))
    return (((q - zp) * d) * s).astype(F), q.astype(F)


# ------------------------------------------------------------------------------------------------ optimiser
def adam_step(p, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch/optim/adam.py _single_tensor_adam (defaults used at quant/block_recon.py:60)"""
    p = _f(p); g = _f(g); m = _f(m); v = _f(v)
    m = m + F(1 - beta1) * (g - m)
    v = v * F(beta2) + F(1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / F(bc2 ** 0.5) + F(eps)
    p = p - F(lr / bc1) * (m / denom)
    return p.astype(F), m.astype(F), v.astype(F)


# ------------------------------------------------------------------------------------------------ integer export
def storage_bits(n_bits: int) -> int:
    return 1 if n_bits <= 1 else (2 if n_bits <= 2 else (4 if n_bits <= 4 else 8))


def pack_rows(codes, qmin, n_bits: int):
    """bit-packing of the integer intermediate `x_quant` the reference's hard forwards compute and discard
    (quant/quant_layer.py:92-96, quant/adaptive_rounding.py:50-58): u = q - qmin, `storage_bits` bits per element,
    little-endian within a byte, every row ([rows, k] view) starting on a byte boundary (include/ssq_b200.h)."""
    q = np.asarray(codes)
    rows = q.shape[0]
    u = (q.reshape(rows, -1).astype(np.int64) - int(qmin)).astype(np.uint8)
    sb = storage_bits(n_bits)
    per = 8 // sb
    k = u.shape[1]
    row_bytes = (k * sb + 7) // 8
    pad = np.zeros((rows, row_bytes * per), dtype=np.uint8)
    pad[:, :k] = u
    pad = pad.reshape(rows, row_bytes, per)
    out = np.zeros((rows, row_bytes), dtype=np.uint8)
    for e in range(per):
        out |= (pad[:, :, e] << (e * sb)).astype(np.uint8)
    return out


def unpack_rows(packed, k: int, qmin, n_bits: int):
    sb = storage_bits(n_bits)
    per = 8 // sb
    packed = np.asarray(packed, dtype=np.uint8)
    rows = packed.shape[0]
    parts = [(packed >> (e * sb)) & ((1 << sb) - 1) for e in range(per)]
    u = np.stack(parts, axis=-1).reshape(rows, -1)[:, :k]
    return (u.astype(F) + F(qmin)).astype(F)


# ------------------------------------------------------------------------------------------------ zero-packed host cache
# No reference counterpart (quant/data_utils.py:29-36 keeps dense CPU tensors): this is the library's own transport format for
# the host-resident cache (include/ssq_b200.h, ssq_pull_rows_host_packed); parity = the round trip is the identity, bit for bit.
SPARSE_CHUNK = 1024


def sparse_pack_rows(x):
    """x [N, P] fp32, P % 1024 == 0 -> (mask uint32 [N, P/32], vals fp32 [nnz], chunk_off int64 [N*P/1024 + 1]).
    An element is dropped iff its 32 bits are all zero (+0.0f); bit e%32 of word e/32 marks element e."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, p = x.shape
    assert p % SPARSE_CHUNK == 0
    nz = x.view(np.uint32) != 0
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    mask = (nz.reshape(n, p // 32, 32).astype(np.uint64) * weights).sum(-1).astype(np.uint32)
    counts = nz.reshape(n, p // SPARSE_CHUNK, SPARSE_CHUNK).sum(-1).astype(np.int64).reshape(-1)
    chunk_off = np.concatenate([np.zeros(1, np.int64), np.cumsum(counts)])
    return mask, x[nz], chunk_off


def sparse_unpack_rows(mask, vals, chunk_off, rows, per_sample):
    """dense [len(rows), per_sample] rows of a zero-packed cache (plain loops over chunks: small cases only)"""
    c = per_sample // SPARSE_CHUNK
    out = np.zeros((len(rows), per_sample), np.float32)
    for j, r in enumerate(rows):
        for ch in range(c):
            bits = np.zeros(SPARSE_CHUNK, bool)
            for wd in range(32):
                m = int(mask[r, ch * 32 + wd])
                bits[wd * 32:(wd + 1) * 32] = [(m >> b) & 1 for b in range(32)]
            base = int(chunk_off[r * c + ch])
            seg = out[j, ch * SPARSE_CHUNK:(ch + 1) * SPARSE_CHUNK]
            seg[bits] = vals[base:base + int(bits.sum())]
    return out
