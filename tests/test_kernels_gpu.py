"""GPU parity suite: every kernel, called through the C-ABI (ctypes -> libssq_b200.so), against
 (a) the committed golden vectors produced by the real reference, and
 (b) the numpy oracle on seeded inputs incl. ragged / misaligned / empty / large cases.
Integer codes must be bit-exact; floats <= 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_exact, golden
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    a = np.asarray(a, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(a)).reshape(a.shape).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from shiftedscalequantization_b200 import ops
    return ops


def rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------------------------------- K1a
@pytest.mark.parametrize("case", golden("uaq").cases())
def test_fq_affine_golden(ops, case):
    g = golden("uaq").case(case)
    bits, sym, cw, _ = (int(v) for v in g["meta"])
    qmin, qmax = (float(v) for v in O.bounds(2 ** bits, bool(sym)))
    x, d, z = dev(g["x"]), dev(g["delta"]), dev(g["zp"])
    y, codes = ops.fq_affine_fwd(x, d, z, qmin, qmax, want_codes=True)
    assert_exact(host(codes), g["codes"], "codes vs reference")
    assert_close(host(y), g["y"], what="dequant vs reference")
    gx, gd, gz = ops.fq_affine_bwd(dev(g["gy"]), x, d, z, qmin, qmax)
    assert_close(host(gx), g["gx"], what="gx")
    assert_close(host(gd), g["gdelta"], rtol=2e-5, what="gdelta")
    assert_close(host(gz), g["gzp"], rtol=2e-5, what="gzp")


@pytest.mark.parametrize("shape,cw,bits,sym", [
    ((64, 64, 3, 3), True, 2, False), ((32, 1, 3, 3), True, 4, False), ((64, 3, 7, 7), True, 8, False),
    ((1000, 512), True, 8, False), ((16, 24, 1, 1), True, 3, True), ((8, 64, 28, 28), False, 4, False),
    ((5, 7, 3, 3), False, 4, False), ((3, 5), False, 2, False)])
def test_fq_affine_oracle(ops, shape, cw, bits, sym):
    r = rng(hash((shape, bits)) % 2 ** 31)
    x = (r.standard_normal(shape) * (0.05 if cw else 1.0)).astype(np.float32)
    if cw:
        d = (np.abs(x.reshape(shape[0], -1)).max(1) / (2 ** bits - 1) * 1.7).astype(np.float32)
        z = np.round(r.uniform(0, 2 ** bits - 1, shape[0])).astype(np.float32)
        dshape = (shape[0],) + (1,) * (len(shape) - 1)
        d, z = d.reshape(dshape), z.reshape(dshape)
    else:
        d = np.float32(np.abs(x).max() / (2 ** bits - 1) * 1.3).reshape(())
        z = np.float32(3.0).reshape(())
    qmin, qmax = (float(v) for v in O.bounds(2 ** bits, sym))
    y_ref, c_ref = O.uaq_forward(x, d, z, qmin, qmax)
    y, codes = ops.fq_affine_fwd(dev(x), dev(d), dev(z), qmin, qmax, want_codes=True)
    assert_exact(host(codes), c_ref, "codes"); assert_exact(host(y), y_ref, "dequant")
    gy = r.standard_normal(shape).astype(np.float32)
    gx_ref, gd_ref, gz_ref = O.uaq_backward(gy, x, d, z, qmin, qmax)
    gx, gd, gz = ops.fq_affine_bwd(dev(gy), dev(x), dev(d), dev(z), qmin, qmax)
    assert_exact(host(gx), gx_ref, "gx")
    assert_close(host(gd), gd_ref, what="gdelta"); assert_close(host(gz), gz_ref, what="gzp")


def test_fq_affine_misaligned_and_empty(ops):
    r = rng(7)
    base = torch.as_tensor(r.standard_normal(4 * 1024 + 1).astype(np.float32)).cuda()
    x = base[1:]                                     # 4-byte aligned only -> scalar path
    d = dev(np.float32(0.11).reshape(())); z = dev(np.float32(5).reshape(()))
    y, c = ops.fq_affine_fwd(x, d, z, 0.0, 15.0, want_codes=True)
    y_ref, c_ref = O.uaq_forward(host(x), 0.11, 5, 0, 15)
    assert_exact(host(c), c_ref); assert_exact(host(y), y_ref)
    e = torch.empty(0, device="cuda")
    assert ops.fq_affine_fwd(e, d, z, 0.0, 15.0).numel() == 0


def test_fq_affine_large_property(ops):
    """BASELINE-size activation [256,256,56,56] is too big for the numpy oracle in seconds: check
    idempotence (fq(fq(x)) == fq(x)), code range, and a checksum against torch-free arithmetic on a slice."""
    torch.manual_seed(0)
    x = torch.relu(torch.randn(64, 256, 56, 56, device="cuda"))
    d = dev(np.float32(0.2).reshape(())); z = dev(np.float32(0).reshape(()))
    y, c = ops.fq_affine_fwd(x, d, z, 0.0, 15.0, want_codes=True)
    assert float(c.min()) >= 0 and float(c.max()) <= 15
    assert torch.equal(c, c.round())
    y2 = ops.fq_affine_fwd(y, d, z, 0.0, 15.0)
    assert torch.equal(y2, y)
    sl = slice(12345, 12345 + 100000)
    y_ref, c_ref = O.uaq_forward(host(x.flatten()[sl]), 0.2, 0, 0, 15)
    assert_exact(host(c.flatten()[sl]), c_ref); assert_exact(host(y.flatten()[sl]), y_ref)


@pytest.mark.parametrize("case", golden("channelquantmse").cases())
def test_fq_affine_in_scale_golden(ops, case):
    g = golden("channelquantmse").case(case)
    L = 2 ** int(g["bits"])
    w, d = dev(g["w"]), dev(g["delta"])
    zero = torch.round(dev(g["raw"]) / d)
    for level in (1, 4, 16, 64):
        s = dev(g[f"inp_scale_l{level}"])
        y, c = ops.fq_affine_fwd(w, d, zero, 0.0, float(L - 1), in_scale=s.flatten(), want_codes=True)
        assert_exact(host(c), g[f"codes_l{level}"], "codes vs reference")
        assert_close(host(y), g[f"y_l{level}"], what="dequant vs reference")


# ------------------------------------------------------------------------------------------- K1b
@pytest.mark.parametrize("case", golden("adaround").cases())
def test_adaround_golden(ops, case):
    g = golden("adaround").case(case)
    L = 2 ** int(g["bits"])
    w, d, z, a = dev(g["w"]), dev(g["delta"]), dev(g["zp"]), dev(g["alpha"])
    assert_close(host(ops.adaround_init_alpha(w, d)), g["alpha0"], rtol=2e-5, what="alpha init")
    wq = ops.adaround_fwd(w, a, d, z, 0.0, float(L - 1), soft=True)
    assert_close(host(wq), g["wq_soft"], what="soft")
    wqh, codes = ops.adaround_fwd(w, a, d, z, 0.0, float(L - 1), soft=False, want_codes=True)
    assert_exact(host(codes), g["codes_hard"], "hard codes vs reference")
    assert_exact(host(wqh), g["wq_hard"], "hard dequant vs reference")
    ga = ops.adaround_bwd(dev(g["gw"]), w, a, d, z, 0.0, float(L - 1))
    assert_close(host(ga), g["galpha"], what="galpha")
    for b in (20, 11.3, 2.0):
        bd = ops.scalar_dev(b, "cuda")
        assert_close(host(ops.round_reg_fwd(a, bd, 0.01))[0], g[f"reg_b{b}"], rtol=2e-5, what=f"reg {b}")
        assert_close(host(ops.round_reg_bwd(a, bd, 0.01)), g[f"greg_b{b}"], rtol=2e-5, what=f"greg {b}")
        # fused: forward + regulariser in one launch, backward with both paths in one launch
        wq2, reg = ops.adaround_fwd(w, a, d, z, 0.0, float(L - 1), soft=True, b_dev=bd, lam=0.01, want_reg=True)
        assert torch.equal(wq2, wq)
        assert_close(host(reg)[0], g[f"reg_b{b}"], rtol=2e-5, what="fused reg")
        gboth = ops.adaround_bwd(dev(g["gw"]), w, a, d, z, 0.0, float(L - 1), b_dev=bd, lam=0.01)
        assert_close(host(gboth), g["galpha"] + g[f"greg_b{b}"], rtol=2e-5, what="fused galpha")


@pytest.mark.parametrize("shape,bits", [((128, 64, 3, 3), 2), ((96, 1, 3, 3), 3), ((1000, 512), 8), ((64, 3, 7, 7), 8)])
def test_adaround_oracle_and_mt(ops, shape, bits):
    r = rng(sum(shape) + bits)
    L = 2 ** bits
    w = (r.standard_normal(shape) * 0.05).astype(np.float32)
    dshape = (shape[0],) + (1,) * (len(shape) - 1)
    d = (np.abs(w.reshape(shape[0], -1)).max(1) / (L - 1) * 1.5).astype(np.float32).reshape(dshape)
    z = np.round(r.uniform(0, L - 1, shape[0])).astype(np.float32).reshape(dshape)
    a = (r.standard_normal(shape) * 3).astype(np.float32)
    wq_ref, _ = O.adaround_forward(w, a, d, z, 0, L - 1, soft=True)
    _, ch_ref = O.adaround_forward(w, a, d, z, 0, L - 1, soft=False)
    W, A, D, Z = dev(w), dev(a), dev(d), dev(z)
    assert_close(host(ops.adaround_fwd(W, A, D, Z, 0.0, float(L - 1), soft=True)), wq_ref, what="soft")
    _, ch = ops.adaround_fwd(W, A, D, Z, 0.0, float(L - 1), soft=False, want_codes=True)
    assert_exact(host(ch), ch_ref, "hard codes")
    # multi-tensor launch == single-tensor launches, and its regulariser == sum of oracle regs
    w2 = (r.standard_normal((32, 16, 3, 3)) * 0.05).astype(np.float32)
    d2 = np.full((32, 1, 1, 1), 0.02, np.float32); z2 = np.full((32, 1, 1, 1), 1.0, np.float32)
    a2 = (r.standard_normal(w2.shape) * 3).astype(np.float32)
    W2, A2, D2, Z2 = dev(w2), dev(a2), dev(d2), dev(z2)
    out1, out2 = torch.empty_like(W), torch.empty_like(W2)
    g1, g2 = torch.empty_like(W), torch.empty_like(W2)
    tab = ops.AdaRoundTable([
        dict(w=W, alpha=A, delta=D, zero_point=Z, wq=out1, galpha=g1, qmin=0.0, qmax=float(L - 1)),
        dict(w=W2, alpha=A2, delta=D2, zero_point=Z2, wq=out2, galpha=g2, qmin=0.0, qmax=3.0)])
    bd = ops.scalar_dev(7.5, "cuda"); reg = torch.zeros(1, device="cuda")
    tab.forward(True, bd, 0.01, reg)
    assert torch.equal(out1, ops.adaround_fwd(W, A, D, Z, 0.0, float(L - 1), soft=True))
    assert torch.equal(out2, ops.adaround_fwd(W2, A2, D2, Z2, 0.0, 3.0, soft=True))
    assert_close(host(reg)[0], np.float32(O.round_reg(a, 7.5, 0.01) + O.round_reg(a2, 7.5, 0.01)), rtol=2e-5, what="mt reg")
    gw1, gw2 = dev(r.standard_normal(shape)), dev(r.standard_normal(w2.shape))
    tab.backward([gw1, gw2], bd, 0.01)
    assert torch.equal(g1, ops.adaround_bwd(gw1, W, A, D, Z, 0.0, float(L - 1), b_dev=bd, lam=0.01))
    assert torch.equal(g2, ops.adaround_bwd(gw2, W2, A2, D2, Z2, 0.0, 3.0, b_dev=bd, lam=0.01))
    ref = O.adaround_backward(host(gw1), w, a, d, z, 0, L - 1) + O.round_reg_grad(a, 7.5, 0.01)
    assert_close(host(g1), ref, rtol=2e-5, what="mt galpha")
    # warm-up: b <= 0 switches the regulariser off (block_recon.py:167)
    b0 = ops.scalar_dev(0.0, "cuda")
    tab.forward(True, b0, 0.01, reg)
    assert float(reg) == 0.0


# ------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("case", ["conv", "fc"])
def test_recon_loss_golden(ops, case):
    g = golden("loss").case(case)
    pred, tgt = dev(g["pred"]), dev(g["tgt"])
    for p in (2.0, 2.4):
        loss, dp = ops.recon_loss(pred, tgt, p)
        assert_close(host(loss)[0], g[f"lp{p}"], what=f"lp{p}")
        assert_close(host(dp), g[f"dlp{p}"], what=f"dlp{p}")
        dp2 = ops.recon_loss_bwd(pred, tgt, ops.scalar_dev(1.0, "cuda"), p)
        assert torch.equal(dp, dp2)
    if case == "conv":
        f = dev(g["fisher"])
        loss, dp = ops.recon_loss(pred, tgt, mode="fisher_diag", fisher=f)
        assert_close(host(loss)[0], g["fdiag"], what="fdiag"); assert_close(host(dp), g["dfdiag"], what="dfdiag")
        loss, dp = ops.recon_loss(pred, tgt, mode="fisher_full", fisher=f)
        assert_close(host(loss)[0], g["ffull"], what="ffull"); assert_close(host(dp), g["dffull"], rtol=2e-5, what="dffull")


def test_recon_loss_gather_and_autograd(ops):
    r = rng(11)
    cache = dev(r.standard_normal((40, 8, 6, 6)))
    idx = torch.as_tensor(r.permutation(40)[:16]).cuda()
    pred = dev(r.standard_normal((16, 8, 6, 6)))
    l1, d1 = ops.recon_loss(pred, cache, 2.0, tgt_index=idx)
    l2, d2 = ops.recon_loss(pred, cache[idx].contiguous(), 2.0)
    assert torch.equal(l1, l2) and torch.equal(d1, d2)
    assert torch.equal(ops.gather_rows(cache, idx), cache[idx])
    l_ref, d_ref = O.lp_loss(host(pred), host(cache[idx]), 2.0)
    assert_close(host(l1)[0], l_ref); assert_close(host(d1), d_ref)
    pr = pred.clone().requires_grad_(True)
    out = ops.ReconLoss.apply(pr, cache[idx].contiguous(), 2.4, "mse", None)
    (out * 3.0).backward()
    _, d_ref = O.lp_loss(host(pred), host(cache[idx]), 2.4)
    assert_close(host(pr.grad), 3.0 * d_ref)


def test_recon_loss_large_property(ops):
    """ResNet-50 layer1-sized batch (25.7 M elements): linearity + determinism."""
    torch.manual_seed(1)
    pred = torch.randn(32, 256, 56, 56, device="cuda"); tgt = torch.randn_like(pred)
    l1, d1 = ops.recon_loss(pred, tgt, 2.0)
    l2, d2 = ops.recon_loss(pred, tgt, 2.0)
    assert torch.equal(l1, l2) and torch.equal(d1, d2)          # fixed-order reduction
    ref = ((pred.double() - tgt.double()) ** 2).sum() / (32 * 56 * 56)
    assert abs(float(l1) - float(ref)) / float(ref) < 1e-6
    assert torch.allclose(d1, 2 * (pred - tgt) / (32 * 56 * 56), rtol=1e-6, atol=0)


# ------------------------------------------------------------------------------------------- K2a
@pytest.mark.parametrize("case", [c for c in golden("uaq").cases() if "max" not in c])
def test_mse_search_golden(ops, case):
    g = golden("uaq").case(case)
    bits, sym, cw, _ = (int(v) for v in g["meta"])
    x = dev(g["x"])
    x2 = x.reshape(x.shape[0], -1) if cw else x.reshape(1, -1)
    d, z, raw, score, idx = ops.mse_scale_search(x2, 2 ** bits, bool(sym))
    assert (host(idx) >= 0).all()
    assert_exact(host(d).reshape(g["delta"].shape), g["delta"], "delta vs reference")
    assert_exact(host(z).reshape(g["zp"].shape), g["zp"], "zero_point vs reference")
    assert_exact(host(raw).reshape(g["raw"].shape), g["raw"], "raw_zero_point vs reference")


@pytest.mark.parametrize("rows,k,bits", [(64, 576, 2), (48, 9, 4), (16, 147, 8), (10, 512, 8), (128, 4608, 2), (1, 70000, 4)])
def test_mse_search_oracle(ops, rows, k, bits):
    r = rng(rows * 131 + k)
    x = (r.standard_normal((rows, k)) * 0.05).astype(np.float32)
    if rows == 1:
        x = np.maximum(x * 20, 0).astype(np.float32)         # activation-like, goes down the grid-wide path
    d, z, raw, score, idx = ops.mse_scale_search(dev(x), 2 ** bits, False)
    flips = 0
    for i in range(rows):
        (dr, zr, rr, ir), scores = O.mse_search_row(x[i], bits, return_scores=True)
        if int(host(idx)[i]) != ir:
            # a flip is only legitimate at an fp32 tie of the reference's own scores
            rel = abs(float(scores[int(host(idx)[i])]) - float(scores[ir])) / float(scores[ir])
            assert rel < 2e-6, f"row {i}: picked {int(host(idx)[i])} vs {ir}, score gap {rel:.2e}"
            flips += 1
        else:
            assert host(d)[i] == dr and host(z)[i] == zr and host(raw)[i] == rr
    assert flips <= max(1, rows // 50)


def test_mse_search_degenerate_row(ops):
    x = torch.zeros(3, 36, device="cuda"); x[1] = torch.randn(36, device="cuda")
    d, z, raw, score, idx = ops.mse_scale_search(x, 4, False)
    assert int(idx[0]) == -1 and int(idx[2]) == -1 and int(idx[1]) >= 0   # all-zero channel: delta stays None upstream
    mn, mx = ops.row_minmax(x)
    assert torch.equal(mn, x.min(1)[0]) and torch.equal(mx, x.max(1)[0])


# ------------------------------------------------------------------------------------------- K2b
@pytest.mark.parametrize("case", golden("channelquantmse").cases())
def test_inp_scale_search_golden(ops, case):
    g = golden("channelquantmse").case(case)
    L = 2 ** int(g["bits"])
    w = dev(g["w"]); oc = w.shape[0]
    for level in (1, 4, 16, 64):
        thr = float(g[f"thr_l{level}"])
        cand = torch.tensor([i / level for i in range(level, 0, -1)], dtype=torch.float32).cuda()
        lo = float(np.float32(0.0 - 0.5 / (L - 1) * thr)); hi = float(np.float32(1.0 + 0.5 / (L - 1) * thr))
        s = torch.ones(w.numel() // oc, device="cuda")
        ops.inp_scale_search(w.reshape(oc, -1), dev(g["delta"]).flatten(), dev(g["raw"]).flatten(), cand, L - 1, lo, hi, s)
        assert_exact(host(s).reshape(g[f"inp_scale_l{level}"].shape), g[f"inp_scale_l{level}"], f"inp_scale l{level}")


def test_inp_scale_search_oracle_large(ops):
    r = rng(5)
    w = (r.standard_normal((512, 256, 3, 3)) * 0.02).astype(np.float32)       # the notebook's layer shape
    d, z, raw = zip(*[O.max_init(row, 2) for row in w.reshape(512, -1)])
    d = np.array(d, np.float32); raw = np.array(raw, np.float32)
    for level, thr in [(8, 1.5), (1024, 1.5)]:
        ref = O.inp_scale_search(w, d.reshape(-1, 1, 1, 1), raw.reshape(-1, 1, 1, 1), 4, level, thr) if level <= 8 else None
        cand = torch.tensor([i / level for i in range(level, 0, -1)], dtype=torch.float32).cuda()
        lo = float(np.float32(0.0 - 0.5 / 3 * thr)); hi = float(np.float32(1.0 + 0.5 / 3 * thr))
        s = torch.ones(256 * 9, device="cuda")
        ops.inp_scale_search(dev(w).reshape(512, -1), dev(d), dev(raw), cand, 3, lo, hi, s)
        if ref is not None:
            assert_exact(host(s), ref.reshape(-1), "inp_scale")
        else:
            # monotone property: every chosen scale is a candidate and the column fits at it
            assert set(np.unique(host(s))).issubset(set(host(cand).tolist()))


@pytest.mark.parametrize("level,thr,bits", [(1, 1.0, 2), (2, 1.5, 2), (7, 1.0, 3), (16, 1.5, 2), (333, 2.0, 4), (1024, 1.5, 2), (1024, 0.3, 8)])
def test_inp_scale_search_one_pass_equals_brute_force(ops, level, thr, bits):
    """the one-pass monotone-prefix search against the brute-force sweep of every (column, candidate) with the reference
    expression, on weights salted with the cases the estimate is weakest on: zeros, values sitting exactly on a candidate
    boundary (w = V * c_j), NaN / Inf, ragged column counts"""
    r = rng(level * 7 + bits)
    L = 2 ** bits
    oc, k = 96, 1000 + (level % 3)                                   # k not a multiple of 4 for some cases
    w = (r.standard_normal((oc, k)) * 0.03).astype(np.float32)
    w[r.integers(0, oc, 200), r.integers(0, k, 200)] = 0.0
    d, z, raw = zip(*[O.max_init(row, bits) for row in w])
    d = np.array(d, np.float32); raw = np.array(raw, np.float32)
    cand_np = np.array([i / level for i in range(level, 0, -1)], dtype=np.float32)
    lo = np.float32(0.0 - 0.5 / (L - 1) * thr); hi = np.float32(1.0 + 0.5 / (L - 1) * thr)
    # boundary salting: w = fl(vmax * c_j) where vmax ~ delta * ((hi * (L-1)) - zero) is close to the row's interval end
    zero = np.rint(raw / d)
    for _ in range(300):
        i, j, c = r.integers(0, oc), r.integers(0, k), cand_np[r.integers(0, level)]
        vmax = d[i] * (hi * (L - 1) - zero[i])
        w[i, j] = np.float32(vmax * c * (1 + (r.integers(-3, 4)) * 6e-8))
    w[3, 5] = np.nan; w[7, 11] = np.inf; w[9, 13] = -np.inf
    w[11, 17] = 3e38; w[12, 19] = -1e20; w[13, 23] = 1e-40; w[14, 29] = -0.0      # huge finite, denormal, negative zero
    cand = torch.from_numpy(cand_np).cuda()
    fast = torch.ones(k, device="cuda"); brute = torch.ones(k, device="cuda")
    ops.inp_scale_search(dev(w), dev(d), dev(raw), cand, L - 1, float(lo), float(hi), fast)
    ops.inp_scale_search(dev(w), dev(d), dev(raw), cand, L - 1, float(lo), float(hi), brute, force_brute=True)
    assert_exact(host(fast), host(brute), f"inp_scale, level {level}")
    assert host(fast)[5] == 1.0 and host(fast)[11] == 1.0 and host(fast)[13] == 1.0     # NaN / Inf columns never fit: untouched
    assert host(fast)[17] == 1.0 and host(fast)[19] == 1.0                               # ... nor do huge finite values
    if level <= 16:
        ref = O.inp_scale_search(w, d.reshape(-1, 1), raw.reshape(-1, 1), L, level, thr)
        assert_exact(host(fast), ref.reshape(-1), "inp_scale vs oracle")


@pytest.mark.parametrize("seed,oc,k,level,bits,thr", [(1, 300, 2052, 16, 2, 1.5), (2, 1100, 516, 100, 3, 1.0), (3, 64, 4100, 1024, 4, 2.0),
                                                         (4, 513, 1030, 3, 2, 1.5), (5, 2048, 260, 64, 2, 0.7)])
def test_inp_scale_search_many_columns_on_candidate_boundaries(ops, seed, oc, k, level, bits, thr):
    """the column-wise decision of the one-pass search under stress: in most columns the tightest element sits within a few ulps
    of a candidate boundary (w = V c_j (1 +- n ulp)), so most columns take the finish kernel's exact settle path — many per CTA —
    and several near-maximal elements compete in one column. Must equal the brute force bit for bit."""
    r = rng(7000 + seed)
    L = 2 ** bits
    w = (r.standard_normal((oc, k)) * 0.01).astype(np.float32)
    d, z, raw = zip(*[O.max_init(row, bits) for row in w])
    d = np.array(d, np.float32); raw = np.array(raw, np.float32)
    w *= np.float32(0.25)                                            # background well inside every row's interval: the salted entries decide
    lo = np.float32(0.0 - 0.5 / (L - 1) * thr); hi = np.float32(1.0 + 0.5 / (L - 1) * thr)
    cand_np = np.array([i / level for i in range(level, 0, -1)], dtype=np.float32)
    zero = np.rint(raw / d)
    vmax = (d * (hi * (L - 1) - zero)).astype(np.float32)            # ~ the positive end of each row's fitting interval
    vmin = (d * (lo * (L - 1) - zero)).astype(np.float32)            # ~ the negative end
    cols = r.permutation(k)[: (3 * k) // 4]                          # three quarters of the columns get boundary elements
    for j in cols:
        c = cand_np[r.integers(0, level)]
        for i in r.integers(0, oc, size=r.integers(1, 4)):           # one to three competing rows
            end = vmax[i] if r.random() < 0.5 else vmin[i]
            w[i, j] = np.float32(end * c) * np.float32(1 + r.integers(-4, 5) * 6e-8)
    cand = torch.from_numpy(cand_np).cuda()
    fast = torch.ones(k, device="cuda"); brute = torch.ones(k, device="cuda")
    ops.inp_scale_search(dev(w), dev(d), dev(raw), cand, L - 1, float(lo), float(hi), fast)
    ops.inp_scale_search(dev(w), dev(d), dev(raw), cand, L - 1, float(lo), float(hi), brute, force_brute=True)
    assert_exact(host(fast), host(brute), f"inp_scale, {oc}x{k}, level {level}")
    assert len(np.unique(host(fast))) > 1                            # the columns really end up on different candidates
    # the same workspace serves the next call: a forced brute force must not leave the switch on
    again = torch.ones(k, device="cuda")
    ops.inp_scale_search(dev(w), dev(d), dev(raw), cand, L - 1, float(lo), float(hi), again)
    assert_exact(host(again), host(fast), "second call on the same workspace")


def test_inp_scale_search_falls_back_outside_its_preconditions(ops):
    """a zero point outside [0, L-1], a foreign candidate list and lo >= 0 take the brute-force path: same answers as the
    forced brute force, and as the oracle"""
    r = rng(77)
    w = (r.standard_normal((40, 260)) * 0.03).astype(np.float32)
    d = np.full(40, 0.02, np.float32)
    raw = (r.uniform(-0.2, 0.4, 40)).astype(np.float32)              # zero = rint(raw/d) in [-10, 20]: outside [0, 3]
    level, L = 16, 4
    cand_np = np.array([i / level for i in range(level, 0, -1)], dtype=np.float32)
    for cand_try, lo, hi in ((cand_np, -0.25, 1.25), (cand_np * np.float32(0.97), -0.25, 1.25), (cand_np, 0.01, 1.25)):
        cand = torch.from_numpy(cand_try.astype(np.float32)).cuda()
        raw_try = raw if cand_try is cand_np and lo < 0 else np.abs(raw) * 0.1
        fast = torch.ones(260, device="cuda"); brute = torch.ones(260, device="cuda")
        ops.inp_scale_search(dev(w), dev(d), dev(raw_try), cand, L - 1, lo, hi, fast)
        ops.inp_scale_search(dev(w), dev(d), dev(raw_try), cand, L - 1, lo, hi, brute, force_brute=True)
        assert_exact(host(fast), host(brute), "fallback path")


@pytest.mark.parametrize("k", [1, 5, 9, 27, 127, 128, 129, 1023])
def test_mse_search_row_variants_agree_with_oracle(ops, k):
    """warp-per-row (k < 128) and CTA-per-row variants, ragged lengths, against the oracle's full 80-candidate libm scan"""
    r = rng(900 + k)
    rows = 37
    x = (r.standard_normal((rows, k)) * 0.05).astype(np.float32)
    for bits in (2, 4):
        d, z, raw, score, idx = ops.mse_scale_search(dev(x), 2 ** bits, False)
        for i in range(rows):
            res, scores = O.mse_search_row(x[i], bits, return_scores=True)
            got = int(host(idx)[i])
            if res is None:                                          # k == 1: delta = 0, every score NaN
                assert got == -1
                continue
            dr, zr, rr, ir = res
            if got != ir:
                rel = abs(float(scores[got]) - float(scores[ir])) / max(float(scores[ir]), 1e-30)
                assert rel < 2e-6, f"k={k} row {i}: picked {got} vs {ir}, score gap {rel:.2e}"
            else:
                assert host(d)[i] == dr and host(z)[i] == zr and host(raw)[i] == rr


def test_mse_search_two_pass_equals_full_scan_on_hostile_rows(ops):
    """rows the ranking pass must hand to the settle pass untouched: NaN / Inf / huge elements, p = 2 (no ranking), and
    plain rows for which every candidate is settled (p = 2.4 vs the same search with the window forced open by p = 2.4000001)"""
    x = torch.randn(6, 300, device="cuda") * 0.1
    x[1, 7] = float("nan"); x[2, 9] = float("inf"); x[3, 11] = 1e30
    d, z, raw, score, idx = ops.mse_scale_search(x, 16, False)
    hx = host(x)
    for i in (0, 3, 4, 5):
        (dr, zr, rr, ir), scores = O.mse_search_row(hx[i], 4, return_scores=True)
        assert int(host(idx)[i]) == ir and host(d)[i] == dr, i
    assert int(host(idx)[1]) == -1                                   # NaN poisons every score: delta stays None upstream
    d2, z2, raw2, score2, idx2 = ops.mse_scale_search(x[:1].repeat(4, 1), 16, False, p_norm=2.0)
    (dr, zr, rr, ir), _ = O.mse_search_row(hx[0], 4, p=2.0, return_scores=True)
    assert (host(idx2) == ir).all() and (host(d2) == dr).all()


# ------------------------------------------------------------------------------------------- K1c
@pytest.mark.parametrize("case", golden("channelquant").cases())
def test_shift_golden(ops, case):
    g = golden("channelquant").case(case)
    L = 2 ** int(g["bits"]); shifts = [float(s) for s in g["shifts"]]
    w, d, z = dev(g["w"]), dev(g["delta"]), dev(g["zp"])
    per_el = w.dim() == 2
    sd = torch.stack([d.flatten() * s for s in shifts])          # delta * python float, fp32 like ATen
    alpha = dev(g["alpha"])
    p = ops.shift_probs_fwd(alpha)
    assert_close(host(p), g["p"], what="p")
    y = ops.fq_shift_fwd(w, sd, d, z, p, None, ops.SHIFT_DEQUANT, False, False, 0.0, float(L - 1), per_el)
    assert_close(host(y), g["y_soft"], what="soft mixture vs reference")
    yh = ops.fq_shift_fwd(w, sd, d, z, dev(g["p"]), None, ops.SHIFT_DEQUANT, True, False, 0.0, float(L - 1), per_el)
    assert_exact(host(yh), g["y_hard"], "hard select vs reference")
    gp, _ = ops.fq_shift_bwd(dev(g["gy"]), w, sd, d, z, p, None, ops.SHIFT_DEQUANT, False, 0.0, float(L - 1), per_el, False)
    assert_close(host(ops.shift_probs_bwd(alpha, gp)), g["galpha_soft"], rtol=3e-5, what="galpha vs reference")
    _, ent = ops.shift_probs_fwd(alpha, 0, None, 0.7, want_reg=True)
    assert_close(host(ent)[0], g["ent"], what="entropy")
    assert_close(host(ops.shift_probs_bwd(alpha, None, 0, None, 0.7)), g["gent"], rtol=3e-5, what="d entropy")
    # adaShift
    a2, b2 = dev(g["as_alpha"]), dev(g["as_beta"])
    p2 = ops.shift_probs_fwd(a2)
    y = ops.fq_shift_fwd(w, sd, d, z, p2, b2, ops.SHIFT_ADASHIFT, False, False, 0.0, float(L - 1), per_el)
    assert_close(host(y), g["as_y"], what="adaShift soft vs reference")
    yh = ops.fq_shift_fwd(w, sd, d, z, p2, b2, ops.SHIFT_ADASHIFT, True, True, 0.0, float(L - 1), per_el)
    assert_exact(host(yh), g["as_y_hard"], "adaShift hard vs reference")
    gp, gb = ops.fq_shift_bwd(dev(g["gy"]), w, sd, d, z, p2, b2, ops.SHIFT_ADASHIFT, False, 0.0, float(L - 1), per_el, True)
    assert_close(host(ops.shift_probs_bwd(a2, gp)), g["as_galpha"], rtol=3e-5, what="adaShift galpha")
    assert_close(host(gb), g["as_gbeta"], what="adaShift gbeta")
    for bb in (20, 7.7):
        bd = ops.scalar_dev(bb, "cuda")
        _, rs = ops.shift_probs_fwd(a2, 1, bd, 0.3, want_reg=True)
        assert_close(host(rs)[0], g[f"as_regS_b{bb}"], rtol=2e-5, what="regS")
        assert_close(host(ops.shift_probs_bwd(a2, None, 1, bd, 0.3)), g[f"as_gregS_b{bb}"], rtol=3e-5, what="d regS")
    # adaround mode on the per-(oc,ic) delta selected by update_delta
    y = ops.adaround_fwd(w, dev(g["ar_beta"]), dev(g["ar_delta"]), z, 0.0, float(L - 1), soft=True)
    assert_close(host(y), g["ar_y"], what="adaround-after-shift")
    yh = ops.adaround_fwd(w, dev(g["ar_beta"]), dev(g["ar_delta"]), z, 0.0, float(L - 1), soft=False)
    assert_exact(host(yh), g["ar_y_hard"], "adaround-after-shift hard")
    gb = ops.adaround_bwd(dev(g["gy"]), w, dev(g["ar_beta"]), dev(g["ar_delta"]), z, 0.0, float(L - 1))
    assert_close(host(gb), g["ar_gbeta"], what="adaround-after-shift gbeta")


def test_shift_oracle_resnet_shape(ops):
    r = rng(23)
    shape = (128, 64, 3, 3); L = 4; shifts = [0.96875, 1.03125, 1.0]
    w = (r.standard_normal(shape) * 0.05).astype(np.float32)
    d = (np.abs(w.reshape(128, -1)).max(1) / 3 * 1.4).astype(np.float32).reshape(128, 1, 1, 1)
    z = np.full((128, 1, 1, 1), 2.0, np.float32)
    alpha = (r.standard_normal((64, 3)) * 1.5).astype(np.float32)
    beta = (r.standard_normal(shape) * 2).astype(np.float32)
    gy = r.standard_normal(shape).astype(np.float32)
    p_ref = O.shift_probs(alpha)
    W, D, Z, A, B, GY = dev(w), dev(d), dev(z), dev(alpha), dev(beta), dev(gy)
    sd = torch.stack([D.flatten() * s for s in shifts])
    P = ops.shift_probs_fwd(A)
    for mode, name in ((ops.SHIFT_DEQUANT, "dequant"), (ops.SHIFT_ADASHIFT, "adashift")):
        y = ops.fq_shift_fwd(W, sd, D, Z, P, B, mode, False, False, 0.0, 3.0, False)
        assert_close(host(y), O.shift_forward(w, d, z, shifts, p_ref, 0, 3, name, beta=beta), what=name)
        gp, gb = ops.fq_shift_bwd(GY, W, sd, D, Z, P, B, mode, False, 0.0, 3.0, False, mode == ops.SHIFT_ADASHIFT)
        gp_ref, gb_ref = O.shift_backward(gy, w, d, z, shifts, p_ref, 0, 3, name, beta=beta)
        assert_close(host(gp), gp_ref, rtol=3e-5, what=name + " gp")
        if gb_ref is not None:
            assert_close(host(gb), gb_ref, what=name + " gbeta")


def _misaligned(t):
    """same values at an address that is 4 (mod 16): forces the scalar K1c kernels"""
    buf = torch.empty(t.numel() + 1, device=t.device, dtype=t.dtype)
    v = buf[1:]
    v.copy_(t.reshape(-1))
    return v.view(t.shape)


def _same_bits(a, b):
    a = a.reshape(-1); b = b.reshape(-1)
    return bool(((a.view(torch.int32) == b.view(torch.int32)) | (torch.isnan(a) & torch.isnan(b))).all())


@pytest.mark.parametrize("shape", [(64, 64, 3, 3), (96, 64, 1, 1), (48, 32, 1, 2), (40, 24, 5, 5), (33, 20, 1, 1), (7, 4, 1, 3), (3, 4, 1, 1)])
@pytest.mark.parametrize("S", [1, 2, 3, 4])
def test_shift_vector_paths_vs_oracle_and_scalar(ops, shape, S):
    """K1c vector kernels (group layouts kk >= 4 / kk == 1 / narrow groups, K = 4 fallback) against the oracle, and bit-for-bit
    against the scalar kernels of the same library, incl. the out-of-line slow path: NaN / Inf / huge / denormal weights, scales
    outside the hoisted-reciprocal range and a non-integer zero point"""
    oc, ic, kh, kw = shape
    r = rng(oc * 1000 + ic * 10 + kh * kw + S)
    shifts = [0.96875, 1.03125, 1.0, 0.9375][:S]
    w = (r.standard_normal(shape) * 0.05).astype(np.float32)
    d = (np.abs(w.reshape(oc, -1)).max(1) / 3 * 1.3).astype(np.float32).reshape(oc, 1, 1, 1)
    z = r.integers(0, 3, (oc, 1, 1, 1)).astype(np.float32)
    alpha = (r.standard_normal((ic, S)) * 1.5).astype(np.float32)
    beta = (r.standard_normal(shape) * 2).astype(np.float32)
    gy = r.standard_normal(shape).astype(np.float32)
    p_ref = O.shift_probs(alpha)
    W, D, Z, A, B, GY = dev(w), dev(d), dev(z), dev(alpha), dev(beta), dev(gy)
    sd = torch.stack([D.flatten() * s for s in shifts])
    # the oracle's own probabilities: with p one ulp apart, a mixture of equal floors lands one ulp either side of the clamp
    # bound and the (legitimately discontinuous) inside-mask of the adaShift gradient flips
    assert_close(host(ops.shift_probs_fwd(A)), p_ref, what="group probabilities")
    P = dev(p_ref)
    for mode, name in ((ops.SHIFT_DEQUANT, "dequant"), (ops.SHIFT_ADASHIFT, "adashift")):
        for ht, hr in ((False, False), (True, True), (False, True)):
            y = ops.fq_shift_fwd(W, sd, D, Z, P, B, mode, ht, hr, 0.0, 3.0, False)
            ref = O.shift_forward(w, d, z, shifts, p_ref, 0, 3, name, hard_targets=ht, beta=beta, hard_round=hr)
            if ht and (hr or mode == ops.SHIFT_DEQUANT):
                assert_exact(host(y), ref, f"{name} hard ht={ht} hr={hr}")
            else:
                assert_close(host(y), ref, what=f"{name} ht={ht} hr={hr}")
            ys = ops.fq_shift_fwd(_misaligned(W), sd, D, Z, P, _misaligned(B), mode, ht, hr, 0.0, 3.0, False)
            assert _same_bits(y, ys), f"{name} ht={ht} hr={hr}: vector and scalar kernels differ"
        gp, gb = ops.fq_shift_bwd(GY, W, sd, D, Z, P, B, mode, False, 0.0, 3.0, False, mode == ops.SHIFT_ADASHIFT)
        gp_ref, gb_ref = O.shift_backward(gy, w, d, z, shifts, p_ref, 0, 3, name, beta=beta)
        assert_close(host(gp), gp_ref, rtol=3e-5, what=name + " gp")
        if gb_ref is not None:
            assert_close(host(gb), gb_ref, what=name + " gbeta")
    # slow path: the same float4 / row mixes ordinary and special values
    if oc * ic * kh * kw >= 80:
        ws = w.copy().reshape(-1)
        ws[[3, 17, 40, 60, 70, 75, 76, 77, 78]] = [np.nan, np.inf, -np.inf, 1e30, -3e38, 0.0, -0.0, 1e-40, -1e-30]
        ds = d.copy(); ds[1] = 1e-25
        if oc > 2:
            ds[2] = 1e25
        zs = z.copy(); zs[0] = 1.25
        WS, DS, ZS = dev(ws.reshape(shape)), dev(ds), dev(zs)
        sds = torch.stack([DS.flatten() * s for s in shifts])
        gs = GY.clone(); gs.view(-1)[[3, 17, 40]] = 0.0
        for mode in (ops.SHIFT_DEQUANT, ops.SHIFT_ADASHIFT):
            y = ops.fq_shift_fwd(WS, sds, DS, ZS, P, B, mode, False, False, 0.0, 3.0, False)
            ys = ops.fq_shift_fwd(_misaligned(WS), sds, DS, ZS, P, _misaligned(B), mode, False, False, 0.0, 3.0, False)
            assert _same_bits(y, ys), f"mode {mode}: vector and scalar kernels differ on special values"
            gp, gb = ops.fq_shift_bwd(gs, WS, sds, DS, ZS, P, B, mode, False, 0.0, 3.0, False, mode == ops.SHIFT_ADASHIFT)
            gps, gbs = ops.fq_shift_bwd(_misaligned(gs), _misaligned(WS), sds, DS, ZS, P, _misaligned(B), mode, False, 0.0, 3.0, False,
                                        mode == ops.SHIFT_ADASHIFT)
            fin = torch.isfinite(gps)
            assert bool((torch.isfinite(gp) == fin).all())
            assert_close(host(gp[fin]), host(gps[fin]), rtol=3e-5, what="gp vector vs scalar (special values)")
            if gb is not None:
                assert bool(((gb == gbs) | (torch.isnan(gb) & torch.isnan(gbs))).all()), "gbeta vector vs scalar (special values)"


# ------------------------------------------------------------------------------------------- Adam / loop / affine
def test_adam_and_loop_advance(ops):
    g = golden("adam")
    p = dev(g["p0"]); m = torch.zeros_like(p); v = torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    lr = ops.scalar_dev(1e-3, "cuda")
    idx_table = torch.arange(5 * 4, dtype=torch.int64, device="cuda").reshape(5, 4)
    idx_live = torch.zeros(4, dtype=torch.int64, device="cuda")
    b_table = torch.arange(5, dtype=torch.float32, device="cuda") + 0.5
    b_live = torch.zeros(1, device="cuda")
    for s in range(1, 6):
        ops.loop_advance(step, idx_table, idx_live, b_table, b_live, None, None, 5)
        assert int(step) == s and torch.equal(idx_live, idx_table[s - 1]) and float(b_live) == s - 0.5
        ops.adam_step(p, dev(g[f"g{s}"]), m, v, lr, step)
        assert_close(host(p), g[f"p{s}"], rtol=2e-6, what=f"adam step {s} vs torch.optim.Adam")


def test_fused_iteration_kernels_equal_the_separate_launches(ops):
    """ssq_iter_prologue == loop_advance + gather_rows + fq_adaround_fwd_mt and ssq_fq_adaround_bwd_adam_mt ==
    fq_adaround_bwd_mt + adam_step, bit for bit, over several iterations of a two-layer unit (one layer with a ragged,
    non-vector size); the iteration counter ends every iteration one higher"""
    r = rng(2024)
    dev_ = torch.device("cuda")
    shapes = [(64, 32, 3, 3), (7, 5, 3, 3)]
    iters, batch, n_img = 6, 8, 20
    cache = dev(r.standard_normal((n_img, 4, 6, 6)))
    idx_table = torch.stack([torch.randperm(n_img)[:batch] for _ in range(iters)]).cuda()
    b_table = torch.tensor([0.0, 0.0, 20.0, 14.0, 8.0, 2.0], device=dev_)
    lr_table = torch.full((iters,), 1e-3, device=dev_)

    def unit():
        torch.manual_seed(3)
        sizes = [int(np.prod(s)) for s in shapes]
        pad = [(n + 3) // 4 * 4 for n in sizes]
        flat = torch.zeros(sum(pad), device=dev_); gflat = torch.zeros_like(flat)
        entries, off = [], 0
        for shp, n, pn in zip(shapes, sizes, pad):
            w = torch.randn(shp, device=dev_) * 0.05
            d = (w.abs().amax(dim=(1, 2, 3), keepdim=True) / 1.5).contiguous(); z = torch.full_like(d, 2.0)
            a = flat[off:off + n].view(shp); a.copy_(ops.adaround_init_alpha(w, d))
            entries.append(dict(w=w, alpha=a, delta=d, zero_point=z, wq=torch.empty_like(w), galpha=gflat[off:off + n].view(shp),
                                qmin=0.0, qmax=3.0))
            off += pn
        return flat, gflat, entries, ops.AdaRoundTable(entries)

    res = {}
    for fused in (False, True):
        flat, gflat, entries, table = unit()
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        step = torch.zeros(1, dtype=torch.int64, device=dev_)
        idx_live = torch.zeros(batch, dtype=torch.int64, device=dev_); b_live = torch.zeros(1, device=dev_); lr_live = torch.zeros(1, device=dev_)
        cur = torch.empty((batch,) + tuple(cache.shape[1:]), device=dev_); reg = torch.zeros(1, device=dev_)
        state = ops.IterationState(step, idx_table, idx_live, b_table, b_live, lr_table, lr_live, iters)
        trace = []
        for it in range(iters + 2):                                  # two replays past the schedule: last row repeats
            if fused:
                ops.iter_prologue(state, cache, cur, table, 0.01, reg)
                assert int(step) == it
            else:
                ops.loop_advance(step, idx_table, idx_live, b_table, b_live, lr_table, lr_live, iters)
                ops.gather_rows(cache, idx_live, out=cur)
                table.forward(True, b_live, 0.01, reg)
            row = min(it, iters - 1)
            assert torch.equal(idx_live, idx_table[row]) and float(b_live) == float(b_table[row]) and torch.equal(cur, cache[idx_table[row]])
            gwqs = [torch.randn_like(e["w"]) * (1 + it) for e in entries] if it == 0 else gwqs
            if fused:
                table.backward_adam(gwqs, b_live, 0.01, flat, m, v, lr_live, step, store_grad=True)
            else:
                table.backward(gwqs, b_live, 0.01)
                ops.adam_step(flat, gflat, m, v, lr_live, step)
            assert int(step) == it + 1
            trace.append([reg.clone(), flat.clone(), m.clone(), v.clone(), gflat.clone()] + [e["wq"].clone() for e in entries])
        res[fused] = trace
    for a, b in zip(res[False], res[True]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # the un-fused Adam that ends an iteration (multi-GPU / activation phase): same numbers, counter handed over
    p0 = dev(golden("adam")["p0"]); g1 = dev(golden("adam")["g1"])
    pa, ma, va = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    pb, mb, vb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    sa = torch.ones(1, dtype=torch.int64, device=dev_); sb = torch.zeros(1, dtype=torch.int64, device=dev_)
    lr = ops.scalar_dev(1e-3, dev_)
    ops.adam_step(pa, g1, ma, va, lr, sa)
    ops.adam_step_end_iteration(pb, g1, mb, vb, lr, sb)
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb) and int(sb) == 1


def test_chan_affine(ops):
    r = rng(3)
    x = r.standard_normal((4, 6, 5, 5)).astype(np.float32)
    a = r.standard_normal(6).astype(np.float32); b = r.standard_normal(6).astype(np.float32)
    y = ops.chan_affine_fwd(dev(x), dev(a), dev(b))
    assert_exact(host(y), x * a.reshape(1, 6, 1, 1) + b.reshape(1, 6, 1, 1))
    gy = r.standard_normal(x.shape).astype(np.float32)
    gx, ga, gb = ops.chan_affine_bwd(dev(gy), dev(x), dev(a))
    assert_close(host(gx), gy * a.reshape(1, 6, 1, 1))
    assert_close(host(ga), (gy.astype(np.float64) * x).sum((0, 2, 3)))
    assert_close(host(gb), gy.astype(np.float64).sum((0, 2, 3)))


def test_pull_rows_host_packed_is_lossless(ops):
    """zero-packed host cache: device-side packing == the oracle's packing, and the pulling kernel rebuilds the dense mini-batch
    rows bit for bit from mapped pinned memory (incl. -0.0, denormals, NaN, empty / full chunks, values straddling the 16-byte
    alignment of the packed stream), for every step of the index table incl. the lookahead clamp"""
    r = rng(77)
    n, shape = 40, (3, 32, 32)                                  # 3072 elements = 3 chunks per row
    x = np.maximum(r.standard_normal((n,) + shape), 0).astype(np.float32)
    flat = x.reshape(n, -1)
    flat[1, :1024] = 0.0
    flat[2, 1024:2048] = r.standard_normal(1024).astype(np.float32) + 3.0
    flat[3, 5] = -0.0; flat[3, 6] = np.float32(1e-42); flat[3, 7] = np.nan; flat[3, 8] = -np.inf
    flat[5] = 0.0
    t = torch.from_numpy(x)
    assert ops.packable(t) and not ops.packable(torch.zeros(4, 1000))
    packed = ops.pack_rows_sparse(t, torch.device("cuda"), rows_per_slice=16)
    mask, vals, off = O.sparse_pack_rows(flat)
    assert np.array_equal(packed.mask.cpu().numpy().view(np.uint32), mask)
    assert np.array_equal(packed.vals[:-8].numpy().view(np.uint32), vals.view(np.uint32)) and packed.nnz == vals.size
    assert np.array_equal(packed.chunk_off.cpu().numpy(), off)
    assert packed.vals.is_pinned() and abs(packed.density - vals.size / flat.size) < 1e-12
    steps, batch = 5, 8
    tab = torch.from_numpy(r.integers(0, n, (steps, batch))).cuda()
    tab[0, :6] = torch.tensor([1, 2, 3, 5, 3, 0], device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    dst = torch.full((batch,) + shape, 7.0, device="cuda")
    for s in range(steps + 2):                                  # two replays past the table keep its last row
        step.fill_(s)
        for look in (0, 1):
            dst.fill_(7.0)
            ops.pull_rows_host_packed(packed, tab, step, look, steps, dst, max_ctas=3)
            rows = tab[min(s + look, steps - 1)].cpu().numpy()
            want = flat[rows]
            got = host(dst).reshape(batch, -1)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (s, look)
            if s == 0 and look == 0:
                assert np.array_equal(got.view(np.uint32), O.sparse_unpack_rows(mask, vals, off, rows, flat.shape[1]).view(np.uint32))


def test_bad_arguments_raise(ops):
    from shiftedscalequantization_b200._lib import SsqError
    with pytest.raises(SsqError):
        ops.fq_affine_fwd(torch.zeros(4), torch.ones(1), torch.zeros(1), 0.0, 3.0)          # CPU tensor: no fallback
    with pytest.raises(SsqError):
        ops.fq_affine_fwd(torch.zeros(4, device="cuda").double(), torch.ones(1, device="cuda"),
                          torch.zeros(1, device="cuda"), 0.0, 3.0)


def test_workspace_reuse_across_channel_counts(ops):
    """regression: reductions with different channel counts share one workspace; the ticket header must not move"""
    r = rng(99)
    for rows, k in [(64, 576), (128, 576), (512, 4608), (3, 36), (1000, 512), (128, 1152)]:
        x = torch.as_tensor(r.standard_normal((rows, k)).astype(np.float32)).cuda()
        mn, mx = ops.row_minmax(x)
        assert torch.equal(mn, x.min(1)[0]) and torch.equal(mx, x.max(1)[0]), (rows, k)
        d = (x.abs().amax(1, keepdim=True) / 7).contiguous(); z = torch.full_like(d, 3.0)
        gy = torch.as_tensor(r.standard_normal((rows, k)).astype(np.float32)).cuda()
        _, gd, gz = ops.fq_affine_bwd(gy, x, d, z, 0.0, 15.0)
        _, gd_ref, gz_ref = O.uaq_backward(host(gy), host(x), host(d), host(z), 0, 15)
        assert_close(host(gd), gd_ref, what=f"gdelta {rows}x{k}"); assert_close(host(gz), gz_ref, what=f"gzp {rows}x{k}")
        # per-tensor reduction on the same workspace in between
        _, gd1, _ = ops.fq_affine_bwd(gy, x, d[:1].reshape(()), z[:1].reshape(()), 0.0, 15.0)
        _, gd1_ref, _ = O.uaq_backward(host(gy), host(x), host(d[:1].reshape(())), host(z[:1].reshape(())), 0, 15)
        assert_close(host(gd1), gd1_ref, what="per-tensor gdelta")


def test_no_writes_outside_the_outputs(ops):
    """guard bands (compute-sanitizer is not available on the GPU pool): every output lives in the middle of a canary-filled
    arena; after running each streaming kernel on shapes whose last tile is partial, the canaries must be untouched"""
    CANARY = 12345.678
    arena = torch.full((1 << 22,), CANARY, device='cuda')
    cursor = [1024]

    def out_like(shape, misalign=0):
        n = int(np.prod(shape))
        start = (cursor[0] + 255) // 256 * 256 + misalign        # 1 KiB-aligned (+ optional 4-byte misalignment)
        cursor[0] = start + n + 1024
        return arena[start:start + n].view(shape), (start, n)

    spans = []

    def intact():
        mask = torch.ones_like(arena, dtype=torch.bool)
        for s, n in spans:
            mask[s:s + n] = False
        return bool((arena[mask] == CANARY).all())

    g = torch.Generator(device='cuda').manual_seed(3)
    for shape in [(37, 5, 3, 3), (64, 33, 3, 3), (10, 1000), (3, 4100)]:       # 1665 / 19008 / 10000 / 12300 elements
        for mis in (0, 1):
            w = torch.randn(shape, device='cuda', generator=g) * 0.05
            ps = (shape[0],) + (1,) * (len(shape) - 1)
            d = (w.abs().reshape(shape[0], -1).amax(1) / 2 + 1e-6).view(ps).contiguous()
            z = torch.full_like(d, 2.0)
            alpha = ops.adaround_init_alpha(w, d)
            gy = torch.randn(shape, device='cuda', generator=g)
            n = w.numel()
            lib = ops._lib.load()
            st = torch.cuda.current_stream().cuda_stream
            inner, nchan = ops.channel_layout(w, d)
            # K1a fwd (y + codes), K1b fwd / bwd, straight through the C ABI into arena views
            y, sp = out_like(shape, mis); spans.append(sp)
            c, sp = out_like(shape, mis); spans.append(sp)
            ops._lib.check(lib.ssq_fq_affine_fwd(w.data_ptr(), d.data_ptr(), z.data_ptr(), None, y.data_ptr(), c.data_ptr(), n, inner, nchan, 0.0, 3.0, st), "fwd")
            y2, sp = out_like(shape, mis); spans.append(sp)
            ops._lib.check(lib.ssq_fq_adaround_fwd(w.data_ptr(), alpha.data_ptr(), d.data_ptr(), z.data_ptr(), y2.data_ptr(), None, n, inner, nchan,
                                                   0.0, 3.0, 1, None, 0.0, None, None, 0, st), "ada fwd")
            ga, sp = out_like(shape, mis); spans.append(sp)
            ops.adaround_bwd(gy, w, alpha, d, z, 0.0, 3.0, out=ga)
            # K3 dpred, gather, Adam, import
            dp, sp = out_like(shape, mis); spans.append(sp)
            ws = ops._ws(w, 1, "loss"); loss = torch.empty(1, device='cuda')
            b, per = shape[0], n // shape[0]
            ops._lib.check(lib.ssq_recon_loss(w.data_ptr(), gy.data_ptr(), None, None, loss.data_ptr(), dp.data_ptr(), b, per, float(n / max(shape[1], 1)),
                                              0, 2.0, None, ws.data_ptr(), ws.numel(), st), "loss")
            go, sp = out_like(shape, mis); spans.append(sp)
            idx = torch.randperm(shape[0], device='cuda')
            ops.gather_rows(w, idx, out=go)
            pa, sp = out_like(shape, mis); spans.append(sp)
            m_, sp2 = out_like(shape, mis); spans.append(sp2)
            v_, sp3 = out_like(shape, mis); spans.append(sp3)
            pa.copy_(w); m_.zero_(); v_.zero_()
            ops.adam_step(pa.view(-1), gy.view(-1), m_.view(-1), v_.view(-1), ops.scalar_dev(1e-3, 'cuda'), torch.ones(1, dtype=torch.int64, device='cuda'))
            packed = ops.export_codes(w, d, z, 0.0, 3.0, 2, alpha=alpha)
            wi, sp = out_like(shape, mis); spans.append(sp)
            k = n // shape[0]
            ops._lib.check(lib.ssq_import_codes(packed.data_ptr(), None, d.data_ptr(), z.data_ptr(), wi.data_ptr(), shape[0], k, inner, nchan, 0.0, 2, st), "import")
            torch.cuda.synchronize()
            assert intact(), (shape, mis)
            assert torch.equal(wi, ops.adaround_fwd(w, alpha, d, z, 0.0, 3.0, soft=False)) and torch.equal(go, w[idx])
    # packed export buffer: bytes after the last row must stay zero-initialised canary (uint8 arena)
    barena = torch.full((1 << 16,), 0xAB, dtype=torch.uint8, device='cuda')
    w = torch.randn(7, 3, 3, 3, device='cuda', generator=g) * 0.05          # rows of 27 -> 7 bytes per row at 2 bits
    d = (w.abs().reshape(7, -1).amax(1) / 2).view(7, 1, 1, 1).contiguous(); z = torch.full_like(d, 2.0)
    rb = ops.packed_row_bytes(27, 2)
    view = barena[4096:4096 + 7 * rb]
    ops._lib.check(ops._lib.load().ssq_export_codes(w.data_ptr(), None, None, d.data_ptr(), z.data_ptr(), view.data_ptr(), 7, 27, 27, 7, 0.0, 3.0, 2,
                                                    torch.cuda.current_stream().cuda_stream), "export")
    torch.cuda.synchronize()
    assert bool((barena[:4096] == 0xAB).all()) and bool((barena[4096 + 7 * rb:] == 0xAB).all())
    assert_exact(host(view.view(7, rb)), O.pack_rows(O.uaq_forward(host(w), host(d), host(z), 0, 3)[1], 0, 2), "packed rows of 27")
