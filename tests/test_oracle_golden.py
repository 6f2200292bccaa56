"""CPU suite: the numpy oracle (oracle/ssq_oracle.py) against golden vectors produced by the real
reference (tests/golden/make_golden.py). Integer codes bit-exact; floats <= 1e-5 relative."""
import numpy as np
import pytest

from conftest import assert_close, assert_exact, golden
from oracle import ssq_oracle as O


@pytest.mark.parametrize("case", golden("uaq").cases())
def test_uaq_forward_backward(case):
    g = golden("uaq").case(case)
    bits, sym, cw, is_mse = (int(v) for v in g["meta"])
    qmin, qmax = O.bounds(2 ** bits, bool(sym))
    y, codes = O.uaq_forward(g["x"], g["delta"], g["zp"], qmin, qmax)
    assert_exact(codes, g["codes"], "codes")
    assert_close(y, g["y"], what="dequant")
    gx, gd, gz = O.uaq_backward(g["gy"], g["x"], g["delta"], g["zp"], qmin, qmax)
    assert_close(gx, g["gx"], what="gx")
    assert_close(gd, g["gdelta"], rtol=2e-5, what="gdelta")
    assert_close(gz, g["gzp"], rtol=2e-5, what="gzp")


@pytest.mark.parametrize("case", golden("uaq").cases())
def test_uaq_scale_init(case):
    g = golden("uaq").case(case)
    bits, sym, cw, is_mse = (int(v) for v in g["meta"])
    x = g["x"]
    if is_mse:
        d, z, raw, _ = O.mse_search(x, bits, bool(sym), channel_wise=bool(cw))
    else:
        rows = x.reshape(x.shape[0], -1)
        res = [O.max_init(r, bits, bool(sym)) for r in rows]
        d, z, raw = (np.array([r[j] for r in res], dtype=np.float32) for j in range(3))
    assert_exact(d.reshape(g["delta"].shape), g["delta"], "delta")
    assert_exact(z.reshape(g["zp"].shape), g["zp"], "zero_point")
    assert_exact(raw.reshape(g["raw"].shape), g["raw"], "raw_zero_point")


@pytest.mark.parametrize("case", golden("adaround").cases())
def test_adaround(case):
    g = golden("adaround").case(case)
    L = 2 ** int(g["bits"])
    assert_close(O.adaround_init_alpha(g["w"], g["delta"]), g["alpha0"], rtol=2e-5, what="alpha init")
    assert_close(O.rect_sigmoid(g["alpha"]), g["h"], what="h(alpha)")
    wq, _ = O.adaround_forward(g["w"], g["alpha"], g["delta"], g["zp"], 0, L - 1, soft=True)
    assert_close(wq, g["wq_soft"], what="soft forward")
    wqh, codes = O.adaround_forward(g["w"], g["alpha"], g["delta"], g["zp"], 0, L - 1, soft=False)
    assert_exact(codes, g["codes_hard"], "hard codes")
    assert_exact(wqh, g["wq_hard"], "hard dequant")
    assert_close(O.adaround_backward(g["gw"], g["w"], g["alpha"], g["delta"], g["zp"], 0, L - 1), g["galpha"], what="galpha")
    for b in (20, 11.3, 2.0):
        assert_close(O.round_reg(g["alpha"], b, 0.01), g[f"reg_b{b}"], rtol=2e-5, what=f"reg b={b}")
        assert_close(O.round_reg_grad(g["alpha"], b, 0.01), g[f"greg_b{b}"], rtol=2e-5, what=f"greg b={b}")


@pytest.mark.parametrize("case", ["conv", "fc"])
def test_losses(case):
    g = golden("loss").case(case)
    for p in (2.0, 2.4):
        l, d = O.lp_loss(g["pred"], g["tgt"], p)
        assert_close(l, g[f"lp{p}"], what=f"lp {p}")
        assert_close(d, g[f"dlp{p}"], what=f"dlp {p}")
    if case == "conv":
        l, d = O.fisher_diag_loss(g["pred"], g["tgt"], g["fisher"])
        assert_close(l, g["fdiag"], what="fisher_diag"); assert_close(d, g["dfdiag"], what="d fisher_diag")
        l, d = O.fisher_full_loss(g["pred"], g["tgt"], g["fisher"])
        assert_close(l, g["ffull"], what="fisher_full"); assert_close(d, g["dffull"], rtol=2e-5, what="d fisher_full")


def test_temperature_schedule():
    g = golden("loss")
    got = [float(O.linear_temp_decay(int(t), 200, 0.2, 20, 2)) for t in g["temp.t"]]
    assert got == [float(v) for v in g["temp.b"]]


@pytest.mark.parametrize("case", golden("channelquant").cases())
def test_channelquant(case):
    g = golden("channelquant").case(case)
    L = 2 ** int(g["bits"]); shifts = [float(s) for s in g["shifts"]]
    w, d, z = g["w"], g["delta"], g["zp"]
    y_none, _ = O.uaq_forward(w, d, z, 0, L - 1)
    assert_exact(y_none, g["y_none"], "'none' forward")
    assert_exact(O.shift_terms(w, d, z, shifts, 0, L - 1, "dequant"), g["xq"], "x_q (init_v)")
    p = O.shift_probs(g["alpha"])
    assert_close(p, g["p"], what="p")
    assert_close(O.shift_forward(w, d, z, shifts, p, 0, L - 1, "dequant"), g["y_soft"], what="soft mixture")
    assert_exact(O.shift_forward(w, d, z, shifts, g["p"], 0, L - 1, "dequant", hard_targets=True), g["y_hard"], "hard select")
    gp, _ = O.shift_backward(g["gy"], w, d, z, shifts, p, 0, L - 1, "dequant")
    assert_close(O.shift_probs_backward(g["alpha"], gp), g["galpha_soft"], rtol=3e-5, what="galpha")
    ent, dent = O.entropy_reg(g["alpha"], 0.7)
    assert_close(ent, g["ent"], what="entropy")
    assert_close(O.shift_probs_backward(g["alpha"], dent), g["gent"], rtol=3e-5, what="d entropy")
    # adaround mode on the selected delta ([OC,IC,1,1] / [OC,IC])
    dd = g["ar_delta"].reshape(g["ar_delta"].shape[:2] + (1,) * (w.ndim - 2))
    zb = np.broadcast_to(z.reshape((-1,) + (1,) * (w.ndim - 1)), w.shape)
    fl = np.floor(w / dd)
    q = np.clip((fl + O.rect_sigmoid(g["ar_beta"])) + zb, 0, L - 1).astype(np.float32)
    assert_close((q - zb) * dd, g["ar_y"], what="adaround-after-shift soft")
    # adaShift
    assert_exact(O.shift_terms(w, d, z, shifts, 0, L - 1, "floor"), g["as_xq"], "x_q (init_v_beta)")
    pa = O.shift_probs(g["as_alpha"])
    assert_close(O.shift_forward(w, d, z, shifts, pa, 0, L - 1, "adashift", beta=g["as_beta"]), g["as_y"], what="adaShift soft")
    assert_exact(O.shift_forward(w, d, z, shifts, pa, 0, L - 1, "adashift", hard_targets=True, beta=g["as_beta"], hard_round=True),
                 g["as_y_hard"], "adaShift hard")
    gp, gbeta = O.shift_backward(g["gy"], w, d, z, shifts, pa, 0, L - 1, "adashift", beta=g["as_beta"])
    assert_close(O.shift_probs_backward(g["as_alpha"], gp), g["as_galpha"], rtol=3e-5, what="adaShift galpha")
    assert_close(gbeta, g["as_gbeta"], what="adaShift gbeta")
    for b2 in (20, 7.7):
        r, dr = O.pow_reg_on_probs(g["as_alpha"], b2, 0.3)
        assert_close(r, g[f"as_regS_b{b2}"], rtol=2e-5, what="regS")
        assert_close(O.shift_probs_backward(g["as_alpha"], dr), g[f"as_gregS_b{b2}"], rtol=3e-5, what="d regS")


@pytest.mark.parametrize("case", golden("channelquantmse").cases())
def test_channelquantmse(case):
    g = golden("channelquantmse").case(case)
    L = 2 ** int(g["bits"])
    for level in (1, 4, 16, 64):
        s = O.inp_scale_search(g["w"], g["delta"], g["raw"], L, level, float(g[f"thr_l{level}"]))
        assert_exact(s, g[f"inp_scale_l{level}"], f"inp_scale level {level}")
        y, codes = O.channelquantmse_forward(g["w"], g["delta"], g["raw"], s, L)
        assert_exact(codes, g[f"codes_l{level}"], "codes")
        assert_close(y, g[f"y_l{level}"], what="dequant")


def test_adam():
    g = golden("adam")
    p = g["p0"]; m = np.zeros_like(p); v = np.zeros_like(p)
    for s in range(1, 6):
        p, m, v = O.adam_step(p, g[f"g{s}"], m, v, s)
        assert_close(p, g[f"p{s}"], rtol=2e-6, what=f"adam step {s}")


@pytest.mark.parametrize("bits", [1, 2, 3, 4, 8])
def test_pack_unpack_round_trip(bits):
    """oracle bit-packing of the integer codes (layout of include/ssq_b200.h, export section): ragged rows, all widths"""
    rng = np.random.default_rng(bits)
    for rows, k in [(5, 9), (3, 27), (2, 147), (4, 64), (1, 1)]:
        q = rng.integers(0, 2 ** bits, size=(rows, k)).astype(np.float32)
        packed = O.pack_rows(q, 0, bits)
        sb = O.storage_bits(bits)
        assert packed.shape == (rows, (k * sb + 7) // 8) and packed.dtype == np.uint8
        assert np.array_equal(O.unpack_rows(packed, k, 0, bits), q)
    # known answer: 2-bit codes 1,2,3,0 -> 0b00_11_10_01; symmetric codes are stored offset by qmin
    assert O.pack_rows(np.array([[1, 2, 3, 0]], np.float32), 0, 2)[0, 0] == 0b00111001
    assert O.pack_rows(np.array([[-2, -1, 0, 1]], np.float32), -2, 2)[0, 0] == 0b11100100
    # adaround codes of the reference (golden) survive the round trip
    g = golden("adaround").case(golden("adaround").cases()[0])
    codes = g["codes_hard"]
    nb = int(np.ceil(np.log2(codes.max() + 1)))
    assert np.array_equal(O.unpack_rows(O.pack_rows(codes, 0, nb), codes[0].size, 0, nb).reshape(codes.shape), codes)


def test_mse_search_vs_reference_numpy_restatement():
    """the reference ships its own NumPy version of the clip search (myQuant.py:6-45, 4 bits); its outputs on a seeded
    weight (tests/golden/make_golden_myquant.py) must agree with the oracle that the kernels are held to: same candidate
    per channel (so the same zero point) and the same step size to fp32 rounding"""
    g = golden("myquant")
    d, z, raw, idx = O.mse_search(g["w"], 4)
    assert_exact(z.astype(np.float64), g["zero_point"], "zero point vs myQuant")
    assert_close(d.astype(np.float64), g["delta"], rtol=2e-7, what="delta vs myQuant")


def test_sparse_row_packing_round_trip():
    """the zero-packed host cache format (no reference counterpart): pack -> unpack is the identity bit for bit, incl. -0.0,
    denormals, NaN payloads, all-zero and all-non-zero chunks"""
    from oracle import ssq_oracle as O
    r = np.random.default_rng(5)
    x = np.maximum(r.standard_normal((6, 3072)), 0).astype(np.float32)       # post-ReLU: about half zeros
    x[1, :1024] = 0.0                                                          # an empty chunk
    x[2, 1024:2048] = r.standard_normal(1024).astype(np.float32) + 3.0        # a full chunk
    x[3, 5] = -0.0; x[3, 6] = np.float32(1e-42); x[3, 7] = np.nan; x[3, 8] = -np.inf
    mask, vals, off = O.sparse_pack_rows(x)
    assert mask.shape == (6, 96) and off.shape == (6 * 3 + 1,) and off[-1] == vals.size == int((x.view(np.uint32) != 0).sum())
    assert off[4] - off[3] == 0 and off[8] - off[7] == 1024
    rows = [4, 1, 3, 3, 0, 2]
    y = O.sparse_unpack_rows(mask, vals, off, rows, 3072)
    assert np.array_equal(y.view(np.uint32), x[rows].view(np.uint32))


def test_oracle_loop_rec_loss_matches_reference_lossfunction():
    """oracle/ref_loop_torch.rec_loss — the reconstruction term the oracle LOOP uses (mse / fisher_diag / fisher_full,
    quant/block_recon.py:154-162) — against the values and gradients the real LossFunction produced (tests/golden/loss.npz);
    the GPU test test_fisher_engine_vs_oracle_loop leans on it"""
    import torch
    from oracle import ref_loop_torch as R
    g = golden("loss").case("conv")
    tgt, fisher = torch.from_numpy(g["tgt"]), torch.from_numpy(g["fisher"])
    for mode, val, grad, p in (("mse", "lp2.0", "dlp2.0", 2.0), ("mse", "lp2.4", "dlp2.4", 2.4),
                               ("fisher_diag", "fdiag", "dfdiag", 2.0), ("fisher_full", "ffull", "dffull", 2.0)):
        pred = torch.from_numpy(g["pred"]).clone().requires_grad_(True)
        loss = R.rec_loss(pred, tgt, fisher if mode != "mse" else None, mode, p)
        loss.backward()
        assert_close(loss.detach().numpy(), g[val], what=f"{mode} p={p}")
        assert_close(pred.grad.numpy(), g[grad], rtol=2e-5, what=f"d {mode} p={p}")


@pytest.mark.parametrize("case", golden("shift_candidates").cases())
def test_oracle_init_shift_candidates_matches_reference(case):
    """oracle restatement of ChannelQuant.init_shift_candidates (quant/channelQuant.py:240-277) against the real class"""
    g = golden("shift_candidates").case(case)
    got = O.init_shift_candidates(g["w"], g["delta"], g["zp"], 2 ** int(g["bits"]))
    assert [float(s) for s in got] == [float(s) for s in g["shiftTarget"]], (got, g["shiftTarget"])


def test_oracle_channelquantact_modes_match_reference():
    """ChannelQuantAct (quant/channelQuantAct.py:45-54): 'none' = the plain UAQ forward, 'adaround' with hard rounding = the AdaRound
    hard forward on the activation with the caller's beta; the oracle's two functions against the real class (channelquantact.npz)"""
    g = golden("channelquantact")
    assert bool(g["soft_raises"])                                    # upstream's soft branch calls an undefined get_soft_round
    y_none, _ = O.uaq_forward(g["x"], g["delta"], g["zp"], 0, 15)
    assert_exact(y_none, g["y_none"], "'none' mode")
    y_hard, _ = O.adaround_forward(g["x"], g["beta"], g["delta"], g["zp"], 0, 15, soft=False)
    assert_exact(y_hard, g["y_hard"], "'adaround' mode, hard rounding")
