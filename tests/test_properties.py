"""Property tests (hypothesis) over ragged shapes, odd bit widths and misaligned views — the edge cases SURVEY.md App. B /
§4(iv) lists. CPU part: the oracle's own invariants. GPU part (-m gpu): kernels against the oracle on generated cases."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import assert_close, assert_exact
from oracle import ssq_oracle as O

SET = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _weights(seed, shape, scale=0.05):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


shapes = st.one_of(
    st.tuples(st.integers(1, 24), st.integers(1, 12), st.sampled_from([1, 3, 5, 7]), st.sampled_from([1, 3, 5, 7])),   # conv OIHW (incl. rows of 9, 27, 147)
    st.tuples(st.integers(1, 40), st.integers(1, 300)))                                                               # linear


# ------------------------------------------------------------------------------------------------ oracle invariants (CPU)
@settings(max_examples=60, **SET)
@given(shape=shapes, bits=st.integers(1, 8), sym=st.booleans(), seed=st.integers(0, 2**16))
def test_oracle_fake_quant_invariants(shape, bits, sym, seed):
    w = _weights(seed, shape)
    L = 2 ** bits
    qmin, qmax = O.bounds(L, sym)
    ps = (shape[0],) + (1,) * (len(shape) - 1)
    d = (np.abs(w).reshape(shape[0], -1).max(1) / max(L / 2, 1) + 1e-6).astype(np.float32).reshape(ps)
    z = (np.zeros_like(d) if sym else np.full_like(d, float(L // 2)))
    y, q = O.uaq_forward(w, d, z, qmin, qmax)
    assert np.all(q == np.rint(q)) and q.min() >= qmin and q.max() <= qmax              # codes are integers inside the grid
    y2, q2 = O.uaq_forward(y, d, z, qmin, qmax)
    assert_exact(q2, q, "re-quantising a dequantised tensor keeps its codes")            # idempotence
    k = int(np.prod(shape[1:]))
    packed = O.pack_rows(q, qmin, bits)
    assert packed.shape == (shape[0], (k * O.storage_bits(bits) + 7) // 8)
    assert_exact(O.unpack_rows(packed, k, qmin, bits).reshape(shape), q, "unpack(pack(codes))")


@settings(max_examples=40, **SET)
@given(shape=shapes, bits=st.integers(2, 8), seed=st.integers(0, 2**16))
def test_oracle_adaround_init_reproduces_the_weight(shape, bits, seed):
    """alpha = init_alpha(w) makes the SOFT forward return w itself wherever it is not clipped (adaptive_rounding.py:66-74),
    and hard rounding of that alpha is round-to-nearest"""
    w = _weights(seed, shape)
    L = 2 ** bits
    ps = (shape[0],) + (1,) * (len(shape) - 1)
    d = (np.abs(w).reshape(shape[0], -1).max(1) / (L / 2) * 1.05 + 1e-6).astype(np.float32).reshape(ps)
    z = np.full_like(d, float(L // 2))
    alpha = O.adaround_init_alpha(w, d)
    y, q = O.adaround_forward(w, alpha, d, z, 0, L - 1, soft=True)
    inside = (q > 0) & (q < L - 1)
    assert np.abs(y - w)[inside].max(initial=0.0) <= 2e-6 * max(np.abs(w).max(), 1e-6) + 1e-7
    _, q_hard = O.adaround_forward(w, alpha, d, z, 0, L - 1, soft=False)
    _, q_near = O.uaq_forward(w, d, z, 0, L - 1)
    frac = np.abs((w / d) - np.floor(w / d) - 0.5)
    assert np.array_equal(q_hard[frac > 1e-3], q_near[frac > 1e-3])                     # away from exact .5 ties


# ------------------------------------------------------------------------------------------------ kernels vs oracle (GPU)
def _dev(a, misalign=False):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if not misalign:
        return t.cuda()
    buf = torch.zeros(t.numel() + 1, device='cuda')          # a view whose pointer is 4- but not 16-byte aligned
    v = buf[1:].view(t.shape)
    v.copy_(t)
    return v


@pytest.mark.gpu
@settings(max_examples=40, **SET)
@given(shape=shapes, bits=st.integers(1, 8), sym=st.booleans(), misalign=st.booleans(), seed=st.integers(0, 2**16))
def test_fq_affine_and_export_match_oracle(shape, bits, sym, misalign, seed):
    from shiftedscalequantization_b200 import ops
    w = _weights(seed, shape)
    L = 2 ** bits
    qmin, qmax = O.bounds(L, sym)
    ps = (shape[0],) + (1,) * (len(shape) - 1)
    d = (np.abs(w).reshape(shape[0], -1).max(1) / max(L / 2, 1) + 1e-6).astype(np.float32).reshape(ps)
    z = (np.zeros_like(d) if sym else np.full_like(d, float(L // 2)))
    y_ref, q_ref = O.uaq_forward(w, d, z, qmin, qmax)
    wd = _dev(w, misalign)
    y, q = ops.fq_affine_fwd(wd, _dev(d), _dev(z), float(qmin), float(qmax), want_codes=True)
    assert_exact(q.cpu().numpy(), q_ref, "codes")
    assert_exact(y.cpu().numpy(), y_ref, "dequantised")
    packed = ops.export_codes(wd, _dev(d), _dev(z), float(qmin), float(qmax), bits)
    assert_exact(packed.cpu().numpy(), O.pack_rows(q_ref, qmin, bits), "packed codes")
    assert_exact(ops.import_codes(packed, shape, _dev(d), _dev(z), float(qmin), bits).cpu().numpy(), y_ref, "import(export)")


@pytest.mark.gpu
@settings(max_examples=30, **SET)
@given(shape=shapes, bits=st.integers(2, 8), misalign=st.booleans(), seed=st.integers(0, 2**16))
def test_adaround_kernels_match_oracle(shape, bits, misalign, seed):
    from shiftedscalequantization_b200 import ops
    w = _weights(seed, shape)
    L = 2 ** bits
    ps = (shape[0],) + (1,) * (len(shape) - 1)
    d = (np.abs(w).reshape(shape[0], -1).max(1) / (L / 2) * 1.05 + 1e-6).astype(np.float32).reshape(ps)
    z = np.full_like(d, float(L // 2))
    alpha = (O.adaround_init_alpha(w, d) + _weights(seed + 1, shape, 0.7)).astype(np.float32)
    g = _weights(seed + 2, shape, 1.0)
    wd, ad = _dev(w, misalign), _dev(alpha, misalign)
    for soft in (True, False):
        y_ref, q_ref = O.adaround_forward(w, alpha, d, z, 0, L - 1, soft=soft)
        y, q = ops.adaround_fwd(wd, ad, _dev(d), _dev(z), 0.0, float(L - 1), soft=soft, want_codes=True)
        if soft:
            assert_close(y.cpu().numpy(), y_ref, rtol=1e-5, what="soft forward")
        else:
            assert_exact(q.cpu().numpy(), q_ref, "hard codes"); assert_exact(y.cpu().numpy(), y_ref, "hard forward")
    ga_ref = O.adaround_backward(g, w, alpha, d, z, 0, L - 1)
    ga = ops.adaround_bwd(_dev(g, misalign), wd, ad, _dev(d), _dev(z), 0.0, float(L - 1))
    assert_close(ga.cpu().numpy(), ga_ref, rtol=2e-5, what="d/d alpha")
