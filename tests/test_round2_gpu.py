"""Round-2 parity tests against goldens produced by the REAL reference (tests/golden/make_golden_round2.py):
  * layer_recon_shiftedScale on a ResNet-50 conv2 layer (BASELINE configs[2]) — shift, AdaRound on top, act=True flavour;
  * block_reconstruction on a ResNet-50 bottleneck and a RegNetX-3200M block (configs[2], configs[4] families);
  * the long horizon: 2 000 and 20 000 iterations (the north_star's per-block budget) of block_reconstruction /
    layer_reconstruction — agreement of the final hard integer codes with the reference's;
  * ChannelQuantAct 'adaround' / 'none' forward modes.
cuDNN and the reference's CPU convolutions differ in the last bits, so trained tensors are compared with a tolerance plus
agreement of the hard decisions (tests/test_shift_modules_gpu.py explains the bounds); everything upstream of the loop
(weights, step sizes, zero points, candidate caches) is bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_close, assert_exact, golden
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu

AQ = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
host = lambda t: t.detach().cpu().numpy()


def _build(arch, bits, cali, **zoo_kw):
    from shiftedscalequantization_b200 import quant as Q, zoo
    torch.manual_seed(1005)
    cnn = zoo.build(arch, **zoo_kw).cuda().eval()
    qnn = Q.QuantModel(cnn, {'n_bits': bits, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).cuda().eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:32].cuda())
    return Q, qnn


def _report(name, payload):
    """numbers worth keeping from a GPU run (merged back through gpurun_out/)"""
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_report.json")
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
    data[name] = payload
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


# ------------------------------------------------------------------------------------------------ layer_recon_shiftedScale
def _layer_shift_setup():
    from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
    g = golden("layer_shift")
    cali = torch.from_numpy(g["cali"])
    Q, qnn = _build("resnet50", 4, cali, num_classes=10)
    layer = qnn.model.layer1[0].conv2
    assert_exact(host(layer.org_weight[:4]), g["probe.conv2_w"], "seeded ResNet-50 weights vs the reference constructor")
    layer.weight_quantizer = ChannelQuant(1.0, uaq=layer.weight_quantizer, weight_tensor=layer.org_weight.data,
                                          shiftTarget=[float(s) for s in g["shifts"]], name=layer.pathName)
    return Q, qnn, layer, cali, g


def _close_params(ours, ref, what, atol=2e-3, lr=1e-3, steps=24):
    ours, ref = host(ours), np.asarray(ref)
    assert ours.shape == ref.shape, what
    err = np.abs(ours - ref)
    bad = int((err > atol).sum())
    allowed = max(int(np.ceil(1e-3 * err.size)), 3)
    assert bad <= allowed, f"{what}: {bad} of {err.size} entries further than {atol} (allowed {allowed})"
    assert err.max() <= 2 * lr * steps, f"{what}: max abs diff {err.max():.3e}"


def test_layer_feature_cache_matches_reference():
    """the 'if' / 'of' cache protocol on a single QuantModule (ShiftedScaleQuant.py:244-255): inputs from the quantised
    prefix, FP outputs — against the reference's own cached tensors"""
    Q, qnn, layer, cali, g = _layer_shift_setup()
    for mode, wq_on in (('if', True), ('of', False)):
        qnn.set_quant_state(wq_on, False)
        layer.cache_features = mode
        with torch.no_grad():
            for i in range(0, cali.shape[0], 16):
                qnn(cali[i:i + 16].cuda())
        layer.cache_features = 'none'
    assert_close(host(torch.cat(layer.cached_inp_features)), g["A.inp"], rtol=1e-4, what="cached inputs")
    assert_close(host(torch.cat(layer.cached_out_features)), g["A.out"], rtol=1e-4, what="cached FP outputs")


@pytest.mark.parametrize("captured", [True, False])
def test_layer_recon_shifted_scale_matches_reference(captured, monkeypatch):
    """quant/layer_recon_shiftedScale.py:262-338 on ResNet-50 layer1.0.conv2, the reference's own cached features and
    torch.randperm stream: shift loop (entropy regulariser), then AdaRound on top (pow regulariser; upstream writes the
    hard switch to `layer.hard_round`, so its 'Hard Round' read-out equals the soft one — reproduced)"""
    from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
    monkeypatch.setattr(LS, "USE_CAPTURED_LOOP", captured)
    Q, qnn, layer, cali, g = _layer_shift_setup()
    iters = int(g["iters"])
    layer.cached_inp_features = [torch.from_numpy(g["A.inp"])]
    layer.cached_out_features = [torch.from_numpy(g["A.out"])]
    qnn.set_quant_state(False, False); layer.set_quant_state(True, False)
    torch.manual_seed(191)
    soft, hard = LS.layer_recon_shiftedScale(layer, iters=iters, lmda=0.01, model=qnn)
    assert_close(np.array([soft, hard]), g["A.shift.losses"], rtol=2e-3, what="[soft, hard] loss after the shift loop")
    q = layer.weight_quantizer
    _close_params(q.alpha, g["A.shift.alpha"], "alpha after the shift loop")
    assert (host(q.alpha).argmax(-1) == g["A.shift.alpha"].argmax(-1)).mean() > 0.98
    assert q.hard_targets and q.shiftedDone
    with torch.no_grad():
        out = host(layer(torch.from_numpy(g["A.inp"][:8]).cuda()))
    assert np.abs(out - g["A.shift.hard_out"]).max() <= 5e-3 * np.abs(g["A.shift.hard_out"]).max()
    torch.manual_seed(192)
    soft, hard = LS.layer_recon_shiftedScale(layer, iters=iters, lmda=0.01, model=qnn, adaround=True)
    assert_close(np.array([soft, hard]), g["A.ada.losses"], rtol=2e-3, what="[soft, hard] loss after AdaRound on the shift")
    assert soft == hard                                                    # the quirk: nothing was hard-rounded
    assert bool(getattr(layer, "hard_round", False)) == bool(g["A.ada.layer_hard_round"])
    assert bool(q.hard_round) == bool(g["A.ada.quantizer_hard_round"])
    # update_delta freezes the argmax of alpha; an argmax that differs from the reference's (a near-tie) moves that input channel
    same = host(q.alpha).argmax(-1) == g["A.shift.alpha"].argmax(-1)
    d_ours, d_ref = host(q.delta), g["A.ada.delta"]
    assert np.array_equal(d_ours[:, same], d_ref[:, same]), "delta after update_delta (channels with the reference's choice)"
    b_ours, b_ref = host(q.beta), g["A.ada.beta"]
    _close_params(torch.from_numpy(b_ours[:, same]), b_ref[:, same], "beta after the AdaRound loop")
    assert (np.sign(b_ours[:, same]) == np.sign(b_ref[:, same])).mean() > 0.999
    with torch.no_grad():
        out = host(layer(torch.from_numpy(g["A.inp"][:8]).cuda()))
    rel = np.abs(out - g["A.ada.out_after"]).max() / np.abs(g["A.ada.out_after"]).max()
    assert rel <= 5e-3 or not same.all(), rel


def test_layer_recon_shifted_scale_act_flag_matches_reference():
    """act=True: upstream builds the loss with round_loss='none' (regulariser off for the whole loop, :283) and still
    optimises the WEIGHT parameters (:270-279)"""
    from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
    Q, qnn, layer, cali, g = _layer_shift_setup()
    layer.cached_inp_features = [torch.from_numpy(g["A.inp"])]
    layer.cached_out_features = [torch.from_numpy(g["A.out"])]
    qnn.set_quant_state(False, False); layer.set_quant_state(True, False)
    torch.manual_seed(191)
    soft, hard = LS.layer_recon_shiftedScale(layer, iters=int(g["iters"]), lmda=0.01, model=qnn, act=True)
    assert_close(np.array([soft, hard]), g["B.shift.losses"], rtol=2e-3, what="[soft, hard] loss, act=True")
    _close_params(layer.weight_quantizer.alpha, g["B.shift.alpha"], "alpha, act=True (no regulariser)")
    # and it is a different trajectory from the regularised one: the flag really switched the regulariser off
    assert np.abs(g["B.shift.alpha"] - g["A.shift.alpha"]).max() > 1e-4


def test_layer_recon_fused_shifted_scale_runs():
    """quant/layer_recon_fused_shiftedScale.py:144-221 raises UnboundLocalError upstream (:156): parity unpinned. The
    single-layer analogue of the (pinned) block function must run, stay finite and leave the hard switches set."""
    from shiftedscalequantization_b200.quant.layer_recon_fused_shiftedScale import layer_recon_fused_shiftedScale
    Q, qnn, layer, cali, g = _layer_shift_setup()
    layer.cached_inp_features = [torch.from_numpy(g["A.inp"])]
    layer.cached_out_features = [torch.from_numpy(g["A.out"])]
    qnn.set_quant_state(False, False); layer.set_quant_state(True, False)
    torch.manual_seed(5)
    soft, hard = layer_recon_fused_shiftedScale(layer, iters=24, lmda=[0.01, 0.01], model=qnn)
    q = layer.weight_quantizer
    assert np.isfinite([soft, hard]).all() and torch.isfinite(q.alpha).all()
    assert q.opt_mode == 'adaShift' and q.hard_targets and q.shiftedDone
    with torch.no_grad():
        assert torch.isfinite(layer(torch.from_numpy(g["A.inp"][:8]).cuda())).all()


# ------------------------------------------------------------------------------------------------ other families
@pytest.mark.parametrize("tag,arch,bits,pick", [("r50", "resnet50", 4, lambda q: q.model.layer1[0]),
                                                ("rx32", "regnetx_3200m", 2, lambda q: q.model.s2.b1)])
def test_block_reconstruction_families_match_reference(tag, arch, bits, pick):
    """the real reference block_reconstruction (16 iterations) on a ResNet-50 bottleneck with downsample and on a
    RegNetX-3200M block with grouped 3x3 convolutions: same seeded network, calibration tensor and index stream"""
    from shiftedscalequantization_b200 import quant as Q
    g = golden("families")
    cali = torch.from_numpy(g[f"{tag}.cali"])
    Q, qnn = _build(arch, bits, cali, **({"num_classes": 10} if arch.startswith("resnet") else {}))
    block = pick(qnn)
    mods = [(n, m) for n, m in block.named_modules() if isinstance(m, Q.QuantModule)]
    assert len(mods) == 4
    for n, m in mods:
        assert_exact(host(m.org_weight.reshape(-1)[:16]), g[f"{tag}.{n}.probe_w"], f"{n}: seeded weights")
        assert_exact(host(m.weight_quantizer.delta), g[f"{tag}.{n}.delta"], f"{n}: delta ('max' init)")
        assert_exact(host(m.weight_quantizer.zero_point), g[f"{tag}.{n}.zp"], f"{n}: zero point")
    torch.manual_seed(277)
    Q.block_reconstruction(qnn, block, cali_data=cali, iters=16, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                           act_quant=False, opt_mode='mse', batch_size=16)
    agree, total = 0, 0
    for n, m in mods:
        a, ref = host(m.weight_quantizer.alpha), g[f"{tag}.{n}.alpha"]
        err = np.abs(a - ref)
        assert (err > 2e-3).sum() <= max(3, int(2e-3 * err.size)), f"{n}: {(err > 2e-3).sum()} of {err.size} alphas off by > 2e-3"
        assert err.max() <= 2 * 16 * 1.05e-3, f"{n}: max alpha distance {err.max():.3e}"
        agree += int((np.sign(a) == np.sign(ref)).sum()); total += a.size
    assert agree / total > 0.9995
    _report(f"families.{tag}", {"alpha_sign_agreement": agree / total, "alphas": total})
    from shiftedscalequantization_b200.quant.data_utils import save_inp_oup_data
    qnn.set_quant_state(False, False); block.set_quant_state(True, False)
    inps, _ = save_inp_oup_data(qnn, block, cali[:16], True, False, 16)
    qnn.set_quant_state(False, False); block.set_quant_state(True, False)
    with torch.no_grad():
        out = host(block(inps[:8]))
    rel = np.linalg.norm(out - g[f"{tag}.hard_out"]) / np.linalg.norm(g[f"{tag}.hard_out"])
    assert rel < 2e-2, f"hard-rounded block output vs reference: relative L2 {rel:.3e}"


# ------------------------------------------------------------------------------------------------ long horizon
def _hard_codes(m):
    """integer codes of the hard AdaRound forward, through the C ABI (ssq_fq_adaround_fwd with codes)"""
    from shiftedscalequantization_b200 import ops
    q = m.weight_quantizer
    _, codes = ops.adaround_fwd(m.org_weight.detach(), q.alpha.detach(), q.delta.detach(), q.zero_point.detach(), 0.0,
                                float(q.n_levels - 1), soft=False, want_codes=True)
    return host(codes).astype(np.uint8)


@pytest.mark.parametrize("iters", [2000, 20000])
def test_long_horizon_code_agreement(iters, monkeypatch):
    """north_star: codes after the full per-block budget vs the reference. The REAL reference loops ran on the CPU for
    `iters` iterations (block_reconstruction on layer1.0, then layer_reconstruction on fc); here the public API runs the
    same flow on the GPU with the same seeds. Bit-identical trajectories are impossible (cuDNN vs CPU convolutions), so the
    figure of merit is the fraction of identical hard integer codes, judged against the reference's agreement with itself
    under another CPU convolution backend (stored in the golden); both are reported (gpurun_out/parity_report.json,
    profiles/r02_parity_report.json, bench.py extra.code_agreement)."""
    from shiftedscalequantization_b200 import quant as Q, zoo
    # a 20 000-iteration trajectory amplifies last-bit differences into different codes, and autotuned cuDNN algorithms differ from
    # run to run (five runs of this test on different boxes: fc 96.8-98.2 %): pin the algorithms so the figure is reproducible
    monkeypatch.setattr(torch.backends.cudnn, "benchmark", False)
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    g = golden("long_horizon")
    torch.manual_seed(1005)
    cnn = zoo.resnet18(num_classes=10).cuda().eval()
    qnn = Q.QuantModel(cnn, {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).cuda().eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(64, 3, 32, 32)
    assert_exact(cali.reshape(-1)[:64].numpy(), g["cali_probe"], "seeded calibration tensor")
    assert_exact(host(qnn.model.conv1.org_weight[:4]), g["probe.conv1_w"], "seeded weights")
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:32].cuda())
    kw = dict(cali_data=cali, iters=iters, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=False,
              opt_mode='mse', batch_size=32)
    block = qnn.model.layer1[0]
    torch.manual_seed(377)
    Q.block_reconstruction(qnn, block, **kw)
    torch.manual_seed(378)
    Q.layer_reconstruction(qnn, qnn.model.fc, **kw)
    rep = {}
    for name, m in (("block.conv1", block.conv1), ("block.conv2", block.conv2), ("fc", qnn.model.fc)):
        codes, ref = _hard_codes(m), g[f"i{iters}.{name}.codes"]
        a_ref = g[f"i{iters}.{name}.alpha16"].astype(np.float32)
        same = codes == ref
        # a reference alpha that finished close to 0 is a decision the reference itself barely made
        rep[name] = {"code_agreement": float(same.mean()), "codes": int(same.size), "differing": int((~same).sum()),
                     "reference_vs_itself": float(g[f"i{iters}.{name}.self_agreement"]),
                     "differing_where_ref_alpha_abs_gt_1": int(((~same) & (np.abs(a_ref) > 1.0)).sum()),
                     "ref_alpha_abs_lt_1": float((np.abs(a_ref) < 1.0).mean())}
    print(f"long horizon, {iters} iterations:", json.dumps(rep))
    _report(f"long_horizon.{iters}", rep)
    # the yardstick: the reference's agreement with ITSELF when only its CPU convolution backend changes (oneDNN -> native),
    # i.e. under the same kind of last-bit differences a cuDNN run has (0.9999 after 2 000 iterations, 0.97-0.985 after 20 000).
    # That yardstick is ONE sample of a chaotic process and so is this run (fc has 5 120 codes: 80 of them are 1.6 %), hence the
    # allowance of 3 points after 20 000 iterations; every differing code must sit where the reference itself had made up its mind
    # (|alpha| > 1), i.e. be a diverged trajectory and not a rounding-boundary disagreement of the kernels
    for name, r in rep.items():
        assert r["code_agreement"] >= r["reference_vs_itself"] - (0.002 if iters == 2000 else 0.03), (name, r)
        if iters == 20000:
            assert r["differing_where_ref_alpha_abs_gt_1"] == r["differing"], (name, r)


# ------------------------------------------------------------------------------------------------ ChannelQuantAct
def test_channelquantact_adaround_mode():
    """quant/channelQuantAct.py:45-54. Hard rounding and 'none' against the real class (golden); the soft branch calls an
    undefined get_soft_round upstream (golden records the AttributeError), here it is AdaRound's soft forward: checked
    against the oracle's restatement of adaptive_rounding.py:50-59 incl. the gradient of beta"""
    from shiftedscalequantization_b200.quant.channelQuantAct import ChannelQuantAct
    from shiftedscalequantization_b200.quant.quant_layer import UniformAffineQuantizer
    g = golden("channelquantact")
    assert bool(g["soft_raises"])
    x = torch.from_numpy(g["x"]).cuda()
    uaq = UniformAffineQuantizer(n_bits=4, channel_wise=False, scale_method='mse', leaf_param=True).cuda()
    uaq(x)
    assert_exact(host(uaq.delta), g["delta"], "activation delta (per-tensor mse search)")
    assert_exact(host(uaq.zero_point), g["zp"], "activation zero point")
    q = ChannelQuantAct(uaq)
    assert_exact(host(q(x)), g["y_none"], "'none' mode")
    q.opt_mode = 'adaround'
    q.beta = torch.nn.Parameter(torch.from_numpy(g["beta"]).cuda())
    q.hard_round = True
    assert_exact(host(q(x)), g["y_hard"], "'adaround' mode, hard rounding")
    q.hard_round = False
    y = q(x)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    d, z = float(g["delta"]), float(g["zp"])
    y_ref, _ = O.adaround_forward(g["x"], g["beta"], np.float32(d), np.float32(z), 0, 15, soft=True)
    assert_close(host(y), y_ref, what="'adaround' mode, soft rounding vs oracle")
    gb_ref = O.adaround_backward(host(gy), g["x"], g["beta"], np.float32(d), np.float32(z), 0, 15)
    assert_close(host(q.beta.grad), gb_ref, what="d/d beta vs oracle")


# ------------------------------------------------------------------------------------------------ feature capture
@pytest.mark.parametrize("act_quant", [False, True])
def test_carried_capture_equals_capture_from_the_image(act_quant):
    """the quantised input and the FP output of every unit carried forward from the previous unit's cached
    tensors (one unit forward each per mini-batch) are bit-identical to the reference's two prefix forwards per mini-batch —
    through a sequential calibration of ResNet-18 (stem, 8 blocks, average-pool + flatten into fc), weight phase and
    activation phase"""
    from test_recon_gpu import build_qnn
    from shiftedscalequantization_b200.quant import data_utils as DU
    Q, qnn, cali = build_qnn(n_cali=64)
    if act_quant:
        qnn.set_quant_state(True, True)
        with torch.no_grad():
            qnn(cali[:32].cuda())
        qnn.disable_network_output_quantization()
    units = [qnn.model.conv1] + [b for l in (qnn.model.layer1, qnn.model.layer2, qnn.model.layer3, qnn.model.layer4) for b in l] + [qnn.model.fc]
    kw = dict(iters=8, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=act_quant, opt_mode='mse', batch_size=16,
              lr=4e-4, p=2.4 if act_quant else 2.0)
    for i, u in enumerate(units):
        qnn.set_quant_state(False, False); u.set_quant_state(True, act_quant)
        ci, co = DU.save_inp_oup_data(qnn, u, cali, True, act_quant, 16)
        mode = DU.capture_stats(qnn).last_mode
        assert mode == ('from_image' if i == 0 else 'carried'), (i, mode)
        DU.reset_capture(qnn, enabled=False)
        ri, ro = DU.save_inp_oup_data(qnn, u, cali, True, act_quant, 16)
        DU.reset_capture(qnn, enabled=True)
        assert torch.equal(ci, ri) and torch.equal(co, ro), f"unit {i}"
        # put the frontier back (the comparison run cleared it), then calibrate the unit as a real run would
        DU._plan(qnn)["frontier"] = DU._Frontier(u, ci, co, DU._cali_key(cali, 16, True, act_quant))
        if i in (1, 2):                                        # a real reconstruction in between, on two of the blocks
            Q.block_reconstruction(qnn, u, cali_data=cali, **kw)


def test_public_api_uses_the_carried_capture_and_host_resident_frees_the_device(monkeypatch):
    """block_reconstruction over consecutive units takes the carried path; host_resident=True (upstream keep_gpu=False) keeps
    the caches in pinned host memory only — peak device memory drops by about the cache size"""
    from test_recon_gpu import build_qnn
    from shiftedscalequantization_b200.quant import data_utils as DU
    kw = dict(iters=8, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=False, opt_mode='mse', batch_size=16)
    peaks = {}
    alphas = {}
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)    # as the reference's drivers (common.py:84-85): no atomics in wgrad
    for host in (False, True):
        Q, qnn, cali = build_qnn(n_cali=256, res=64)
        torch.cuda.synchronize(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        torch.manual_seed(5)
        for u in (qnn.model.layer1[0], qnn.model.layer1[1], qnn.model.layer2[0]):
            Q.block_reconstruction(qnn, u, cali_data=cali, host_resident=host, **kw)
        peaks[host] = torch.cuda.max_memory_allocated() - base
        st = DU.capture_stats(qnn)
        assert st.carried == 2 and st.from_image == 1, (st.carried, st.from_image)
        alphas[host] = qnn.model.layer2[0].conv1.weight_quantizer.alpha.detach().clone()
    cache_bytes = 256 * 64 * 16 * 16 * 4 * 2                   # layer1.x: inputs + outputs of 256 images
    assert peaks[True] < peaks[False] - 0.5 * cache_bytes, peaks
    assert torch.equal(alphas[True], alphas[False])            # same trajectory either way
