"""Integer export / import: packed codes bit-exact against the oracle's restatement of the reference's
`x_quant` intermediates, import(export(w)) bit-identical to the hard forward, model-level round trip."""
import numpy as np
import pytest
import torch

from conftest import assert_exact, golden
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from shiftedscalequantization_b200 import ops as o
    return o


@pytest.mark.parametrize("shape,bits,sym", [((128, 64, 3, 3), 2, False), ((96, 1, 3, 3), 3, False), ((1000, 512), 8, False),
                                            ((64, 3, 7, 7), 8, False), ((32, 16, 3, 3), 4, True), ((16, 64), 1, False),
                                            ((40, 27), 2, False), ((8, 4096), 4, False)])
@pytest.mark.parametrize("rounding", ["nearest", "adaround"])
def test_export_import_vs_oracle(ops, shape, bits, sym, rounding):
    rng = np.random.default_rng(hash((shape, bits, rounding)) % 2**32)
    w = (rng.standard_normal(shape) * 0.05).astype(np.float32)
    L = 2 ** bits
    qmin, qmax = O.bounds(L, sym)
    d = (np.abs(w).reshape(shape[0], -1).max(1) / (L / 2) * 1.1 + 1e-6).astype(np.float32)
    z = np.zeros_like(d) if sym else np.full_like(d, float(L // 2))
    pshape = (shape[0],) + (1,) * (len(shape) - 1)
    if rounding == "adaround":
        alpha = O.adaround_init_alpha(w, d.reshape(pshape)) + (rng.standard_normal(shape) * 0.5).astype(np.float32)
        y_ref, q_ref = O.adaround_forward(w, alpha, d.reshape(pshape), z.reshape(pshape), qmin, qmax, soft=False)
    else:
        alpha = None
        y_ref, q_ref = O.uaq_forward(w, d.reshape(pshape), z.reshape(pshape), qmin, qmax)
    packed = ops.export_codes(dev(w), dev(d.reshape(pshape)), dev(z.reshape(pshape)), float(qmin), float(qmax), bits,
                              alpha=None if alpha is None else dev(alpha))
    ref_packed = O.pack_rows(q_ref, qmin, bits)
    assert packed.shape == ref_packed.shape == (shape[0], ops.packed_row_bytes(int(np.prod(shape[1:])), bits))
    assert_exact(host(packed), ref_packed, "packed codes vs oracle")
    assert_exact(O.unpack_rows(ref_packed, int(np.prod(shape[1:])), qmin, bits).reshape(shape), q_ref, "oracle unpack(pack)")
    y = ops.import_codes(packed, shape, dev(d.reshape(pshape)), dev(z.reshape(pshape)), float(qmin), bits)
    assert_exact(host(y), y_ref, "import(export(w)) vs the reference's hard forward")
    # and against the product's own fake-quant kernels (what the quantised model computes)
    if rounding == "adaround":
        y_k = ops.adaround_fwd(dev(w), dev(alpha), dev(d.reshape(pshape)), dev(z.reshape(pshape)), float(qmin), float(qmax), soft=False)
    else:
        y_k = ops.fq_affine_fwd(dev(w), dev(d.reshape(pshape)), dev(z.reshape(pshape)), float(qmin), float(qmax))
    assert torch.equal(y, y_k)


def test_export_per_input_channel_delta_and_misaligned(ops):
    """delta of shape [OC,IC,1,1] (after ChannelQuant.update_delta, quant/channelQuant.py:221-237,296-298): channels change
    every kh*kw = 9 elements inside one packed word; plus a weight pointer that is not 16-byte aligned (byte path)"""
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((24, 32, 3, 3)) * 0.05).astype(np.float32)
    d = (0.02 + 0.01 * rng.random((24, 32, 1, 1))).astype(np.float32)
    z = np.full((24, 32, 1, 1), 2.0, np.float32)
    beta = rng.standard_normal(w.shape).astype(np.float32)
    fl = np.floor(w / d)
    q_ref = np.clip(fl + (beta >= 0) + z, 0, 3).astype(np.float32)
    y_ref = ((q_ref - z) * d).astype(np.float32)
    packed = ops.export_codes(dev(w), dev(d), dev(z), 0.0, 3.0, 2, alpha=dev(beta))
    assert_exact(host(packed), O.pack_rows(q_ref, 0, 2), "packed codes, per-(oc,ic) delta")
    assert_exact(host(ops.import_codes(packed, w.shape, dev(d), dev(z), 0.0, 2)), y_ref, "import, per-(oc,ic) delta")
    buf = torch.zeros(w.size + 1, device='cuda')
    wm = buf[1:].view(w.shape); wm.copy_(dev(w))
    assert not wm.data_ptr() % 16 == 0
    p2 = ops.export_codes(wm, dev(d), dev(z), 0.0, 3.0, 2, alpha=dev(beta))
    assert torch.equal(p2, packed)


@pytest.mark.parametrize("case", golden("channelquantmse").cases())
@pytest.mark.parametrize("level", [1, 4, 16, 64])
def test_export_channelquantmse_golden(ops, case, level):
    """codes of ChannelQuantMSE.quant / forward (quant/channelQuantMSE.py:126-143) from the real reference"""
    g = golden("channelquantmse").case(case)
    w, d, raw, s = g["w"], g["delta"], g["raw"], g[f"inp_scale_l{level}"]
    bits = int(g["bits"])
    L = 2 ** bits
    zero = np.rint(raw / d).astype(np.float32)
    packed = ops.export_codes(dev(w), dev(d), dev(zero), 0.0, float(L - 1), bits, in_scale=dev(s.reshape(-1)))
    assert_exact(host(packed), O.pack_rows(g[f"codes_l{level}"], 0, bits), "packed ChannelQuantMSE codes vs reference")
    y = ops.import_codes(packed, w.shape, dev(d), dev(zero), 0.0, bits, in_scale=dev(s.reshape(-1)))
    assert_exact(host(y), g[f"y_l{level}"], "import vs reference forward")


def test_export_large_idempotent(ops):
    """BASELINE-size tensor (ResNet-18 layer4 conv, 2.36 M weights) through a size-independent property:
    export(import(export(w))) == export(w), and the packed size is n*bits/8"""
    torch.manual_seed(0)
    w = torch.randn(512, 512, 3, 3, device='cuda') * 0.03
    d = (w.abs().amax(dim=(1, 2, 3), keepdim=True) / 2).contiguous()
    z = torch.full_like(d, 2.0)
    alpha = ops.adaround_init_alpha(w, d) + torch.randn_like(w)
    p1 = ops.export_codes(w, d, z, 0.0, 3.0, 2, alpha=alpha)
    assert p1.numel() == w.numel() // 4
    y = ops.import_codes(p1, w.shape, d, z, 0.0, 2)
    assert torch.equal(y, ops.adaround_fwd(w, alpha, d, z, 0.0, 3.0, soft=False))
    p2 = ops.export_codes(y, d, z, 0.0, 3.0, 2)           # a dequantised weight re-quantises (nearest) to the same codes
    assert torch.equal(p1, p2)


def test_model_round_trip():
    """calibrate two units, export every layer, write the dequantised integers back as plain weights: the float forward
    must reproduce the quantised forward bit for bit; 2-bit layers pack 16 weights per 4 bytes"""
    from test_recon_gpu import build_qnn
    from shiftedscalequantization_b200 import export as E
    Q, qnn, cali = build_qnn()
    kw = dict(cali_data=cali, iters=20, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=False, opt_mode='mse', batch_size=16)
    Q.block_reconstruction(qnn, qnn.model.layer1[0], **kw)
    Q.layer_reconstruction(qnn, qnn.model.fc, **kw)
    qnn.set_quant_state(True, False)
    x = cali[:8].cuda()
    with torch.no_grad():
        ref = qnn(x)
    blob = E.export_int_weights(qnn)
    n_layers = sum(isinstance(m, Q.QuantModule) for m in qnn.modules())
    assert len(blob) - 1 == n_layers == 21
    e = blob['model.layer1.0.conv1']
    assert e['n_bits'] == 2 and e['codes'].shape == (64, 64 * 9 // 4) and e['codes'].dtype == torch.uint8
    assert blob['model.conv1']['n_bits'] == 8 and blob['model.conv1']['codes'].shape == (64, 147)
    total_w = sum(m.weight.numel() for m in qnn.modules() if isinstance(m, Q.QuantModule))
    assert E.packed_bytes(blob) < total_w            # < 1 byte per weight overall (2-bit body, 8-bit stem/head)
    assert E.import_int_weights(qnn, blob) == n_layers
    qnn.set_quant_state(False, False)
    with torch.no_grad():
        out = qnn(x)
    assert torch.equal(out, ref)
    with pytest.raises(Exception):
        q = qnn.model.layer1[0].conv1.weight_quantizer
        q.soft_targets = True
        E.export_int_weights(qnn)


def test_export_shifted_modes():
    """ChannelQuant hard modes (adaShift after the fused loop; adaround after shift) export to one integer grid"""
    from test_shift_modules_gpu import _cache_block_features, _shift_qnn
    from shiftedscalequantization_b200 import export as E
    from shiftedscalequantization_b200.quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale
    Q, qnn, block, cali = _shift_qnn()
    _cache_block_features(Q, qnn, block, cali)
    block_recon_fused_shiftedScale(block, iters=12, lmda=(0.01, 0.01), model=qnn)
    for m in (block.conv1, block.conv2, block.downsample):
        q = m.weight_quantizer
        d_code, d_deq, zp, qmin, qmax, beta, _ = E._describe(q, m.weight)
        from shiftedscalequantization_b200 import ops
        packed = ops.export_codes(m.weight.detach(), d_code.contiguous(), zp.detach(), qmin, qmax, q.n_bits, alpha=beta.detach())
        y = ops.import_codes(packed, m.weight.shape, d_deq.contiguous(), zp.detach(), qmin, q.n_bits)
        assert torch.equal(y, q(m.weight)), m.pathName


def test_export_carries_the_trained_output_affine():
    """bias_cal trains alpha_out / beta_out (gamma^z, varphi^z); the export must carry them: the import folds them into
    weight and bias of the target (QuantModule FP path or plain float model; equal up to fp32 rounding)"""
    from test_recon_gpu import build_qnn
    from shiftedscalequantization_b200 import export as E, zoo
    Q, qnn, cali = build_qnn()
    block = qnn.model.layer1[0]
    Q.block_reconstruction(qnn, block, cali_data=cali, iters=20, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                           act_quant=False, opt_mode='mse', batch_size=16, bias_cal=True)
    assert not block.conv1._output_affine_is_identity()
    qnn.set_quant_state(True, False)
    x = cali[:8].cuda()
    with torch.no_grad():
        ref = qnn(x)
    blob = E.export_int_weights(qnn)
    assert "out_scale" in blob["model.layer1.0.conv1"] and "out_scale" not in blob["model.layer2.0.conv1"]
    # (a) back into a fresh QuantModel of the same network, quantisers off: the affine is folded into org_weight / org_bias
    Q2, qnn2, _ = build_qnn()
    E.import_int_weights(qnn2, blob)
    qnn2.set_quant_state(False, False)
    with torch.no_grad():
        out2 = qnn2(x)
    assert torch.allclose(out2, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
    # (b) into a plain float network (BN folded the same way): affine folded into weight / bias
    torch.manual_seed(1005)
    from shiftedscalequantization_b200.quant.fold_bn import search_fold_and_remove_bn
    cnn = zoo.resnet18().cuda().eval()
    search_fold_and_remove_bn(cnn)
    wrapper = torch.nn.Module(); wrapper.model = cnn
    # the FP network keeps the shortcut as Sequential(conv, folded bn): the quantised block's `downsample` layer is its `.0`
    plain = {(k + ".0" if k.endswith(".downsample") else k): v for k, v in blob.items()}
    E.import_int_weights(wrapper, plain)
    with torch.no_grad():
        out = cnn(x)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
