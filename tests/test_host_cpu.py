"""CPU suite for the host logic: C-ABI surface, model refactoring, schedules, RNG stream, loud failure
without CUDA, the oracle's loop restatement against the real reference's loops, 2-rank gloo plumbing."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_close, assert_exact, golden

WQ = {'n_bits': 2, 'channel_wise': True, 'scale_method': 'mse'}
AQ = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}


# ------------------------------------------------------------------------------------------- C-ABI
def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ssq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ssq_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from shiftedscalequantization_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ssq_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert _lib.load().ssq_abi_version() == 1
    assert _lib.load().ssq_ws_bytes(64) > 0
    assert b"NULL" in _lib.load().ssq_status_string(-1)


def test_argument_errors_are_reported_without_a_gpu():
    """argument validation happens before any CUDA call, so it can be exercised here"""
    from shiftedscalequantization_b200 import _lib
    lib = _lib.load()
    assert lib.ssq_fq_affine_fwd(None, None, None, None, None, None, 16, 4, 4, 0.0, 3.0, None) == -1      # NULL
    assert lib.ssq_fq_affine_fwd(None, None, None, None, None, None, 0, 4, 4, 0.0, 3.0, None) == 0        # empty is a no-op
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.ssq_fq_affine_fwd(p, p, p, None, p, None, 15, 4, 4, 0.0, 3.0, None) == -2                  # n % inner
    assert lib.ssq_mse_scale_search(p, 1, 16, 1000, 0, 2.4, p, p, p, None, None, None, 0, None) == -2     # n_levels
    assert lib.ssq_recon_loss(p, p, None, None, p, None, 2, 8, 4.0, 7, 2.0, None, None, 0, None) == -4     # mode


def test_host_side_sizing_functions_of_the_abi():
    """the ABI's pure host functions (no CUDA call inside): the exchange kernel's shard plan covers the flat buffer exactly once,
    workspaces grow with the problem, packed rows are byte-aligned, argument errors of the round-2 entry points"""
    from shiftedscalequantization_b200 import _lib
    lib = _lib.load()
    pad = lib.ssq_exchange_pad_bytes()
    assert pad > 0 and pad % 4 == 0
    for n in (0, 4, 8, 1000, 4096, 73728, 4718592, 15232128):
        for world in (1, 2, 3, 4, 8, 16):
            shard = lib.ssq_exchange_shard_elems(n, world)
            assert shard % 4 == 0 and shard * world >= n                   # every vector has an owner
            if n:
                assert shard * (world - 1) < n + 4 * world                 # ... and the last rank's shard starts inside (or at the end of) the buffer
            # ranks' ranges [r*shard, min((r+1)*shard, n)) are disjoint and cover [0, n)
            covered = sum(max(0, min((r + 1) * shard, n) - min(r * shard, n)) for r in range(world))
            assert covered == n
    assert lib.ssq_exchange_shard_elems(-4, 2) == 0 and lib.ssq_exchange_shard_elems(16, 0) == 0
    assert lib.ssq_inp_scale_search_ws_bytes2(512, 2304) > 2 * 2304 * 4 + 512 * 16
    assert lib.ssq_inp_scale_search_ws_bytes2(4096, 36864) > lib.ssq_inp_scale_search_ws_bytes2(512, 2304)
    assert lib.ssq_inp_scale_search_ws_bytes(2304) >= lib.ssq_inp_scale_search_ws_bytes2(65536, 2304)
    assert lib.ssq_mse_scale_search_ws_bytes(1, 51380224) > 0
    for bits, per in ((1, 8), (2, 4), (4, 2), (8, 1)):
        for k in (1, 9, 147, 576, 4608):
            assert lib.ssq_packed_row_bytes(k, bits) == (k + per - 1) // per
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    # K2b: NULL pointers, a workspace that is too small, empty problems
    assert lib.ssq_inp_scale_search(None, p, p, p, 16, 3.0, -0.25, 1.25, p, 4, 4, p, 1 << 20, None) == -1
    assert lib.ssq_inp_scale_search(p, p, p, p, 16, 3.0, -0.25, 1.25, p, 4, 4, p, 8, None) != 0
    assert lib.ssq_inp_scale_search(p, p, p, p, 16, 3.0, -0.25, 1.25, p, 0, 4, None, 0, None) == 0
    # exchange: world out of range, a buffer length that is not a multiple of four floats
    arr = (ctypes.c_void_p * 2)(p, p)
    assert lib.ssq_grad_exchange_adam(arr, arr, arr, 0, 17, 16, p, p, p, p, 0.9, 0.999, 1e-8, None, p, p, p, 1 << 20, None) == -2
    assert lib.ssq_grad_exchange_adam(arr, arr, arr, 0, 2, 6, p, p, p, p, 0.9, 0.999, 1e-8, None, p, p, p, 1 << 20, None) == -2
    assert lib.ssq_grad_exchange_adam(None, arr, arr, 0, 2, 8, p, p, p, p, 0.9, 0.999, 1e-8, None, p, p, p, 1 << 20, None) == -1


def test_quantiser_refuses_cpu_tensors():
    from shiftedscalequantization_b200 import ops
    from shiftedscalequantization_b200._lib import SsqError
    from shiftedscalequantization_b200.quant.quant_layer import UniformAffineQuantizer
    q = UniformAffineQuantizer(n_bits=4, channel_wise=True, scale_method='mse')
    with pytest.raises(SsqError, match="no CPU fallback"):
        q(torch.randn(4, 3, 3, 3))
    with pytest.raises(SsqError):
        ops.recon_loss(torch.randn(2, 3), torch.randn(2, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shiftedscalequantization_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r"#.*", "", src).replace("ssq_oracle", "oracle") or "import" not in \
                    "".join(l for l in src.splitlines() if "oracle" in l), f"{f} references oracle/"


# ------------------------------------------------------------------------------------------- model graph
def _qnn(arch="resnet18", **kw):
    from shiftedscalequantization_b200 import quant as Q, zoo
    torch.manual_seed(1005)
    cnn = zoo.build(arch, **kw).eval()
    return Q, Q.QuantModel(cnn, dict(WQ), dict(AQ))


def test_resnet18_refactor_matches_survey_counts():
    Q, qnn = _qnn()
    from shiftedscalequantization_b200.quant.quant_block import QuantBasicBlock
    from shiftedscalequantization_b200.quant.quant_layer import StraightThrough
    mods = [m for m in qnn.modules() if isinstance(m, Q.QuantModule)]
    blocks = [m for m in qnn.modules() if isinstance(m, Q.BaseQuantBlock)]
    assert len(mods) == 21 and len(blocks) == 8 and all(isinstance(b, QuantBasicBlock) for b in blocks)
    assert sum(m.weight.numel() for m in mods) == 11_678_912 and sum(m.weight.shape[0] for m in mods) == 5800
    assert isinstance(qnn.model.bn1, StraightThrough) and isinstance(qnn.model.relu, StraightThrough)
    assert isinstance(qnn.model.conv1.activation_function, torch.nn.ReLU)
    assert qnn.model.layer1[0].conv1.pathName == '.layer1.0.conv1' and qnn.model.fc.pathName == '.fc'
    qnn.set_first_last_layer_to_8bit()
    assert mods[0].ignore_reconstruction and mods[0].weight_quantizer.n_bits == 8 and mods[-1].weight_quantizer.n_levels == 256
    assert mods[-2].act_quantizer.n_bits == 8
    qnn.disable_network_output_quantization()
    assert mods[-1].disable_act_quant
    # FP forward works on CPU (no quantiser kernel involved) and BN folding kept the function
    torch.manual_seed(1005)
    from shiftedscalequantization_b200 import zoo
    ref = zoo.resnet18().eval()
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        assert torch.allclose(qnn(x), ref(x), atol=1e-4)
    qnn.set_quant_state(True, True)
    assert all(m.use_weight_quant and m.use_act_quant for m in mods)
    qnn.store_quantization_state(); qnn.set_quant_state(False, False); qnn.restore_quantization_state()
    assert all(m.use_weight_quant for m in mods) and not any(m.use_act_quant for m in mods)


@pytest.mark.parametrize("arch,n_units,n_qm", [("resnet50", 18, 54), ("mobilenetv2", 20, 53), ("regnetx_600m", 18, 54), ("regnetx_3200m", 27, 81)])
def test_other_families_build(arch, n_units, n_qm):
    Q, qnn = _qnn(arch)
    mods = [m for m in qnn.modules() if isinstance(m, Q.QuantModule)]
    blocks = [m for m in qnn.modules() if isinstance(m, Q.BaseQuantBlock)]
    in_blocks = {id(m) for b in blocks for m in b.modules() if isinstance(m, Q.QuantModule)}
    units = len(blocks) + sum(1 for m in mods if id(m) not in in_blocks)
    assert (units, len(mods)) == (n_units, n_qm)
    assert all(b.pathName for b in blocks)


def test_zoo_seeded_init_equals_reference_probe():
    from shiftedscalequantization_b200 import quant as Q, zoo
    g = golden("recon_loop")
    torch.manual_seed(1005)
    qnn = Q.QuantModel(zoo.resnet18(num_classes=10).eval(), dict(WQ, scale_method='max'), dict(AQ))
    assert_exact(qnn.model.conv1.org_weight[:4].numpy(), g["probe.conv1_w"], "stem weights")
    assert_exact(qnn.model.layer4[1].conv2.org_weight[:2].numpy(), g["probe.l4_w"], "layer4 weights")
    assert_exact(torch.randn(32, 3, 16, 16).numpy(), g["cali"], "calibration tensor")


# ------------------------------------------------------------------------------------------- schedules / RNG
def test_schedules_follow_the_reference():
    from shiftedscalequantization_b200.engine import brecq_b_table, cosine_lr_table, index_table, temperature
    g = golden("loss")
    assert [float(temperature(int(t), 200, 0.2, 20, 2)) for t in g["temp.t"]] == [float(v) for v in g["temp.b"]]
    tab = brecq_b_table(200, 0.2, (20, 2), True)
    assert float(tab[38]) == 0.0 and float(tab[39]) == 20.0                 # count = i+1 reaches loss_start = 40 at i = 39
    assert float(tab[199]) == 2.0 and brecq_b_table(10, 0.2, (20, 2), False).abs().sum() == 0
    assert_close(cosine_lr_table(4e-4, 50).numpy(), golden("adam")["cosine_lr"], rtol=1e-6, what="cosine lr")
    torch.manual_seed(5); a = index_table(100, 32, 6)
    torch.manual_seed(5); b = torch.stack([torch.randperm(100)[:32] for _ in range(6)])
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------- oracle loop vs reference loop
def _golden_block_unit(g):
    L = {}
    for n, act in (("conv1", "relu"), ("conv2", None)):
        L[n] = dict(weight=torch.from_numpy(g[f"block.{n}.weight"]), bias=torch.from_numpy(g[f"block.{n}.bias"]),
                    conv=dict(stride=1, padding=1, dilation=1, groups=1), act=act, delta=torch.from_numpy(g[f"block.{n}.delta"]),
                    zero_point=torch.from_numpy(g[f"block.{n}.zp"]), n_levels=4)
    return {"kind": "basic", "layers": L, "tail_act": "relu"}


def test_oracle_loop_reproduces_reference_block_loop():
    from oracle import ref_loop_torch as R
    g = golden("recon_loop")
    unit = _golden_block_unit(g)
    alphas, losses = R.recon_weight_loop(unit, torch.from_numpy(g["block.inps"]), torch.from_numpy(g["block.outs"]),
                                         torch.from_numpy(g["block.idx"]), 12, weight=0.01, b_range=(20, 2), warmup=0.2)
    assert_close(np.array(losses), g["block.losses"], rtol=1e-6, what="loss trace")
    for n in ("conv1", "conv2"):
        assert_close(alphas[n].detach().numpy(), g[f"block.{n}.alpha"], rtol=1e-6, what=f"alpha {n}")


def test_oracle_loop_reproduces_reference_layer_reconstruction():
    """also proves the index stream: the table is regenerated from the seed, not stored"""
    from oracle import ref_loop_torch as R
    from shiftedscalequantization_b200.engine import index_table
    g = golden("recon_loop")
    unit = {"kind": "layer", "layers": {"fc": dict(weight=torch.from_numpy(g["fc.weight"]), bias=torch.from_numpy(g["fc.bias"]),
            conv=None, act=None, delta=torch.from_numpy(g["fc.delta"]), zero_point=torch.from_numpy(g["fc.zp"]), n_levels=256)}}
    # inputs of fc = pooled features; outputs = FP logits: regenerate them with the FP weights of the golden run
    from shiftedscalequantization_b200 import quant as Q, zoo
    torch.manual_seed(1005)
    cnn = zoo.resnet18(num_classes=10).eval()
    feats = []
    h = cnn.avgpool.register_forward_hook(lambda m, i, o: feats.append(torch.flatten(o, 1)))
    # the golden run's fc input comes from the quantised prefix (asym=True), which needs the GPU path; here we only
    # check the loop mechanics on the FP features (GPU test test_full_api_matches_reference_loops covers the rest)
    with torch.no_grad():
        logits = cnn(torch.from_numpy(g["cali"]))
    h.remove()
    torch.manual_seed(78)
    tab = index_table(32, 16, 12)
    alphas, losses = R.recon_weight_loop(unit, feats[0], logits, tab, 12, weight=0.01, b_range=(20, 2), warmup=0.2)
    assert np.isfinite(losses).all() and alphas["fc"].shape == (10, 512)


# ------------------------------------------------------------------------------------------- 2-rank gloo
_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["SSQ_ROOT"])
from shiftedscalequantization_b200 import dist as D
rk, local, world = D.init_from_env("gloo")
assert (world, D.world_size(), D.rank()) == (2, 2, rk)
lo, hi = D.shard_range(1024)
assert (lo, hi) == ((0, 512) if rk == 0 else (512, 1024))
cali = torch.arange(10.).reshape(10, 1)
assert D.shard_calibration(cali).shape[0] == 5
# gradient bucket: SUM all-reduce keeps replicas identical (block_recon.py:100-102 semantics)
g = torch.full((7,), float(rk + 1)); D.all_reduce_sum_(g); assert torch.equal(g, torch.full((7,), 3.0))
d = torch.tensor([1.0 + rk]); D.all_average_(d); assert float(d) == 1.5
# output-channel-sharded scale search: every rank contributes its rows, all ranks see all rows
rows = torch.arange(9.).reshape(9, 1)
lo, hi = D.shard_range(9)
full = D.all_gather_rows(rows[lo:hi] * 2, 9)
assert torch.equal(full, rows * 2)
# the peer-memory exchange needs symmetric (P2P) device memory: where it cannot be set up — here: CPU tensors under gloo —
# EVERY rank must agree to fall back to the NCCL/gloo all-reduce path (a MIN all-reduce of the per-rank outcome), never hang
assert D.EXCHANGE == "p2p"
unit = D.symmetric_unit_or_none(64, torch.device("cpu"))
assert unit is None
D.EXCHANGE = "nccl"
assert D.symmetric_unit_or_none(64, torch.device("cpu")) is None      # switched off: no collective, no allocation
# activation statistics: every initialised activation quantiser — the QuantModules' AND the blocks' own — ends up identical on
# all ranks (Brecq/main_imagenet_dist.py:211; a block quantiser left out keeps the replicas apart for the whole act phase)
from shiftedscalequantization_b200 import quant as Q, zoo
torch.manual_seed(1005)
qnn = Q.QuantModel(zoo.resnet18(num_classes=10), {'n_bits': 2, 'channel_wise': True, 'scale_method': 'mse'},
                   {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True})
qs = [m.act_quantizer for m in qnn.modules() if isinstance(m, (Q.QuantModule, Q.BaseQuantBlock))]
assert any(isinstance(m, Q.BaseQuantBlock) for m in qnn.modules()) and len(qs) > 21
for i, q in enumerate(qs):
    q.delta = torch.nn.Parameter(torch.tensor(0.1 * (i + 1) * (1 + rk)))
    q.zero_point = torch.nn.Parameter(torch.tensor(float(rk * (i % 2))))
    q.inited = True
qnn.synchorize_activation_statistics()
for i, q in enumerate(qs):
    assert abs(float(q.delta) - 0.15 * (i + 1)) < 1e-6, (i, float(q.delta))
    assert float(q.zero_point) in (0.0, 1.0) and float(q.zero_point) == round(0.5 * (i % 2))    # averaged, then rounded (half to even)
sys.stdout.write(f"rank {rk} ok\n"); sys.stdout.flush()      # one write per rank: the two ranks share the pipe
'''


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SSQ_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


_FASTDIV_CHECK = r'''
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "ssq_fastdiv.h"
// the K1c kernels' multiply-shift division (host constants + the very function the device code calls) against the built-in
// division: every divisor up to 20000 plus the shapes of real layers and random / extreme ones, numerators around every
// multiple boundary and across the whole 31-bit index range
static int check(uint32_t d) {
    const ssq::FastDiv f = ssq::make_fastdiv(d);
    auto one = [&](uint64_t n) { if (n >= (1ull << 31)) return 0; return ssq::fastdiv((uint32_t)n, f) == (uint32_t)n / d ? 0 : 1; };
    int bad = 0;
    const uint64_t top = (1ull << 31) - 1;
    for (uint64_t q = 0; q < 64; ++q) for (int e = -2; e <= 2; ++e) { const int64_t n = (int64_t)(q * d) + e; if (n >= 0) bad += one((uint64_t)n); }
    for (uint64_t q = top / d; q + 4 > top / d && q <= top / d; --q) { for (int e = -2; e <= 2; ++e) { const int64_t n = (int64_t)(q * d) + e; if (n >= 0) bad += one((uint64_t)n); } if (q == 0) break; }
    for (int k = 0; k < 31; ++k) for (int e = -1; e <= 1; ++e) { const int64_t n = (int64_t)(1ull << k) + e; if (n >= 0) bad += one((uint64_t)n); }
    uint64_t x = 0x9E3779B97F4A7C15ull * (d + 1);
    for (int i = 0; i < 2000; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; bad += one(x & top); }
    bad += one(top);
    return bad;
}
int main() {
    long bad = 0, n = 0;
    for (uint32_t d = 2; d <= 20000; ++d) { bad += check(d); ++n; }
    const uint32_t shapes[] = {144, 288, 576, 1152, 2304, 4608, 9216, 16, 32, 64, 128, 256, 512, 36864, 18432, 9, 25, 49, 3, 5, 7, 27, 147,
                               (1u << 30) - 1, 1u << 30, (1u << 30) + 1, (1u << 31) - 1, 1u << 31, 0x7fffffffu, 0xfffffffu, 1000003u, 2147483629u};
    for (uint32_t d : shapes) { bad += check(d); ++n; }
    uint64_t x = 88172645463325252ull;
    for (int i = 0; i < 20000; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; uint32_t d = (uint32_t)(x >> 33); if (d < 2) d = 2; bad += check(d); ++n; }
    std::printf("%ld divisors, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
'''


def test_fastdiv_matches_integer_division(tmp_path):
    """host logic of the K1c kernels: ssq_fastdiv.h (compiled here with g++) == n / d over the 31-bit index space"""
    src = tmp_path / "fastdiv_check.cpp"
    src.write_text(_FASTDIV_CHECK)
    exe = tmp_path / "fastdiv_check"
    inc = os.path.join(ROOT, "shiftedscalequantization_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), str(src)], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout


_SLAB_CHECK = r'''
#include <cstdio>
#include <cstdint>
#include "ssq_slab_plan.h"
// K1c backward grid plan: the slabs cover every row exactly once, and the vector plan never exceeds ONE wave of resident CTAs
// (the property that took the kernel from 0.66 to 0.85 of the HBM peak) unless the column blocks alone already do
int main() {
    long bad = 0, n = 0;
    const int64_t ocs[] = {1, 3, 4, 7, 16, 33, 64, 96, 128, 256, 300, 512, 1000, 1024, 2048, 4096, 8192};
    const int64_t ks[] = {8, 12, 20, 64, 144, 576, 600, 1152, 2304, 4608, 9216, 18432, 36864, 147456, 1048576};
    for (int64_t oc : ocs) for (int64_t K : ks) for (int ctas = 0; ctas <= 5; ++ctas) {
        int nslab = 0; int64_t rps = 0;
        ssq::slab_plan(oc, K, nslab, rps, ctas);
        ++n;
        if (nslab < 1 || rps < 1 || (int64_t)nslab * rps < oc || (int64_t)(nslab - 1) * rps >= oc) { ++bad; continue; }
        if (ctas > 0) {
            const int64_t colblocks = (K / 4 + 255) / 256, slots = 148 * (int64_t)ctas;
            if (colblocks <= slots && colblocks * nslab > slots) ++bad;           // more than one wave
            if (nslab > 1 && rps < 4 && oc >= 4) ++bad;                           // slabs of at least 4 rows
        }
    }
    std::printf("%ld plans, %ld bad\n", n, bad);
    return bad ? 1 : 0;
}
'''


def test_k1c_backward_slab_plan_is_one_wave(tmp_path):
    """host logic of ssq_fq_shift_bwd: ssq_slab_plan.h compiled here with g++; plus the exported workspace size for the bench shape"""
    src = tmp_path / "slab_check.cpp"
    src.write_text(_SLAB_CHECK)
    exe = tmp_path / "slab_check"
    inc = os.path.join(ROOT, "shiftedscalequantization_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), str(src)], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    from shiftedscalequantization_b200 import _lib
    # [4096,4096,3,3], S = 3: 36 column blocks -> 16 slabs at 4 CTAs/SM (576 of 592 slots), 8 at 2; the workspace holds the larger plan
    assert _lib.load().ssq_shift_bwd_ws_bytes(4096, 4096, 9, 3, 0) == 16 * 36864 * 3 * 4 + 16


_SURFACE_CHECK = r'''
import os, sys
root = os.environ["SSQ_ROOT"]
sys.path[:0] = [os.path.join(root, "compat"), root]          # what run_driver.py arranges
# main_cifar10.py:4-7, Brecq/main_imagenet.py:8-10, ShiftedScaleQuant.py:4-10 (the import lines of the upstream drivers)
ns = {}
exec("from quant import *", ns)
exec("from common import *", ns)
exec("from myScaledMethods import *", ns)
exec("from quant.layer_recon_shiftedScale import *", ns)
exec("from quant.layer_recon_fused_shiftedScale import *", ns)
import hubconf
from data.cifar10 import build_cifar10_data
from data.imagenet import build_imagenet_data
from pretrained.PyTorch_CIFAR10.cifar10_models.resnet import resnet18
from quant.quant_layer import QuantModule, UniformAffineQuantizer
from quant.quant_block import BaseQuantBlock, QuantBasicBlock
from quant.channelQuant import ChannelQuant
from quant.channelQuantMSE import ChannelQuantMSE
from quant.channelQuantAct import ChannelQuantAct
from quant.adaptive_rounding import AdaRoundQuantizer
from quant.data_utils import save_inp_oup_data, save_grad_data
need = ["QuantModel", "QuantModule", "BaseQuantBlock", "block_reconstruction", "layer_reconstruction",      # quant/__init__.py
        "loadArgments", "seed_all", "validate_model", "get_train_samples",                                    # common.py
        "block_recon_shiftedScale", "layer_recon_shiftedScale", "block_recon_fused_shiftedScale"]            # the shifted loops
missing = [n for n in need if n not in ns]
assert not missing, missing
import inspect
sig = inspect.signature(ns["block_reconstruction"])
# upstream's positional order (quant/block_recon.py:10-14); the extra keywords come after it
up = ["model", "block", "cali_data", "batch_size", "iters", "weight", "opt_mode", "asym", "include_act_func", "b_range", "warmup",
      "act_quant", "lr", "p", "multi_gpu"]
assert list(sig.parameters)[:len(up)] == up, list(sig.parameters)
d = {k: v.default for k, v in sig.parameters.items()}
assert (d["batch_size"], d["iters"], d["weight"], d["opt_mode"], d["asym"], d["b_range"], d["warmup"], d["lr"], d["p"]) == \
       (32, 20000, 0.01, "mse", False, (20, 2), 0.0, 4e-5, 2.0)
sig = inspect.signature(ns["layer_reconstruction"])
assert list(sig.parameters)[:3] == ["model", "layer", "cali_data"] and sig.parameters["weight"].default == 0.001
print("surface ok")
'''


def test_upstream_import_surface_resolves_through_compat(tmp_path):
    """the import lines of the reference's drivers (main_cifar10.py:4-7, Brecq/main_imagenet.py:8-10, ShiftedScaleQuant.py:4-10)
    resolve against compat/ the way run_driver.py arranges sys.path, and the reconstruction entry points keep upstream's
    positional order and defaults (quant/block_recon.py:10-14, quant/layer_recon.py:10-13)"""
    script = tmp_path / "surface.py"
    script.write_text(_SURFACE_CHECK)
    out = subprocess.run([sys.executable, str(script)], env=dict(os.environ, SSQ_ROOT=ROOT), capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "surface ok" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


# ------------------------------------------------------------------------------------------- oracle loop, other families
@pytest.mark.parametrize("tag,arch,bits,pick", [("r50", "resnet50", 4, lambda q: q.model.layer1[0]),
                                                ("rx32", "regnetx_3200m", 2, lambda q: q.model.s2.b1)])
def test_oracle_loop_reproduces_reference_bottleneck_blocks(tag, arch, bits, pick):
    """the oracle's loop restatement (bottleneck kind, grouped 3x3 convolutions, downsample branch) against the REAL reference's
    block_reconstruction on a ResNet-50 bottleneck and a RegNetX-3200M block (tests/golden/families.npz; 16 iterations, asym=True).
    Everything runs on the CPU: the FP block outputs come from the seeded zoo model, the block inputs from the same model with every
    weight replaced by the oracle's nearest-rounding fake-quant ('max' init, 8-bit stem) — what the reference's quantised prefix
    computes — and the index stream is regenerated from the seed."""
    from oracle import ref_loop_torch as R
    from oracle import ssq_oracle as O
    from shiftedscalequantization_b200 import quant as Q, zoo
    from shiftedscalequantization_b200.engine import index_table
    g = golden("families")
    cali = torch.from_numpy(g[f"{tag}.cali"])
    torch.manual_seed(1005)
    cnn = zoo.build(arch, **({"num_classes": 10} if arch.startswith("resnet") else {})).eval()
    qnn = Q.QuantModel(cnn, {'n_bits': bits, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).eval()
    qnn.set_first_last_layer_to_8bit()
    block = pick(qnn)
    named = [(n, m) for n, m in block.named_modules() if isinstance(m, Q.QuantModule)]
    for n, m in named:
        assert_exact(m.org_weight.reshape(-1)[:16].numpy(), g[f"{tag}.{n}.probe_w"], f"{n}: seeded weights")

    def capture(model_input):
        seen = {}
        h = block.register_forward_hook(lambda m, i, o: seen.update(inp=i[0].detach().clone(), out=o.detach().clone()))
        qnn.set_quant_state(False, False)                    # QuantModule.forward then uses org_weight: plain torch on the CPU
        with torch.no_grad():
            qnn(model_input)
        h.remove()
        return seen["inp"], seen["out"]

    _, outs = capture(cali)                                  # FP targets
    # quantised prefix: nearest-rounding fake-quant of every layer's weights, by the oracle ('max' init in Python doubles)
    mods = [m for m in qnn.modules() if isinstance(m, Q.QuantModule)]
    saved = [m.org_weight for m in mods]
    for m in mods:
        nb = m.weight_quantizer.n_bits
        w = m.weight.detach().numpy()
        rows = w.reshape(w.shape[0], -1)
        d, z, _raw = zip(*[O.max_init(r, nb) for r in rows])
        shape = (-1,) + (1,) * (w.ndim - 1)
        d = np.array(d, np.float32).reshape(shape); z = np.array(z, np.float32).reshape(shape)
        if m in [mm for _n, mm in named]:
            n = [nn for nn, mm in named if mm is m][0]
            assert_exact(d, g[f"{tag}.{n}.delta"], f"{n}: delta of the oracle's max init"); assert_exact(z, g[f"{tag}.{n}.zp"], f"{n}: zero point")
        y, _codes = O.uaq_forward(w, d, z, 0, 2 ** nb - 1)
        m.org_weight = torch.from_numpy(y)
    inps, _ = capture(cali)
    for m, w in zip(mods, saved):
        m.org_weight = w
    layers = {}
    for n, m in named:
        act = {"ReLU": "relu", "ReLU6": "relu6"}.get(type(m.activation_function).__name__)
        layers[n] = dict(weight=m.org_weight.detach(), bias=None if m.org_bias is None else m.org_bias.detach(), conv=dict(m.fwd_kwargs),
                         act=act, delta=torch.from_numpy(g[f"{tag}.{n}.delta"]), zero_point=torch.from_numpy(g[f"{tag}.{n}.zp"]),
                         n_levels=2 ** bits)
    unit = {"kind": "bottleneck", "layers": layers, "tail_act": "relu"}
    torch.manual_seed(277)
    tab = index_table(32, 16, 16)
    alphas, losses = R.recon_weight_loop(unit, inps, outs, tab, 16, weight=0.01, b_range=(20, 2), warmup=0.2)
    assert np.isfinite(losses).all()
    for n, _m in named:
        a, ref = alphas[n].detach().numpy(), g[f"{tag}.{n}.alpha"]
        moved = np.abs(ref - R.init_alpha(layers[n]["weight"], layers[n]["delta"]).numpy()).max()
        assert moved > 5e-3
        assert_close(a, ref, rtol=1e-5, what=f"{tag} alpha {n}: oracle loop vs the reference's block_reconstruction")


def test_oracle_loop_reproduces_reference_bias_cal_trajectory():
    """README --bias_cal on the CPU: the oracle loop with every layer's output affine (gamma^z, varphi^z = alpha_out, beta_out) in the
    optimiser against the real reference's forward / autograd / LossFunction / Adam (tests/golden/bias_cal.npz: ResNet-18 layer2.0 with
    its downsample branch, 16 iterations, the reference's own cached features and index stream)"""
    from oracle import ref_loop_torch as R
    from oracle import ssq_oracle as O
    from shiftedscalequantization_b200 import quant as Q, zoo
    g = golden("bias_cal")
    torch.manual_seed(1005)
    qnn = Q.QuantModel(zoo.resnet18(num_classes=10).eval(), {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}, dict(AQ)).eval()
    qnn.set_first_last_layer_to_8bit()
    assert_exact(torch.randn(32, 3, 16, 16).numpy(), g["cali"], "seeded calibration tensor (same RNG position as the golden run)")
    block = qnn.model.layer2[0]
    layers = {}
    for n, m in [(n, m) for n, m in block.named_modules() if isinstance(m, Q.QuantModule)]:
        w = m.org_weight.detach().numpy()
        d, z, _raw = zip(*[O.max_init(r, 2) for r in w.reshape(w.shape[0], -1)])
        mk = lambda v: torch.from_numpy(np.array(v, np.float32).reshape(-1, 1, 1, 1))
        act = {"ReLU": "relu"}.get(type(m.activation_function).__name__)
        layers[n] = dict(weight=m.org_weight.detach(), bias=None if m.org_bias is None else m.org_bias.detach(), conv=dict(m.fwd_kwargs), act=act,
                         delta=mk(d), zero_point=mk(z), n_levels=4, alpha_out=torch.ones(1, w.shape[0], 1, 1), beta_out=torch.zeros(1, w.shape[0], 1, 1))
    assert sorted(layers) == ["conv1", "conv2", "downsample"]
    unit = {"kind": "basic", "layers": layers, "tail_act": "relu"}
    iters = int(g["iters"])
    alphas, losses = R.recon_weight_loop(unit, torch.from_numpy(g["inps"]), torch.from_numpy(g["outs"]), torch.from_numpy(g["idx"]), iters,
                                         weight=0.01, b_range=(20, 2), warmup=0.2, train_affine=True)
    assert_close(np.array(losses), g["losses"], rtol=1e-5, what="loss trace with the affine in the optimiser")
    for n in layers:
        assert_close(alphas[n].detach().numpy(), g[f"{n}.alpha"], rtol=1e-5, what=f"alpha {n}")
        assert_close(layers[n]["alpha_out"].detach().numpy(), g[f"{n}.alpha_out"], rtol=1e-5, what=f"gamma {n}")
        assert_close(layers[n]["beta_out"].detach().numpy(), g[f"{n}.beta_out"], rtol=1e-5, what=f"varphi {n}")
        assert float((layers[n]["alpha_out"].detach() - 1).abs().max()) > 5e-3                     # the affine really moved


def test_oracle_act_loop_reproduces_reference_activation_phase():
    """the oracle's activation-phase loop (LSQ step sizes, Adam lr 4e-4 + cosine annealing, lp p = 2.4; hard-rounded weights) against
    the REAL reference's block_reconstruction(act_quant=True) on the CPU (tests/golden/act_phase.npz, make_golden_act_phase.py):
    same cached features, same index stream, same initial step sizes -> same learned step sizes"""
    from oracle import ref_loop_torch as R
    g = golden("act_phase")
    layers, alphas = {}, {}
    for n, act in (("conv1", "relu"), ("conv2", None)):
        layers[n] = dict(weight=torch.from_numpy(g[f"{n}.weight"]), bias=torch.from_numpy(g[f"{n}.bias"]),
                         conv=dict(stride=1, padding=1, dilation=1, groups=1), act=act, delta=torch.from_numpy(g[f"{n}.delta"]),
                         zero_point=torch.from_numpy(g[f"{n}.zp"]), n_levels=4)
        alphas[n] = torch.from_numpy(g[f"{n}.alpha"])
    unit = {"kind": "basic", "layers": layers, "tail_act": "relu"}
    act_state = {k: (torch.from_numpy(g[f"{k}.act_delta0"]).clone().requires_grad_(True), torch.from_numpy(g[f"{k}.act_zp"]),
                     int(g[f"{k}.act_levels"])) for k in ("conv1", "__block__")}
    losses = R.recon_act_loop(unit, torch.from_numpy(g["inps"]), torch.from_numpy(g["outs"]), torch.from_numpy(g["idx"]), int(g["iters"]),
                              act_state, alphas, lr=4e-4, p=2.4)
    assert np.isfinite(losses).all()
    for k in act_state:
        d0, d1, got = float(g[f"{k}.act_delta0"]), float(g[f"{k}.act_delta1"]), float(act_state[k][0])
        assert abs(d1 - d0) > 1e-3                                   # the reference's step size moved ...
        assert abs(got - d1) <= 1e-5 * abs(d1), (k, d0, d1, got)     # ... and the oracle's ends up in the same place
