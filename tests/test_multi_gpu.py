"""Multi-GPU reconstruction on real devices: 2 ranks, NCCL. Needs >= 2 GPUs (skipped on the 1-GPU box;
run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`). The CPU-side plumbing (sharding,
all-reduce, all-gather) is covered by the gloo test in test_host_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as td
sys.path.insert(0, os.environ["SSQ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SSQ_ROOT"], "tests"))
from shiftedscalequantization_b200 import dist as D
from shiftedscalequantization_b200.engine import ReconEngine, index_table
from shiftedscalequantization_b200.quant.adaptive_rounding import AdaRoundQuantizer
from shiftedscalequantization_b200.quant.data_utils import save_inp_oup_data
from test_recon_gpu import build_qnn
rk, local, world = D.init_from_env()
torch.cuda.set_device(local)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.deterministic = True          # as the reference's drivers (common.py:84-85): runs are comparable bit for bit
iters, bs = 20, 16

def run(mode):
    Q, qnn, cali = build_qnn()                       # same seed on every rank => identical replicas
    block = qnn.model.layer2[0]
    qnn.set_quant_state(False, False); block.set_quant_state(True, False)
    mods = [m for m in block.modules() if isinstance(m, Q.QuantModule)]
    for m in mods:
        m.weight_quantizer = AdaRoundQuantizer(uaq=m.weight_quantizer, round_mode='learned_hard_sigmoid', weight_tensor=m.org_weight.data)
        m.weight_quantizer.soft_targets = True
    data = D.shard_calibration(cali) if mode == 'weak' else cali
    inps, outs = save_inp_oup_data(qnn, block, data, True, False, bs)
    torch.manual_seed(7)
    tab = index_table(inps.shape[0], bs, iters)
    eng = ReconEngine(block, mods, inps, outs, None, act_quant=False, iters=iters, weight=0.01, b_range=(20, 2), warmup=0.2,
                      p=2.0, batch_size=bs, use_graph=True, idx_table=tab, verbose=False, multi_gpu=(mode != 'single'),
                      scaling='strong' if mode == 'strong' else 'weak')
    assert (eng.sym is not None) == (mode != 'single' and D.EXCHANGE == 'p2p'), "symmetric memory must be available on a multi-GPU box"
    eng.run(); eng.close()
    return torch.cat([m.weight_quantizer.alpha.detach().reshape(-1) for m in mods])

# the exchange step itself: reduce-scatter + Adam + all-gather in one peer-memory kernel against torch / NCCL
from shiftedscalequantization_b200 import ops
n = 4 * 1037                                                     # not a multiple of the world size or of the tile
sym = D.SymmetricUnit(n, torch.device("cuda", local))
torch.manual_seed(11)
p0 = torch.randn(n, device="cuda")                               # same on every rank
sym.flat.copy_(p0)
shard = sym.shard
m = torch.zeros(shard, device="cuda"); v = torch.zeros(shard, device="cuda"); red = torch.zeros(shard, device="cuda")
step = torch.zeros(1, dtype=torch.int64, device="cuda"); lr = ops.scalar_dev(1e-3, "cuda")
pr, mr, vr = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
sr = torch.zeros(1, dtype=torch.int64, device="cuda")
for it in range(4):
    torch.manual_seed(100 * it + rk)
    g = torch.randn(n, device="cuda")                            # a different gradient on every rank
    sym.gflat.copy_(g)
    ops.grad_exchange_adam(sym, m, v, lr, step, reduced_out=red)
    gs = g.clone(); td.all_reduce(gs)                            # world == 2: a + b in either order is the same float
    ops.adam_step_end_iteration(pr, gs, mr, vr, lr, sr)
    torch.cuda.synchronize(); td.barrier()
    lo = rk * shard; hi = min(lo + shard, n)
    assert torch.equal(red[:hi - lo], gs[lo:hi]), f"iteration {it}: reduced shard"
    assert torch.equal(sym.flat, pr), f"iteration {it}: parameters after the fused exchange vs all-reduce + Adam"
    assert int(step) == it + 1
sym.check()

single = run('single')
D.EXCHANGE = 'nccl'
weak_nccl = run('weak')
D.EXCHANGE = 'p2p'
for mode in ('strong', 'weak'):
    a = run(mode)
    both = [torch.empty_like(a) for _ in range(world)]
    td.all_gather(both, a)
    assert torch.equal(both[0], both[1]), f"{mode}: replicas diverged"          # same reduced gradient, same Adam => bit-identical
    if mode == 'strong':
        err = (a - single).abs()
        frac = float((err <= 2e-3).float().mean())
        assert frac >= 0.999 and float(err.max()) <= 2 * 1e-3 * iters, (frac, float(err.max()))
        assert float((torch.sign(a) == torch.sign(single)).float().mean()) > 0.9995
    else:
        assert not torch.equal(a, single)                                       # other mini-batches: a different trajectory
        assert torch.equal(a, weak_nccl), "peer-memory exchange vs NCCL all-reduce + Adam (two ranks: identical sums)"
td.barrier()
sys.stdout.write(f"rank {rk} ok\n"); sys.stdout.flush()
td.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_strong_matches_single_and_replicas_identical(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SSQ_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29621", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


_WORKER_SHIFT = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as td
sys.path.insert(0, os.environ["SSQ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SSQ_ROOT"], "tests"))
from shiftedscalequantization_b200 import dist as D
from shiftedscalequantization_b200 import quant as Q, zoo
from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
rk, local, world = D.init_from_env()
torch.cuda.set_device(local)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.deterministic = True

def run(multi):
    """BASELINE configs[2]: ResNet-50 W4A4 shifted-scale LAYER reconstruction, calibration batch sharded across the ranks"""
    torch.manual_seed(1005)
    cnn = zoo.resnet50(num_classes=10).cuda().eval()
    qnn = Q.QuantModel(cnn, {'n_bits': 4, 'channel_wise': True, 'scale_method': 'max'},
                       {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}).cuda().eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(96, 3, 32, 32)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(cali[:32].cuda())
    layer = qnn.model.layer2[0].conv2
    layer.weight_quantizer = ChannelQuant(1.0, uaq=layer.weight_quantizer, weight_tensor=layer.org_weight.data,
                                          shiftTarget=[0.96875, 1.03125, 1.0], name=layer.pathName)
    data = D.shard_calibration(cali) if multi else cali
    for mode, wq_on in (('if', True), ('of', False)):
        qnn.set_quant_state(wq_on, False)
        layer.cache_features = mode
        with torch.no_grad():
            for i in range(0, data.shape[0], 16):
                qnn(data[i:i + 16].cuda())
        layer.cache_features = 'none'
    qnn.set_quant_state(False, False); layer.set_quant_state(True, False)
    LS.MULTI_GPU = multi
    torch.manual_seed(31)
    soft, hard = LS.layer_recon_shiftedScale(layer, iters=40, lmda=0.01, model=qnn)
    assert np.isfinite([soft, hard]).all()
    LS.MULTI_GPU = False
    return layer.weight_quantizer.alpha.detach().reshape(-1).clone(), LS.LAST_LOOP_STATS.get("captured")

single, _ = run(False)
a, captured = run(True)
assert captured, "the sharded loop must run as the captured graph"
D.EXCHANGE = "nccl"                                                   # same run through NCCL all-reduce + ssq_adam_step
a_nccl, _ = run(True)
D.EXCHANGE = "p2p"
assert torch.equal(a, a_nccl), "peer-memory exchange vs NCCL all-reduce + Adam in the shifted loop (two ranks: identical sums)"
both = [torch.empty_like(a) for _ in range(world)]
td.all_gather(both, a)
assert torch.equal(both[0], both[1]), "replicas diverged"            # summed gradients => identical Adam steps
assert not torch.equal(a, single)                                     # 2 x 32 images per step instead of 32
assert float((a - single).abs().max()) < 0.2                          # ... but the same problem: a nearby trajectory
td.barrier()
sys.stdout.write(f"rank {rk} ok\n"); sys.stdout.flush()
td.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_shifted_layer_reconstruction(tmp_path):
    """quant/layer_recon_shiftedScale.py:262-338 under torchrun with the MULTI_GPU module switch (configs[2])"""
    script = tmp_path / "worker_shift.py"
    script.write_text(_WORKER_SHIFT)
    env = dict(os.environ, SSQ_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29622", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
