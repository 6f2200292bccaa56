"""N-GPU probe of the exchange step: step time with no exchange / NCCL all-reduce + Adam / the peer-memory kernel, and the
per-kernel device time of one eager iteration per unit (CUDA events). torchrun --nproc-per-node N scratch/exchange_probe.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                                    # noqa: E402
from shiftedscalequantization_b200 import dist as D, ops            # noqa: E402

rank, local, world = D.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
n_total = 1024
lo, hi = D.shard_range(n_total, rank, world)
Q, qnn, cali = B.build_model(dev, n_total)
cali = cali[lo:hi]
qnn.set_quant_state(True, False)
with torch.no_grad():
    qnn(cali[:64].to(dev))
out, feats = {}, None
for mode in ("none", "nccl", "p2p"):
    D.EXCHANGE = mode
    engines, feats = B.make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=world > 1, feats=feats)
    ms = B.timed_steps(engines, 30, 5, dev, world)
    rec = {"ms_per_step": ms}
    if mode != "none":
        graphs = [e.graph for e in engines]
        for e in engines:
            e.graph = None
        for e in engines:
            e.step()
        torch.cuda.synchronize(dev)
        ops.profile_begin()
        for e in engines:
            e.step()
        prof = ops.profile_end()
        rec["eager_kernels_ms"] = {k: [n, round(v, 4)] for k, (n, v) in prof.items()}
        for e, g in zip(engines, graphs):
            e.graph = g
    B.release(engines)
    del engines
    torch.cuda.empty_cache()
    out[mode] = rec
if rank == 0:
    B.emit(json.dumps(out))
import torch.distributed as td
td.barrier()
td.destroy_process_group()
