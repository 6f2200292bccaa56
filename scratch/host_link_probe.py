"""N-rank probe of the host link the host-resident (e2e) mode depends on: NUMA layout of the box, and host->device bandwidth of
every rank ALONE versus ALL ranks at once, for the copy engine (cudaMemcpyAsync from pinned memory) and for the SM pull kernel
(ssq_pull_rows_host). torchrun --nproc-per-node N scratch/host_link_probe.py   (rank 0 prints one JSON line)"""
import glob
import json
import os
import subprocess
import sys

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shiftedscalequantization_b200 import dist as D, ops            # noqa: E402

rank, local, world = D.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
info = {}
if rank == 0:
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node*")):
        try:
            nodes[os.path.basename(p)] = {"cpulist": open(p + "/cpulist").read().strip(),
                                          "MemTotal_kB": int(open(p + "/meminfo").read().split("MemTotal:")[1].split()[0])}
        except Exception as e:
            nodes[os.path.basename(p)] = str(e)
    info["numa_nodes"] = nodes
    info["cpus_allowed"] = len(os.sched_getaffinity(0))
    try:
        info["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as e:
        info["topo"] = str(e)
    gp = {}
    for i in range(torch.cuda.device_count()):
        try:
            bus = subprocess.run(["nvidia-smi", "-i", str(i), "--query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=30).stdout.strip()
            b = bus.split(",")[0].strip().lower()
            b = b[4:] if len(b) > 12 else b
            node = "?"
            try:
                node = open(f"/sys/bus/pci/devices/{b}/numa_node").read().strip()
            except Exception:
                pass
            gp[i] = {"pci": bus, "numa_node": node}
        except Exception as e:
            gp[i] = str(e)
    info["gpus"] = gp

n, per, batch, steps = 512, 64 * 56 * 56, 32, 16
x = torch.randn(n, per).pin_memory()
tab = torch.stack([torch.randperm(n)[:batch] for _ in range(steps)]).to(dev)
step = torch.zeros(1, dtype=torch.int64, device=dev)
dst = torch.empty(batch, per, device=dev)
big = torch.empty(64 * per, device=dev)
nbytes_pull = 4 * batch * per


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dma():
    big.copy_(x[:64].reshape(-1), non_blocking=True)


def pull():
    ops.pull_rows_host(x, tab, step, 0, steps, dst.view(batch, 64, 56, 56), max_ctas=24)


def gbs(fn, nbytes):
    return nbytes / timeit(fn) / 1e6


res = {"alone": {}, "together": {}}
for r in range(world):                       # one rank at a time
    if world > 1:
        td.barrier()
    if r == rank:
        res["alone"] = {"dma_gbs": round(gbs(dma, 4 * 64 * per), 1), "pull_gbs": round(gbs(pull, nbytes_pull), 1)}
if world > 1:
    td.barrier()
d = gbs(dma, 4 * 64 * per)
if world > 1:
    td.barrier()
p = gbs(pull, nbytes_pull)
res["together"] = {"dma_gbs": round(d, 1), "pull_gbs": round(p, 1)}
allres = [None] * world
if world > 1:
    td.all_gather_object(allres, res)
else:
    allres = [res]
if rank == 0:
    info["ranks"] = allres
    info["aggregate_together_gbs"] = {"dma": round(sum(a["together"]["dma_gbs"] for a in allres), 1),
                                      "pull": round(sum(a["together"]["pull_gbs"] for a in allres), 1)}
    print(json.dumps(info), flush=True)
if world > 1:
    td.barrier()
    td.destroy_process_group()
