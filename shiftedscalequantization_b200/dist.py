"""Multi-GPU plumbing for the calibration path: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch; gloo in the CPU tests). The path has exactly one exchange step per reconstruction
iteration — a SUM all-reduce of the unit's flat gradient buffer (reference intent: link.allreduce at
quant/block_recon.py:100-102) — plus an all-average of the activation step sizes after their init
(quant/quant_model.py:78-83) and an all-gather when the scale search is sharded by output channel.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as td


def is_initialized() -> bool:
    return td.is_available() and td.is_initialized()


def world_size() -> int:
    return td.get_world_size() if is_initialized() else 1


def rank() -> int:
    return td.get_rank() if is_initialized() else 0


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world). Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* as torchrun sets them."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            td.init_process_group(backend, rank=rk, world_size=world, device_id=torch.device("cuda", local))
        else:
            td.init_process_group(backend, rank=rk, world_size=world)
    return rk, local, world


def bind_to_gpu_cpus(local_rank: int) -> int:
    """pin this process to the CPU cores NVML reports as local to its GPU, so that the pinned host cache of the
    host-resident mode (first touch) and the copy/launch threads sit on the GPU's NUMA node; 8 ranks pulling
    mini-batches over 8 PCIe links otherwise meet in one socket's memory controllers. Returns the number of cores
    kept (0 = left unchanged: NVML missing, or the container's cpuset has no overlap)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        h = pynvml.nvmlDeviceGetHandleByIndex(int(visible[local_rank]) if local_rank < len(visible) else local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        keep = cores & allowed
        if keep and keep != allowed:
            os.sched_setaffinity(0, keep)
            return len(keep)
    except Exception:
        pass
    return 0


def all_reduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        td.all_reduce(flat, op=td.ReduceOp.SUM)
    return flat


def all_average_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
        t.div_(world_size())
    return t


def shard_range(n: int, r: int = None, w: int = None) -> Tuple[int, int]:
    """contiguous shard [lo, hi) of n items for rank r of w (remainder spread over the first ranks)"""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def shard_calibration(cali_data: torch.Tensor) -> torch.Tensor:
    """this rank's images (reference intent: num_samples/ngpus per rank, Brecq/main_imagenet_dist.py:165)"""
    lo, hi = shard_range(cali_data.shape[0])
    return cali_data[lo:hi]


def all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """concatenate per-rank row shards produced with shard_range (output-channel-sharded scale search)"""
    if world_size() == 1:
        return local
    w = world_size()
    sizes = [shard_range(n_total, r, w) for r in range(w)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(w)]
    td.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)])


# ------------------------------------------------------------------------------------------- symmetric memory
EXCHANGE = os.environ.get("SSQ_EXCHANGE", "p2p")     # "p2p": the fused peer-memory kernel (csrc/exchange.cu); "nccl": all-reduce + Adam


class SymmetricUnit:
    """The flat parameter / gradient buffers of one reconstruction unit in symmetric memory, plus the flag pad the
    exchange kernel synchronises through: one allocation per rank at the same offset everywhere, peer pointers exchanged
    once (torch.distributed._symmetric_memory — plumbing only; the data path is ssq_grad_exchange_adam)."""

    def __init__(self, n_floats: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        self.n = int(n_floats)
        assert self.n % 4 == 0
        pad_floats = int(lib.ssq_exchange_pad_bytes()) // 4
        self.world, self.rank = world_size(), rank()
        total = 2 * self.n + pad_floats
        self.buf = symm.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.buf, td.group.WORLD)
        self.flat = self.buf[:self.n]
        self.gflat = self.buf[self.n:2 * self.n]
        bases = [self.handle.get_buffer(r, (total,), torch.float32, 0).data_ptr() for r in range(self.world)]
        import ctypes as C
        arr = C.c_void_p * self.world
        self.flat_ptrs = arr(*bases)
        self.grad_ptrs = arr(*[b + 4 * self.n for b in bases])
        self.pad_ptrs = arr(*[b + 8 * self.n for b in bases])
        self.shard = int(lib.ssq_exchange_shard_elems(self.n, self.world))
        self.timeouts = torch.zeros(1, dtype=torch.int32, device=device)
        self.epochs = torch.zeros(pad_floats // 8 + 8, dtype=torch.int32, device=device)     # per-CTA launch counters (local)
        td.barrier()                                   # every rank has zeroed its pad before anybody polls it

    def check(self):
        """raise if a rendezvous of the exchange kernel ever timed out (one device read; call after a run)"""
        if int(self.timeouts) != 0:
            raise RuntimeError("gradient exchange: a peer rank did not arrive at a rendezvous (results are invalid)")


def symmetric_unit_or_none(n_floats: int, device: torch.device):
    """a SymmetricUnit when the peer-memory exchange is selected AND available on every rank, else None (NCCL path)"""
    if world_size() <= 1 or EXCHANGE != "p2p":
        return None
    unit, ok = None, 1
    try:
        unit = SymmetricUnit(n_floats, device)
    except Exception as e:             # symmetric memory unsupported here (no P2P, fabric handles unavailable, ...)
        ok = 0
        if rank() == 0:
            import sys
            print(f"[ssq] symmetric memory unavailable ({type(e).__name__}: {e}); the gradient exchange uses NCCL", file=sys.stderr)
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    return unit if int(flag) == 1 else None
