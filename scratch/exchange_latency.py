"""exchange kernel in isolation: both ranks launch it back to back (no other work), so the time per launch is the kernel's own
cost: two rendezvous over NVLink + the data movement. Also checks the result against all-reduce + Adam every launch."""
import json, os, sys, time
import torch, torch.distributed as td
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shiftedscalequantization_b200 import dist as D, ops
rank, local, world = D.init_from_env()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
out = {}
for n in (4 * 64, 4 * 18432, 4 * 131072, 4 * 1179648):
    sym = D.SymmetricUnit(n, dev)
    torch.manual_seed(1); p0 = torch.randn(n, device=dev); sym.flat.copy_(p0)
    m = torch.zeros(sym.shard, device=dev); v = torch.zeros(sym.shard, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev); lr = ops.scalar_dev(1e-3, dev)
    pr, mr, vr, sr = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
    bad = 0
    for it in range(6):
        torch.manual_seed(10 * it + rank); g = torch.randn(n, device=dev); sym.gflat.copy_(g)
        torch.cuda.synchronize(); td.barrier()
        ops.grad_exchange_adam(sym, m, v, lr, step)
        gs = g.clone(); td.all_reduce(gs)
        ops.adam_step_end_iteration(pr, gs, mr, vr, lr, sr)
        torch.cuda.synchronize(); td.barrier()
        if world == 2 and not torch.equal(sym.flat, pr):
            bad += 1
            d = (sym.flat != pr).nonzero().flatten()
            if rank == 0: print(f"n={n} it={it}: {d.numel()} differ, first {d[:4].tolist()} last {d[-4:].tolist()} shard={sym.shard}", file=sys.stderr)
    torch.cuda.synchronize(); td.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for _ in range(reps):
        ops.grad_exchange_adam(sym, m, v, lr, step)
    e1.record(); torch.cuda.synchronize()
    out[n] = {"us_per_launch": round(1e3 * e0.elapsed_time(e1) / reps, 2), "mismatching_iterations": bad, "timeouts": int(sym.timeouts)}
    td.barrier()
if rank == 0:
    print(json.dumps(out))
td.barrier(); td.destroy_process_group()
