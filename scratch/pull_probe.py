"""dense vs zero-packed pull of one mini-batch (32 rows of [64,56,56] post-ReLU features) out of pinned host memory"""
import sys, torch
sys.path.insert(0, '.')
from shiftedscalequantization_b200 import ops
dev = torch.device('cuda', 0)
torch.manual_seed(0)
n, shape, batch, steps = 256, (64, 56, 56), 32, 16
x = torch.relu(torch.randn((n,) + shape))
per = x[0].numel()
xp = x.pin_memory()
packed = ops.pack_rows_sparse(xp, dev)
tab = torch.stack([torch.randperm(n)[:batch] for _ in range(steps)]).to(dev)
step = torch.zeros(1, dtype=torch.int64, device=dev)
dst = torch.empty((batch,) + shape, device=dev)
def timeit(fn, reps=12):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        step.fill_(i % steps); fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
dense_bytes = 4 * batch * per
print(f"density {packed.density:.3f}; dense {dense_bytes / 1e6:.1f} MB, packed {batch * packed.host_bytes_per_row() / 1e6:.1f} MB per mini-batch", flush=True)
for ctas in (8, 16, 24, 32, 48, 64, 96):
    md = timeit(lambda: ops.pull_rows_host(xp, tab, step, 0, steps, dst, max_ctas=ctas))
    mp = timeit(lambda: ops.pull_rows_host_packed(packed, tab, step, 0, steps, dst, max_ctas=ctas))
    pb = batch * packed.host_bytes_per_row()
    print(f"ctas {ctas:3d}: dense {md:.3f} ms ({dense_bytes / md / 1e6:.1f} GB/s PCIe) | packed {mp:.3f} ms ({pb / mp / 1e6:.1f} GB/s PCIe, {dense_bytes / mp / 1e6:.1f} GB/s dense-equivalent)", flush=True)
