import sys, torch
sys.path.insert(0, '.')
from shiftedscalequantization_b200 import ops
dev = torch.device('cuda')
def timeit(fn, nbytes, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return nbytes / ms / 1e6
torch.manual_seed(0)
for name, shape in (("w 151M", (4096, 4096, 3, 3)), ("a 205M", (256, 256, 56, 56))):
    for dist in ("randn", "relu"):
        x = torch.randn(shape, device=dev) * (0.02 if name[0] == 'w' else 1.0)
        if dist == "relu": x = torch.relu(x)
        n = x.numel()
        y = torch.empty_like(x)
        dpc = (x.abs().amax(dim=(1, 2, 3), keepdim=True) / 3 * 1.2 + 1e-6).contiguous(); zpc = torch.full_like(dpc, 2.0)
        ds, zs = ops.scalar_dev(0.25, dev).reshape(()), ops.scalar_dev(0.0, dev).reshape(())
        print(name, dist, "per-channel fwd %.0f GB/s" % timeit(lambda: ops.fq_affine_fwd(x, dpc, zpc, 0.0, 3.0), 8 * n),
              "per-tensor fwd %.0f GB/s" % timeit(lambda: ops.fq_affine_fwd(x, ds, zs, 0.0, 15.0), 8 * n),
              "torch copy %.0f GB/s" % timeit(lambda: y.copy_(x), 8 * n),
              "gather %.0f" % timeit(lambda: ops.gather_rows(x, torch.arange(shape[0], device=dev), out=y), 8 * n), flush=True)
        del x, y
        torch.cuda.empty_cache()
