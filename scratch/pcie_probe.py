import torch, time, sys
sys.path.insert(0, '.')
from shiftedscalequantization_b200 import ops
dev = torch.device('cuda', 0)
def bw(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return nbytes / dt / 1e9, dt * 1e3
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dbig = torch.empty_like(big, device=dev)
print("one 256MB H2D: %.1f GB/s (%.2f ms)" % bw(lambda: dbig.copy_(big, non_blocking=True), big.numel()))
print("one 256MB D2H: %.1f GB/s (%.2f ms)" % bw(lambda: big.copy_(dbig, non_blocking=True), big.numel()))
for per in (25088, 50176, 100352, 200704):   # floats per row: layer4 .. layer1
    N = 1024
    src = torch.randn(N, per).pin_memory()
    dst = torch.empty(32, per, device=dev)
    rows = torch.randperm(N)[:32]
    cs = torch.cuda.Stream()
    def f():
        for _ in range(8):
            ops.stage_rows_h2d(src, rows, dst, cs)
    g, ms = bw(f, 8 * 32 * per * 4)
    t0 = time.perf_counter(); f(); t1 = time.perf_counter()
    print(f"rows of {per*4/1024:.0f} KB: {g:.1f} GB/s; host issue {1e6*(t1-t0)/256:.2f} us/row")
    torch.cuda.synchronize()

print("--- SM pull from mapped pinned memory")
for per in (25088, 200704):
    N = 1024
    src = torch.randn(N, per).pin_memory()
    dst = torch.empty(32, per, device=dev)
    tab = torch.stack([torch.randperm(N)[:32] for _ in range(4)]).to(dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for ctas in (8, 16, 32, 64, 148, 296):
        def f():
            for _ in range(8):
                ops.pull_rows_host(src, tab, step, 1, 4, dst, max_ctas=ctas)
        g, ms = bw(f, 8 * 32 * per * 4)
        ok = torch.equal(dst.cpu(), src[tab[1].cpu()])
        print(f"rows of {per*4/1024:.0f} KB, {ctas} CTAs: {g:.1f} GB/s  correct={ok}")
