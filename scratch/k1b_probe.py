"""K1b forward (+ regulariser) GB/s for each libssq variant given on the command line; results checked against the first."""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
from shiftedscalequantization_b200 import _lib as L
dev = torch.device('cuda', 0)
def load(path):
    lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND | os.RTLD_NOW)
    for name in ("ssq_fq_adaround_fwd", "ssq_ws_bytes", "ssq_adaround_init_alpha"):
        res, args = L.PROTOTYPES[name]; fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib
torch.manual_seed(0)
oc, K = 4096, 4096 * 9; n = oc * K
w = torch.randn(oc, K, device=dev) * 0.02
d = (w.abs().amax(1) / 3 * 1.2).contiguous(); z = torch.full_like(d, 2.0)
alpha = torch.empty_like(w); y = torch.empty_like(w); reg = torch.zeros(1, device=dev); b = torch.full((1,), 11.0, device=dev)
ref = None
for path in sys.argv[1:]:
    lib = load(path)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.ssq_adaround_init_alpha(w.data_ptr(), d.data_ptr(), alpha.data_ptr(), n, K, oc, st) == 0
    ws = torch.zeros(lib.ssq_ws_bytes(1) // 4 + 4, device=dev)
    run = lambda: lib.ssq_fq_adaround_fwd(w.data_ptr(), alpha.data_ptr(), d.data_ptr(), z.data_ptr(), y.data_ptr(), None, n, K, oc, 0.0, 3.0, 1,
                                          b.data_ptr(), 0.01, reg.data_ptr(), ws.data_ptr(), ws.numel() * 4, st)
    for _ in range(3): assert run() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    cur = (y.clone(), float(reg))
    same = True if ref is None else (torch.equal(cur[0], ref[0]) and cur[1] == ref[1])
    ref = ref or cur
    print(f"{os.path.basename(path)}: {ms:.4f} ms {12 * n / ms / 1e6:.0f} GB/s ({12 * n / ms / 1e6 / 6554.9:.3f}) same={same}", flush=True)
