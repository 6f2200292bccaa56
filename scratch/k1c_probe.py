"""K1c probe: for each libssq variant given on the command line, (1) vector path vs the scalar path of the same library on
identical inputs (misaligned views force the scalar kernels) incl. NaN/Inf/huge weights, (2) GB/s on the bench shape."""
import ctypes as C
import sys
import torch

sys.path.insert(0, '.')
from shiftedscalequantization_b200 import _lib as L

dev = torch.device('cuda', 0)
PEAK = 6554.9


import os


def load(path):
    lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND | os.RTLD_NOW)   # full builds; keep each variant's symbols to itself
    for name in ("ssq_fq_shift_fwd", "ssq_fq_shift_bwd", "ssq_shift_bwd_ws_bytes"):
        res, args = L.PROTOTYPES[name]
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib


def ptr(t):
    return None if t is None else t.data_ptr()


def fwd(lib, w, sd, d, z, p, beta, y, oc, ic, kk, S, mode, ht, hr, qmin, qmax):
    st = torch.cuda.current_stream().cuda_stream
    e = lib.ssq_fq_shift_fwd(ptr(w), ptr(sd), ptr(d), ptr(z), ptr(p), ptr(beta), ptr(y), oc, ic, kk, S, 0, mode, ht, hr, qmin, qmax, st)
    assert e == 0, e


def bwd(lib, g, w, sd, d, z, p, beta, gp, gbeta, oc, ic, kk, S, mode, hr, qmin, qmax, ws):
    st = torch.cuda.current_stream().cuda_stream
    e = lib.ssq_fq_shift_bwd(ptr(g), ptr(w), ptr(sd), ptr(d), ptr(z), ptr(p), ptr(beta), ptr(gp), ptr(gbeta), oc, ic, kk, S, 0, mode, hr,
                             qmin, qmax, ptr(ws), ws.numel() * 4, st)
    assert e == 0, e


def misaligned(t):
    buf = torch.empty(t.numel() + 1, device=dev, dtype=t.dtype)
    v = buf[1:]
    v.copy_(t.reshape(-1))
    return v


def check(lib, oc, ic, kk, S, mode, ht, hr, special, zint=True):
    torch.manual_seed(oc * 131 + ic * 7 + kk + S)
    K = ic * kk
    w = torch.randn(oc, K, device=dev) * 0.05
    if special:
        flat = w.view(-1)
        vals = [(3, float('nan')), (17, float('inf')), (40, -float('inf')), (60, 1e30), (70, -3e38), (75, 0.0), (76, -0.0), (77, 1e-40), (78, -1e-30)]
        for ix, v in vals:
            flat[ix] = v
    d = (w.abs().amax(1).clamp_min(1e-3) / 3).contiguous()
    if special:
        d[1] = 1e-25; d[2] = 1e25
    z = torch.randint(0, 3, (oc,), device=dev).float()
    if not zint:
        z = z + 0.25
    shifts = [0.96875, 1.03125, 1.0, 0.9375][:S]
    sd = torch.stack([d * s for s in shifts]).contiguous()
    p = torch.softmax(torch.randn(ic, S, device=dev), -1).mul(1.2).sub(0.1).clamp(0, 1).contiguous()
    beta = torch.randn(oc, K, device=dev) * 2 if mode == 1 else None
    g = torch.randn(oc, K, device=dev)
    y1 = torch.empty(oc, K, device=dev); y2 = misaligned(y1)
    fwd(lib, w, sd, d, z, p, beta, y1, oc, ic, kk, S, mode, ht, hr, 0.0, 3.0)
    wm = misaligned(w); bm = misaligned(beta) if beta is not None else None
    fwd(lib, wm, sd, d, z, p, bm, y2, oc, ic, kk, S, mode, ht, hr, 0.0, 3.0)
    a, b = y1.view(-1), y2
    same = (a.view(torch.int32) == b.view(torch.int32)) | (torch.isnan(a) & torch.isnan(b))
    ok_f = bool(same.all())
    # backward (soft targets only)
    ok_b = True; err = 0.0
    if not ht:
        nb = lib.ssq_shift_bwd_ws_bytes(oc, ic, kk, S, 0)
        ws = torch.zeros(nb // 4 + 4, device=dev)
        gp1 = torch.full((ic, S), 7.0, device=dev); gp2 = torch.full((ic, S), -7.0, device=dev)
        gb1 = torch.empty(oc, K, device=dev) if mode == 1 else None
        gb2 = misaligned(gb1) if mode == 1 else None
        gs = g.clone()
        if special:   # keep the non-finite weights out of the sums (NaN * 0 would poison both the same way anyway)
            gs.view(-1)[[3, 17, 40]] = 0.0
        bwd(lib, gs, w, sd, d, z, p, beta, gp1, gb1, oc, ic, kk, S, mode, hr, 0.0, 3.0, ws)
        bwd(lib, misaligned(gs), wm, sd, d, z, p, bm, gp2, gb2, oc, ic, kk, S, mode, hr, 0.0, 3.0, ws)
        fin = torch.isfinite(gp2)
        err = float(((gp1 - gp2).abs()[fin] / (gp2.abs()[fin] + 1e-3 * gp2.abs()[fin].max().clamp_min(1e-20))).max()) if fin.any() else 0.0
        ok_b = err < 2e-5 and bool((torch.isfinite(gp1) == fin).all())
        if mode == 1:
            x, yv = gb1.view(-1), gb2
            sameb = (x.view(torch.int32) == yv.view(torch.int32)) | (torch.isnan(x) & torch.isnan(yv)) | ((x == 0) & (yv == 0))
            ok_b = ok_b and bool(sameb.all())
    tag = f"oc={oc} ic={ic} kk={kk} S={S} mode={mode} ht={ht} hr={hr} special={special} zint={zint}"
    print(("OK  " if ok_f and ok_b else "FAIL"), tag, f"fwd_bitexact={ok_f} bwd_relerr={err:.2e}", flush=True)
    if not ok_b and mode == 0 and not special:
        t = [((torch.round(w / sd[i][:, None]) + z[:, None]).clamp(0, 3) - z[:, None]) * sd[i][:, None] for i in range(S)]
        ref = torch.stack([(g * ti).view(oc, ic, kk).sum((0, 2)) for ti in t], -1)
        print("   ref", ref.view(-1)[:6].tolist(), "\n   vec", gp1.view(-1)[:6].tolist(), "\n   sca", gp2.view(-1)[:6].tolist())
    if not ok_f:
        bad = (~same).nonzero().view(-1)[:5]
        print("   first mismatches", bad.tolist(), a[bad].tolist(), b[bad].tolist())
    return ok_f and ok_b


NCU = "--ncu" in sys.argv


def timeit(fn, nbytes, reps=10):
    if NCU:
        fn(); fn(); torch.cuda.synchronize()
        return 0.0, 0.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    return nbytes / ms / 1e6, ms


def speed(lib, tag):
    oc, ic, kk, S = 4096, 4096, 9, 3
    K = ic * kk; n = oc * K
    torch.manual_seed(0)
    w = torch.randn(oc, K, device=dev) * 0.02
    d = (w.abs().amax(1) / 3).contiguous()
    z = torch.zeros(oc, device=dev)
    sd = torch.stack([d * s for s in (0.96875, 1.03125, 1.0)]).contiguous()
    p = torch.softmax(torch.randn(ic, S, device=dev), -1).mul(1.2).sub(0.1).clamp(0, 1).contiguous()
    beta = torch.randn(oc, K, device=dev)
    g = torch.randn(oc, K, device=dev)
    y = torch.empty(oc, K, device=dev); gb = torch.empty(oc, K, device=dev); gp = torch.empty(ic, S, device=dev)
    ws = torch.zeros(lib.ssq_shift_bwd_ws_bytes(oc, ic, kk, S, 0) // 4 + 4, device=dev)
    out = {}
    out["ada_fwd"] = timeit(lambda: fwd(lib, w, sd, d, z, p, beta, y, oc, ic, kk, S, 1, 0, 0, 0.0, 3.0), 12 * n)
    out["ada_bwd"] = timeit(lambda: bwd(lib, g, w, sd, d, z, p, beta, gp, gb, oc, ic, kk, S, 1, 0, 0.0, 3.0, ws), 16 * n)
    out["deq_fwd"] = timeit(lambda: fwd(lib, w, sd, d, z, p, None, y, oc, ic, kk, S, 0, 0, 0, 0.0, 3.0), 8 * n)
    out["deq_bwd"] = timeit(lambda: bwd(lib, g, w, sd, d, z, p, None, gp, None, oc, ic, kk, S, 0, 0, 0.0, 3.0, ws), 8 * n)
    # 1x1 convolution shape (kk == 1)
    oc2, ic2 = 8192, 18432
    n2 = oc2 * ic2
    w2 = w.view(-1)[:n2].view(oc2, ic2); d2 = (w2.abs().amax(1) / 3).contiguous(); z2 = torch.zeros(oc2, device=dev)
    sd2 = torch.stack([d2 * s for s in (0.96875, 1.03125, 1.0)]).contiguous()
    p2 = torch.softmax(torch.randn(ic2, S, device=dev), -1).mul(1.2).sub(0.1).clamp(0, 1).contiguous()
    ws2 = torch.zeros(lib.ssq_shift_bwd_ws_bytes(oc2, ic2, 1, S, 0) // 4 + 4, device=dev); gp2 = torch.empty(ic2, S, device=dev)
    out["ada_fwd_1x1"] = timeit(lambda: fwd(lib, w2, sd2, d2, z2, p2, beta, y, oc2, ic2, 1, S, 1, 0, 0, 0.0, 3.0), 12 * n2)
    out["ada_bwd_1x1"] = timeit(lambda: bwd(lib, g, w2, sd2, d2, z2, p2, beta, gp2, gb, oc2, ic2, 1, S, 1, 0, 0.0, 3.0, ws2), 16 * n2)
    print(tag, "  ".join(f"{k}: {v[0]:.0f} GB/s ({v[0] / PEAK:.3f}) {v[1]:.3f} ms" for k, v in out.items()), flush=True)


if __name__ == "__main__":
    libs = [a for a in sys.argv[1:] if not a.startswith("--")] or [str(L.LIB_PATH)]
    for i, path in enumerate(libs):
        lib = load(path)
        if (i == 0 or "--check-all" in sys.argv) and not NCU:
            allok = True
            for (oc, ic, kk) in ((64, 64, 9), (256, 64, 1), (128, 32, 2), (96, 24, 25), (40, 8, 4), (33, 20, 1), (7, 4, 3)):
                for S in (1, 2, 3, 4):
                    for mode in (0, 1):
                        for ht, hr in ((0, 0), (1, 1), (0, 1)):
                            allok &= check(lib, oc, ic, kk, S, mode, ht, hr, special=(S == 3))
            allok &= check(lib, 64, 64, 9, 3, 0, 0, 0, special=False, zint=False)
            allok &= check(lib, 300, 512, 9, 3, 1, 0, 0, special=True)
            allok &= check(lib, 300, 512, 9, 3, 0, 0, 0, special=True)
            print("ALL OK" if allok else "SOME FAILED", flush=True)
        speed(lib, path.split('/')[-1])
