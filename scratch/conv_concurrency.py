import torch, time
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda')
def t(fn, reps=50):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
cb = torch.ops.aten.convolution_backward
for name, c, hw in (("layer1", 64, 56), ("layer2", 128, 28), ("layer3", 256, 14), ("layer4", 512, 7)):
    x = torch.randn(32, c, hw, hw, device=dev); w = torch.randn(c, c, 3, 3, device=dev) * 0.05
    g = torch.randn(32, c, hw, hw, device=dev)
    args = (g, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1)
    fwd = lambda: torch.nn.functional.conv2d(x, w, None, 1, 1)
    both = lambda: cb(*args, [True, True, False])
    dg = lambda: cb(*args, [True, False, False])
    wg = lambda: cb(*args, [False, True, False])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def par():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1): dg()
        with torch.cuda.stream(s2): wg()
        cur.wait_stream(s1); cur.wait_stream(s2)
    # graph-captured versions (what the engine replays)
    def graphed(fn):
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3): fn()
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr): fn()
        return gr.replay
    seq = lambda: (dg(), wg())
    print(f"{name}: fwd {t(fwd):.0f} us | bwd both {t(both):.0f} | dgrad {t(dg):.0f} + wgrad {t(wg):.0f} | graph seq {t(graphed(seq)):.0f} | graph parallel {t(graphed(par)):.0f}", flush=True)
