"""reproduces tests/test_kernels_gpu.py::test_inp_scale_search_many_columns_on_candidate_boundaries[5-...] and prints the mismatching columns"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ssq_oracle as O
from shiftedscalequantization_b200 import ops
seed, oc, k, level, bits, thr = [type(d)(a) for a, d in zip(sys.argv[1:], (5, 2048, 260, 64, 2, 0.7))] if len(sys.argv) > 6 else (5, 2048, 260, 64, 2, 0.7)
r = np.random.default_rng(7000 + seed)
L = 2 ** bits
w = (r.standard_normal((oc, k)) * 0.01).astype(np.float32)
d, z, raw = zip(*[O.max_init(row, bits) for row in w])
d = np.array(d, np.float32); raw = np.array(raw, np.float32)
w *= np.float32(0.25)
lo = np.float32(0.0 - 0.5 / (L - 1) * thr); hi = np.float32(1.0 + 0.5 / (L - 1) * thr)
cand_np = np.array([i / level for i in range(level, 0, -1)], dtype=np.float32)
zero = np.rint(raw / d)
vmax = (d * (hi * (L - 1) - zero)).astype(np.float32); vmin = (d * (lo * (L - 1) - zero)).astype(np.float32)
cols = r.permutation(k)[: (3 * k) // 4]
for j in cols:
    c = cand_np[r.integers(0, level)]
    for i in r.integers(0, oc, size=r.integers(1, 4)):
        end = vmax[i] if r.random() < 0.5 else vmin[i]
        w[i, j] = np.float32(end * c) * np.float32(1 + r.integers(-4, 5) * 6e-8)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
cand = cu(cand_np)
fast = torch.ones(k, device="cuda"); brute = torch.ones(k, device="cuda")
ops.inp_scale_search(cu(w), cu(d), cu(raw), cand, L - 1, float(lo), float(hi), fast)
ops.inp_scale_search(cu(w), cu(d), cu(raw), cand, L - 1, float(lo), float(hi), brute, force_brute=True)
f, b = fast.cpu().numpy(), brute.cpu().numpy()
bad = np.nonzero(f != b)[0]
print("zero range", zero.min(), zero.max(), "mismatching columns", len(bad), "of", k, "unique fast", len(np.unique(f)), "unique brute", len(np.unique(b)))
ref = O.inp_scale_search(w, d.reshape(-1, 1), raw.reshape(-1, 1), L, level, thr).reshape(-1)
print("brute == oracle:", np.array_equal(b, ref), " fast == oracle:", np.array_equal(f, ref))
for j in bad[:8]:
    # per-row exact prefix of this column (numpy restatement of the predicate)
    col = w[:, j]
    pre = np.full(oc, level)
    for jj, c in enumerate(cand_np):
        g = ((col / c) / d + zero) / np.float32(L - 1)
        fits = (g > lo) & (g < hi)
        pre = np.where((pre == level) & ~fits, jj, pre)            # first non-fitting candidate index = prefix length
    rows = np.argsort(pre)[:3]
    print(f"col {j}: fast {f[j]} brute {b[j]} oracle {ref[j]}; min prefix {pre.min()} at rows {rows.tolist()} prefixes {pre[rows].tolist()} w {col[rows].tolist()} d {d[rows].tolist()} zero {zero[rows].tolist()}")
    # is the fitting set a prefix for those rows?
    for i in rows[:2]:
        g = ((np.float32(col[i]) / cand_np) / d[i] + zero[i]) / np.float32(L - 1)
        fits = (g > lo) & (g < hi)
        print("   row", i, "fits pattern:", "".join("1" if x else "0" for x in fits))
