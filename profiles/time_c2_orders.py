"""C2 (256^3 f64 cubic valgrad, 1e7 queries) under different ORDERS of the same query set: stage times of the tiled path
(gb_profile_*), whether the probe found the batch coherent (plain kernel, in place), and the plain path next to it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grid_jl_b200 as gb

n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
fl = X.floor().to(torch.int64)
def morton(f):
    k = torch.zeros_like(f[:, 0])
    for b in range(9):
        for d in range(3):
            k |= ((f[:, d] >> b) & 1) << (3 * b + d)
    return k
orders = {
    "random": None,
    "blocks 16x16x8 (user cells)": (fl[:, 2] >> 3) * 65536 + (fl[:, 1] >> 4) * 256 + (fl[:, 0] >> 4),
    "blocks 8x8x4": (fl[:, 2] >> 2) * 65536 + (fl[:, 1] >> 3) * 256 + (fl[:, 0] >> 3),
    "blocks 32x32x32": (fl[:, 2] >> 5) * 65536 + (fl[:, 1] >> 5) * 256 + (fl[:, 0] >> 5),
    "morton (cell)": morton(fl),
    "raster (cell)": fl[:, 2] * 65536 + fl[:, 1] * 256 + fl[:, 0],
}
v = torch.empty(nq, dtype=torch.float64, device="cuda"); g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for name, key in orders.items():
    Xs = X if key is None else X[torch.argsort(key)].contiguous()
    for fast in (False, True):
        t = timed(lambda: gb.valgrad_(v, g, G, Xs, fast=fast))
        tp = timed(lambda: gb.valgrad_(v, g, G, Xs, fast=fast, tiled=False))
        gb.profile_enable(True); gb.valgrad_(v, g, G, Xs, fast=fast); st = gb.profile_read(); gb.profile_enable(False)
        print("%-30s %-5s tiled %.3f ms  plain %.3f ms  coherent %d  stages %s" % (name, "FAST" if fast else "EXACT", t, tp, st.get("coherent", -1),
              " ".join("%s %.3f" % (k[:6], x) for k, x in st.items() if k != "coherent")), flush=True)
