"""C2 step with spatially pre-sorted queries (same 1e7 uniform points, ordered by grid tile) through the
ordinary API: what the standard pipeline gives a caller whose queries are coherent."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
cell = X.floor().long()
key = (cell[:, 2] // 8) * 4096 + (cell[:, 1] // 16) * 64 + (cell[:, 0] // 16)
Xs = X[torch.argsort(key)].contiguous()
Xm = X[torch.argsort(cell[:, 2] * 65536 + cell[:, 1] * 256 + cell[:, 0])].contiguous()  # cell order (finer)
v = torch.empty(nq, dtype=torch.float64, device="cuda"); g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
for name, Q in (("random", X), ("tile-sorted", Xs), ("cell-sorted", Xm)):
    for fast in (False, True):
        for tiled in (None, False):
            for _ in range(3): gb.valgrad_(v, g, G, Q, fast=fast, tiled=tiled)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): gb.valgrad_(v, g, G, Q, fast=fast, tiled=tiled)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            st = ""
            if tiled is None:
                gb.profile_enable(True); gb.valgrad_(v, g, G, Q, fast=fast); st = str({k: round(x, 3) for k, x in gb.profile_read().items()}); gb.profile_enable(False)
            print("%-12s %-5s %-6s %.3f ms  %.1f Gpoints/s  %s" % (name, "FAST" if fast else "EXACT", "brick" if tiled is None else "plain", ms, nq / ms / 1e6, st), flush=True)
