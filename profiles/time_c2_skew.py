import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
v = torch.empty(nq, dtype=torch.float64, device="cuda"); g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
for name, lo, hi in (("uniform", 1.0, 256.0), ("one tile", 100.0, 108.0), ("one cell", 100.0, 100.9), ("half domain", 1.0, 128.0)):
    X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=lo, hi=hi)
    for tiled in (None, False):
        for _ in range(2): gb.valgrad_(v, g, G, X, tiled=tiled)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): gb.valgrad_(v, g, G, X, tiled=tiled)
        e1.record(); torch.cuda.synchronize()
        print("%-12s %-6s %.3f ms" % (name, "auto" if tiled is None else "plain", e0.elapsed_time(e1) / 3), flush=True)
