"""Time the C2 step (valgrad, 1e7 device-resident queries, 258^3 cubic grid) with CUDA events.
Usage: python profiles/time_c2.py [fast] [tiled|plain] -- prints median / min ms over 10 steps.
Environment knobs of the library (GB_TILED_BATCH, GB_TILE_KB) are read once per process."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import grid_jl_b200 as gb

fast = "fast" in sys.argv[1:]
tiled = True if "tiled" in sys.argv[1:] else (False if "plain" in sys.argv[1:] else None)
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64)
gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
v = torch.empty(nq, dtype=torch.float64, device="cuda")
g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
for _ in range(3):
    gb.valgrad_(v, g, G, X, fast=fast, tiled=tiled)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    gb.valgrad_(v, g, G, X, fast=fast, tiled=tiled)
    ev[i + 1].record()
torch.cuda.synchronize()
t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
print("%s batch=%s tile_kb=%s : median %.3f ms  min %.3f ms  checksum %.6f" % (
    "FAST " if fast else "EXACT", os.environ.get("GB_TILED_BATCH", "-"), os.environ.get("GB_TILE_KB", "-"), t[5], t[0], float(v[:1000].sum())))
