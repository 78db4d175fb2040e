"""Minimal C2 driver for ncu captures: build the 256^3 cubic grid (pad + 3 prefilter sweeps), run a
few valgrad steps.  Usage: python profiles/run_c2_once.py [steps] [fast] [tiled|plain]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import grid_jl_b200 as gb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fast = "fast" in sys.argv[2:]
tiled = True if "tiled" in sys.argv[2:] else (False if "plain" in sys.argv[2:] else None)
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64)
gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
v = torch.empty(nq, dtype=torch.float64, device="cuda")
g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
for _ in range(steps):
    gb.valgrad_(v, g, G, X, fast=fast, tiled=tiled)
torch.cuda.synchronize()
print("ok", float(v[:1000].sum()))
