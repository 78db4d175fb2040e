"""Minimal C5 driver for ncu captures: one fused restrict 1024^3 -> 513^3 and one fused prolong back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = gb.jl_empty((n, n, n), torch.float64)
gb.synth_uniform(A, 5)
for _ in range(2):
    R = gb.restrict(A, [True, True, True])
    P = gb.prolong(R, [n, n, n])
torch.cuda.synchronize()
print("ok", tuple(R.shape), tuple(P.shape))
