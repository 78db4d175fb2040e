"""Minimal C1 driver for ncu captures: 1024^2 f64 quadratic reflect, 1e8 getindex queries (plain path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
n, nq = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
A = gb.jl_empty((n, n), torch.float64); gb.synth_uniform(A, 1)
G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
X = torch.empty((nq, 2), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 11, lo=1.0, hi=float(n))
v = torch.empty(nq, dtype=torch.float64, device="cuda")
for _ in range(2):
    gb.getindex_(v, G, X)
torch.cuda.synchronize()
print("ok", float(v[:100].sum()))
