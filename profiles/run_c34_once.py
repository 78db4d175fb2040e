"""Minimal C3 / C4 drivers for ncu captures (plain path): python profiles/run_c34_once.py c3|c4"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    n, nq = 8192, 100_000_000
    A = gb.jl_empty((n, n), torch.float32); gb.synth_uniform(A, 3)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    X = torch.empty((nq, 2), dtype=torch.float32, device="cuda"); gb.synth_uniform(X, 14, lo=1.0, hi=float(n))
    v = torch.empty(nq, dtype=torch.float32, device="cuda"); g = torch.empty((nq, 2), dtype=torch.float32, device="cuda"); h = torch.empty((nq, 2, 2), dtype=torch.float32, device="cuda")
    for _ in range(2): gb.valgradhess_(v, g, h, G, X)
else:
    n, nq = 64, 100_000_000
    A = gb.jl_empty((n,) * 4, torch.float64); gb.synth_uniform(A, 4)
    rng = gb.FloatRange(-63.0, 2.0, 64, 63.0)
    G = gb.InterpGrid(A, gb.BCperiodic, gb.InterpLinear); CG = gb.CoordInterpGrid((rng,) * 4, G)
    X = torch.empty((nq, 4), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, 15, lo=-1.0 + (-16.0) * 2.0 / 63.0, hi=1.0 + 16.0 * 2.0 / 63.0)
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    for _ in range(2): gb.getindex_(v, CG, X)
torch.cuda.synchronize()
print("ok")
