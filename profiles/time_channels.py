"""Multi-valued grids (f-1): an RGB image (4096^2 x 3, f32, quadratic, reflect) evaluated at 2e7 random points,
InterpGridChannels (taps and weights once per point) vs three separate InterpGrids."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

n, nq, C = 4096, 20_000_000, 3
A = gb.jl_empty((n, n, C), torch.float32); gb.synth_uniform(A, 7)
G = gb.InterpGridChannels(A, gb.BCreflect, gb.InterpQuadratic)
X = torch.empty((nq, 2), dtype=torch.float32, device="cuda"); gb.synth_uniform(X, 8, lo=1.0, hi=float(n))
v = torch.empty((nq, C), dtype=torch.float32, device="cuda")
ms = t(lambda: gb.getindex_(v, G, X))
singles = [gb.InterpGrid(A[:, :, c].clone(memory_format=torch.contiguous_format) if False else gb.jl_from_numpy(gb.jl_to_numpy(A)[:, :, c].copy(order="F")), gb.BCreflect, gb.InterpQuadratic) for c in range(C)]
v1 = [torch.empty(nq, dtype=torch.float32, device="cuda") for _ in range(C)]
ms3 = t(lambda: [gb.getindex_(v1[c], singles[c], X) for c in range(C)])
for c in range(C):
    assert torch.equal(v[:, c], v1[c])
print("RGB 4096^2 f32 quadratic, 2e7 points: channels %.3f ms (%.1f Gpoints/s, %.1f Gvalues/s) vs 3 grids %.3f ms; bit-identical" % (ms, nq / ms / 1e6, C * nq / ms / 1e6, ms3))
