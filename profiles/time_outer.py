"""Tensor-product getindex G[xs, ys] (src/interp.jl:273-306) on the separable path: 8192^2 f32 quadratic
image resampled on an 8192 x 8192 output lattice, and a 256^3 f64 cubic volume on 400^3."""
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

n = 8192
A = gb.jl_empty((n, n), torch.float32); gb.synth_uniform(A, 3)
G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
xs = torch.linspace(1.3, n - 0.7, 8192, device="cuda", dtype=torch.float32)
ys = torch.linspace(2.1, n - 1.9, 8192, device="cuda", dtype=torch.float32)
ms = t(lambda: G[xs, ys])
print("2-D quadratic f32 8192^2 -> 8192x8192 lattice: %.3f ms, %.1f Gpoints/s" % (ms, 8192 * 8192 / ms / 1e6))
X = torch.stack(torch.meshgrid(xs, ys, indexing="ij"), dim=-1).permute(1, 0, 2).reshape(-1, 2).contiguous()
v = torch.empty(X.shape[0], dtype=torch.float32, device="cuda")
ms2 = t(lambda: gb.getindex_(v, G, X))
print("  same points as an explicit query list (plain path): %.3f ms, %.1f Gpoints/s" % (ms2, X.shape[0] / ms2 / 1e6))
assert torch.equal(G[xs, ys].permute(1, 0).reshape(-1), v)
del A, G, X, v
m = 256
B = gb.jl_empty((m, m, m), torch.float64); gb.synth_uniform(B, 2)
H = gb.InterpGrid(B, gb.BCnan, gb.InterpCubic)
zs = [torch.linspace(1.0, m, 400, device="cuda", dtype=torch.float64) for _ in range(3)]
ms3 = t(lambda: H[zs[0], zs[1], zs[2]])
print("3-D cubic f64 256^3 -> 400^3 lattice: %.3f ms, %.1f Gpoints/s" % (ms3, 400 ** 3 / ms3 / 1e6))
