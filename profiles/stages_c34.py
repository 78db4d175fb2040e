"""Stage timings (gb_profile_*) of the tiled path on the C3 and C4 workloads, and the plain path beside it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

which = sys.argv[1:] or ["c3", "c4"]
if "c3" in which:
    n, nq = 8192, 100_000_000
    A = gb.jl_empty((n, n), torch.float32); gb.synth_uniform(A, 3)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    X = torch.empty((nq, 2), dtype=torch.float32, device="cuda"); gb.synth_uniform(X, 14, lo=1.0, hi=float(n))
    v = torch.empty(nq, dtype=torch.float32, device="cuda"); g = torch.empty((nq, 2), dtype=torch.float32, device="cuda"); h = torch.empty((nq, 2, 2), dtype=torch.float32, device="cuda")
    for tiled in (True, False):
        gb.profile_enable(True)
        ms = t(lambda: gb.valgradhess_(v, g, h, G, X, tiled=tiled))
        print("C3 tiled=%s: %.2f ms" % (tiled, ms), gb.profile_read() if tiled else "")
        gb.profile_enable(False)
    del A, G, X, v, g, h
if "c4" in which:
    n, nq = 64, 100_000_000
    A = gb.jl_empty((n,) * 4, torch.float64); gb.synth_uniform(A, 4)
    rng = gb.FloatRange(-63.0, 2.0, 64, 63.0)
    X = torch.empty((nq, 4), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, 15, lo=-1.0 + (-16.0) * 2.0 / 63.0, hi=1.0 + 16.0 * 2.0 / 63.0)
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    for name, bc in (("periodic", gb.BCperiodic), ("fill", -1.0)):
        G = gb.InterpGrid(A, bc, gb.InterpLinear); CG = gb.CoordInterpGrid((rng,) * 4, G)
        for tiled in (True, False):
            gb.profile_enable(True)
            ms = t(lambda: gb.getindex_(v, CG, X, tiled=tiled))
            print("C4 %s tiled=%s: %.2f ms" % (name, tiled, ms), gb.profile_read() if tiled else "")
            gb.profile_enable(False)
