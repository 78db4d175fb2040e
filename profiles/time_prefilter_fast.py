"""interp_invert! exact vs GB_FAST on the shapes of the BASELINE configs: event timings of the in-place call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
def timed(fn, reps=7):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for name, shape, dt, bc, it in (("C3 8192^2 f32 quadratic reflect", (8192, 8192), torch.float32, gb.BCreflect, gb.InterpQuadratic),
                                ("C1 1024^2 f64 quadratic reflect", (1024, 1024), torch.float64, gb.BCreflect, gb.InterpQuadratic),
                                ("4096^2 f64 cubic nan", (4096, 4096), torch.float64, gb.BCnan, gb.InterpCubic),
                                ("8192^2 f32 quadratic periodic", (8192, 8192), torch.float32, gb.BCperiodic, gb.InterpQuadratic),
                                ("C2 258^3 f64 cubic nan (many short lines: exact kernels either way)", (258, 258, 258), torch.float64, gb.BCnan, gb.InterpCubic)):
    A = gb.jl_empty(shape, dt); gb.synth_uniform(A, 3)
    B = A.clone()
    tc = timed(lambda: B.copy_(A))
    for fast in (False, True):
        t = timed(lambda: (B.copy_(A), gb.interp_invert_(B, bc, it, fast=fast)))
        el = 1
        for s in shape: el *= s
        by = el * A.element_size() * 2 * len(shape)
        print("%-75s %s  %.3f ms (copy %.3f subtracted)  %.0f GB/s algorithmic" % (name, "FAST " if fast else "EXACT", t - tc, tc, by / ((t - tc) * 1e-3) / 1e9), flush=True)
