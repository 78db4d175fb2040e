"""interp_invert! sweeps on a device-resident array, per dimension (CUDA events, best of 9): the C2 prefilter
(258^3 f64 cubic, Woodbury), plus quadratic variants.

    python profiles/time_prefilter.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import grid_jl_b200 as gb


def timed(fn, reps=9):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def run(shape, dt, bc, it, name):
    A = gb.jl_empty(shape, dt)
    gb.synth_uniform(A, 7, -1.0, 1.0)
    per = [timed(lambda d=d: gb.interp_invert_(A, bc, it, [d + 1])) for d in range(len(shape))]
    gb.synth_uniform(A, 7, -1.0, 1.0)
    allt = timed(lambda: gb.interp_invert_(A, bc, it))
    n = 1
    for s in shape:
        n *= s
    esz = 8 if dt == torch.float64 else 4
    print(f"{name}: per-dim sweeps {' / '.join(f'{t:.3f}' for t in per)} ms, all dims {allt:.3f} ms "
          f"({2 * len(shape) * n * esz / allt / 1e6:.0f} GB/s of one read + one write per sweep)", flush=True)


if __name__ == "__main__":
    run((258, 258, 258), torch.float64, gb.BCnan, gb.InterpCubic, "258^3 f64 cubic (Woodbury)")
    run((258, 258, 258), torch.float32, gb.BCnan, gb.InterpCubic, "258^3 f32 cubic (Woodbury)")
    run((256, 256, 256), torch.float64, gb.BCreflect, gb.InterpQuadratic, "256^3 f64 quadratic reflect")
    run((256, 256, 256), torch.float64, gb.BCperiodic, gb.InterpQuadratic, "256^3 f64 quadratic periodic (Woodbury)")
    run((8192, 8192), torch.float32, gb.BCreflect, gb.InterpQuadratic, "8192^2 f32 quadratic reflect (C3)")
