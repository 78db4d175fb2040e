"""2-D image pyramid: restrict(A, [true,true]) 8192^2 -> 4097^2 -> ... -> 17^2 and prolong back, fused (gb_restrict12 /
gb_prolong12: one pass per level) against the per-dimension calls.  CUDA events, best of 7."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for dt in (torch.float32, torch.float64):
    A = gb.jl_empty((8192, 8192), dt)
    gb.synth_uniform(A, 5)
    for fused in (True, False):
        levels = []
        def down():
            levels.clear()
            a = A
            for _ in range(9):
                levels.append(a)
                a = gb.restrict(a, [True, True], fused=fused)
            levels.append(a)
        t_r = timed(down)
        def up():
            a = levels[-1]
            for k in range(8, -1, -1):
                a = gb.prolong(a, list(levels[k].shape), fused=fused)
        t_p = timed(up)
        print(f"{dt} 8192^2 pyramid, {'fused' if fused else 'per-dimension'}: restrict chain {t_r:.3f} ms, prolong chain {t_p:.3f} ms", flush=True)
