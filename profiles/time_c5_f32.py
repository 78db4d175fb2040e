"""C5 chain in Float32 (1024^3 -> 17^3 and back), CUDA events, best of 5."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

for dt in (torch.float32, torch.float64):
    A = gb.jl_empty((1024, 1024, 1024), dt)
    gb.synth_uniform(A, 5)
    levels = []
    def down():
        levels.clear()
        a = A
        for _ in range(6):
            levels.append(a)
            a = gb.restrict(a, [True, True, True])
        levels.append(a)
    t_r = timed(down)
    def up():
        a = levels[-1]
        for k in range(5, -1, -1):
            a = gb.prolong(a, list(levels[k].shape))
    t_p = timed(up)
    print(f"{dt}: restrict chain {t_r:.3f} ms, prolong chain {t_p:.3f} ms", flush=True)
    del A, levels[:]
    torch.cuda.empty_cache()
