"""Blocked (4^N Morton blocks) vs column-major resident layout on the plain path: C3 (2-D quadratic f32 8192^2,
valgradhess) and C4 (4-D linear f64 64^4, getindex), 1e8 device-resident queries, random and coherent
(queries sorted along a raster scan of the grid) -- run once per layout:

    python profiles/time_blocked.py                      # blocked (default for these grids)
    GB_BLOCK_MIN_MB=1e9 python profiles/time_blocked.py  # column-major
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import grid_jl_b200 as gb


def timed(fn, reps=7):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def run(name, shape, bc, it, mask, dt=torch.float32, nq=100_000_000):
    N = len(shape)
    A = gb.jl_empty(shape, dt)
    gb.synth_uniform(A, 7, -1.0, 1.0)
    g = gb.InterpGrid(A, bc, it)
    del A
    X = torch.empty((nq, N), dtype=dt, device="cuda")
    gb.synth_uniform(X, 11, 0.0, 1.0)
    X = 1.0 + X * (torch.tensor(shape, dtype=dt, device="cuda") - 1.0)
    v = torch.empty(nq, dtype=dt, device="cuda")
    gr = torch.empty((nq, N), dtype=dt, device="cuda") if mask & 2 else None
    h = torch.empty((nq, N, N), dtype=dt, device="cuda") if mask & 4 else None
    for label in ("random", "coherent"):
        if label == "coherent":  # raster order: last dimension slowest, 1st fastest
            key = torch.zeros(nq, dtype=torch.int64, device="cuda")
            for d in range(N - 1, -1, -1):
                key = key * shape[d] + X[:, d].long()
            X = X[torch.argsort(key)].contiguous()
            del key
        best, med = timed(lambda: g._eval(X, mask, val=v, grad=gr, hess=h))
        print(f"{name:4s} {label:9s} layout={'column-major' if os.environ.get('GB_BLOCK_MIN_MB') else 'blocked':12s} best {best:7.3f} ms  median {med:7.3f} ms  "
              f"{nq / best / 1e6:6.2f} Gpoints/s", flush=True)


if __name__ == "__main__":
    run("C3", (8192, 8192), gb.BCreflect, gb.InterpQuadratic, 7)
    torch.cuda.empty_cache()
    run("C4", (64, 64, 64, 64), gb.BCnil, gb.InterpLinear, 1, torch.float64)
