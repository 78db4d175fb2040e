"""Grids with more tiles than the shared-memory sort histograms hold (> 12000): 3-D cubic f64 valgrad on 384^3 and
512^3 (stored 386^3 / 514^3: 26 k / 71 k tiles), uniformly random device-resident queries.  Brick path (global-histogram
sort) against the plain path it used to fall back to; both bit-identical to each other and, on a subsample, to the
CPU oracle.

    python profiles/time_big3d.py [n ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import grid_jl_b200 as gb
from oracle import oracle as O


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def run(n, nq):
    A = gb.jl_empty((n, n, n), torch.float64)
    gb.synth_uniform(A, 7, -1.0, 1.0)
    g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    del A
    X = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, 11, 1.0, float(n))
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    gr = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    v2, gr2 = torch.empty_like(v), torch.empty_like(gr)
    t_auto = timed(lambda: g._eval(X, 3, val=v, grad=gr))
    gb.profile_enable(True)
    g._eval(X, 3, val=v, grad=gr)
    stages = gb.profile_read()
    gb.profile_enable(False)
    t_plain = timed(lambda: g._eval(X, 3, val=v2, grad=gr2, tiled=False))
    t_fast = timed(lambda: g._eval(X, 3, val=v2, grad=gr2, fast=True))
    g._eval(X, 3, val=v2, grad=gr2, tiled=False)
    same = bool(torch.equal(v.view(torch.int64), v2.view(torch.int64)) and torch.equal(gr.view(torch.int64), gr2.view(torch.int64)))
    sub = slice(0, nq, max(1, nq // 20000))
    ov, og, _, _ = O.evaluate(g.coefs, "nan", "cubic", X[sub].cpu().numpy(), 3)
    par = bool(np.array_equal(v[sub].cpu().numpy().view(np.int64), ov.view(np.int64)) and np.array_equal(gr[sub].cpu().numpy().view(np.int64), og.view(np.int64)))
    print(f"{n}^3 f64 cubic valgrad, {nq:.0e} queries: default {t_auto:7.3f} ms ({nq / t_auto / 1e6:5.2f} Gpoints/s)  FAST {t_fast:7.3f} ms  plain {t_plain:7.3f} ms  "
          f"brick==plain {same}  oracle subsample {par}\n    stages {({k: round(x, 3) for k, x in stages.items()})}", flush=True)


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [256, 384, 512]
    for n in sizes:
        for nq in (10_000_000, 40_000_000):
            run(n, nq)
            torch.cuda.empty_cache()
