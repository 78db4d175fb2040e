"""All five BASELINE.json configs on one B200: device-resident throughput (CUDA events, best and
median of `reps`), % of the measured HBM roofline from SURVEY 8d's algorithmic bytes per unit, and a
parity check of a strided subsample against the CPU oracle (bit-identical, EXACT mode).

    python profiles/bench_configs.py [c1 c2 c3 c4 c5 pref] [--reps 5] [--json out.json]

This is the per-config table of profiles/README.md; bench.py stays the single-line C2 contract."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import grid_jl_b200 as gb
from oracle import oracle as O

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def timed(fn, reps):
    fn()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def row(name, units, unit_name, bytes_per_unit, best_ms, med_ms, parity, note=""):
    gups = units / (best_ms * 1e-3) / 1e9
    frac = gups * bytes_per_unit / PEAK
    r = {"config": name, "units": units, "unit": unit_name, "bytes_per_unit": bytes_per_unit, "best_ms": best_ms, "median_ms": med_ms,
         "G%s_per_s" % unit_name: gups, "achieved_GBs": gups * bytes_per_unit, "roofline_frac": frac, "parity_vs_oracle": parity, "note": note}
    print(json.dumps(r), flush=True)
    return r


def sub(t, idx):
    return t[idx].cpu().numpy()


def c1(reps):
    n = 1024
    A = gb.jl_empty((n, n), torch.float64)
    gb.synth_uniform(A, 1)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    out = []
    for nq in (1_000_000, 100_000_000):
        X = torch.empty((nq, 2), dtype=torch.float64, device="cuda")
        gb.synth_uniform(X, 11, lo=1.0, hi=float(n))
        v = torch.empty(nq, dtype=torch.float64, device="cuda")
        best, med = timed(lambda: gb.getindex_(v, G, X), reps)
        idx = torch.arange(0, nq, max(1, nq // 100_000), device="cuda")
        ov, _, _, _ = O.evaluate(G.coefs, "reflect", "quadratic", sub(X, idx), 1)
        par = bool(np.array_equal(ov, sub(v, idx), equal_nan=True))
        # the wide parity set of SURVEY 8d (mirror images)
        Xw = torch.empty((200_000, 2), dtype=torch.float64, device="cuda")
        gb.synth_uniform(Xw, 13, lo=-1536.0, hi=2560.0)
        vw = gb.getindex_batch(G, Xw)
        ow, _, _, _ = O.evaluate(G.coefs, "reflect", "quadratic", Xw.cpu().numpy(), 1)
        par = par and bool(np.array_equal(ow, vw.cpu().numpy(), equal_nan=True))
        out.append(row("C1 2-D Quadratic BCreflect 1024^2 f64 getindex, %.0e queries" % nq, nq, "points", 24, best, med, par))
        del X, v
    return out


def c2(reps):
    n, nq = 256, 10_000_000
    A = gb.jl_empty((n, n, n), torch.float64)
    gb.synth_uniform(A, 2)
    G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    X = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    out = []
    for fast in (False, True):
        best, med = timed(lambda: gb.valgrad_(v, g, G, X, fast=fast), reps)
        par = None
        if not fast:
            idx = torch.arange(0, nq, 100, device="cuda")
            ov, og, _, _ = O.evaluate(G.coefs, "nan", "cubic", sub(X, idx), 3)
            par = bool(np.array_equal(ov, sub(v, idx), equal_nan=True) and np.array_equal(og, sub(g, idx), equal_nan=True))
        out.append(row("C2 3-D Cubic BCnan 256^3 f64 valgrad, 1e7 queries, %s" % ("FAST" if fast else "EXACT"), nq, "points", 56, best, med, par))
    best, med = timed(lambda: gb.InterpGrid(A, gb.BCnan, gb.InterpCubic), reps)
    co = O.make_coefs(A.cpu().numpy(), "nan", "cubic")
    par = bool(np.array_equal(co, G.coefs))
    out.append(row("C2 prefilter: pad + 3 interp_invert! sweeps of 258^3 f64 (incl. allocation)", (n + 2) ** 3, "points", 64, best, med, par))
    return out


def c3(reps):
    n, nq = 8192, 100_000_000
    A = gb.jl_empty((n, n), torch.float32)
    gb.synth_uniform(A, 3)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    X = torch.empty((nq, 2), dtype=torch.float32, device="cuda")
    gb.synth_uniform(X, 14, lo=1.0, hi=float(n))
    v = torch.empty(nq, dtype=torch.float32, device="cuda")
    g = torch.empty((nq, 2), dtype=torch.float32, device="cuda")
    h = torch.empty((nq, 2, 2), dtype=torch.float32, device="cuda")
    best, med = timed(lambda: gb.valgradhess_(v, g, h, G, X), reps)
    idx = torch.arange(0, nq, 1000, device="cuda")
    ov, og, oh, _ = O.evaluate(G.coefs, "reflect", "quadratic", sub(X, idx), 7)
    par = bool(np.array_equal(ov, sub(v, idx), equal_nan=True) and np.array_equal(og, sub(g, idx), equal_nan=True) and np.array_equal(oh, sub(h, idx), equal_nan=True))
    return [row("C3 2-D Quadratic BCreflect 8192^2 f32 valgradhess, 1e8-query slice of the 1e9 batch (x10 per GPU-pass)", nq, "points", 36, best, med, par)]


def c4(reps):
    n, nq = 64, 100_000_000
    A = gb.jl_empty((n,) * 4, torch.float64)
    gb.synth_uniform(A, 4)
    rng = gb.FloatRange(-63.0, 2.0, 64, 63.0)  # -1.0 : 2/63 : 1.0
    out = []
    X = torch.empty((nq, 4), dtype=torch.float64, device="cuda")
    # index coordinates U[1-16, 64+16] -> user coordinates through the range's inverse map
    gb.synth_uniform(X, 15, lo=-1.0 + (-16.0) * 2.0 / 63.0, hi=1.0 + 16.0 * 2.0 / 63.0)
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    for name, bc in (("periodic", gb.BCperiodic), ("nearest", gb.BCnearest), ("fill", -1.0)):
        G = gb.InterpGrid(A, bc, gb.InterpLinear)
        CG = gb.CoordInterpGrid((rng,) * 4, G)
        best, med = timed(lambda: gb.getindex_(v, CG, X), reps)
        idx = torch.arange(0, nq, 1000, device="cuda")
        ov, _, _, _ = O.evaluate(G.coefs, name, "linear", sub(X, idx), 1, affine=[rng.row()] * 4)
        par = bool(np.array_equal(ov, sub(v, idx), equal_nan=True))
        out.append(row("C4 4-D Linear BC%s 64^4 f64 getindex via CoordInterpGrid, 1e8 queries" % name, nq, "points", 40, best, med, par))
        del G, CG
    return out


def c5(reps):
    n = 1024
    A = gb.jl_empty((n, n, n), torch.float64)
    gb.synth_uniform(A, 5)
    sizes = [n]
    while sizes[-1] > 17:
        sizes.append(gb.restrict_size(sizes[-1]))
    levels = [A]
    torch.cuda.synchronize()

    def down():
        cur = A
        lv = [A]
        for _ in sizes[1:]:
            cur = gb.restrict(cur, [True, True, True])
            lv.append(cur)
        return lv

    best_d, med_d = timed(lambda: down(), reps)
    levels = down()
    coarse = levels[-1]

    def up():
        cur = coarse
        for s in reversed(sizes[:-1]):
            cur = gb.prolong(cur, [s, s, s])
        return cur

    best_u, med_u = timed(lambda: up(), reps)
    # parity on the small end of the chain (the oracle handles 129^3 in seconds)
    k = sizes.index(129)
    a129 = gb.jl_to_numpy(levels[k])
    par = bool(np.array_equal(O.restrict_flags(a129, [True] * 3), gb.jl_to_numpy(levels[k + 1])))
    par = par and bool(np.array_equal(O.prolong_sizes(gb.jl_to_numpy(levels[k + 1]), [129] * 3), gb.jl_to_numpy(gb.prolong(levels[k + 1], [129] * 3))))
    units = n ** 3
    bpu = 8.0 * sum(a ** 3 + b ** 3 for a, b in zip(sizes[:-1], sizes[1:])) / units
    return [row("C5 restrict chain %s f64 (fused 3-D)" % "->".join(map(str, sizes)), units, "points", bpu, best_d, med_d, par),
            row("C5 prolong chain back up (fused 3-D)", units, "points", bpu, best_u, med_u, par)]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rows = []
    for w in a.which:
        t0 = time.time()
        rows += {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}[w](a.reps)
        torch.cuda.empty_cache()
        print("# %s done in %.1f s" % (w, time.time() - t0), file=sys.stderr, flush=True)
    if a.json:
        json.dump({"peak_GBs": PEAK, "rows": rows}, open(a.json, "w"), indent=1)
