"""FAST (separable FMA) vs EXACT (reference arithmetic, bit-identical to the oracle) vs the
extended-precision truth on the full C2 workload, in units of ulp(S), S = sum_i |c_i a_i| per output
(oracle.evaluate_truth).  Usage: python profiles/fast_error_c2.py [sample]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import grid_jl_b200 as gb
from oracle import oracle as O

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64)
gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
v, g = gb.valgrad_batch(G, X)
vf, gf = gb.valgrad_batch(G, X, fast=True)
tv, tg, sv, sg = O.evaluate_truth(G.coefs, "nan", "cubic", X[:ns].cpu().numpy())
eps = 2.220446049250313e-16
ve, ge, vfn, gfn = (a[:ns].cpu().numpy() for a in (v, g, vf, gf))


def rep(name, a, b, s):
    r = np.abs(a - b) / (eps * s)
    print("%-28s max %.2f  p99.9 %.2f  p99 %.2f  mean %.3f  ulp(S);  > 4 ulp: %d of %d" % (name, r.max(), np.quantile(r, 0.999), np.quantile(r, 0.99), r.mean(), int((r > 4).sum()), r.size))


rep("value    FAST  vs EXACT", vfn, ve, sv)
rep("value    FAST  vs truth", vfn, tv, sv)
rep("value    EXACT vs truth", ve, tv, sv)
rep("gradient FAST  vs EXACT", gfn, ge, sg)
rep("gradient FAST  vs truth", gfn, tg, sg)
rep("gradient EXACT vs truth", ge, tg, sg)
