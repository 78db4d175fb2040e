import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
v = torch.empty(nq, dtype=torch.float64, device="cuda"); g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
for fast in (False, True):
    for _ in range(3): gb.valgrad_(v, g, G, X, fast=fast)
    gb.profile_enable(True)
    acc = {}
    for _ in range(5):
        gb.valgrad_(v, g, G, X, fast=fast)
        for k, val in gb.profile_read().items(): acc[k] = acc.get(k, 0) + val / 5
    gb.profile_enable(False)
    print("BRICK_CTAS=%s fast=%s brick %.3f ms total %.3f" % (os.environ.get("GB_BRICK_CTAS", "-"), fast, acc["eval_brick"], sum(acc.values())))
