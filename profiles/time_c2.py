"""C2 step (256^3 f64 cubic valgrad, 1e7 uniform queries): median step time and stage times, EXACT and FAST."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb

n, nq = 256, int(float(os.environ.get("NQ", "1e7")))
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
v = torch.empty(nq, dtype=torch.float64, device="cuda"); g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
for fast in (False, True):
    for _ in range(3): gb.valgrad_(v, g, G, X, fast=fast)
    torch.cuda.synchronize(); ts = []
    for _ in range(11):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gb.valgrad_(v, g, G, X, fast=fast); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    gb.profile_enable(True); acc = {}
    for _ in range(5):
        gb.valgrad_(v, g, G, X, fast=fast)
        for k, x in gb.profile_read().items(): acc[k] = acc.get(k, 0) + x / 5
    gb.profile_enable(False)
    print(os.environ.get("TAG", ""), "FAST" if fast else "EXACT", "median %.4f ms min %.4f" % (sorted(ts)[5], min(ts)), " ".join("%s %.3f" % (k[:7], x) for k, x in acc.items()), flush=True)
    print("   checksum", float(v.sum()), float(g.sum()))
