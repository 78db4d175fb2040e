"""Timeline of one host-buffer C2 call (pinned buffers): run with GB_EVAL_TRACE=1 [GB_EVAL_CHUNK=...]; the library prints
per-chunk copy-in / evaluation / copy-out times on the slot streams to stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
Xh = torch.empty((nq, 3), dtype=torch.float64).pin_memory(); Xh.copy_(X)
vh = torch.empty(nq, dtype=torch.float64).pin_memory(); gh = torch.empty((nq, 3), dtype=torch.float64).pin_memory()
Xn, vn, gn = Xh.numpy(), vh.numpy(), gh.numpy()
for i in range(3):
    t0 = time.perf_counter()
    gb.valgrad_(vn, gn, G, Xn)
    print("call %d: %.2f ms (wall)" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
