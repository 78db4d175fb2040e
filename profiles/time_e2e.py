"""Host-buffer (pinned) valgrad of the C2 batch through the C-ABI: ms per call for tiled=None/True/False."""
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
for tiled in (None, True, False):
    for _ in range(2):
        gb.valgrad_(vn, gn, G, Xn, tiled=tiled)
    t0 = time.perf_counter()
    for _ in range(5):
        gb.valgrad_(vn, gn, G, Xn, tiled=tiled)
    print("chunk=%s tiled=%s: %.2f ms" % (os.environ.get("GB_EVAL_CHUNK", "-"), tiled, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
# pure copies for reference
s = torch.cuda.Stream()
Xd = torch.empty_like(X); vd = torch.empty(nq, dtype=torch.float64, device="cuda"); gd = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    Xd.copy_(Xh, non_blocking=True); torch.cuda.synchronize()
print("H2D 240 MB alone: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for _ in range(5):
    vh.copy_(vd, non_blocking=True); gh.copy_(gd, non_blocking=True); torch.cuda.synchronize()
print("D2H 320 MB alone: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
# both directions at once on two streams: the floor of any pipelined host-buffer call
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        Xd.copy_(Xh, non_blocking=True)
    with torch.cuda.stream(s2):
        vh.copy_(vd, non_blocking=True); gh.copy_(gd, non_blocking=True)
    torch.cuda.synchronize()
print("H2D 240 MB + D2H 320 MB concurrently: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
