"""C2 batch (1e7 queries, valgrad) through gb_eval with HOST buffers: pinned vs pageable (numpy) memory.
Run once per GB_HOST_REGISTER setting (the library reads it once): unset (auto), 0 (never), 1 (always)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grid_jl_b200 as gb

n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
def run(Xn, vn, gn, k=6):
    for _ in range(2): gb.valgrad_(vn, gn, G, Xn)
    t0 = time.perf_counter()
    for _ in range(k): gb.valgrad_(vn, gn, G, Xn)
    return (time.perf_counter() - t0) / k * 1e3
Xh = torch.empty((nq, 3), dtype=torch.float64).pin_memory(); Xh.copy_(X)
vh = torch.empty(nq, dtype=torch.float64).pin_memory(); gh = torch.empty((nq, 3), dtype=torch.float64).pin_memory()
tag = "GB_HOST_REGISTER=%s" % os.environ.get("GB_HOST_REGISTER", "auto")
print(tag, "pinned   %.2f ms" % run(Xh.numpy(), vh.numpy(), gh.numpy()), flush=True)
Xp, vp, gp = np.array(Xh.numpy(), copy=True), np.empty(nq), np.empty((nq, 3))
print(tag, "pageable %.2f ms" % run(Xp, vp, gp), flush=True)
assert np.array_equal(vp, vh.numpy())
# smaller batches: where does registering stop paying?
for m in (100_000, 1_000_000):
    print(tag, "pageable, %d queries: %.3f ms" % (m, run(Xp[:m], vp[:m], gp[:m], 10)), flush=True)
