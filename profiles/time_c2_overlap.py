"""C2 step split into K sub-batches issued round-robin on S streams: does the (atomics / latency bound) sort of one
sub-batch overlap the (FP64-pipe bound) brick evaluation of another?  Prints the median step time per (K, S)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb

n, nq = 256, 10_000_000
A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
X = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
v = torch.empty(nq, dtype=torch.float64, device="cuda"); g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
main = torch.cuda.current_stream()

def step(K, S, fast):
    if K == 1:
        gb.valgrad_(v, g, G, X, fast=fast); return
    per = (nq + K - 1) // K
    fork = torch.cuda.Event(); fork.record(main)
    for i in range(K):
        st = streams[i % S]
        if i < S: st.wait_event(fork)
        a, b = i * per, min(nq, (i + 1) * per)
        with torch.cuda.stream(st):
            gb.valgrad_(v[a:b], g[a:b], G, X[a:b], fast=fast)
    for st in streams[:S]:
        e = torch.cuda.Event(); e.record(st); main.wait_event(e)

for fast in (False, True):
    for K, S in ((1, 1), (2, 1), (2, 2), (4, 1), (4, 2), (4, 4), (8, 2), (8, 4)):
        for _ in range(3): step(K, S, fast)
        torch.cuda.synchronize(); ts = []
        for _ in range(11):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(K, S, fast); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(os.environ.get("TAG", ""), "FAST" if fast else "EXACT", "K=%d S=%d median %.4f ms min %.4f" % (K, S, sorted(ts)[5], min(ts)), "checksum", float(v.sum()), float(g.sum()), flush=True)
