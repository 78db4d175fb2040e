"""Raw pinned-memory copies of the C2 step's traffic (H2D 240 MB, D2H 320 MB) on two streams at once: whole buffers vs chunks."""
import sys, time, torch
nq = 10_000_000
Xh = torch.empty((nq, 3), dtype=torch.float64).pin_memory(); Xh.zero_()
vh = torch.empty(nq, dtype=torch.float64).pin_memory(); gh = torch.empty((nq, 3), dtype=torch.float64).pin_memory(); vh.zero_(); gh.zero_()
Xd = torch.empty_like(Xh, device="cuda"); vd = torch.zeros(nq, dtype=torch.float64, device="cuda"); gd = torch.zeros((nq, 3), dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(nchunk, h2d=True, d2h=True):
    per = (nq + nchunk - 1) // nchunk
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        for c in range(nchunk):
            a, b = c * per, min(nq, (c + 1) * per)
            if h2d:
                with torch.cuda.stream(s1): Xd[a:b].copy_(Xh[a:b], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): vh[a:b].copy_(vd[a:b], non_blocking=True); gh[a:b].copy_(gd[a:b], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 5 * 1e3
for nchunk in (1, 4, 10, 20, 40):
    run(nchunk)
    print("chunks %2d: H2D alone %.2f ms, D2H alone %.2f ms, both at once %.2f ms" % (nchunk, run(nchunk, True, False), run(nchunk, False, True), run(nchunk)), flush=True)
