"""C5 chain (1024^3 f64, restrict 1024 -> 17 and prolong back), slab-sharded over all GPUs of the box from ONE process
(gb_comm_init_all -- what a Julia host uses): peer mode (boundary kernels read the neighbours' planes over NVLink, no
exchange) vs the NCCL halo exchange.  Wall time per chain with every device synchronised, median of 9; plus the per-level
split of rank 0 (gb_comm_stats).  usage: python profiles/time_c5_single_process.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
import grid_jl_b200.dist as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ndev = torch.cuda.device_count()
comm = D.Comm.all_devices(ndev)
fine = []
for r in range(ndev):
    f0, f1 = D.shard_range(n, r, ndev)
    with torch.cuda.device(r):
        a = gb.jl_empty((n, n, f1 - f0), torch.float64, "cuda:%d" % r)
        gb.synth_uniform(a, 7, first=f0 * n * n)
    fine.append(a)
sizes = [n]
while sizes[-1] > 17:
    sizes.append(D.restrict_size(sizes[-1]))

def sync():
    for r in range(ndev):
        torch.cuda.synchronize(r)

def chain(levels_stats=None):
    cur, pyr = fine, []
    for L in sizes[:-1]:
        pyr.append(cur)
        cur = comm.restrict3(cur, (L, L, L), 1.0)
        if levels_stats is not None:
            sync(); levels_stats.append(("restrict %d" % L, comm.last_path(), comm.stats()[0]))
    sync(); t1 = time.perf_counter()
    for L, Lf in zip(reversed(sizes[1:]), reversed(sizes[:-1])):
        cur = comm.prolong3(cur, (L, L, L), (Lf, Lf, Lf))
        if levels_stats is not None:
            sync(); levels_stats.append(("prolong %d" % L, comm.last_path(), comm.stats()[0]))
    sync()
    return t1

def snapshot():
    """every level of one chain (restrict down, prolong up), as CPU tensors"""
    cur, out = fine, []
    for L in sizes[:-1]:
        cur = comm.restrict3(cur, (L, L, L), 1.0)
        if L <= 257: out += [c.cpu() for c in cur]
    for L, Lf in zip(reversed(sizes[1:]), reversed(sizes[:-1])):
        cur = comm.prolong3(cur, (L, L, L), (Lf, Lf, Lf))
        if Lf <= 257: out += [c.cpu() for c in cur]
    out += [c[:, :, :2].cpu() for c in cur] + [c[:, :, -2:].cpu() for c in cur]  # the finest level: the planes next to the slab faces
    return out

def chain_one_call():
    """the same chains as ONE library call each (gb_restrict3_sharded_levels / gb_prolong3_sharded_levels)"""
    low = comm.restrict3_levels(fine, (n, n, n), len(sizes) - 1, 1.0)
    sync(); t1 = time.perf_counter()
    up = comm.prolong3_levels(low, (sizes[-1],) * 3, [(L,) * 3 for L in reversed(sizes[:-1])])
    sync()
    return t1, up

print("devices", ndev, "peer capable", comm.peer_halo(True))
if comm.peer_halo(True):
    a = snapshot(); assert comm.last_path() == "peer"
    comm.peer_halo(False); b = snapshot(); assert comm.last_path() == "exchange"
    assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
    print("peer mode == exchange mode bitwise on", len(a), "arrays")
    comm.peer_halo(True)
    cur = fine
    for L in sizes[:-1]: cur = comm.restrict3(cur, (L, L, L), 1.0)
    low = comm.restrict3_levels(fine, (n, n, n), len(sizes) - 1, 1.0)
    assert all(torch.equal(x.cpu(), y.cpu()) for x, y in zip(cur, low))
    for L, Lf in zip(reversed(sizes[1:]), reversed(sizes[:-1])): cur = comm.prolong3(cur, (L, L, L), (Lf, Lf, Lf))
    _, up = chain_one_call()
    assert all(torch.equal(x[:, :, ::37].cpu(), y[:, :, ::37].cpu()) for x, y in zip(cur, up))
    print("one call per chain == one call per level bitwise")
for peer in ([True, False] if comm.peer_halo(True) else [False]):
    comm.peer_halo(peer)
    for _ in range(3): chain()
    tr, tp = [], []
    for _ in range(9):
        sync(); t0 = time.perf_counter(); t1 = chain(); t2 = time.perf_counter()
        tr.append((t1 - t0) * 1e3); tp.append((t2 - t1) * 1e3)
    tr1, tp1 = [], []
    for _ in range(3): chain_one_call()
    for _ in range(9):
        sync(); t0 = time.perf_counter(); t1, up = chain_one_call(); t2 = time.perf_counter()
        tr1.append((t1 - t0) * 1e3); tp1.append((t2 - t1) * 1e3)
    print("%-8s ONE CALL per chain: restrict %.3f ms  prolong %.3f ms" % ("peer" if peer else "exchange", sorted(tr1)[4], sorted(tp1)[4]))
    st = []
    chain(st)
    print("%-8s restrict chain %.3f ms  prolong chain %.3f ms (median of 9, wall, all devices synchronised)" % ("peer" if peer else "exchange", sorted(tr)[4], sorted(tp)[4]))
    for name, path, s in st:
        print("   %-14s %-8s total %.3f ms interior %.3f halo wait %.3f halo bytes %d" % (name, path, s["total_ms"], s["interior_ms"], s["halo_wait_ms"], s["halo_bytes"]))
comm.close()
