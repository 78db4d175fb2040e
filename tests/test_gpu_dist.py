"""Multi-GPU tests with real NCCL traffic (need >= 2 CUDA devices; skipped otherwise -- the same code paths run on one
GPU with virtual ranks in test_gpu_round2.py::test_comm_loopback_*):
  * one process per GPU (torch.multiprocessing): gb_comm_init_rank through torch.distributed, gb_grid_replicate
    (ncclBroadcast of the coefficients), query sharding, gb_restrict3_sharded / gb_prolong3_sharded (grouped
    ncclSend/ncclRecv halo exchange overlapped with the interior planes);
  * one process driving every GPU (gb_comm_init_all): the same through the single-process entry points, plus
    gb_eval_sharded on host buffers."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import grid_jl_b200 as gb
    import grid_jl_b200.dist as D
    from oracle import oracle as O

    shape = (40, 36, 50)
    A = O.synth(np.float64, int(np.prod(shape)), seed=3).reshape(shape, order="F")
    # replicated grid: prefilter on rank 0, one broadcast
    G = gb.InterpGrid(gb.jl_from_numpy(A), gb.BCnan, gb.InterpCubic) if rank == 0 else None
    G = D.broadcast_grid(G, gb.BCnan, gb.InterpCubic, src=0)
    co = O.make_coefs(A, "nan", "cubic")
    ok = [bool(np.array_equal(G.coefs, co))]
    X = 1.0 + O.synth(np.float64, 3 * 20000, seed=9).reshape(20000, 3) * (np.array(shape) - 1.0)
    q0, q1 = D.shard_range(20000, rank, world)
    v, g = gb.valgrad_batch(G, torch.from_numpy(X[q0:q1]).cuda())
    ov, og, _, _ = O.evaluate(co, "nan", "cubic", X, 3)
    ok.append(bool(np.array_equal(v.cpu().numpy(), ov[q0:q1]) and np.array_equal(g.cpu().numpy(), og[q0:q1])))
    # slab-sharded restrict / prolong with NCCL halo exchange
    L3 = shape[2]
    f0, f1 = D.shard_range(L3, rank, world)
    mine = D.slab_restrict3(gb.jl_from_numpy(np.asfortranarray(A[:, :, f0:f1])), L3, rank, world)
    ref = O.restrict_flags(A, [True, True, True])
    n3 = D.restrict_size(L3)
    k0, k1 = D.shard_range(n3, rank, world)
    ok.append(bool(np.array_equal(gb.jl_to_numpy(mine), ref[:, :, k0:k1])))
    up = D.slab_prolong3(mine, n3, shape, rank, world)
    refp = O.prolong_sizes(ref, list(shape))
    ok.append(bool(np.array_equal(gb.jl_to_numpy(up), refp[:, :, f0:f1])))
    out = [None] * world
    dist.all_gather_object(out, ok)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_multi_gpu_broadcast_shard_and_halo():
    import torch

    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = min(world, 4)
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(t) for t in res), res


def test_single_process_all_devices():
    """gb_comm_init_all: what a Julia host uses -- one process, every GPU of the box."""
    import ctypes as C

    import torch

    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    ndev = min(ndev, 8)
    sys.path.insert(0, ROOT)
    import grid_jl_b200 as gb
    import grid_jl_b200.dist as D
    from oracle import oracle as O

    shape = (48, 40, 70)
    A = O.synth(np.float64, int(np.prod(shape)), seed=4).reshape(shape, order="F")
    ref = O.restrict_flags(A, [True, True, True])
    refp = O.prolong_sizes(ref, list(shape))
    comm = D.Comm.all_devices(ndev)
    try:
        L3, n3 = shape[2], D.restrict_size(shape[2])
        fine = [gb.jl_from_numpy(np.asfortranarray(A[:, :, slice(*D.shard_range(L3, r, ndev))]), device="cuda:%d" % r) for r in range(ndev)]
        # both transports: peer pointers (boundary kernels read the neighbours' slabs over NVLink) and the NCCL exchange
        capable = comm.peer_halo(True)
        for peer in ([True, False] if capable else [False]):
            assert comm.peer_halo(peer) == peer
            coarse = comm.restrict3(fine, shape, 1.0, use_streams=False)
            for r in range(ndev):
                k0, k1 = D.shard_range(n3, r, ndev)
                assert np.array_equal(gb.jl_to_numpy(coarse[r]), ref[:, :, k0:k1])
            if min(L3, n3) >= 4 * ndev:
                assert comm.last_path() == ("peer" if peer else "exchange")
            up = comm.prolong3(coarse, ref.shape, shape, use_streams=False)
            for r in range(ndev):
                f0, f1 = D.shard_range(L3, r, ndev)
                assert np.array_equal(gb.jl_to_numpy(up[r]), refp[:, :, f0:f1])
            assert sum(s["halo_bytes"] for s in comm.stats()) > 0
            # two levels in one call each way (intermediate level in the communicator's scratch arrays)
            ref2 = O.restrict_flags(ref, [True, True, True])
            low = comm.restrict3_levels(fine, shape, 2, 1.0, use_streams=False)
            for r in range(ndev):
                k0, k1 = D.shard_range(ref2.shape[2], r, ndev)
                assert np.array_equal(gb.jl_to_numpy(low[r]), ref2[:, :, k0:k1])
            back = comm.prolong3_levels(low, ref2.shape, [ref.shape, shape], use_streams=False)
            refb = O.prolong_sizes(O.prolong_sizes(ref2, list(ref.shape)), list(shape))
            for r in range(ndev):
                f0, f1 = D.shard_range(L3, r, ndev)
                assert np.array_equal(gb.jl_to_numpy(back[r]), refb[:, :, f0:f1])
        # replicate + host batch split over the devices
        co = O.make_coefs(A, "nan", "cubic")
        with torch.cuda.device(0):
            g0 = gb.InterpGrid(gb.jl_from_numpy(A, device="cuda:0"), gb.BCnan, gb.InterpCubic)
        grids = comm.replicate([g0] + [None] * (ndev - 1), 0, (g0.stored_shape, np.float64, gb.BCnan, gb.InterpCubic, float("nan")))
        assert all(np.array_equal(g.coefs, co) for g in grids)
        X = 1.0 + O.synth(np.float64, 3 * 300_000, seed=6).reshape(300_000, 3) * (np.array(shape) - 1.0)
        ov, og, _, _ = O.evaluate(co, "nan", "cubic", X, 3)
        v, gr = np.empty(len(X)), np.empty((len(X), 3))
        L = gb._lib
        hs = (C.c_void_p * ndev)(*[g._h.value for g in grids])
        L.check(L.lib().gb_eval_sharded(comm._h, hs, 3, len(X), X.ctypes.data, L.GB_F64, None, v.ctypes.data, gr.ctypes.data, None, 0, None))
        assert np.array_equal(v, ov) and np.array_equal(gr, og)
        # a batch large enough for every device's share to take the staged pageable path (>= 2^19 queries each), all at once
        nbig = ndev * ((1 << 19) + 4097)
        Xb = np.tile(X, (nbig // len(X) + 1, 1))[:nbig].copy()
        vb, gb_ = np.empty(nbig), np.empty((nbig, 3))
        L.check(L.lib().gb_eval_sharded(comm._h, hs, 3, nbig, Xb.ctypes.data, L.GB_F64, None, vb.ctypes.data, gb_.ctypes.data, None, 0, None))
        idx = np.arange(nbig) % len(X)
        assert np.array_equal(vb, ov[idx]) and np.array_equal(gb_, og[idx])
    finally:
        comm.close()
