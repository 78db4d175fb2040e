"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): NCCL broadcast of the coefficients,
query sharding and the slab-sharded restrict/prolong with a real halo exchange.  Launched in-process
with torch.multiprocessing, one process per GPU."""
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
