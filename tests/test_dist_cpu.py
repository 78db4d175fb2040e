"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU host logic: query sharding, the slab
plan / halo exchange of restrict-prolong, and that assembling the slabs reproduces the unsharded
result bit for bit.  The local compute engine here is the CPU oracle (test infrastructure); on the
GPU box the same code path runs with the libgridb200 engine (tests/test_gpu_dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """Same interface as grid_jl_b200.dist.GpuEngine, computing with the CPU oracle / numpy."""

    def __init__(self):
        from oracle import oracle as O

        self.O = O

    @staticmethod
    def _np(t):
        n = t.dim()
        return np.asfortranarray(t.permute(*range(n - 1, -1, -1)).contiguous().numpy().T)

    @staticmethod
    def _t(a):
        import torch

        a = np.asarray(a)
        return torch.from_numpy(np.ascontiguousarray(a.T)).permute(*range(a.ndim - 1, -1, -1))

    def restrict_dims12(self, a, scale):
        return self._t(self.O.restrict_flags(self._np(a), [True, True, False], scale))

    def prolong_dims12(self, a, m1, m2):
        A = self._np(a)
        return self._t(self.O.prolong_sizes(A, [m1, m2, A.shape[2]]))

    # the fused slab operators of GpuEngine, restated as dims 1,2 on the slab followed by the dim-3 slab operator
    fused = True

    def restrict3_slab(self, a, scale, L3, in_first, out_first, count):
        return self.restrict_slab(self.restrict_dims12(a, scale), scale, L3, in_first, out_first, count)

    def prolong3_slab(self, a, target, n3, in_first, out_first, count):
        m1, m2, m3 = (int(v) for v in target)
        return self.prolong_slab(self.prolong_dims12(a, m1, m2), n3, m3, in_first, out_first, count)

    def restrict_slab(self, a, scale, L, in_first, out_first, count):
        A = self._np(a)
        T = A.dtype.type
        n = (L + 1) // 2 if L % 2 else L // 2 + 1
        s1, s2, s75, s25 = T(scale), T(scale / 2), T(0.75 * scale), T(0.25 * scale)
        ld = lambda k: A[:, :, k - in_first]
        out = np.zeros(A.shape[:2] + (count,), dtype=A.dtype, order="F")
        for m in range(count):
            j = out_first + m
            acc = np.zeros(A.shape[:2], dtype=A.dtype)
            if L % 2:
                acc = acc + s1 * ld(2 * j)
                if j <= n - 2: acc = acc + s2 * ld(2 * j + 1)
                if j >= 1: acc = acc + s2 * ld(2 * j - 1)
            else:
                if j <= n - 2:
                    acc = acc + s75 * ld(2 * j)
                    acc = acc + s25 * ld(2 * j + 1)
                if j >= 1:
                    acc = acc + s75 * ld(2 * j - 1)
                    acc = acc + s25 * ld(2 * j - 2)
            out[:, :, m] = acc
        return self._t(out)

    def prolong_slab(self, a, n, length, in_first, out_first, count):
        A = self._np(a)
        T = A.dtype.type
        ld = lambda i: A[:, :, i - in_first]
        out = np.zeros(A.shape[:2] + (count,), dtype=A.dtype, order="F")
        for m in range(count):
            k = out_first + m
            i = k >> 1
            if length % 2:
                if k % 2 == 0:
                    out[:, :, m] = ld(i)
                    continue
                acc = np.zeros(A.shape[:2], dtype=A.dtype) + T(0.5) * ld(i)
                acc = acc + T(0.5) * ld(i + 1)
            elif k % 2 == 0:
                acc = np.zeros(A.shape[:2], dtype=A.dtype) + T(0.75) * ld(i)
                acc = acc + T(0.25) * ld(i + 1)
            else:
                acc = np.zeros(A.shape[:2], dtype=A.dtype) + T(0.25) * ld(i)
                acc = acc + T(0.75) * ld(i + 1)
            out[:, :, m] = acc
        return self._t(out)


def _worker(rank, world, port, shape, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import grid_jl_b200.dist as D  # noqa: F401  (package import through the shim)
    except Exception:
        import grid_jl_b200  # noqa: F401
        import grid_jl_b200.dist as D
    from oracle import oracle as O

    eng = OracleEngine()
    A = O.synth(np.float64, int(np.prod(shape)), seed=77).reshape(shape, order="F")
    L3 = shape[2]
    f0, f1 = D.shard_range(L3, rank, world)
    local = eng._t(A[:, :, f0:f1])
    n3 = D.restrict_size(L3)
    k0, k1 = D.shard_range(n3, rank, world)
    ref = O.restrict_flags(A, [True, True, True])
    refp = O.prolong_sizes(ref, list(shape))
    ok_r = ok_p = True
    for fused in (True, False):  # fused slab operators (halo of raw planes) and the per-dimension route
        eng.fused = fused
        # restrict
        mine = D.slab_restrict3(local, L3, rank, world, 1.0, engine=eng)
        ok_r = ok_r and np.array_equal(eng._np(mine), ref[:, :, k0:k1])
        # prolong the (reference) coarse array back to the fine size, sharded
        coarse_local = eng._t(ref[:, :, k0:k1])
        up = D.slab_prolong3(coarse_local, n3, shape, rank, world, engine=eng)
        ok_p = ok_p and np.array_equal(eng._np(up), refp[:, :, f0:f1])
    # query sharding: a shard of the batch evaluated alone equals that slice of the full batch
    co = O.make_coefs(A[:, :, :12], "nan", "cubic")
    X = 1.0 + O.synth(np.float64, 3 * 1000, seed=5).reshape(1000, 3) * (np.array([shape[0], shape[1], 12]) - 1.0)
    q0, q1 = D.shard_range(1000, rank, world)
    v, g, _, _ = O.evaluate(co, "nan", "cubic", X[q0:q1], 3)
    vf, gf, _, _ = O.evaluate(co, "nan", "cubic", X, 3)
    ok_q = np.array_equal(v, vf[q0:q1]) and np.array_equal(g, gf[q0:q1])
    out = [None] * world
    dist.all_gather_object(out, (ok_r, ok_p, ok_q))
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("shape", [(9, 8, 20), (8, 9, 21), (6, 7, 13)])
def test_slab_multigrid_and_query_sharding_gloo(world, shape):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(t) for t in res), res


def test_shard_range_and_plan_cover():
    import grid_jl_b200  # noqa: F401
    import grid_jl_b200.dist as D

    for n in (0, 1, 7, 16, 1024, 1025):
        for w in (1, 2, 3, 4, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    for L in (5, 8, 16, 17, 64, 65, 1024, 513):
        for w in (1, 2, 4, 8):
            n = D.restrict_size(L)
            for kind, n_in, n_out in (("restrict", L, n), ("prolong", n, L)):
                plan = D.SlabPlan(kind, n_in, n_out, w)
                for r in range(w):
                    lo, hi = plan.need[r]
                    have = set(range(*plan.own[r]))
                    for (src, dst, first, count) in plan.transfers:
                        if dst == r:
                            got = set(range(first, first + count))
                            assert got <= set(range(*plan.own[src])) and not (got & have)
                            have |= got
                    assert set(range(lo, hi + 1)) <= have, (kind, L, w, r)
                    e0, e1 = plan.ext_range(r)
                    assert e0 <= min(lo, plan.own[r][0]) and e1 >= max(hi + 1, plan.own[r][1]) or hi < lo
            # halo is at most 2 planes per side for restrict at healthy slab thickness
            if L // w >= 8:
                plan = D.SlabPlan("restrict", L, n, w)
                assert all(c <= 3 for (_, _, _, c) in plan.transfers)


def test_c_plan_matches_python_plan():
    """gb_slab_plan / gb_shard_range (host-only C ABI, csrc/comm.cu) against the Python plan the gloo tests exercise."""
    import grid_jl_b200  # noqa: F401
    import grid_jl_b200.dist as D

    for L in (5, 8, 16, 17, 64, 65, 1024, 513, 33):
        for w in (1, 2, 3, 4, 8, 16):
            n = D.restrict_size(L)
            for kind, n_in, n_out in (("restrict", L, n), ("prolong", n, L)):
                plan = D.SlabPlan(kind, n_in, n_out, w)
                for r in range(w):
                    own, out, need, ext, tr = D.c_slab_plan(kind, n_in, n_out, r, w)
                    assert own == plan.own[r] and out == plan.out[r] and ext == plan.ext_range(r)
                    if plan.need[r][1] >= plan.need[r][0]:
                        assert need == plan.need[r]
                    assert sorted(tr) == sorted(plan.transfers)
