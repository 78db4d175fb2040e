"""Multi-GPU host layer.  Nothing here is on the single-GPU hot path.

The product path is the C ABI (include/gridb200.h, csrc/comm.cu): a `Comm` wraps a gb_comm -- one
process per GPU (`Comm.from_torch()`: the 128-byte NCCL id travels through torch.distributed), one
process driving several GPUs (`Comm.all_devices(n)`, what a Julia host uses) or `Comm.loopback(n)`
(virtual ranks on one GPU, tests) -- and `slab_restrict3` / `slab_prolong3` / `broadcast_grid` below are
thin callers of gb_restrict3_sharded / gb_prolong3_sharded / gb_grid_replicate.

* Point evaluation shards trivially: contiguous query ranges per rank, the coefficient array is
  prefiltered once and replicated with ONE broadcast; no per-call collective.
* restrict / prolong of a 3-D array shard by slabs along the last (slowest) dimension.  One exchange
  step per level: the interior output planes are computed straight from the rank's own slab while the
  few boundary planes ("halo") travel on the communication stream; the boundary output planes follow.
  The per-element arithmetic is unchanged, so the gathered result equals the single-device one bit for bit.

The same plan (who owns / needs / sends which planes) exists here in Python (`SlabPlan`) for the CPU
tests, which run the exchange over gloo with a caller-supplied compute engine, and for degenerate
extents / partial prolongations, which take the per-dimension slab kernels (gb_restrict_slab /
gb_prolong_slab) with a torch.distributed halo exchange."""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced split of range(n): rank r owns [lo, hi)  (== gb_shard_range)."""
    n, rank, world = int(n), int(rank), int(world)
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def restrict_size(n):
    n = int(n)
    return (n + 1) // 2 if n % 2 else n // 2 + 1


def restrict_needs(L, k0, k1):
    """Inclusive range of fine planes that coarse planes [k0,k1) of restrict(len L) read
    (src/restrict_prolong.jl:117-136)."""
    n = restrict_size(L)
    lo = 0 if k0 == 0 else (2 * k0 - 1 if L % 2 else 2 * k0 - 2)
    hi = 2 * (k1 - 1) + 1 if k1 - 1 <= n - 2 else L - 1
    return max(lo, 0), min(hi, L - 1)


def prolong_needs(n, length, k0, k1):
    """Inclusive range of coarse planes that fine planes [k0,k1) of prolong(n -> length) read
    (src/restrict_prolong.jl:26-44)."""
    kl = k1 - 1
    copy_only = (length % 2 == 1) and (kl % 2 == 0)
    return k0 >> 1, min(n - 1, (kl >> 1) + (0 if copy_only else 1))


class SlabPlan:
    """Who sends which planes to whom for one level along the sharded dimension (== gb_slab_plan).

    own[r]  = planes of the INPUT array rank r holds,
    need[r] = inclusive input range rank r's OUTPUT planes read,
    out[r]  = output planes rank r produces.
    transfers: list of (src, dst, first_plane, count) covering need[r] minus own[r]."""

    def __init__(self, kind, n_in, n_out, world):
        self.kind, self.n_in, self.n_out, self.world = kind, int(n_in), int(n_out), int(world)
        self.own = [shard_range(n_in, r, world) for r in range(world)]
        self.out = [shard_range(n_out, r, world) for r in range(world)]
        self.need = []
        for r in range(world):
            k0, k1 = self.out[r]
            if k1 <= k0:
                self.need.append((0, -1))
            elif kind == "restrict":
                self.need.append(restrict_needs(n_in, k0, k1))
            else:
                self.need.append(prolong_needs(n_in, n_out, k0, k1))
        self.transfers = []
        for dst in range(world):
            lo, hi = self.need[dst]
            for src in range(world):
                if src == dst:
                    continue
                a = max(lo, self.own[src][0])
                b = min(hi + 1, self.own[src][1])
                # only planes outside dst's own range travel
                for s0, s1 in ((a, min(b, self.own[dst][0])), (max(a, self.own[dst][1]), b)):
                    if s1 > s0:
                        self.transfers.append((src, dst, s0, s1 - s0))

    def ext_range(self, rank):
        """Planes held after the exchange: own range widened to cover need (always contiguous)."""
        lo, hi = self.need[rank]
        o0, o1 = self.own[rank]
        if hi < lo:
            return o0, o1
        return min(lo, o0), max(hi + 1, o1)


def c_slab_plan(kind, n_in, n_out, rank, world):
    """The same plan as computed by the library (gb_slab_plan): (own, out, need, ext, transfers)."""
    from . import _lib as L

    own, out, need, ext = ((C.c_int64 * 2)() for _ in range(4))
    nt = C.c_int(0)
    k = 0 if kind == "restrict" else 1
    L.check(L.lib().gb_slab_plan(k, int(n_in), int(n_out), int(rank), int(world), own, out, need, ext, None, 0, C.byref(nt)))
    tr = (C.c_int64 * (4 * max(1, nt.value)))()
    L.check(L.lib().gb_slab_plan(k, int(n_in), int(n_out), int(rank), int(world), own, out, need, ext, tr, nt.value, C.byref(nt)))
    return tuple(own), tuple(out), tuple(need), tuple(ext), [tuple(tr[4 * i:4 * i + 4]) for i in range(nt.value)]


def _flat(t):
    """Contiguous 1-D view of a column-major (Julia layout) tensor's storage."""
    n = t.dim()
    return t.permute(*range(n - 1, -1, -1)).reshape(-1)


def jl_empty_like(t, shape):
    import torch

    shape = tuple(int(s) for s in shape)
    base = torch.empty(shape[::-1], dtype=t.dtype, device=t.device)
    return base.permute(*range(len(shape) - 1, -1, -1))


# ---- the C-ABI communicator -----------------------------------------------------------------------------------------------
class Comm:
    """gb_comm: `nranks` ranks, `nlocal` of them driven by this process (rank order, starting at `first_rank`)."""

    def __init__(self, handle):
        from . import _lib as L

        self._L, self._h = L, handle
        nr, nl, fr = C.c_int(), C.c_int(), C.c_int()
        L.check(L.lib().gb_comm_size(handle, C.byref(nr), C.byref(nl), C.byref(fr)))
        self.nranks, self.nlocal, self.first_rank = nr.value, nl.value, fr.value

    @classmethod
    def from_torch(cls, group=None):
        """One process per GPU: rank 0 creates the NCCL id, torch.distributed ships it; the current CUDA device is used."""
        import torch.distributed as dist

        from . import _lib as L

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [None]
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            L.check(L.lib().gb_comm_unique_id(buf, 128))
            box[0] = bytes(buf)
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        h = C.c_void_p()
        L.check(L.lib().gb_comm_init_rank(C.byref(h), world, rank, ident, 128))
        return cls(h)

    @classmethod
    def all_devices(cls, ndev, dev_ids=None):
        """One process, `ndev` GPUs (ncclCommInitAll)."""
        from . import _lib as L

        h = C.c_void_p()
        ids = None if dev_ids is None else (C.c_int * ndev)(*[int(d) for d in dev_ids])
        L.check(L.lib().gb_comm_init_all(C.byref(h), int(ndev), ids))
        return cls(h)

    @classmethod
    def loopback(cls, nranks):
        """`nranks` virtual ranks on the current device (no NCCL): the sharded code path on a one-GPU box."""
        from . import _lib as L

        h = C.c_void_p()
        L.check(L.lib().gb_comm_init_loopback(C.byref(h), int(nranks)))
        return cls(h)

    def close(self):
        if self._h is not None and self._h.value:
            self._L.lib().gb_comm_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def peer_halo(self, enable=None):
        """Peer-pointer halo reads (boundary kernels read the neighbours' slabs in place; single-process communicators with
        peer access): enable True / False switches them, None only queries.  Returns the setting in force."""
        on = C.c_int(0)
        self._L.check(self._L.lib().gb_comm_peer_halo(self._h, -1 if enable is None else int(bool(enable)), C.byref(on)))
        return bool(on.value)

    def last_path(self):
        """'peer' if the last sharded restrict / prolong read its halo planes through peer pointers, else 'exchange'."""
        v = C.c_int(0)
        self._L.check(self._L.lib().gb_comm_last_path(self._h, C.byref(v)))
        return "peer" if v.value else "exchange"

    def stats(self):
        """Last sharded restrict / prolong, per local rank: dict(total_ms, interior_ms, halo_wait_ms, halo_bytes)."""
        out = (C.c_double * (4 * self.nlocal))()
        n = C.c_int(0)
        self._L.check(self._L.lib().gb_comm_stats(self._h, out, 4 * self.nlocal, C.byref(n)))
        return [dict(total_ms=out[4 * i], interior_ms=out[4 * i + 1], halo_wait_ms=out[4 * i + 2], halo_bytes=out[4 * i + 3]) for i in range(n.value // 4)]

    # -- sharded multigrid levels on lists of per-local-rank slabs (column-major CUDA tensors) --
    def _level(self, kind, slabs, global_in, global_out, scale, use_streams):
        import torch

        from .api import _Arr, _i64s, jl_empty

        L = self._L
        assert len(slabs) == self.nlocal
        outs, ip, op, sp = [], [], [], []
        for i, a in enumerate(slabs):
            r = self.first_rank + i
            f0, f1 = shard_range(global_in[2], r, self.nranks)
            k0, k1 = shard_range(global_out[2], r, self.nranks)
            assert tuple(a.shape) == (global_in[0], global_in[1], f1 - f0), "slab %d has the wrong shape" % r
            arr = _Arr(a, name="slab") if f1 > f0 else None
            with torch.cuda.device(a.device):
                o = jl_empty((global_out[0], global_out[1], k1 - k0), a.dtype, a.device)
            outs.append(o)
            ip.append(arr.ptr if arr is not None else None)
            op.append(o.data_ptr() if k1 > k0 else None)
            sp.append(torch.cuda.current_stream(a.device).cuda_stream)
        dt = L.GB_F32 if slabs[0].dtype == torch.float32 else L.GB_F64
        ips, ops = (C.c_void_p * self.nlocal)(*ip), (C.c_void_p * self.nlocal)(*op)
        sps = (C.c_void_p * self.nlocal)(*sp) if use_streams else None
        if kind == "restrict":
            L.check(L.lib().gb_restrict3_sharded(self._h, dt, _i64s(global_in), ips, float(scale), ops, sps))
        else:
            L.check(L.lib().gb_prolong3_sharded(self._h, dt, _i64s(global_in), ips, _i64s(global_out), ops, sps))
        return outs

    def restrict3(self, slabs, global_dims, scale=1.0, use_streams=True):
        gd = tuple(int(v) for v in global_dims)
        return self._level("restrict", slabs, gd, tuple(restrict_size(v) for v in gd), scale, use_streams)

    def prolong3(self, slabs, global_dims, out_dims, use_streams=True):
        return self._level("prolong", slabs, tuple(int(v) for v in global_dims), tuple(int(v) for v in out_dims), 1.0, use_streams)

    def restrict3_levels(self, slabs, global_dims, nlevels, scale=1.0, use_streams=True):
        """restrict(A, flags) with `nlevels` all-true columns in ONE library call (intermediate levels stay in the
        communicator's scratch arrays): returns the slabs of the last level."""
        import torch

        from .api import _Arr, _i64s, jl_empty

        L = self._L
        gd = tuple(int(v) for v in global_dims)
        od = gd
        for _ in range(int(nlevels)):
            od = tuple(restrict_size(v) for v in od)
        outs, ip, op, sp = self._levels_io(slabs, gd, od, torch, _Arr, jl_empty)
        sps = (C.c_void_p * self.nlocal)(*[torch.cuda.current_stream(a.device).cuda_stream for a in slabs]) if use_streams else None
        dt = L.GB_F32 if slabs[0].dtype == torch.float32 else L.GB_F64
        L.check(L.lib().gb_restrict3_sharded_levels(self._h, dt, _i64s(gd), int(nlevels), ip, float(scale), op, sps))
        return outs

    def prolong3_levels(self, slabs, global_dims, sizes, use_streams=True):
        """prolong(A, sizes) with one column of target sizes per level (`sizes`: list of 3-tuples) in ONE library call."""
        import torch

        from .api import _Arr, _i64s, jl_empty

        L = self._L
        gd = tuple(int(v) for v in global_dims)
        cols = [tuple(int(v) for v in col) for col in sizes]
        outs, ip, op, sp = self._levels_io(slabs, gd, cols[-1], torch, _Arr, jl_empty)
        sps = (C.c_void_p * self.nlocal)(*[torch.cuda.current_stream(a.device).cuda_stream for a in slabs]) if use_streams else None
        dt = L.GB_F32 if slabs[0].dtype == torch.float32 else L.GB_F64
        L.check(L.lib().gb_prolong3_sharded_levels(self._h, dt, _i64s(gd), len(cols), _i64s([v for col in cols for v in col]), ip, op, sps))
        return outs

    def _levels_io(self, slabs, gd, od, torch, _Arr, jl_empty):
        assert len(slabs) == self.nlocal
        outs, ip, op = [], [], []
        for i, a in enumerate(slabs):
            r = self.first_rank + i
            f0, f1 = shard_range(gd[2], r, self.nranks)
            k0, k1 = shard_range(od[2], r, self.nranks)
            assert tuple(a.shape) == (gd[0], gd[1], f1 - f0), "slab %d has the wrong shape" % r
            with torch.cuda.device(a.device):
                o = jl_empty((od[0], od[1], k1 - k0), a.dtype, a.device)
            outs.append(o)
            ip.append(_Arr(a, name="slab").ptr if f1 > f0 else None)
            op.append(o.data_ptr() if k1 > k0 else None)
        return outs, (C.c_void_p * self.nlocal)(*ip), (C.c_void_p * self.nlocal)(*op), None

    def replicate(self, grids, src_rank, meta):
        """grids: list (one per local rank; the source rank's entry an InterpGrid, the others None) -> list of InterpGrids.
        meta = (stored_shape, eltype, BC, IT, fill) -- every process passes the same."""
        from . import api
        from .api import _i64s

        L = self._L
        stored, eltype, bc, it, fill = meta
        hs = (C.c_void_p * self.nlocal)(*[(g._h.value if g is not None else None) for g in grids])
        bcode = 6 if bc is api.BCfill else api._bc(bc)
        dt = L.GB_F32 if np.dtype(eltype) == np.float32 else L.GB_F64
        L.check(L.lib().gb_grid_replicate(self._h, int(src_rank), hs, dt, len(stored), _i64s(stored), bcode, api._it(it), float(fill)))
        out = []
        for i, g in enumerate(grids):
            if g is not None:
                out.append(g)
            else:
                out.append(api.InterpGrid._adopt(C.c_void_p(hs[i]), bc, it, fill))
        return out


_default_comm = None


def default_comm(group=None):
    """The process-wide communicator over torch.distributed's world (created on first use)."""
    global _default_comm
    if _default_comm is None:
        _default_comm = Comm.from_torch(group)
    return _default_comm


def exchange_halo(local, plan, rank, group=None):
    """torch.distributed halo exchange (gloo in the CPU tests; NCCL for the per-dimension fallback route).
    local: column-major tensor holding plan.own[rank] planes along its LAST dimension.
    Returns (ext, first): the tensor widened by the received halo planes and its first plane.  Planes of `ext`
    that are neither owned nor received (a gap between an empty own range and the needed planes) are zero."""
    import torch
    import torch.distributed as dist

    o0, o1 = plan.own[rank]
    e0, e1 = plan.ext_range(rank)
    plane = int(np.prod(local.shape[:-1])) if local.dim() > 1 else 1
    ext = jl_empty_like(local, tuple(local.shape[:-1]) + (e1 - e0,))
    fe, fl = _flat(ext), _flat(local)
    lo, hi = plan.need[rank]
    if o1 == o0 and hi >= lo and (lo > o0 or hi + 1 < o0):
        fe.zero_()  # gap planes
    if o1 > o0:
        fe[(o0 - e0) * plane:(o1 - e0) * plane].copy_(fl)
    ops, keep = [], []
    for src, dst, first, count in plan.transfers:
        if src == rank:
            buf = fl[(first - o0) * plane:(first - o0 + count) * plane].contiguous()
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, dst, group=group))
        elif dst == rank:
            ops.append(dist.P2POp(dist.irecv, fe[(first - e0) * plane:(first - e0 + count) * plane], src, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if local.is_cuda:
        torch.cuda.current_stream(local.device).synchronize()
    return ext, e0


class GpuEngine:
    """libgridb200's slab kernels through the host API (device-resident tensors): the per-dimension route and tests."""

    def __init__(self):
        import grid_jl_b200 as gb

        self.gb = gb

    def restrict_dims12(self, a, scale):
        return self.gb.restrict(a, [True, True, False], scale)

    def prolong_dims12(self, a, m1, m2):
        return self.gb.prolong(a, [m1, m2, a.shape[2]])

    def restrict_slab(self, a, scale, L, in_first, out_first, count):
        return self._slab("gb_restrict_slab", a, (float(scale), int(L)), in_first, out_first, count)

    def prolong_slab(self, a, n, length, in_first, out_first, count):
        return self._slab("gb_prolong_slab", a, (int(n), int(length)), in_first, out_first, count)

    # fused 3-D operators on a slab (gb_restrict3_slab / gb_prolong3_slab: the marching kernels, one pass over the slab)
    fused = True

    def restrict3_slab(self, a, scale, L3, in_first, out_first, count):
        from . import _lib as L
        from .api import _Arr, _i64s

        arr = _Arr(a, name="slab")
        n1, n2 = (restrict_size(int(v)) for v in arr.shape[:2])
        out = arr.empty_like_shape((n1, n2, int(count)))
        if int(count) == 0:
            return out
        o = _Arr(out, name="out")
        L.check(L.lib().gb_restrict3_slab(arr.dtype, _i64s(arr.shape), arr.ptr, float(scale), int(L3), int(in_first), int(out_first), int(count), o.ptr,
                                          arr.mem, arr.stream()))
        return out

    def prolong3_slab(self, a, target, n3, in_first, out_first, count):
        from . import _lib as L
        from .api import _Arr, _i64s

        arr = _Arr(a, name="slab")
        m1, m2, m3 = (int(v) for v in target)
        out = arr.empty_like_shape((m1, m2, int(count)))
        if int(count) == 0:
            return out
        o = _Arr(out, name="out")
        L.check(L.lib().gb_prolong3_slab(arr.dtype, _i64s(arr.shape), arr.ptr, _i64s((m1, m2, m3)), int(n3), int(in_first), int(out_first), int(count), o.ptr,
                                         arr.mem, arr.stream()))
        return out

    def _slab(self, fn, a, head, in_first, out_first, count):
        from . import _lib as L
        from .api import _Arr, _i64s

        arr = _Arr(a, name="slab")
        out = arr.empty_like_shape(tuple(arr.shape[:-1]) + (int(count),))
        if int(count) == 0 or 0 in arr.shape:
            return out
        o = _Arr(out, name="out")
        f = getattr(L.lib(), fn)
        L.check(f(arr.dtype, arr.ndim, _i64s(arr.shape), arr.ptr, arr.ndim - 1, *head, int(in_first), int(out_first), int(count), o.ptr, arr.mem, arr.stream()))
        return out


def slab_restrict3(local, global_len3, rank, world, scale=1.0, engine=None, group=None, comm=None, stats=None):
    """restrict(A, [true,true,true], scale) of a 3-D array sharded by slabs along dim 3.
    local holds planes shard_range(global_len3, rank, world); returns this rank's planes
    shard_range(restrict_size(global_len3), rank, world) of the result.
    engine=None (the product path): gb_restrict3_sharded through `comm` (default: the torch.distributed world)."""
    if engine is None and min(int(local.shape[0]), int(local.shape[1]), int(global_len3)) >= 2:
        comm = comm or default_comm(group)
        assert comm.nranks == world and comm.nlocal == 1 and comm.first_rank == rank
        out = comm.restrict3([local], (local.shape[0], local.shape[1], global_len3), scale)[0]
        if stats is not None:
            stats.append(dict(level=int(global_len3), **comm.stats()[0]))
        return out
    engine = engine or GpuEngine()
    plan = SlabPlan("restrict", global_len3, restrict_size(global_len3), world)
    if getattr(engine, "fused", False) and min(int(local.shape[0]), int(local.shape[1]), int(global_len3)) >= 2:
        # fused: the halo is the raw fine planes (<= 2 per side), then ONE pass over the slab
        ext, first = exchange_halo(local, plan, rank, group)
        k0, k1 = plan.out[rank]
        return engine.restrict3_slab(ext, scale, global_len3, first, k0, k1 - k0)
    r12 = engine.restrict_dims12(local, scale)                  # dims 1,2: planes are independent
    ext, first = exchange_halo(r12, plan, rank, group)          # halo of the 4x smaller planes
    k0, k1 = plan.out[rank]
    return engine.restrict_slab(ext, scale, global_len3, first, k0, k1 - k0)


def slab_prolong3(local, global_n3, target, rank, world, engine=None, group=None, comm=None, stats=None):
    """prolong(A, [m1,m2,m3]) of a 3-D array sharded by slabs along dim 3 (coarse planes
    shard_range(global_n3, ...)); returns planes shard_range(m3, rank, world) of the result."""
    m1, m2, m3 = (int(v) for v in target)
    grows = m1 > int(local.shape[0]) and m2 > int(local.shape[1]) and m3 > int(global_n3)
    if engine is None and grows:
        comm = comm or default_comm(group)
        assert comm.nranks == world and comm.nlocal == 1 and comm.first_rank == rank
        out = comm.prolong3([local], (local.shape[0], local.shape[1], global_n3), (m1, m2, m3))[0]
        if stats is not None:
            stats.append(dict(level=int(m3), **comm.stats()[0]))
        return out
    engine = engine or GpuEngine()
    plan = SlabPlan("prolong", global_n3, m3, world)
    ext, first = exchange_halo(local, plan, rank, group)        # halo of the coarse (small) planes
    if getattr(engine, "fused", False) and grows:
        k0, k1 = plan.out[rank]
        return engine.prolong3_slab(ext, (m1, m2, m3), global_n3, first, k0, k1 - k0)   # one pass (marching kernel)
    p12 = engine.prolong_dims12(ext, m1, m2)                    # dims 1,2 on own + halo planes
    k0, k1 = plan.out[rank]
    return engine.prolong_slab(p12, global_n3, m3, first, k0, k1 - k0)


def broadcast_grid(G, bc, it, src=0, group=None, fill=float("nan"), comm=None):
    """Replicate a prefiltered grid: rank `src` passes its InterpGrid, the others None.  One
    broadcast of the padded coefficient array (gb_grid_replicate: ncclBroadcast over NVLink on the box); every rank
    returns an InterpGrid over its own device copy."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta[0] = (tuple(G.stored_shape), "f32" if G.eltype == np.float32 else "f64")
    dist.broadcast_object_list(meta, src=src, group=group)
    stored, dt = meta[0]
    comm = comm or default_comm(group)
    assert comm.nlocal == 1
    return comm.replicate([G if rank == src else None], src, (stored, np.float32 if dt == "f32" else np.float64, bc, it, fill))[0]
