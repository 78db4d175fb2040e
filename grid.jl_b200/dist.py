"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in
the CPU tests).  Nothing here is on the single-GPU hot path.

* Point evaluation shards trivially: contiguous query ranges per rank, the coefficient array is
  prefiltered once and replicated with ONE broadcast (`broadcast_grid`); no per-call collective.
* restrict / prolong of a 3-D array shard by slabs along the last (slowest) dimension.  One
  exchange step per level: each rank receives the few boundary planes ("halo") its outputs read
  from its neighbours, then runs the fused slab kernels (gb_restrict3_slab / gb_prolong3_slab: the marching
  kernels of multigrid.cu with plane offsets, one pass over the slab).  The per-element arithmetic is
  unchanged, so the gathered result equals the single-device one bit for bit.  Degenerate extents and partial
  prolongations take the per-dimension slab kernels (gb_restrict_slab / gb_prolong_slab), where the halo travels
  after the local dim-1/2 restriction / before the dim-1/2 prolongation.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced split of range(n): rank r owns [lo, hi)."""
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
    """Who sends which planes to whom for one level along the sharded dimension.

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


def _flat(t):
    """Contiguous 1-D view of a column-major (Julia layout) tensor's storage."""
    n = t.dim()
    return t.permute(*range(n - 1, -1, -1)).reshape(-1)


def jl_empty_like(t, shape):
    import torch

    shape = tuple(int(s) for s in shape)
    base = torch.empty(shape[::-1], dtype=t.dtype, device=t.device)
    return base.permute(*range(len(shape) - 1, -1, -1))


def exchange_halo(local, plan, rank, group=None):
    """local: column-major tensor holding plan.own[rank] planes along its LAST dimension.
    Returns (ext, first): the tensor widened by the received halo planes and its first plane."""
    import torch
    import torch.distributed as dist

    o0, o1 = plan.own[rank]
    e0, e1 = plan.ext_range(rank)
    plane = int(np.prod(local.shape[:-1])) if local.dim() > 1 else 1
    ext = jl_empty_like(local, tuple(local.shape[:-1]) + (e1 - e0,))
    fe, fl = _flat(ext), _flat(local)
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
    """The product engine: libgridb200 through the host API (device-resident tensors)."""

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
        o = _Arr(out, name="out")
        L.check(L.lib().gb_prolong3_slab(arr.dtype, _i64s(arr.shape), arr.ptr, _i64s((m1, m2, m3)), int(n3), int(in_first), int(out_first), int(count), o.ptr,
                                         arr.mem, arr.stream()))
        return out

    def _slab(self, fn, a, head, in_first, out_first, count):
        import ctypes as C

        from . import _lib as L
        from .api import _Arr, _i64s

        arr = _Arr(a, name="slab")
        out = arr.empty_like_shape(tuple(arr.shape[:-1]) + (int(count),))
        o = _Arr(out, name="out")
        f = getattr(L.lib(), fn)
        L.check(f(arr.dtype, arr.ndim, _i64s(arr.shape), arr.ptr, arr.ndim - 1, *head, int(in_first), int(out_first), int(count), o.ptr, arr.mem, arr.stream()))
        return out


def slab_restrict3(local, global_len3, rank, world, scale=1.0, engine=None, group=None):
    """restrict(A, [true,true,true], scale) of a 3-D array sharded by slabs along dim 3.
    local holds planes shard_range(global_len3, rank, world); returns this rank's planes
    shard_range(restrict_size(global_len3), rank, world) of the result."""
    engine = engine or GpuEngine()
    plan = SlabPlan("restrict", global_len3, restrict_size(global_len3), world)
    if getattr(engine, "fused", False) and min(int(local.shape[0]), int(local.shape[1]), int(global_len3)) >= 2:
        # fused: the halo is the raw fine planes (<= 2 per side; a few MB over NVLink), then ONE pass over the slab
        # (10 B per fine point of HBM traffic instead of ~24 for the three per-dimension passes)
        ext, first = exchange_halo(local, plan, rank, group)
        k0, k1 = plan.out[rank]
        return engine.restrict3_slab(ext, scale, global_len3, first, k0, k1 - k0)
    r12 = engine.restrict_dims12(local, scale)                  # dims 1,2: planes are independent
    ext, first = exchange_halo(r12, plan, rank, group)          # halo of the 4x smaller planes
    k0, k1 = plan.out[rank]
    return engine.restrict_slab(ext, scale, global_len3, first, k0, k1 - k0)


def slab_prolong3(local, global_n3, target, rank, world, engine=None, group=None):
    """prolong(A, [m1,m2,m3]) of a 3-D array sharded by slabs along dim 3 (coarse planes
    shard_range(global_n3, ...)); returns planes shard_range(m3, rank, world) of the result."""
    engine = engine or GpuEngine()
    m1, m2, m3 = (int(v) for v in target)
    plan = SlabPlan("prolong", global_n3, m3, world)
    ext, first = exchange_halo(local, plan, rank, group)        # halo of the coarse (small) planes
    if getattr(engine, "fused", False) and m1 > int(local.shape[0]) and m2 > int(local.shape[1]) and m3 > int(global_n3):
        k0, k1 = plan.out[rank]
        return engine.prolong3_slab(ext, (m1, m2, m3), global_n3, first, k0, k1 - k0)   # one pass (marching kernel)
    p12 = engine.prolong_dims12(ext, m1, m2)                    # dims 1,2 on own + halo planes
    k0, k1 = plan.out[rank]
    return engine.prolong_slab(p12, global_n3, m3, first, k0, k1 - k0)


def broadcast_grid(G, bc, it, src=0, group=None, fill=float("nan")):
    """Replicate a prefiltered grid: rank `src` passes its InterpGrid, the others None.  One
    broadcast of the padded coefficient array (NCCL over NVLink on the box); every rank returns an
    InterpGrid over its own device copy."""
    import torch
    import torch.distributed as dist

    import grid_jl_b200 as gb

    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta[0] = (tuple(G.stored_shape), "f32" if G.eltype == np.float32 else "f64")
    dist.broadcast_object_list(meta, src=src, group=group)
    stored, dt = meta[0]
    tdt = torch.float32 if dt == "f32" else torch.float64
    if rank == src:
        buf = G.coefs_device().clone()
    else:
        buf = gb.jl_empty(stored, tdt, torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(_flat(buf), src=src, group=group)
    if rank == src:
        return G
    return gb.InterpGrid.from_coefs(buf, bc, it, fill=fill)
