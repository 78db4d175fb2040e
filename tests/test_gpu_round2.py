"""GPU tests of the round-2 paths, through the C ABI, against the CPU oracle:
grouped ("sorted") batches evaluated in place, TMA brick staging vs the cp.async fallback, the GB_FAST prefilter
(kernels of prefilter_fast.cu), filledges!, stream ordering of device-built grids, and the multi-GPU entry points
(gb_comm_*) on virtual ranks of one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS64, EPS32 = 2.220446049250313e-16, 1.1920929e-07


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def tile_key(X, shape):
    """a spatial key that groups queries by blocks of 8 x 8 x 4 cells (2-D: 32 x 16), dim 1 fastest: any block whose
    first taps span at most one brick tile (16 x 16 x 8 cells in 3-D, 64 x 64 in 2-D) will do"""
    fl = np.floor(X).astype(np.int64)
    if X.shape[1] == 3:
        return ((fl[:, 2] >> 2) * 4096 + (fl[:, 1] >> 3)) * 4096 + (fl[:, 0] >> 3)
    return (fl[:, 1] >> 4) * 65536 + (fl[:, 0] >> 5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order,shape", [("nan", "cubic", (70, 66, 50)), ("reflect", "quadratic", (300, 200)), ("nan", "quadratic", (64, 40, 33)),
                                            ("periodic", "quadratic", (48, 40, 36))])
def test_coherent_batches_take_the_plain_kernel(gb, oracle, dtype, bc, order, shape):
    """The tiled path probes the batch as it arrived: spatially coherent orders (sorted by block, by cell in raster order) are
    evaluated in place by the plain kernel (profile: coherent == 1), shuffled ones go through the bins (coherent == 0);
    the oracle's bits either way."""
    import torch

    BC = dict(nan=gb.BCnan, reflect=gb.BCreflect, periodic=gb.BCperiodic)[bc]
    IT = dict(cubic=gb.InterpCubic, quadratic=gb.InterpQuadratic)[order]
    rng = np.random.default_rng(5)
    A = rng.standard_normal(shape).astype(dtype)
    g = gb.InterpGrid(A, BC, IT)
    co = oracle.make_coefs(A, bc, order)
    nq = 400_000
    X = (1.0 + rng.random((nq, len(shape))) * (np.array(shape) - 1.0)).astype(dtype)
    X[::1000] = -3.0  # a few invalid / wrapped locations
    mask = 7
    orders = {"blocks": np.argsort(tile_key(X, shape), kind="stable"),
              "raster": np.lexsort(tuple(np.floor(X[:, d]) for d in range(len(shape)))),
              "random": np.arange(nq)}
    gb.profile_enable(True)
    try:
        for name, perm in orders.items():
            Xs = np.ascontiguousarray(X[perm])
            ov, og, oh, _ = oracle.evaluate(co, bc, order, Xs, mask, raise_invalid=False)
            for fast in (False, True):
                v, gr, h = g._eval(torch.from_numpy(Xs).cuda(), mask, tiled=True, sorted=True, fast=fast)  # forced onto the tiled path, probe on
                st = gb.profile_read()
                assert st["coherent"] == (0.0 if name == "random" else 1.0), (name, st)
                if not fast:
                    assert same(v.cpu().numpy(), ov) and same(gr.cpu().numpy(), og) and same(h.cpu().numpy(), oh), name
        v, gr, h = g._eval(torch.from_numpy(np.ascontiguousarray(X[orders["blocks"]])).cuda(), mask, tiled=True)  # GB_TILED alone: the bins, no probe
        assert gb.profile_read()["coherent"] == 0.0
    finally:
        gb.profile_enable(False)


def test_coherence_probe_corner_cases(gb, oracle):
    """tiny batches, batches that are not a multiple of anything, every query invalid, BCnil's first invalid index on the
    coherent route"""
    import torch

    rng = np.random.default_rng(9)
    A = rng.standard_normal((40, 40, 40))
    g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    co = oracle.make_coefs(A, "nan", "cubic")
    gb.profile_enable(True)
    try:
        for nq in (1, 31, 2049, 70_001):
            X = 5.0 + rng.random((nq, 3)) * 3.0  # all in one cell block
            v, gr, _ = g._eval(torch.from_numpy(X).cuda(), 3, tiled=True, sorted=True)
            assert gb.profile_read()["coherent"] == 1.0
            ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", X, 3)
            assert same(v.cpu().numpy(), ov) and same(gr.cpu().numpy(), og)
        X = np.full((5000, 3), -7.0)
        v, gr, _ = g._eval(torch.from_numpy(X).cuda(), 3, tiled=True, sorted=True)
        assert bool(torch.isnan(v).all()) and bool(torch.isnan(gr).all())
    finally:
        gb.profile_enable(False)
    gn = gb.InterpGrid(A, gb.BCnil, gb.InterpCubic)
    X = 5.0 + rng.random((5000, 3)) * 3.0
    X[3210, 1] = 77.0
    with pytest.raises(gb.ErrorException) as ei:
        gn._eval(torch.from_numpy(X).cuda(), 3, tiled=True, sorted=True)
    assert "query 3210" in str(ei.value)


def test_tma_and_cp_async_staging_agree(gb, oracle):
    """Bricks staged by TMA (row pitch a multiple of 16 bytes) and by the element-wise fallback (odd pitches; GB_NO_TMA=1
    in a child process) give the oracle's bits."""
    import torch

    rng = np.random.default_rng(21)
    for dtype, shapes in ((np.float64, [(66, 41, 37), (67, 41, 37), (130, 90)]), (np.float32, [(68, 33, 30), (66, 33, 30), (131, 77)])):
        for shape in shapes:
            A = rng.standard_normal(shape).astype(dtype)
            bc, order = ("nan", "cubic") if len(shape) == 3 else ("reflect", "quadratic")
            g = gb.InterpGrid(A, gb.BCnan if bc == "nan" else gb.BCreflect, gb.InterpCubic if order == "cubic" else gb.InterpQuadratic)
            co = oracle.make_coefs(A, bc, order)
            X = (0.5 + rng.random((120_000, len(shape))) * (np.array(shape))).astype(dtype)
            ov, og, _, _ = oracle.evaluate(co, bc, order, X, 3, raise_invalid=False)
            v, gr, _ = g._eval(torch.from_numpy(X).cuda(), 3, tiled=True)
            assert same(v.cpu().numpy(), ov) and same(gr.cpu().numpy(), og), (dtype, shape)
    code = r"""
import sys; sys.path.insert(0, %r)
import numpy as np, torch
import grid_jl_b200 as gb
from oracle import oracle as O
rng = np.random.default_rng(22)
A = rng.standard_normal((66, 41, 37))
g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
co = O.make_coefs(A, "nan", "cubic")
X = 0.5 + rng.random((100000, 3)) * np.array(A.shape)
ov, og, _, _ = O.evaluate(co, "nan", "cubic", X, 3, raise_invalid=False)
v, gr, _ = g._eval(torch.from_numpy(X).cuda(), 3, tiled=True)
assert np.array_equal(v.cpu().numpy(), ov, equal_nan=True) and np.array_equal(gr.cpu().numpy(), og, equal_nan=True)
print("ok")
""" % ROOT
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GB_NO_TMA="1"), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_multi_level_restrict_prolong_one_call(gb, oracle, dtype):
    """restrict(A, flags) with several all-true columns and prolong(A, sizes) with several columns run as ONE library call
    (gb_restrict3_levels / gb_prolong3_levels; the intermediate levels stay on the device): same bits as the oracle's
    column-by-column mapdim, device and host arrays, odd / even / degenerate extents."""
    rng = np.random.default_rng(47)
    for shape, nlev in (((40, 33, 26), 3), ((17, 16, 9), 2), ((9, 5, 3), 3), ((64, 64, 64), 4)):
        A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
        flags = np.ones((3, nlev), dtype=bool)
        ref = oracle.restrict_flags(A, flags, 0.5)
        sizes, cur = [list(shape)], list(shape)
        for _ in range(nlev):
            cur = [gb.restrict_size(v) for v in cur]
            sizes.append(cur)
        for dev in (True, False):
            a = gb.jl_from_numpy(A) if dev else A
            r = gb.restrict(a, flags, 0.5)
            got = gb.jl_to_numpy(r) if dev else r
            assert got.shape == tuple(sizes[-1]) and same(got, ref), (shape, dev)
            r1 = a
            for _ in range(nlev):
                r1 = gb.restrict(r1, [True, True, True], 0.5)
            assert same(gb.jl_to_numpy(r1) if dev else r1, ref)
            if min(sizes[-1]) >= 2:
                sz = np.array(list(reversed(sizes[:-1]))).T  # (3, nlev): the way back up
                refp = oracle.prolong_sizes(ref, sz)
                p = gb.prolong(r, sz)
                assert same(gb.jl_to_numpy(p) if dev else p, refp), (shape, dev)
    with pytest.raises(gb.ErrorException):
        gb.prolong(gb.jl_from_numpy(np.asfortranarray(np.zeros((5, 5, 5), dtype=dtype))), np.array([[9, 18], [9, 17], [9, 17]]))


def test_small_host_batches_fast_path(gb, oracle):
    """Host batches up to 2^15 queries (the scalar API: one query per C call) run through a pooled context that is
    created once (abi.cu: eval_host_small).  Bits as the oracle's for scalar calls, short batches of every output kind,
    column inputs, Float32, packed records, a multi-valued grid, growing and shrinking sizes, calls from several threads, and
    BCnil's error with the first invalid query."""
    import threading

    rng = np.random.default_rng(43)
    A = rng.standard_normal((30, 28, 26))
    g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    co = oracle.make_coefs(A, "nan", "cubic")
    for nq in (1, 2, 777, 1, 32768, 5):
        X = 1.0 + rng.random((nq, 3)) * (np.array(A.shape) - 1.0)
        ov, og, oh, _ = oracle.evaluate(co, "nan", "cubic", X, 7)
        v, gr, h = g._eval(X, 7)
        assert same(v, ov) and same(gr, og) and same(h, oh), nq
        rec = gb.valgrad_packed(g, X)
        assert same(rec[:, 0], ov) and same(rec[:, 1:4], og)
        cv = gb.getindex_batch(g, [np.ascontiguousarray(X[:, d]) for d in range(3)])
        assert same(cv, ov)
    x = (3.25, 7.5, 11.125)
    ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", np.array([x]), 3)
    assert g[x] == ov[0]
    val, grad = gb.valgrad(g, *x)
    assert val == ov[0] and same(np.asarray(grad), og[0])
    # Float32 grid and queries, 2-D, reflect
    B = rng.standard_normal((50, 40)).astype(np.float32)
    g2 = gb.InterpGrid(B, gb.BCreflect, gb.InterpQuadratic)
    c2 = oracle.make_coefs(B, "reflect", "quadratic")
    X2 = (rng.random((900, 2)) * np.array(B.shape) * 1.3).astype(np.float32)
    ov2, og2, oh2, _ = oracle.evaluate(c2, "reflect", "quadratic", X2, 7)
    v2, gr2, h2 = g2._eval(X2, 7)
    assert same(v2, ov2) and same(gr2, og2) and same(h2, oh2)
    # several host threads, each with its own context
    errs = []

    def work(seed):
        try:
            r = np.random.default_rng(seed)
            for _ in range(20):
                Xt = 1.0 + r.random((int(r.integers(1, 400)), 3)) * (np.array(A.shape) - 1.0)
                e, _, _, _ = oracle.evaluate(co, "nan", "cubic", Xt, 1)
                vt, _, _ = g._eval(Xt, 1)
                assert same(vt, e)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(s,)) for s in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    gn = gb.InterpGrid(A, gb.BCnil, gb.InterpLinear)
    Xn = 2.0 + rng.random((100, 3))
    Xn[41] = (0.5, 2.0, 2.0)
    with pytest.raises(gb.ErrorException) as ei:
        gn._eval(Xn, 1)
    assert "query 41" in str(ei.value)
    v, _, _ = gn._eval(Xn[:41], 1)
    assert np.isfinite(v).all()


def test_pageable_host_buffers_staged_path(gb):
    """Pageable (numpy) host buffers above 2^19 queries go through the library's pinned staging slots with several host threads
    (abi.cu: eval_host_staged).  Same bits as the device-resident call: AoS and column inputs, Float32 queries, value /
    gradient / Hessian / packed outputs, a multi-valued grid, a ragged last chunk, and BCnil's first invalid query."""
    import torch

    rng = np.random.default_rng(41)
    nq = (1 << 19) + 70_001
    A = rng.standard_normal((40, 36, 30))
    g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    X = 1.0 + rng.random((nq, 3)) * (np.array(A.shape) - 1.0)
    Xd = torch.from_numpy(X).cuda()
    dv, dg, _ = g._eval(Xd, 3)
    hv, hg, _ = g._eval(X, 3)
    assert isinstance(hv, np.ndarray) and same(hv, dv.cpu().numpy()) and same(hg, dg.cpu().numpy())
    rec = gb.valgrad_packed(g, X)
    assert same(rec[:, 0], hv) and same(rec[:, 1:4], hg)
    # columns (one vector per dimension), Float32 queries
    cols = [np.ascontiguousarray(X[:, d].astype(np.float32)) for d in range(3)]
    cv = gb.getindex_batch(g, cols)
    dvc = gb.getindex_batch(g, [torch.from_numpy(c).cuda() for c in cols])
    assert same(cv, dvc.cpu().numpy())
    # 2-D quadratic Float32 grid, valgradhess
    B = rng.standard_normal((300, 200)).astype(np.float32)
    g2 = gb.InterpGrid(B, gb.BCreflect, gb.InterpQuadratic)
    X2 = (rng.random((nq, 2)) * np.array(B.shape) * 1.2).astype(np.float32)
    hv2, hg2, hh2 = g2._eval(X2, 7)
    dv2, dg2, dh2 = g2._eval(torch.from_numpy(X2).cuda(), 7)
    assert same(hv2, dv2.cpu().numpy()) and same(hg2, dg2.cpu().numpy()) and same(hh2, dh2.cpu().numpy())
    # BCnil: the first invalid query of the whole batch is reported whichever worker met it
    gn = gb.InterpGrid(A, gb.BCnil, gb.InterpLinear)
    Xn = X.copy()
    Xn[590_123] = (0.25, 2.0, 2.0)
    Xn[333_333] = (2.0, 99.0, 2.0)
    with pytest.raises(gb.ErrorException) as ei:
        gn._eval(Xn, 1)
    assert "query 333333" in str(ei.value)


def test_single_pass_scatter(gb, oracle):
    """GB_SORT_PASSES=1 (child process): bins sized from the probe's sample, one scatter pass, overflow list.  A uniform
    batch (no or few overflows), and a batch whose SAMPLED positions (the first 256 of every 1024 queries, see
    coherence_probe_kernel) all sit in one corner while the rest is uniform -- nearly every query finds its bin full and
    goes through the overflow kernel.  Both bit-identical to the oracle, packed output included."""
    code = r"""
import sys; sys.path.insert(0, %r)
import numpy as np, torch
import grid_jl_b200 as gb
from oracle import oracle as O
rng = np.random.default_rng(23)
for dtype, shape, bc, order in ((np.float64, (66, 41, 37), "nan", "cubic"), (np.float32, (130, 90), "reflect", "quadratic"), (np.float64, (40, 36, 70), "periodic", "quadratic")):
    A = rng.standard_normal(shape).astype(dtype)
    BC = dict(nan=gb.BCnan, reflect=gb.BCreflect, periodic=gb.BCperiodic)[bc]
    g = gb.InterpGrid(A, BC, gb.InterpCubic if order == "cubic" else gb.InterpQuadratic)
    co = O.make_coefs(A, bc, order)
    nq = 200_000
    Xu = (0.5 + rng.random((nq, len(shape))) * np.array(shape)).astype(dtype)
    Xs = Xu.copy()
    Xs[(np.arange(nq) %% 1024) < 256] = 1.25
    for X in (Xu, Xs, Xu[:777], Xu[:4097]):
        ov, og, _, _ = O.evaluate(co, bc, order, X, 3, raise_invalid=False)
        v, gr, _ = g._eval(torch.from_numpy(X).cuda(), 3, tiled=True)
        assert np.array_equal(v.cpu().numpy(), ov, equal_nan=True) and np.array_equal(gr.cpu().numpy(), og, equal_nan=True)
        rec = gb.valgrad_packed(g, torch.from_numpy(X).cuda(), tiled=True) if hasattr(gb, "valgrad_packed") else None
        if rec is not None:
            r = rec.cpu().numpy()
            assert np.array_equal(r[:, 0], ov, equal_nan=True) and np.array_equal(r[:, 1:1 + len(shape)], og, equal_nan=True)
print("ok")
""" % ROOT
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GB_SORT_PASSES="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order", [("nan", "quadratic"), ("reflect", "quadratic"), ("periodic", "quadratic"), ("nearest", "quadratic"), ("nan", "cubic")])
def test_fast_prefilter_kernels(gb, oracle, dtype, bc, order):
    """interp_invert!(..., fast=True): TMA-tiled dim-1 sweeps, strided sweeps, dense end rows, ping-pong buffers, and the
    sweeps that fall back to the exact kernel -- against the oracle's exact solve: <= 6 ulp of max|A| per sweep chain
    (Float64; the oracle's own rounding included), <= 1e-5 relative (Float32)."""
    import torch

    BC = dict(nan=gb.BCnan, reflect=gb.BCreflect, periodic=gb.BCperiodic, nearest=gb.BCnearest)[bc]
    IT = gb.InterpCubic if order == "cubic" else gb.InterpQuadratic
    rng = np.random.default_rng(31)
    eps = EPS64 if dtype == np.float64 else EPS32
    for shape in [(1024, 640), (512, 300), (300, 512), (4096,), (260, 40, 288), (8192, 96), (1000, 33)]:
        A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
        ref = oracle.invert(A, bc, order)
        tol = (6 * len(shape) * eps if dtype == np.float64 else 1e-5) * float(np.abs(ref).max())
        for dev in (True, False):
            B = gb.jl_from_numpy(A) if dev else A.copy(order="F")
            gb.interp_invert_(B, BC, IT, fast=True)
            got = gb.jl_to_numpy(B) if dev else B
            assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= tol, (shape, dev)
        # a dimlist subset (odd number of flipping sweeps in place) and exactness of the default mode next to it
        if len(shape) >= 2:
            B = gb.jl_from_numpy(A)
            gb.interp_invert_(B, BC, IT, dimlist=[1], fast=True)
            r1 = oracle.invert(A, bc, order, dimlist=[0])
            assert np.abs(gb.jl_to_numpy(B).astype(np.float64) - r1.astype(np.float64)).max() <= tol
        B = gb.jl_from_numpy(A)
        gb.interp_invert_(B, BC, IT)
        assert same(gb.jl_to_numpy(B), ref)


def test_fast_prefilter_grid_construction(gb, oracle):
    """InterpGrid(..., fast_prefilter=True) on an image-like grid: coefficients within tolerance of the oracle's, on-grid
    reproduction (test/grid.jl:50-98), and evaluation identical to a grid adopted from those same coefficients."""
    import torch

    rng = np.random.default_rng(33)
    for dtype, bc, order, shape in ((np.float32, "reflect", "quadratic", (2048, 1024)), (np.float64, "nan", "cubic", (700, 300)), (np.float64, -2.5, "quadratic", (600, 512))):
        A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
        BC = bc if not isinstance(bc, str) else dict(nan=gb.BCnan, reflect=gb.BCreflect)[bc]
        IT = gb.InterpCubic if order == "cubic" else gb.InterpQuadratic
        g = gb.InterpGrid(gb.jl_from_numpy(A), BC, IT, fast_prefilter=True)
        co = oracle.make_coefs(A, "fill" if not isinstance(bc, str) else bc, order, bc if not isinstance(bc, str) else None)
        eps = EPS64 if dtype == np.float64 else EPS32
        assert np.abs(g.coefs.astype(np.float64) - co.astype(np.float64)).max() <= 16 * eps * float(np.abs(co).max())
        idx = np.stack([rng.integers(1, n + 1, 20000) for n in shape], axis=1).astype(np.float64)
        v = gb.getindex_batch(g, idx)
        a = A[tuple((idx[:, d] - 1).astype(int) for d in range(2))]
        assert np.abs(v - a).max() < (1e-12 if dtype == np.float64 else 2e-5)
        g2 = gb.InterpGrid.from_coefs(g.coefs, BC if isinstance(bc, str) else gb.BCfill, IT)
        X = 1.0 + rng.random((50000, 2)) * (np.array(shape) - 1.0)
        assert same(gb.getindex_batch(g, X), gb.getindex_batch(g2, X))


def test_filledges(gb):
    """filledges!(A, val), src/utilities.jl:22-33"""
    rng = np.random.default_rng(41)
    for shape in [(7,), (5, 6), (4, 5, 6), (3, 4, 2, 5), (1, 9), (2, 2, 2)]:
        for dtype in (np.float64, np.float32):
            A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
            ref = A.copy(order="F")
            for d in range(A.ndim):
                sl = [slice(None)] * A.ndim
                sl[d] = [0, shape[d] - 1]
                ref[tuple(sl)] = -7.25
            B = A.copy(order="F")
            gb.filledges_(B, -7.25)
            assert same(B, ref)
            D = gb.jl_from_numpy(A)
            gb.filledges_(D, -7.25)
            assert same(gb.jl_to_numpy(D), ref)


def test_device_built_grid_is_ordered_before_host_queries(gb, oracle):
    """A grid built from a device tensor on a non-default stream, then queried through the HOST path right away (other
    streams inside the library): the readers wait for the construction (abi.cu eval_host)."""
    import torch

    n = 200
    A = oracle.synth(np.float64, n ** 3, 8).reshape((n, n, n), order="F")
    co = oracle.make_coefs(A, "nan", "cubic")
    X = 1.0 + oracle.synth(np.float64, 3 * 5000, 18).reshape(5000, 3) * (n - 1.0)
    ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", X, 3)
    Ad = gb.jl_from_numpy(A)
    st = torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(st):
            big = torch.empty(1 << 27, device="cuda").normal_()  # keep the stream busy ahead of the construction
            g = gb.InterpGrid(Ad, gb.BCnan, gb.InterpCubic)
        v, gr = gb.valgrad_batch(g, X)  # numpy queries: host path, library streams
        assert same(v, ov) and same(gr, og)
        assert gb.valgrad(g, *X[7])[0] == ov[7]
        assert same(g.coefs, co)
        del big, g
    torch.cuda.synchronize()


def test_interp_irregular_batched_entry_points(gb):
    rng = np.random.default_rng(43)
    knots = np.sort(rng.random(50)) * 10
    vals = rng.standard_normal(50)
    G = gb.InterpIrregular(knots, vals, gb.BCnan, gb.InterpLinear)
    x = rng.random(1000) * 10
    assert same(gb.getindex_batch(G, x), G[x])
    out = np.empty(1000)
    gb.getindex_(out, G, x)
    assert same(out, G[x])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,world", [((40, 36, 50), 2), ((33, 20, 64), 3), ((20, 18, 129), 5), ((12, 10, 17), 8), ((9, 8, 20), 4), ((66, 40, 128), 8),
                                         ((12, 10, 5), 8), ((9, 8, 6), 8)])  # (the last two: more ranks than planes -- ranks that own or produce nothing)
@pytest.mark.parametrize("peer", [1, 0])
def test_comm_loopback_sharded_multigrid(gb, oracle, dtype, shape, world, peer):
    """gb_restrict3_sharded / gb_prolong3_sharded on `world` virtual ranks of one GPU (gb_comm_init_loopback): plan, halo
    planes, interior / boundary split and the slab kernels -- assembled slabs equal the unsharded oracle result bitwise.
    peer = 1: the boundary kernel reads the neighbours' planes in place through (here: same-device) peer pointers, falling
    back to the exchange where a stencil reaches past the immediate neighbours; peer = 0: halo exchange + assembled buffers."""
    import torch

    import grid_jl_b200.dist as D

    A = oracle.synth(dtype, int(np.prod(shape)), seed=51).reshape(shape, order="F")
    ref = oracle.restrict_flags(A, [True, True, True])
    refp = oracle.prolong_sizes(ref, list(shape))
    comm = D.Comm.loopback(world)
    try:
        assert (comm.nranks, comm.nlocal, comm.first_rank) == (world, world, 0)
        assert comm.peer_halo(peer) == bool(peer) and comm.peer_halo() == bool(peer)
        L3, n3 = shape[2], D.restrict_size(shape[2])
        fine = [gb.jl_from_numpy(np.asfortranarray(A[:, :, slice(*D.shard_range(L3, r, world))])) for r in range(world)]
        for use_streams in (True, False):
            coarse = comm.restrict3(fine, shape, 1.0, use_streams=use_streams)
            torch.cuda.synchronize()
            for r in range(world):
                k0, k1 = D.shard_range(n3, r, world)
                assert same(gb.jl_to_numpy(coarse[r]), np.asfortranarray(ref[:, :, k0:k1])), ("restrict", r)
            st = comm.stats()
            assert len(st) == world and all(s["total_ms"] >= 0 for s in st)
            up = comm.prolong3(coarse, ref.shape, shape, use_streams=use_streams)
            torch.cuda.synchronize()
            for r in range(world):
                f0, f1 = D.shard_range(L3, r, world)
                assert same(gb.jl_to_numpy(up[r]), np.asfortranarray(refp[:, :, f0:f1])), ("prolong", r)
        halo = sum(s["halo_bytes"] for s in comm.stats())
        assert (halo > 0) == (world > 1)
        # slabs at least a few planes thick on every level: the peer route is really taken when it is on
        if min(L3, n3) >= 4 * world:
            assert comm.last_path() == ("peer" if peer else "exchange")
        # several levels in one call (restrict(A, flags) with one column per level, prolong(A, sizes) likewise): same bits
        sizes, cur_ref = [tuple(shape)], A
        while min(sizes[-1]) >= 5 and len(sizes) < 4:
            cur_ref = oracle.restrict_flags(cur_ref, [True, True, True])
            sizes.append(tuple(cur_ref.shape))
        nlev = len(sizes) - 1
        if nlev >= 2:
            for rep in range(2):  # (twice: the communicator's scratch arrays are reused)
                low = comm.restrict3_levels(fine, shape, nlev, 1.0)
                torch.cuda.synchronize()
                for r in range(world):
                    k0, k1 = D.shard_range(sizes[-1][2], r, world)
                    assert same(gb.jl_to_numpy(low[r]), np.asfortranarray(cur_ref[:, :, k0:k1])), ("restrict levels", r)
                back = comm.prolong3_levels(low, sizes[-1], list(reversed(sizes[:-1])))
                torch.cuda.synchronize()
                pref = cur_ref
                for tgt in reversed(sizes[:-1]):
                    pref = oracle.prolong_sizes(pref, list(tgt))
                for r in range(world):
                    f0, f1 = D.shard_range(L3, r, world)
                    assert same(gb.jl_to_numpy(back[r]), np.asfortranarray(pref[:, :, f0:f1])), ("prolong levels", r)

    finally:
        comm.close()


def test_comm_loopback_replicate_and_sharded_eval(gb, oracle):
    import ctypes as C

    import grid_jl_b200.dist as D

    rng = np.random.default_rng(61)
    A = rng.standard_normal((30, 28, 26))
    co = oracle.make_coefs(A, "nan", "cubic")
    comm = D.Comm.loopback(3)
    try:
        g0 = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
        grids = comm.replicate([None, g0, None], 1, (g0.stored_shape, np.float64, gb.BCnan, gb.InterpCubic, float("nan")))
        assert grids[1] is g0 and all(same(g.coefs, co) for g in grids)
        X = 1.0 + rng.random((100_001, 3)) * (np.array(A.shape) - 1.0)
        X[77_777, 2] = 99.0
        ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", X, 3, raise_invalid=False)
        v, gr = np.empty(len(X)), np.empty((len(X), 3))
        L = gb._lib
        hs = (C.c_void_p * 3)(*[g._h.value for g in grids])
        bad = C.c_int64(-5)
        L.check(L.lib().gb_eval_sharded(comm._h, hs, 3, len(X), X.ctypes.data, L.GB_F64, None, v.ctypes.data, gr.ctypes.data, None, 0, C.byref(bad)))
        assert same(v, ov) and same(gr, og) and bad.value == -1  # BCnan: NaN, no error
    finally:
        comm.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order,shape", [("nan", "cubic", (40, 37, 35)), ("reflect", "quadratic", (300, 200)), ("periodic", "linear", (9, 20, 17, 12)),
                                            ("nearest", "nearest", (50,)), ("nan", "quadratic", (33, 31, 30))])
def test_packed_output_layout(gb, oracle, dtype, bc, order, shape):
    """GB_OUT_PACKED: one (value, gradient[, Hessian]) record per query -- the same bits as the separate arrays, on the plain
    path, through the bins, on the probe's coherent route and through host buffers."""
    import torch

    BC = dict(nan=gb.BCnan, reflect=gb.BCreflect, periodic=gb.BCperiodic, nearest=gb.BCnearest)[bc]
    IT = dict(cubic=gb.InterpCubic, quadratic=gb.InterpQuadratic, linear=gb.InterpLinear, nearest=gb.InterpNearest)[order]
    rng = np.random.default_rng(71)
    A = rng.standard_normal(shape).astype(dtype)
    g = gb.InterpGrid(A, BC, IT)
    N = len(shape)
    X = (0.5 + rng.random((90_000, N)) * np.array(shape)).astype(dtype)
    L = gb._lib
    for mask in ((L.GB_OUT_VAL, L.GB_OUT_VAL | L.GB_OUT_GRAD) + ((L.GB_OUT_VAL | L.GB_OUT_GRAD | L.GB_OUT_HESS,) if order in ("quadratic", "cubic") else ())):
        v, gr, h = g._eval(X, mask)
        ref = np.concatenate([v[:, None]] + ([gr] if mask & L.GB_OUT_GRAD else []) + ([h.reshape(len(X), N * N)] if mask & L.GB_OUT_HESS else []), axis=1)
        assert same(g._eval_packed(X, mask), ref)                                        # host buffers
        Xd = torch.from_numpy(X).cuda()
        for tiled, srt in ((False, False), (True, False), (True, True)):
            assert same(g._eval_packed(Xd, mask, tiled=tiled, sorted=srt).cpu().numpy(), ref), (mask, tiled, srt)
        Xs = np.ascontiguousarray(X[np.lexsort(tuple(np.floor(X[:, d]) for d in range(N)))])   # a coherent order: the probe's plain route
        vs, gs, hs = g._eval(Xs, mask)
        refs = np.concatenate([vs[:, None]] + ([gs] if mask & L.GB_OUT_GRAD else []) + ([hs.reshape(len(X), N * N)] if mask & L.GB_OUT_HESS else []), axis=1)
        assert same(g._eval_packed(torch.from_numpy(Xs).cuda(), mask, tiled=True, sorted=True).cpu().numpy(), refs)
    vg = gb.valgrad_packed(g, torch.from_numpy(X).cuda())
    v, gr = gb.valgrad_batch(g, X)
    assert same(vg.cpu().numpy(), np.concatenate([v[:, None], gr], axis=1))
