"""BASELINE.json configs C1, C3, C4, C5 (and C2's oracle equality) at FULL size on the GPU, through
the C-ABI, bitwise against the CPU oracle.

Evaluation configs: the whole batch is evaluated on the device; a strided >= 1e5-query subsample (plus
the wide / edge parity sets of SURVEY 8d) is replayed by the oracle on the same coefficients and must
match bit for bit.  The coefficient arrays themselves (prefilter at full size, incl. the blocked
resident layout C3 / C4 ship with) are compared with oracle.make_coefs in full.
C5: every level of the 1024^3 chain; levels >= 129^3 in full, the big levels through corner and centre
sub-blocks (restrict / prolong are local: an even-offset sub-block of the fine array restricts to the
matching sub-block of the coarse one, edges included when the block touches the array edge)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def sub(t, idx):
    return t[idx].cpu().numpy()


def _assert_blocked(gb, G):
    """The handle keeps its coefficients in 4^N-element Morton blocks: the zero-copy view is refused."""
    import ctypes as C

    p = C.c_void_p()
    assert gb._lib.lib().gb_grid_coefs_device(G._h, C.byref(p)) == gb._lib.GB_ERR_UNSUPPORTED


def test_full_size_c1(gb, oracle):
    """C1: 2-D InterpQuadratic BCreflect 1024^2 f64, getindex; 1e6 and 1e8 uniform queries."""
    import torch

    n = 1024
    A = gb.jl_empty((n, n), torch.float64)
    gb.synth_uniform(A, 1)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    co = oracle.make_coefs(oracle.synth(np.float64, n * n, 1).reshape((n, n), order="F"), "reflect", "quadratic")
    assert same(G.coefs, co), "prefiltered 1024^2 coefficients differ from the oracle"
    for nq in (1_000_000, 100_000_000):
        X = torch.empty((nq, 2), dtype=torch.float64, device="cuda")
        gb.synth_uniform(X, 11, lo=1.0, hi=float(n))
        v = torch.empty(nq, dtype=torch.float64, device="cuda")
        gb.getindex_(v, G, X)
        idx = torch.arange(0, nq, max(1, nq // 200_000), device="cuda")
        Xs = sub(X, idx)
        # the device generator is replayable on the CPU: query q, dim d <- splitmix64(seed + q*N + d)
        if nq == 1_000_000:
            assert same(X.cpu().numpy(), 1.0 + oracle.synth(np.float64, 2 * nq, 11).reshape(nq, 2) * (n - 1.0))
        ov, _, _, _ = oracle.evaluate(co, "reflect", "quadratic", Xs, 1)
        assert same(ov, sub(v, idx))
        del X, v
    # the wide parity set (mirror images) + edge vectors
    Xw = torch.empty((400_000, 2), dtype=torch.float64, device="cuda")
    gb.synth_uniform(Xw, 13, lo=-1536.0, hi=2560.0)
    edges = torch.tensor([[1.0, 1.0], [1.5, 2.5], [0.5, 1024.5], [1024.0, 1024.0], [-0.0, 0.0], [1024.5, 0.5], [2048.5, -1023.5], [3.0, 1023.5]], dtype=torch.float64,
                         device="cuda")
    Xw = torch.cat([Xw, edges])
    vw = gb.getindex_batch(G, Xw)
    ow, _, _, _ = oracle.evaluate(co, "reflect", "quadratic", Xw.cpu().numpy(), 1)
    assert same(ow, vw.cpu().numpy())


def test_full_size_c2_oracle_equality(gb, oracle):
    """C2: 3-D InterpCubic BCnan 256^3 f64 incl. the prefilter, 1e7 queries, valgrad: the 258^3 coefficient
    array in full and a 1e5-query strided subsample, bitwise (the brick path; properties are in
    test_gpu_parity.py::test_full_size_properties_c2)."""
    import torch

    n, nq = 256, 10_000_000
    A = gb.jl_empty((n, n, n), torch.float64)
    gb.synth_uniform(A, 2)
    G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    co = oracle.make_coefs(oracle.synth(np.float64, n ** 3, 2).reshape((n, n, n), order="F"), "nan", "cubic")
    assert same(G.coefs, co), "prefiltered 258^3 coefficients differ from the oracle"
    X = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, 12, lo=1.0, hi=float(n))
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    g = torch.empty((nq, 3), dtype=torch.float64, device="cuda")
    gb.valgrad_(v, g, G, X)
    idx = torch.arange(0, nq, 100, device="cuda")
    ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", sub(X, idx), 3)
    assert same(ov, sub(v, idx)) and same(og, sub(g, idx))
    # out-of-range points: NaN in the value and every gradient component (SURVEY 8d parity set)
    Xo = X[:200_000].clone()
    Xo[::3, 0] = 0.5
    Xo[1::3, 1] = n + 1e-9
    vo, go = gb.valgrad_batch(G, Xo)
    oo, ogo, _, _ = oracle.evaluate(co, "nan", "cubic", Xo.cpu().numpy(), 3)
    assert same(oo, vo.cpu().numpy()) and same(ogo, go.cpu().numpy())


def test_full_size_c3(gb, oracle):
    """C3: 2-D InterpQuadratic BCreflect 8192^2 f32 image, valgradhess; a 1e8-query slice of the 1e9 batch
    (query q of slice k <- splitmix64(seed + (k*1e8 + q)*2 + d)), on the blocked resident layout."""
    import torch

    n, nq = 8192, 100_000_000
    A = gb.jl_empty((n, n), torch.float32)
    gb.synth_uniform(A, 3)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    _assert_blocked(gb, G)  # 268 MB, 36-byte neighbourhoods: the shipped layout is the blocked one
    co = oracle.make_coefs(oracle.synth(np.float32, n * n, 3).reshape((n, n), order="F"), "reflect", "quadratic")
    assert same(G.coefs, co), "prefiltered 8192^2 f32 coefficients differ from the oracle"
    X = torch.empty((nq, 2), dtype=torch.float32, device="cuda")
    v = torch.empty(nq, dtype=torch.float32, device="cuda")
    g = torch.empty((nq, 2), dtype=torch.float32, device="cuda")
    h = torch.empty((nq, 2, 2), dtype=torch.float32, device="cuda")
    for k in (0, 9):  # first and last slice of the 1e9-query batch
        gb.synth_uniform(X, 14, lo=1.0, hi=float(n), first=k * nq * 2)
        gb.valgradhess_(v, g, h, G, X)
        idx = torch.arange(0, nq, 500, device="cuda")
        ov, og, oh, _ = oracle.evaluate(co, "reflect", "quadratic", sub(X, idx), 7)
        assert same(ov, sub(v, idx)) and same(og, sub(g, idx)) and same(oh, sub(h, idx))
    # mirror images and edge vectors
    Xw = torch.empty((300_000, 2), dtype=torch.float32, device="cuda")
    gb.synth_uniform(Xw, 16, lo=-12288.0, hi=20480.0)
    edges = torch.tensor([[1.0, 1.0], [1.5, 8192.5], [0.5, 8192.0], [8192.5, 0.5], [-0.0, 16384.5], [8191.5, 2.5]], dtype=torch.float32, device="cuda")
    Xw = torch.cat([Xw, edges])
    vw, gw, hw = gb.valgradhess_batch(G, Xw)
    ow, ogw, ohw, _ = oracle.evaluate(co, "reflect", "quadratic", Xw.cpu().numpy(), 7)
    assert same(ow, vw.cpu().numpy()) and same(ogw, gw.cpu().numpy()) and same(ohw, hw.cpu().numpy())


@pytest.mark.parametrize("bc", ["periodic", "nearest", "fill"])
def test_full_size_c4(gb, oracle, bc):
    """C4: 4-D InterpLinear 64^4 f64 (66^4 stored for BCfill) through CoordInterpGrid with FloatRange
    coordinates, 1e8 queries with index coordinates U[1-16, 64+16] (25 % outside the grid)."""
    import torch

    n, nq = 64, 100_000_000
    A = gb.jl_empty((n,) * 4, torch.float64)
    gb.synth_uniform(A, 4)
    An = oracle.synth(np.float64, n ** 4, 4).reshape((n,) * 4, order="F")
    rng = gb.FloatRange(-63.0, 2.0, 64, 63.0)  # -1.0 : 2/63 : 1.0
    G = gb.InterpGrid(A, -1.0 if bc == "fill" else dict(periodic=gb.BCperiodic, nearest=gb.BCnearest)[bc], gb.InterpLinear)
    co = oracle.make_coefs(An, bc, "linear", -1.0 if bc == "fill" else None)
    assert same(G.coefs, co)
    _assert_blocked(gb, G)
    CG = gb.CoordInterpGrid((rng,) * 4, G)
    X = torch.empty((nq, 4), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, 15, lo=-1.0 + (-16.0) * 2.0 / 63.0, hi=1.0 + 16.0 * 2.0 / 63.0)
    v = torch.empty(nq, dtype=torch.float64, device="cuda")
    gb.getindex_(v, CG, X)
    idx = torch.arange(0, nq, 500, device="cuda")
    ov, _, _, _ = oracle.evaluate(co, bc, "linear", sub(X, idx), 1, affine=[rng.row()] * 4)
    assert same(ov, sub(v, idx))
    # valgrad through the coordinate map (coordiscale, src/coord.jl:54-56) on the head of the batch
    vv, gg = gb.valgrad_batch(CG, X[:400_000])
    ov, og, _, _ = oracle.evaluate(co, bc, "linear", X[:400_000].cpu().numpy(), 3, affine=[rng.row()] * 4)
    assert same(ov, vv.cpu().numpy()) and same(og, gg.cpu().numpy())


def _blocks(L, w):
    """(offset, length) of low-corner, centre and high-corner sub-blocks along one dimension: even offsets,
    lengths with the parity of L (so a block touching the upper edge applies the same edge rule)."""
    w = min(w, L)
    w -= (w - L) % 2
    mid = ((L - w) // 2) & ~1
    hi = L - w  # even: L and w share parity
    return [(0, w), (mid, w), (hi, w)]


def test_full_size_c5_chain(gb, oracle):
    """C5: restrict(A, [true,true,true]) 1024 -> 513 -> 257 -> 129 -> 65 -> 33 -> 17 and prolong back up, f64.
    Every level is checked against the oracle: small levels in full, big ones on sub-blocks."""
    import torch

    n = 1024
    A = gb.jl_empty((n, n, n), torch.float64)
    gb.synth_uniform(A, 5)
    sizes = [n]
    while sizes[-1] > 17:
        sizes.append(gb.restrict_size(sizes[-1]))
    assert sizes == [1024, 513, 257, 129, 65, 33, 17]
    levels = [A]
    for _ in sizes[1:]:
        levels.append(gb.restrict(levels[-1], [True, True, True]))
    for k in range(len(sizes) - 1):
        fine, coarse, L = levels[k], levels[k + 1], sizes[k]
        assert tuple(coarse.shape) == (sizes[k + 1],) * 3
        if L <= 129:
            assert same(oracle.restrict_flags(gb.jl_to_numpy(fine), [True] * 3), gb.jl_to_numpy(coarse)), f"restrict level {L}"
            continue
        for (o, w) in _blocks(L, 66):
            f = np.asfortranarray(fine[o:o + w, o:o + w, o:o + w].cpu().numpy())
            r = oracle.restrict_flags(f, [True] * 3)
            m = r.shape[0]
            lo = 0 if o == 0 else 1              # the block's own edge rows are edge-rule artefacts unless it touches the array edge
            hi = m if o + w == L else m - 1
            got = coarse[o // 2 + lo:o // 2 + hi, o // 2 + lo:o // 2 + hi, o // 2 + lo:o // 2 + hi].cpu().numpy()
            assert same(np.ascontiguousarray(r[lo:hi, lo:hi, lo:hi]), np.ascontiguousarray(got)), f"restrict level {L} block {o}"
    cur = levels[-1]
    for k in range(len(sizes) - 2, -1, -1):
        s, c = sizes[k], sizes[k + 1]
        up = gb.prolong(cur, [s, s, s])
        assert tuple(up.shape) == (s, s, s)
        if s <= 129:
            assert same(oracle.prolong_sizes(gb.jl_to_numpy(cur), [s] * 3), gb.jl_to_numpy(up)), f"prolong to {s}"
        else:
            for c0 in (0, (c // 2) - 16, c - 34):
                wlen = 34
                blk = np.asfortranarray(cur[c0:c0 + wlen, c0:c0 + wlen, c0:c0 + wlen].cpu().numpy())
                t = 2 * wlen - 1 if s % 2 else 2 * (wlen - 1)
                p = oracle.prolong_sizes(blk, [t] * 3)
                got = up[2 * c0:2 * c0 + t, 2 * c0:2 * c0 + t, 2 * c0:2 * c0 + t].cpu().numpy()
                assert same(np.ascontiguousarray(p), np.ascontiguousarray(got)), f"prolong to {s} block {c0}"
        del cur
        cur = up
