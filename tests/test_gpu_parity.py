"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  EXACT mode must be BIT-IDENTICAL (values, gradients, Hessians, NaN placement,
first invalid index); FAST mode is held to a stated tolerance.  Run with `-m gpu` on a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BCS = ["nil", "nan", "na", "reflect", "periodic", "nearest", "fill"]
ORDERS = ["nearest", "linear", "quadratic", "cubic"]
V, G, H = 1, 2, 4


def tags(gb):
    bc = dict(nil=gb.BCnil, nan=gb.BCnan, na=gb.BCna, reflect=gb.BCreflect, periodic=gb.BCperiodic, nearest=gb.BCnearest, fill=gb.BCfill)
    it = dict(nearest=gb.InterpNearest, linear=gb.InterpLinear, quadratic=gb.InterpQuadratic, cubic=gb.InterpCubic)
    return bc, it


def supported(bc, order):
    return not (order == "cubic" and bc not in ("nil", "nan", "na"))


def same(a, b):
    """bitwise equality with NaN == NaN"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def make_queries(rng, shape, nq, dtype, wide=True):
    """uniform in a window wider than the grid + edge vectors (integers, half-integers, ends)."""
    N = len(shape)
    lo = np.array([-1.5 * n if wide else 1.0 for n in shape])
    hi = np.array([2.5 * n if wide else float(n) for n in shape])
    X = lo + (hi - lo) * rng.random((nq, N))
    edges = []
    for d, n in enumerate(shape):
        for v in (0.0, 0.5, 1.0, 1.5, 2.0, 2.5, n - 1.0, n - 0.5, n, n + 0.5, n + 1.0, -0.5, -1.0, 1.0 - 1e-9, n + 1e-9, n / 2, 3.0, -n, 2 * n, 2 * n + 0.5):
            x = 1.0 + (np.array(shape) - 1.0) * rng.random(N)
            x[d] = v
            edges.append(x)
    X = np.concatenate([X, np.array(edges)])
    return np.ascontiguousarray(X.astype(dtype))


def build(gb, oracle, A, bc, order, fill=-1.25):
    BC, IT = tags(gb)
    if bc == "fill":
        g = gb.InterpGrid(A, fill, IT[order])
        co = oracle.make_coefs(A, "fill", order, fill)
    else:
        g = gb.InterpGrid(A, BC[bc], IT[order])
        co = oracle.make_coefs(A, bc, order)
    return g, co


def eval_both(gb, oracle, g, co, bc, order, X, mask, affine=None, fast=False):
    rows = None if affine is None else [r.row() for r in affine]
    ov, og, oh, obad = oracle.evaluate(co, bc, order, X, mask, affine=rows, raise_invalid=False)
    try:
        v, gr, h = g._eval(X, mask, fast=fast, affine=rows)
        bad = -1
    except gb.ErrorException as e:
        assert bc == "nil" and "Invalid location" in str(e)
        bad = int(str(e).split("query")[1].strip(" )"))
        # outputs are still produced for inspection through a nan grid of the same data
        v = gr = h = None
    return (ov, og, oh, obad), (v, gr, h, bad)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(9,), (7, 6), (6, 5, 7), (5, 6, 4, 5)])
@pytest.mark.parametrize("bc", BCS)
@pytest.mark.parametrize("order", ORDERS)
def test_eval_exact_bitwise(gb, oracle, dtype, shape, bc, order):
    if not supported(bc, order):
        pytest.skip("no method in the reference")
    rng = np.random.default_rng(abs(hash((str(dtype), shape, bc, order))) % 2 ** 32)
    A = rng.standard_normal(shape).astype(dtype)
    g, co = build(gb, oracle, A, bc, order)
    assert same(g.coefs, co), "prefiltered coefficients differ from the oracle"
    X = make_queries(rng, shape, 400, dtype)
    masks = [V, V | G] + ([V | G | H] if order in ("quadratic", "cubic") else [])
    for mask in masks:
        (ov, og, oh, obad), (v, gr, h, bad) = eval_both(gb, oracle, g, co, bc, order, X, mask)
        assert bad == obad
        if obad >= 0:
            # BCnil raises at the first invalid query (src/interp.jl:402-404): same index as the
            # oracle; then compare values on the valid subset (the oracle marks invalid rows NaN)
            assert bc == "nil"
            Xv = np.ascontiguousarray(X[~np.isnan(ov)])
            (ov, og, oh, obad), (v, gr, h, bad) = eval_both(gb, oracle, g, co, bc, order, Xv, mask)
            assert obad == -1 and bad == -1
        assert same(v, ov)
        if mask & G:
            assert same(gr, og)
        if mask & H:
            assert same(h, oh)


@pytest.mark.parametrize("bc,order,shape", [("reflect", "quadratic", (40, 30)), ("nan", "cubic", (12, 11, 13)),
                                            ("periodic", "linear", (6, 7, 5, 6)), ("fill", "quadratic", (9, 8)),
                                            ("nearest", "linear", (8, 9, 7))])
def test_query_dtype_and_layouts(gb, oracle, bc, order, shape):
    """Float32 queries on Float64 grids and vice versa (the +1 of setx happens in the query's type,
    src/interp.jl:97-139), AoS vs tuple-of-vectors, host vs device residency."""
    import torch

    rng = np.random.default_rng(5)
    for gdt in (np.float64, np.float32):
        A = rng.random(shape).astype(gdt)
        g, co = build(gb, oracle, A, bc, order)
        for xdt in (np.float64, np.float32):
            X = make_queries(rng, shape, 300, xdt, wide=(bc not in ("nan",)))
            ov, og, _, _ = oracle.evaluate(co, bc, order, X, V | G, raise_invalid=False)
            v, gr, _ = g._eval(X, V | G)
            assert same(v, ov) and same(gr, og)
            cols = tuple(np.ascontiguousarray(X[:, d]) for d in range(len(shape)))
            v2, gr2, _ = g._eval(cols, V | G)
            assert same(v2, ov) and same(gr2, og)
            Xd = torch.from_numpy(X).cuda()
            v3, gr3 = gb.valgrad_batch(g, Xd)
            assert same(v3.cpu().numpy(), ov) and same(gr3.cpu().numpy(), og)
            v4, gr4 = gb.valgrad_batch(g, tuple(Xd[:, d].contiguous() for d in range(len(shape))))
            assert same(v4.cpu().numpy(), ov) and same(gr4.cpu().numpy(), og)


def test_host_chunking_matches_single_launch(gb, oracle, monkeypatch):
    """The host path streams chunks over several streams; chunk boundaries must not show."""
    rng = np.random.default_rng(11)
    A = rng.random((33, 29))
    g, co = build(gb, oracle, A, "reflect", "quadratic")
    X = make_queries(rng, A.shape, 3_000_000, np.float64)[:2_500_001]
    v, gr, h = g._eval(X, V | G | H)
    ov, og, oh, _ = oracle.evaluate(co, "reflect", "quadratic", X, V | G | H)
    assert same(v, ov) and same(gr, og) and same(h, oh)


@pytest.mark.parametrize("bc", ["nan", "reflect", "periodic", "nearest", "fill"])
@pytest.mark.parametrize("kind", ["float", "step", "unit"])
def test_coord_grid_bitwise(gb, oracle, bc, kind):
    """CoordInterpGrid: coordlookup in Float64, gradient scaled by coordiscale (src/coord.jl:30-69)."""
    BC, IT = tags(gb)
    rng = np.random.default_rng(3)
    shape = (21, 17)
    A = rng.random(shape)
    if kind == "float":
        ranges = (gb.frange(-1.0, 0.1, 1.0), gb.frange(2.0, 0.5, 10.0))
        lo, hi = np.array([-1.4, 1.0]), np.array([1.4, 11.0])
    elif kind == "step":
        ranges = (gb.StepRange(3, 2, 43), gb.StepRange(-5, 3, 43))
        lo, hi = np.array([0.0, -10.0]), np.array([46.0, 47.0])
    else:
        ranges = (gb.UnitRange(5, 25), gb.UnitRange(-3, 13))
        lo, hi = np.array([3.0, -5.0]), np.array([27.0, 15.0])
    for order in ("linear", "quadratic"):
        g, co = build(gb, oracle, A, bc, order, fill=0.5)
        cg = gb.CoordInterpGrid(ranges, g)
        for xdt in (np.float64, np.float32):
            X = (lo + (hi - lo) * rng.random((500, 2))).astype(xdt)
            rows = [r.row() for r in ranges]
            ov, og, _, obad = oracle.evaluate(co, bc, order, X, V | G, affine=rows, raise_invalid=False)
            assert obad == -1
            v, gr = gb.valgrad_batch(cg, X)
            assert same(v, ov) and same(gr, og)
            assert same(gb.getindex_batch(cg, X), ov)
    with pytest.raises(gb.DimensionMismatch):
        gb.CoordInterpGrid((ranges[1], ranges[0]), A, BC["nan"], IT["linear"])
    with pytest.raises(gb.MethodError):
        gb.valgradhess_batch(cg, X)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order", [("nan", "quadratic"), ("reflect", "quadratic"), ("periodic", "quadratic"),
                                      ("nearest", "quadratic"), ("nan", "cubic"), ("nil", "cubic")])
@pytest.mark.parametrize("shape", [(5,), (300,), (17, 9), (6, 40, 7), (5, 6, 7, 8)])
def test_prefilter_bitwise(gb, oracle, dtype, bc, order, shape):
    BC, IT = tags(gb)
    if order == "cubic" and min(shape) < 4:
        pytest.skip("too small")
    rng = np.random.default_rng(len(shape) * 100 + shape[0])
    A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
    ref = oracle.invert(A, bc, order)
    B = A.copy(order="F")
    out = gb.interp_invert_(B, BC[bc], IT[order])
    assert out is B and same(B, ref)
    # dimlist subsets, in the given order (src/interp.jl:916)
    if len(shape) >= 2:
        dl = [2, 1]
        ref2 = oracle.invert(A, bc, order, dimlist=[d - 1 for d in dl])
        B2 = A.copy(order="F")
        gb.interp_invert_(B2, BC[bc], IT[order], dl)
        assert same(B2, ref2)
    # device-resident, in place
    Bd = gb.jl_from_numpy(A)
    gb.interp_invert_(Bd, BC[bc], IT[order])
    assert same(gb.jl_to_numpy(Bd), ref)
    # Nearest / Linear are no-ops (:909-912)
    C_ = A.copy(order="F")
    gb.interp_invert_(C_, BC[bc], IT["linear"])
    assert same(C_, A)


def test_pad1_bitwise(gb, oracle):
    rng = np.random.default_rng(1)
    for shape in [(5,), (4, 3), (3, 4, 5), (2, 3, 4, 3)]:
        for dt in (np.float64, np.float32):
            A = rng.random(shape).astype(dt)
            assert same(gb.pad1(A, 0.75), oracle.pad1(A, 0.75))
            assert same(gb.jl_to_numpy(gb.pad1(gb.jl_from_numpy(A), -2.0)), oracle.pad1(A, -2.0))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(9,), (8,), (1,), (2,), (9, 6), (7, 8, 9), (10, 9, 8), (33, 34, 35), (4, 5, 6, 7)])
def test_restrict_prolong_bitwise(gb, oracle, dtype, shape):
    rng = np.random.default_rng(sum(shape))
    A = rng.standard_normal(shape).astype(dtype)
    for dim in range(len(shape)):
        for scale in (1.0, 0.5, 1.0 / 3.0):
            assert same(gb.restrict(A, dim + 1, scale), oracle.restrict(A, dim, scale))
        n = shape[dim]
        for ln in (2 * n - 1, 2 * (n - 1)):
            if ln >= 1 and n >= 2:
                assert same(gb.prolong(A, dim + 1, ln), oracle.prolong(A, dim, ln))
        with pytest.raises(gb.ErrorException):
            gb.prolong(A, dim + 1, 2 * n + 1)
    Ad = gb.jl_from_numpy(A)
    assert same(gb.jl_to_numpy(gb.restrict(Ad, 1)), oracle.restrict(A, 0))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(9, 9, 9), (8, 8, 8), (10, 7, 33), (65, 34, 17), (5, 1, 2), (130, 9, 12), (259, 21, 40), (256, 22, 41), (64, 64, 64),
                                   (127, 5, 6), (2, 2, 2), (3, 2, 5)])
def test_fused_3d_equals_staged(gb, oracle, dtype, shape):
    """gb_restrict3 / gb_prolong3 (marching kernels: several 64-wide tiles, several plane chunks, odd and even extents,
    edge tiles) are bit-identical to the dim-by-dim reference (mapdim)."""
    rng = np.random.default_rng(sum(shape))
    A = rng.standard_normal(shape).astype(dtype)
    flags = [True, True, True]
    ref = oracle.restrict_flags(A, flags, 1.0)
    assert same(gb.restrict(A, flags), ref)
    assert same(gb.restrict(A, flags, fused=False), ref)
    assert same(gb.restrict(A, flags, 0.7), oracle.restrict_flags(A, flags, 0.7))
    assert same(gb.jl_to_numpy(gb.restrict(gb.jl_from_numpy(A), flags)), ref)
    two = np.array([[1, 1], [1, 1], [1, 1]], dtype=bool)
    assert same(gb.restrict(A, two), oracle.restrict_flags(A, two, 1.0))
    if min(shape) >= 2:
        for tgt in ([2 * s - 1 for s in shape], [2 * (s - 1) for s in shape], [2 * shape[0] - 1, shape[1], 2 * (shape[2] - 1)]):
            if min(tgt) < 1:
                continue
            refp = oracle.prolong_sizes(A, tgt)
            assert same(gb.prolong(A, tgt), refp)
            assert same(gb.prolong(A, tgt, fused=False), refp)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(9, 9), (8, 8), (130, 21), (259, 40), (256, 41), (2, 2), (3, 2), (64, 64), (127, 6), (70, 9, 5), (65, 34, 3), (8, 8, 1),
                                   (4, 600001), (3, 300000)])  # (tall arrays: more row tiles than gridDim.y allows)
def test_fused_dims12_equals_staged(gb, oracle, dtype, shape):
    """gb_restrict12 / gb_prolong12 (the marching kernels without their dim-3 stage: 2-D arrays -- image pyramids -- and
    flags [true,true,false] on 3-D arrays) are bit-identical to the per-dimension reference, host and device arrays."""
    rng = np.random.default_rng(sum(shape) + 1)
    A = rng.standard_normal(shape).astype(dtype)
    flags = [True, True] + [False] * (len(shape) - 2)
    for scale in (1.0, 0.7):
        ref = oracle.restrict_flags(A, flags, scale)
        assert same(gb.restrict(A, flags, scale), ref)
        assert same(gb.restrict(A, flags, scale, fused=False), ref)
    assert same(gb.jl_to_numpy(gb.restrict(gb.jl_from_numpy(A), flags)), oracle.restrict_flags(A, flags, 1.0))
    if len(shape) == 2:
        for tgt in ([2 * s - 1 for s in shape], [2 * (s - 1) for s in shape], [2 * shape[0] - 1, 2 * (shape[1] - 1)]):
            if min(tgt) < 1 or min(shape) < 2:
                continue
            refp = oracle.prolong_sizes(A, tgt)
            assert same(gb.prolong(A, tgt), refp)
            assert same(gb.prolong(A, tgt, fused=False), refp)
            assert same(gb.jl_to_numpy(gb.prolong(gb.jl_from_numpy(A), tgt)), refp)
        with pytest.raises(gb.ErrorException):
            gb.prolong(A, [2 * shape[0] + 1, 2 * shape[1] - 1])
    else:  # [m1, m2, n3]: every plane prolonged along dims 1 and 2
        for tgt in ([2 * shape[0] - 1, 2 * shape[1] - 1, shape[2]], [2 * (shape[0] - 1), 2 * shape[1] - 1, shape[2]]):
            refp = oracle.prolong_sizes(A, tgt)
            assert same(gb.prolong(A, tgt), refp)
            assert same(gb.prolong(A, tgt, fused=False), refp)


def test_restrict_prolong_chain_roundtrip(gb, oracle):
    """C5 in miniature: 66^3 -> 34 -> 18 -> 10 and back up the same size list."""
    rng = np.random.default_rng(0)
    A = rng.random((66, 65, 64))
    flags = np.ones((3, 3), dtype=bool)
    sched = gb.restrict_size(A.shape, flags)
    down = gb.restrict(A, flags)
    assert down.shape == tuple(sched[:, -1])
    assert same(down, oracle.restrict_flags(A, flags))
    up_sizes = gb.prolong_size(A.shape, flags)
    up = gb.prolong(down, up_sizes)
    assert up.shape == A.shape
    assert same(up, oracle.prolong_sizes(down, up_sizes))


EPS64 = 2.220446049250313e-16


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order,shape", [("nan", "cubic", (20, 21, 22)), ("reflect", "quadratic", (50, 60)),
                                            ("periodic", "linear", (8, 9, 7, 8)), ("fill", "quadratic", (12, 11, 10)),
                                            ("nearest", "nearest", (9, 9)), ("nan", "cubic", (30,)), ("nan", "cubic", (7, 8, 6, 7))])
def test_fast_mode_tolerance(gb, oracle, dtype, bc, order, shape):
    """GB_FAST (separable FMA contraction): same taps, weights to 1 ulp, different rounding order.
    north_star's tolerance -- 4 ulp (Float64) / 1e-5 relative (Float32) -- is taken relative to the
    magnitude scale S = sum_i |c_i a_i| of each output (a cancelling sum has no other meaningful
    ulp), S from the oracle's extended-precision evaluator.  FAST is held to it against the
    long-double truth, and to twice it against the reference arithmetic (EXACT, bit-identical to
    the oracle), which itself sits up to ~4.6 ulp(S) from the truth."""
    rng = np.random.default_rng(9)
    A = rng.random(shape).astype(dtype)
    g, co = build(gb, oracle, A, bc, order)
    X = make_queries(rng, shape, 2000, dtype, wide=(bc not in ("nan", "nil")))
    e = g._eval(X, V | G)
    f = g._eval(X, V | G, fast=True)
    tv, tg, sv, sg = oracle.evaluate_truth(co.astype(np.float64), bc, order, X.astype(np.float64))
    tol = 4 * EPS64 if dtype == np.float64 else 1e-5
    for a, b, t, s in ((e[0], f[0], tv, sv), (e[1], f[1], tg, sg)):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(a) & ~np.isnan(s)
        # FAST vs the reference arithmetic: 2 x tol, because the reference's own rounding noise reaches
        # 4.6 ulp(S) on large batches (profiles/README.md) while FAST stays within 2.7 ulp(S) of the truth
        assert np.all(np.abs(a[ok].astype(np.float64) - b[ok]) <= 2 * tol * s[ok])
        if dtype == np.float64:  # (a Float32 grid rounds the coordinates differently from the double truth)
            assert np.array_equal(np.isnan(a), np.isnan(t))
            assert np.all(np.abs(b[ok] - t[ok]) <= tol * s[ok])                  # FAST vs truth
    if order in ("quadratic", "cubic"):  # Hessians: bounded through the coefficient magnitude
        eh, fh = g._eval(X, V | G | H)[2], g._eval(X, V | G | H, fast=True)[2]
        ok = ~np.isnan(eh)
        assert np.array_equal(np.isnan(eh), np.isnan(fh))
        assert np.max(np.abs(eh[ok].astype(np.float64) - fh[ok]), initial=0.0) <= tol * np.max(np.abs(co)) * 4.0 ** len(shape)


def test_nan_under_zero_weight_does_not_poison(gb, oracle):
    """c == 0 ? 0 : c*A[i] (src/interp.jl:504): a NaN sample under a zero weight is never read (Q6)."""
    A = np.arange(1.0, 8.0)
    A[4] = np.nan
    g, co = build(gb, oracle, A, "nan", "linear")
    X = np.array([[1.0], [2.0], [3.0], [4.0], [3.5], [6.0], [7.0]])
    for fast in (False, True):
        v, gr, _ = g._eval(X, V | G, fast=fast)
        ov, og, _, _ = oracle.evaluate(co, "nan", "linear", X, V | G)
        assert same(v, ov) and same(gr, og)
        assert v[3] == 4.0 and not np.isnan(v[:4]).any()


def test_bcnil_reports_first_invalid(gb, oracle):
    A = np.arange(1.0, 9.0)
    g = gb.InterpGrid(A, gb.BCnil, gb.InterpQuadratic)
    X = np.full((5000, 1), 3.3)
    X[4321, 0] = -0.8
    X[4800, 0] = 100.0
    with pytest.raises(gb.ErrorException, match=r"Invalid location \(query 4321\)"):
        gb.getindex_batch(g, X)
    import torch

    with pytest.raises(gb.ErrorException, match=r"Invalid location \(query 4321\)"):
        gb.getindex_batch(g, torch.from_numpy(X).cuda())


def test_golden_fixtures(gb):
    """Committed oracle outputs (tests/golden/make_golden.py) -- no oracle needed at run time."""
    import glob
    import os

    BC, IT = tags(gb)
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eval_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        bc, order = str(z["bc"]), str(z["order"])
        A = z["A"]
        g = gb.InterpGrid(A, float(z["fill"]), IT[order]) if bc == "fill" else gb.InterpGrid(A, BC[bc], IT[order])
        assert same(g.coefs, z["coefs"])
        mask = int(z["mask"])
        v, gr, h = g._eval(z["X"], mask)
        assert same(v, z["val"])
        if mask & G:
            assert same(gr, z["grad"])
        if mask & H:
            assert same(h, z["hess"])
    for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mg_*.npz"))):
        z = np.load(f)
        assert same(gb.restrict(z["A"], [True] * z["A"].ndim), z["R"])
        assert same(gb.prolong(z["R"], list(z["A"].shape)), z["P"])
    # widening rows
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mgb_*.npz")))
    assert len(files) == 4
    for f in files:
        z = np.load(f)
        A = z["A"]
        for d in range(A.ndim):
            assert same(gb.restrictb(A, d + 1), z[f"restrictb_{d}"])
            assert same(gb.restrict_extrap(A, d + 1), z[f"extrap_{d}"])
            assert same(gb.prolongb(z[f"restrictb_{d}"], d + 1, A.shape[d]), z[f"prolongb_{d}"])
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "channels_quadratic_reflect_2d.npz"))
    g = gb.InterpGridChannels(np.asfortranarray(z["A"]), gb.BCreflect, gb.InterpQuadratic)
    assert same(g.coefs, np.asfortranarray(z["coefs"]))
    v, gr, _ = g._eval(z["X"], V | G)
    assert same(v, z["val"]) and same(gr, z["grad"])
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "irregular_1d.npz"))
    for it, IT in (("forward", gb.InterpForward), ("backward", gb.InterpBackward), ("nearest", gb.InterpNearest), ("linear", gb.InterpLinear)):
        assert same(gb.InterpIrregular(z["knots"], z["samples"], gb.BCnan, IT)[z["x"]], z[f"nan_{it}"])
        assert same(gb.InterpIrregular(z["knots"], z["samples"], gb.BCnearest, IT)[z["x"]], z[f"nearest_{it}"])
        assert same(gb.InterpIrregular(z["knots"], z["samples"], -7.5, IT)[z["x"]], z[f"fill_{it}"])


def test_full_size_properties_c2(gb):
    """BASELINE config 2 at full size (256^3 f64, cubic, BCnan, 1e7 queries), size-independent
    properties: on-grid reproduction, NaN exactly outside [1,n], value pass identical in
    getindex / valgrad, linearity of the whole pipeline in the samples, FAST within tolerance."""
    import torch

    n = 256
    A = gb.jl_empty((n, n, n), torch.float64)
    gb.synth_uniform(A, seed=2)
    g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    assert g.shape == (n, n, n) and g.stored_shape == (n + 2,) * 3
    # on-grid reproduction (test/grid.jl:50-98 at scale)
    idx = torch.randint(1, n + 1, (200_000, 3), device="cuda")
    v = gb.getindex_batch(g, idx.to(torch.float64))
    a = A[idx[:, 0] - 1, idx[:, 1] - 1, idx[:, 2] - 1]
    assert float((v - a).abs().max()) < 1.5e-8
    # 1e7 uniform queries in [1,n]^3: finite; outside: NaN in value and every gradient component
    X = torch.empty((10_000_000, 3), dtype=torch.float64, device="cuda")
    gb.synth_uniform(X, seed=12, lo=1.0, hi=float(n))
    val, grad = gb.valgrad_batch(g, X)
    assert bool(torch.isfinite(val).all()) and bool(torch.isfinite(grad).all())
    v_only = gb.getindex_batch(g, X)
    assert torch.equal(v_only, val)
    # FAST within north_star's 4 ulp of the magnitude scale S = sum|c_i a_i| (extended-precision oracle)
    # on a 200k-query sample, against the reference arithmetic and against the truth
    vf, gf = gb.valgrad_batch(g, X, fast=True)
    ns = 200_000
    from oracle import oracle as O

    tv, tg, sv, sg = O.evaluate_truth(g.coefs, "nan", "cubic", X[:ns].cpu().numpy())
    ve, ge, vfn, gfn = val[:ns].cpu().numpy(), grad[:ns].cpu().numpy(), vf[:ns].cpu().numpy(), gf[:ns].cpu().numpy()
    # measured on 1e6 queries (profiles/README.md): FAST vs truth max 2.6 / 2.3 ulp(S) (value / gradient), the
    # reference arithmetic (EXACT) vs truth max 4.6 / 3.7 -- so FAST is held to 4 ulp(S) of the TRUTH and
    # to 8 ulp(S) of the reference arithmetic, whose own rounding noise exceeds 4 for a few queries in a million
    assert np.all(np.abs(vfn - tv) <= 4 * EPS64 * sv) and np.all(np.abs(gfn - tg) <= 4 * EPS64 * sg)
    assert np.all(np.abs(vfn - ve) <= 8 * EPS64 * sv) and np.all(np.abs(gfn - ge) <= 8 * EPS64 * sg)
    assert np.all(np.abs(ve - tv) <= 8 * EPS64 * sv) and np.all(np.abs(ge - tg) <= 8 * EPS64 * sg)
    Xo = X[:100_000].clone()
    Xo[::2, 1] = n + 0.25
    Xo[1::2, 2] = 0.75
    vo, go = gb.valgrad_batch(g, Xo)
    assert bool(torch.isnan(vo).all()) and bool(torch.isnan(go).all())
    # linearity: coefficients of (2A) == 2*coefficients(A) exactly (power-of-two scaling)
    g2 = gb.InterpGrid(2.0 * A, gb.BCnan, gb.InterpCubic)
    v2 = gb.getindex_batch(g2, X[:1_000_000])
    assert torch.equal(v2, 2.0 * val[:1_000_000])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order,shape", [("nan", "cubic", (40, 37, 35)), ("reflect", "quadratic", (300, 200)), ("periodic", "quadratic", (70, 80)),
                                            ("periodic", "linear", (9, 20, 17, 12)), ("fill", "linear", (30, 30, 30)), ("nearest", "quadratic", (50, 3, 40)),
                                            ("reflect", "linear", (33, 65)), ("nil", "cubic", (36, 5, 5, 6)), ("reflect", "nearest", (64, 64)),
                                            ("nan", "quadratic", (100,))])
def test_tiled_path_is_bit_identical(gb, oracle, dtype, bc, order, shape):
    """The binned shared-memory path (GB_TILED) only reschedules work: bit-identical to the plain
    path and to the oracle, including wrapped taps at the boundaries, invalid queries and NaN."""
    import torch

    rng = np.random.default_rng(17)
    A = rng.standard_normal(shape).astype(dtype)
    g, co = build(gb, oracle, A, bc, order)
    X = make_queries(rng, shape, 60_000, dtype, wide=(bc != "nil"))
    mask = V | G | (H if order in ("quadratic", "cubic") else 0)
    ov, og, oh, obad = oracle.evaluate(co, bc, order, X, mask, raise_invalid=False)
    if bc == "nil":
        X = np.ascontiguousarray(X[~np.isnan(ov)])
        ov, og, oh, obad = oracle.evaluate(co, bc, order, X, mask)
    Xd = torch.from_numpy(X).cuda()
    for fast in (False, True):
        outs = {}
        for tiled in (False, True):
            v, gr, h = g._eval(Xd, mask, fast=fast, tiled=tiled)
            outs[tiled] = [None if a is None else a.cpu().numpy() for a in (v, gr, h)]
        for a, b in zip(outs[False], outs[True]):
            assert (a is None and b is None) or same(a, b)
        if not fast:
            assert same(outs[True][0], ov) and same(outs[True][1], og) and (oh is None or same(outs[True][2], oh))


@pytest.mark.parametrize("bc,order", [("nil", "linear"), ("reflect", "quadratic"), ("nan", "cubic"), ("periodic", "nearest"), ("fill", "linear")])
def test_generic_nd_kernel(gb, oracle, bc, order):
    """N >= 5 (and 4-D Quadratic/Cubic) run on the rolled generic kernel (src/interp.jl:1189-1283);
    test/grid.jl:260-264 is the reference's own 5-D smoke test."""
    rng = np.random.default_rng(23)
    B = rng.random((4, 4, 4, 4, 4))
    Bi = gb.InterpGrid(B, gb.BCnil, gb.InterpLinear)
    assert Bi[2, 2, 2, 2, 2] == B[1, 1, 1, 1, 1]
    for shape in [(5, 4, 6, 5, 4), (4, 5, 4, 4, 5, 4)]:
        A = rng.standard_normal(shape)
        g, co = build(gb, oracle, A, bc, order)
        assert same(g.coefs, co)
        X = make_queries(rng, shape, 300, np.float64, wide=(bc != "nil"))
        mask = V | G | (H if order in ("quadratic", "cubic") else 0)
        ov, og, oh, obad = oracle.evaluate(co, bc, order, X, mask, raise_invalid=False)
        if bc == "nil":
            X = np.ascontiguousarray(X[~np.isnan(ov)])
            ov, og, oh, obad = oracle.evaluate(co, bc, order, X, mask)
        v, gr, h = g._eval(X, mask)
        assert same(v, ov) and same(gr, og) and (oh is None or same(h, oh))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,world", [((9, 8, 20), 2), ((8, 9, 21), 3), ((12, 6, 64), 4), ((5, 7, 33), 8), ((130, 21, 95), 3), ((66, 40, 128), 5)])
def test_slab_kernels_assemble_to_unsharded(gb, oracle, dtype, shape, world):
    """gb_restrict_slab / gb_prolong_slab and the fused gb_restrict3_slab / gb_prolong3_slab: every rank's slab
    (own planes + halo, cut straight out of the global array here instead of exchanged) reproduces its planes of the
    unsharded result, odd and even plane offsets included."""
    import grid_jl_b200.dist as D

    eng = D.GpuEngine()
    A = oracle.synth(dtype, int(np.prod(shape)), seed=31).reshape(shape, order="F")
    L3 = shape[2]
    n3 = D.restrict_size(L3)
    ref = oracle.restrict_flags(A, [True, True, True])
    refp = oracle.prolong_sizes(ref, list(shape))
    rplan, pplan = D.SlabPlan("restrict", L3, n3, world), D.SlabPlan("prolong", n3, L3, world)
    for r in range(world):
        e0, e1 = rplan.ext_range(r)
        k0, k1 = rplan.out[r]
        r12 = eng.restrict_dims12(gb.jl_from_numpy(np.asfortranarray(A[:, :, e0:e1])), 1.0)
        mine = eng.restrict_slab(r12, 1.0, L3, e0, k0, k1 - k0)
        assert same(gb.jl_to_numpy(mine), ref[:, :, k0:k1])
        e0, e1 = pplan.ext_range(r)
        k0, k1 = pplan.out[r]
        p12 = eng.prolong_dims12(gb.jl_from_numpy(np.asfortranarray(ref[:, :, e0:e1])), shape[0], shape[1])
        up = eng.prolong_slab(p12, n3, L3, e0, k0, k1 - k0)
        assert same(gb.jl_to_numpy(up), refp[:, :, k0:k1])
        # the fused slab operators (marching kernels with plane offsets) on the same slabs
        e0, e1 = rplan.ext_range(r)
        k0, k1 = rplan.out[r]
        mine = eng.restrict3_slab(gb.jl_from_numpy(np.asfortranarray(A[:, :, e0:e1])), 1.0, L3, e0, k0, k1 - k0)
        assert same(gb.jl_to_numpy(mine), ref[:, :, k0:k1])
        e0, e1 = pplan.ext_range(r)
        k0, k1 = pplan.out[r]
        up = eng.prolong3_slab(gb.jl_from_numpy(np.asfortranarray(ref[:, :, e0:e1])), shape, n3, e0, k0, k1 - k0)
        assert same(gb.jl_to_numpy(up), refp[:, :, k0:k1])
    # a slab that lacks a needed plane is refused
    with pytest.raises(gb.ErrorException):
        eng.restrict_slab(gb.jl_from_numpy(np.asfortranarray(A[:, :, 4:8])), 1.0, L3, 4, 1, 3)
    with pytest.raises(gb.ErrorException):
        eng.restrict3_slab(gb.jl_from_numpy(np.asfortranarray(A[:, :, 4:8])), 1.0, L3, 4, 1, 3)
    with pytest.raises(gb.ErrorException):
        eng.prolong3_slab(gb.jl_from_numpy(np.asfortranarray(ref[:, :, 2:4])), shape, n3, 2, 1, 6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order,shape,nchan", [("reflect", "quadratic", (40, 30), 3), ("nan", "cubic", (10, 10, 10), 7), ("periodic", "linear", (6, 7, 5, 6), 2),
                                                  ("fill", "quadratic", (12, 9), 4), ("nil", "linear", (20,), 5), ("nearest", "nearest", (9, 8, 7), 3)])
def test_multivalued_grid_matches_per_channel_grids(gb, oracle, dtype, bc, order, shape, nchan):
    """(f)-1: multi-valued evaluation (README.md:151-209).  One InterpGridChannels == nchan InterpGrids
    on A[..., c]: coefficients, values, gradients and Hessians bit-identical to the oracle per channel,
    host and device queries."""
    import torch

    rng = np.random.default_rng(41)
    A = np.asfortranarray(rng.standard_normal(shape + (nchan,)).astype(dtype))
    BC, IT = tags(gb)
    fill = -1.25
    g = gb.InterpGridChannels(A, fill if bc == "fill" else BC[bc], IT[order])
    assert g.nchan == nchan and g.shape == shape
    co = g.coefs
    X = make_queries(rng, shape, 1500, dtype, wide=(bc != "nil"))
    mask = V | G | (H if order in ("quadratic", "cubic") else 0)
    if bc == "nil":
        ov0 = oracle.evaluate(oracle.make_coefs(A[..., 0], bc, order), bc, order, X, V, raise_invalid=False)[0]
        X = np.ascontiguousarray(X[~np.isnan(ov0)])
    v, gr, h = g._eval(X, mask)
    vd, grd, hd = g._eval(torch.from_numpy(X).cuda(), mask)
    for c in range(nchan):
        oc = oracle.make_coefs(np.ascontiguousarray(A[..., c]), bc, order, fill if bc == "fill" else None)
        assert same(co[..., c], oc)
        ov, og, oh, _ = oracle.evaluate(oc, bc, order, X, mask, raise_invalid=False)
        assert same(v[:, c], ov) and same(gr[:, c], og) and (oh is None or same(h[:, c], oh))
        assert same(vd.cpu().numpy()[:, c], ov) and same(grd.cpu().numpy()[:, c], og)


def test_multivalued_lowlevel_spectra_example(gb, oracle):
    """README.md:178-209: a 10x10x10 grid of 200-pixel spectra, interp_invert! along dims 1:3, then one
    set_position and interp(ic, grid, index) per pixel -- here all pixels of many positions at once,
    with array index coordinates (raw) like set_position takes."""
    rng = np.random.default_rng(43)
    npix = 200
    grid = np.asfortranarray(rng.random((10, 10, 10, npix)))
    co = grid.copy(order="F")
    gb.interp_invert_(co, gb.BCnil, gb.InterpCubic, dimlist=[1, 2, 3])
    for c in (0, 17, npix - 1):
        assert same(co[..., c], oracle.invert(np.ascontiguousarray(grid[..., c]).copy(order="F"), "nil", "cubic"))
    g = gb.InterpGridChannels.from_coefs(co, gb.BCnil, gb.InterpCubic, raw=True)
    # (the README's own x = 1.5 reads index 0 of the unpadded array under InterpCubic -- Q1 / D2; use 2.5)
    X = np.concatenate([[[2.5, 2.5, 2.5]], 2.0 + 6.9 * rng.random((300, 3))])
    spec = g._eval(X, V)[0]
    assert spec.shape == (301, npix)
    for c in (0, 1, 99, npix - 1):
        # the low-level call on channel c: raw coordinates on the unpadded array == oracle.lowlevel
        for qi in (0, 5, 300):
            ref = oracle.lowlevel(np.ascontiguousarray(co[..., c]), "cubic", "nil", X[qi], hess=False)
            assert spec[qi, c] == ref[0]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(9,), (8,), (2,), (3,), (7, 10), (6, 5, 9), (4, 5, 6, 3)])
def test_balanced_and_extrapolating_variants(gb, oracle, dtype, shape):
    """(f)-2: restrictb / restrict_extrap / prolongb (src/restrict_prolong.jl:187-320) -- the plain operator
    with the reference's edge planes; bit-identical to the oracle's restatement, every dim, host and device."""
    import torch

    rng = np.random.default_rng(47)
    A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
    for d in range(len(shape)):
        for scale in (1.0, 0.5):
            assert same(gb.restrictb(A, d + 1, scale), oracle.restrictb(A, d, scale))
            assert same(gb.restrict_extrap(A, d + 1, scale), oracle.restrict_extrap(A, d, scale))
        R = oracle.restrictb(A, d, 1.0)
        assert same(gb.prolongb(R, d + 1, shape[d]), oracle.prolongb(R, d, shape[d]))
    # device arrays and the mapdim forms
    Ad = gb.jl_from_numpy(A)
    flags = [True] * len(shape)
    ref = A
    for d in range(len(shape)):
        ref = oracle.restrictb(ref, d, 1.0)
    assert same(gb.jl_to_numpy(gb.restrictb(Ad, flags)), ref) and same(gb.restrictb(A, flags), ref)
    up = ref
    for d in range(len(shape)):
        if shape[d] > up.shape[d]:  # :274 -- prolong only where the target is larger
            up = oracle.prolongb(up, d, shape[d])
    assert same(gb.prolongb(ref, list(shape)), up)
    with pytest.raises(gb.ErrorException):
        gb.prolongb(ref, 1, shape[0] + 5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("it", ["forward", "backward", "nearest", "linear"])
@pytest.mark.parametrize("bc", ["nil", "nan", "na", "fill", "nearest"])
def test_interp_irregular(gb, oracle, dtype, it, bc):
    """(f)-4: InterpIrregular (src/interp.jl:318-377): binary search on sorted knots, bit-identical to the
    oracle's restatement for every (BC, IT), including hits on knots, the x == g[1] rule, duplicates,
    out-of-range points, NaN queries and signed zeros."""
    import torch

    rng = np.random.default_rng(53)
    IT = dict(forward=gb.InterpForward, backward=gb.InterpBackward, nearest=gb.InterpNearest, linear=gb.InterpLinear)[it]
    BC = tags(gb)[0]
    knots = np.sort(np.concatenate([rng.uniform(-3, 7, 40), [0.0, -0.0, 2.5, 2.5]])).astype(dtype)
    vals = rng.standard_normal(knots.shape[0]).astype(dtype)
    x = np.concatenate([rng.uniform(-4, 8, 3000), knots.astype(np.float64), [knots[0], knots[-1], -0.0, 0.0, np.nan, np.inf, -np.inf,
                                                                           np.nextafter(knots[0], -np.inf), np.nextafter(knots[-1], np.inf)]]).astype(dtype)
    fill = 4.5
    G = gb.InterpIrregular(knots, vals, fill if bc == "fill" else BC[bc], IT)
    ov, obad = oracle.irregular(knots, vals, bc, it, x, fill=fill, raise_invalid=False)
    if bc == "nil":
        with pytest.raises(gb.BoundsError, match=rf"query {obad}\b"):
            G[x]
        inside = ~np.isnan(ov)
        x = x[inside]
        ov, obad = oracle.irregular(knots, vals, bc, it, x, fill=fill)
    assert same(G[x], ov)
    assert same(G[torch.from_numpy(x).cuda()].cpu().numpy(), ov)
    assert G[float(x[3])] == ov[3]
    with pytest.raises(gb.ErrorException):
        gb.InterpIrregular(knots[::-1].copy(), vals, gb.BCnan, IT)
    with pytest.raises(gb.ErrorException):
        gb.InterpIrregular(knots, vals, gb.BCperiodic, IT)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("bc,order,shape", [("reflect", "quadratic", (30, 20)), ("nan", "cubic", (12, 11, 10)), ("periodic", "linear", (6, 7, 5, 6)),
                                            ("fill", "linear", (9, 8, 7)), ("nearest", "quadratic", (15,)), ("periodic", "nearest", (8, 9)),
                                            ("nil", "quadratic", (10, 12)), ("nan", "quadratic", (5, 6, 4, 5)), ("reflect", "linear", (4, 5, 4, 3, 4))])
def test_tensor_product_getindex(gb, oracle, dtype, bc, order, shape):
    """(f)-3: G[xs, ys, ...] (src/interp.jl:273-306) on the separable path (taps and weights once per axis
    point; N <= 4) and on the expanded path (N = 5, 4-D quadratic): bit-identical to the oracle's outer
    product of scalar calls, host vectors and CUDA tensors, including wrapped / out-of-range axis points."""
    import torch

    rng = np.random.default_rng(59)
    A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
    g, co = build(gb, oracle, A, bc, order)
    og = oracle.OracleGrid(A, bc, order, fill=-1.25 if bc == "fill" else None)
    wide = bc not in ("nil",)
    vecs = []
    for n in shape:
        lo, hi = (-1.5 * n, 2.5 * n) if wide else (1.0, float(n))
        v = np.concatenate([lo + (hi - lo) * rng.random(7), [1.0, float(n), 1.5, n / 2]]).astype(dtype)
        vecs.append(v)
    ref = og.outer(*vecs)
    got = g[tuple(vecs)]
    assert got.shape == ref.shape and same(got, ref.astype(dtype))
    gotd = g[tuple(torch.from_numpy(v).cuda() for v in vecs)]
    assert same(gb.jl_to_numpy(gotd), ref.astype(dtype))
    if bc == "nil":
        bad = [v.copy() for v in vecs]
        bad[-1][3] = shape[-1] + 2.0
        with pytest.raises(gb.ErrorException, match="Invalid location"):
            g[tuple(bad)]


def test_brick_path_in_slices(gb, oracle):
    """Batches above the slice limit are evaluated slice by slice (GB_TILED_BATCH, default 2^26): run a
    child process with a 20000-query limit so that a 70000-query batch takes four slices, with BCnil's
    first-invalid index falling in the third one."""
    import os
    import subprocess
    import sys
    import tempfile

    rng = np.random.default_rng(61)
    A = np.asfortranarray(rng.standard_normal((33, 20, 18)))
    X = make_queries(rng, A.shape, 70_000, np.float64, wide=False)[:70_000]
    X[45_678, 1] = 25.0  # invalid under BCnil
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "A.npy"), A)
        np.save(os.path.join(d, "X.npy"), X)
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
            "import grid_jl_b200 as gb\n"
            f"A = np.load({os.path.join(d, 'A.npy')!r}); X = np.load({os.path.join(d, 'X.npy')!r})\n"
            "g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)\n"
            "v, gr, _ = g._eval(torch.from_numpy(X).cuda(), 3, tiled=True)\n"
            f"np.save({os.path.join(d, 'v.npy')!r}, v.cpu().numpy()); np.save({os.path.join(d, 'g.npy')!r}, gr.cpu().numpy())\n"
            "n = gb.InterpGrid(A, gb.BCnil, gb.InterpQuadratic)\n"
            "try:\n"
            "    n._eval(torch.from_numpy(X).cuda(), 1, tiled=True)\n"
            "    print('no error')\n"
            "except gb.ErrorException as e:\n"
            "    print(str(e))\n"
        )
        env = dict(os.environ, GB_TILED_BATCH="20000")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "Invalid location (query 45678)" in out.stdout
        co = oracle.make_coefs(A, "nan", "cubic")
        ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", X, 3)
        assert same(np.load(os.path.join(d, "v.npy")), ov) and same(np.load(os.path.join(d, "g.npy")), og)


def test_brick_path_global_histogram(gb, oracle):
    """Grids with more tiles than the shared-memory histograms of the counting sort hold (12000; e.g. 512^3) are
    sorted through one global histogram with warp-aggregated atomics.  GB_SORT_SMEM_BINS=8 in a child process
    sends small grids down that route: a 3-D cubic grid (19 bins), a 2-D cubic grid with 13001 bins (two segments
    of the bin scan; GB_TILE_KB=4 keeps the bricks tiny), clustered queries (every lane of a warp in one tile),
    and BCnil's first-invalid index."""
    import os
    import subprocess
    import sys
    import tempfile

    rng = np.random.default_rng(73)
    A3 = np.asfortranarray(rng.standard_normal((33, 20, 18)))
    X3 = make_queries(rng, A3.shape, 70_000, np.float64, wide=False)[:70_000]
    X3[:5000] = 7.0 + rng.random((5000, 3))  # clustered
    X3[45_678, 1] = 25.0  # invalid under BCnil
    A2 = np.asfortranarray(rng.standard_normal((640, 5200)))
    X2 = make_queries(rng, A2.shape, 150_000, np.float64, wide=True)
    with tempfile.TemporaryDirectory() as d:
        np.savez(os.path.join(d, "in.npz"), A3=A3, X3=X3, A2=A2, X2=X2)
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
            "import grid_jl_b200 as gb\n"
            f"z = np.load({os.path.join(d, 'in.npz')!r})\n"
            "g = gb.InterpGrid(np.asfortranarray(z['A3']), gb.BCnan, gb.InterpCubic)\n"
            "X3 = torch.from_numpy(z['X3']).cuda()\n"
            "v, gr, _ = g._eval(X3, 3, tiled=True)\n"
            "vf = g._eval(X3, 1, tiled=True, fast=True)[0]\n"
            "g2 = gb.InterpGrid(np.asfortranarray(z['A2']), gb.BCna, gb.InterpCubic)\n"
            "v2, gr2, h2 = g2._eval(torch.from_numpy(z['X2']).cuda(), 7, tiled=True)\n"
            f"np.savez({os.path.join(d, 'out.npz')!r}, v=v.cpu().numpy(), g=gr.cpu().numpy(), vf=vf.cpu().numpy(), v2=v2.cpu().numpy(), g2=gr2.cpu().numpy(), h2=h2.cpu().numpy())\n"
            "n = gb.InterpGrid(np.asfortranarray(z['A3']), gb.BCnil, gb.InterpQuadratic)\n"
            "try:\n"
            "    n._eval(X3, 1, tiled=True)\n"
            "    print('no error')\n"
            "except gb.ErrorException as e:\n"
            "    print(str(e))\n"
        )
        env = dict(os.environ, GB_SORT_SMEM_BINS="8", GB_TILE_KB="4")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "Invalid location (query 45678)" in out.stdout
        z = np.load(os.path.join(d, "out.npz"))
        co = oracle.make_coefs(A3, "nan", "cubic")
        ov, og, _, _ = oracle.evaluate(co, "nan", "cubic", X3, 3)
        assert same(z["v"], ov) and same(z["g"], og)
        tv, _, sv, _ = oracle.evaluate_truth(co, "nan", "cubic", X3)
        ok = ~np.isnan(ov)
        assert np.all(np.abs(z["vf"][ok] - tv[ok]) <= 4 * np.spacing(sv[ok]))
        co2 = oracle.make_coefs(A2, "na", "cubic")
        ov, og, oh, _ = oracle.evaluate(co2, "na", "cubic", X2, 7)
        assert same(z["v2"], ov) and same(z["g2"], og) and same(z["h2"], oh)


def test_blocked_layout(gb, oracle):
    """Large 2-D to 4-D grids with light neighbourhoods are kept in 4^N-element Morton blocks (a pure addressing
    change, eval.cuh tap_offset).  GB_BLOCK_MIN_MB=0 in a child process puts small grids in that layout: every
    order and BC, all output modes, ragged extents (not multiples of 4), queries exactly on the upper edges
    (the D1 / offset-table taps that carry into the next column), device and host queries, the tensor-product
    call, G.coefs back in the reference layout, and a second grid adopted from those coefficients (the multi-GPU
    broadcast path) -- all bit-identical to the oracle."""
    import os
    import subprocess
    import sys
    import tempfile

    rng = np.random.default_rng(71)
    shapes2 = [(37, 22), (8, 4), (5, 6), (4, 9)]
    shapes3 = [(9, 6, 5), (4, 8, 7), (13, 4, 6), (5, 4, 6, 5)]  # (the 4-D one: Nearest / Linear are blocked, the generic kernel's orders are not)
    cases = [(bc, order) for order in ORDERS for bc in BCS if not (order == "cubic" and bc in ("reflect", "periodic", "nearest", "fill"))]
    with tempfile.TemporaryDirectory() as d:
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
            "import grid_jl_b200 as gb\n"
            "BC = dict(nil=gb.BCnil, nan=gb.BCnan, na=gb.BCna, reflect=gb.BCreflect, periodic=gb.BCperiodic, nearest=gb.BCnearest)\n"
            "IT = dict(nearest=gb.InterpNearest, linear=gb.InterpLinear, quadratic=gb.InterpQuadratic, cubic=gb.InterpCubic)\n"
            f"d = {d!r}\n"
            "for line in open(d + '/cases.txt'):\n"
            "    tag, bc, order, mode = line.split()\n"
            "    z = np.load(f'{d}/in_{tag}.npz'); A = np.asfortranarray(z['A']); X = z['X']\n"
            "    g = gb.InterpGrid(A, -1.25, IT[order]) if bc == 'fill' else gb.InterpGrid(A, BC[bc], IT[order])\n"
            "    v, gr, h = g._eval(torch.from_numpy(X).cuda(), int(mode))\n"
            "    out = dict(v=v.cpu().numpy(), g=gr.cpu().numpy(), hv=gb.getindex_batch(g, X), c=g.coefs, cd=gb.jl_to_numpy(g.coefs_device()))\n"
            "    if h is not None: out['h'] = h.cpu().numpy()\n"
            "    out['o'] = g[tuple(z[f'x{k}'] for k in range(A.ndim))]\n"
            "    g2 = gb.InterpGrid.from_coefs(g.coefs_device(), gb.BCfill if bc == 'fill' else BC[bc], IT[order], fill=-1.25)\n"
            "    out['v2'] = g2._eval(torch.from_numpy(X).cuda(), 1)[0].cpu().numpy()\n"
            "    np.savez(f'{d}/out_{tag}.npz', **out)\n"
        )
        keep = []
        with open(os.path.join(d, "cases.txt"), "w") as f:
            k = 0
            for bc, order in cases:
                for shape in (shapes2[k % len(shapes2)], shapes3[k % len(shapes3)]):
                    dtype = np.float32 if k % 3 == 2 else np.float64
                    A = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
                    wide = bc != "nil"
                    X = make_queries(rng, shape, 3000, dtype, wide=wide)
                    if bc == "nil":  # keep the valid locations (same validity rule as BCnan, whose answer there is NaN)
                        ok = ~np.isnan(oracle.evaluate(oracle.make_coefs(A, "nan", order), "nan", order, X, 1)[0])
                        X = np.ascontiguousarray(X[ok])
                    vecs = {}
                    for j, n in enumerate(shape):
                        lo, hi = (-1.5 * n, 2.5 * n) if wide else (2.0, n - 1.0)
                        vecs[f"x{j}"] = np.concatenate([lo + rng.random(4) * (hi - lo), [float(n - 1), 2.0]]).astype(dtype)
                    mode = 7 if order in ("quadratic", "cubic") else 3
                    tag = f"{k}_{len(shape)}"
                    np.savez(os.path.join(d, f"in_{tag}.npz"), A=A, X=X, **vecs)
                    f.write(f"{tag} {bc} {order} {mode}\n")
                    keep.append((tag, bc, order, A, X, vecs, dtype, mode))
                k += 1
        env = dict(os.environ, GB_BLOCK_MIN_MB="0")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-3000:]
        for tag, bc, order, A, X, vecs, dtype, mode in keep:
            fill = -1.25 if bc == "fill" else None
            co = oracle.make_coefs(A, bc, order, fill)
            ov, og, oh, _ = oracle.evaluate(co, bc, order, X, mode)
            z = np.load(os.path.join(d, f"out_{tag}.npz"))
            assert same(z["v"], ov) and same(z["g"], og) and same(z["hv"], ov), (tag, bc, order)
            if mode == 7:
                assert same(z["h"], oh), (tag, bc, order)
            assert same(z["c"], co) and same(z["cd"], co) and same(z["v2"], ov), (tag, bc, order)
            ref = oracle.OracleGrid(A, bc, order, fill=fill).outer(*[vecs[f"x{j}"] for j in range(A.ndim)])
            assert same(z["o"], ref.astype(dtype)), (tag, bc, order)


def test_stage_profiler(gb):
    """gb_profile_enable / gb_profile_read (bench.py's roofline line): six positive stage times after a brick-path
    call, nothing after a plain one."""
    import torch

    rng = np.random.default_rng(67)
    A = np.asfortranarray(rng.standard_normal((40, 30, 20)))
    g = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    X = torch.from_numpy(make_queries(rng, A.shape, 50_000, np.float64, wide=False)).cuda()
    gb.profile_enable(True)
    try:
        g._eval(X, 3, tiled=False)
        assert gb.profile_read() == {}
        g._eval(X, 3, tiled=True)
        st = gb.profile_read()
        assert tuple(st) == gb.PROFILE_STAGES + ("coherent",) and all(st[k] > 0 for k in gb.PROFILE_STAGES)
        assert st["coherent"] == 0.0  # GB_TILED alone: the bins, no coherence probe
    finally:
        gb.profile_enable(False)
