"""Pins the CPU oracle against the reference's own known-answer tests and README values.

Every block cites the reference test it replays (paths under /root/reference/).  These run on
CPU (`-m "not gpu"`) and are what "oracle pinned" in DESIGN.md refers to.
"""
import numpy as np
import pytest

EPS = np.sqrt(np.finfo(float).eps)
eps = np.finfo(float).eps
BCS = ["nil", "nan", "na", "reflect", "periodic", "nearest", "fill"]
ORDERS = ["nearest", "linear", "quadratic", "cubic"]


def mk(O, A, bc, order):
    return O.OracleGrid(A, bc, order, fill=0.0 if bc == "fill" else None)


def supported(bc, order):
    return not (order == "cubic" and bc not in ("nil", "nan", "na"))


# --- test/counter.jl:4-10 : enumeration order is dim-1-fastest (offset table order) -------------
def test_enumeration_order(oracle):
    sp = oracle.set_position("linear", (2, 3), (1, 2), "nearest", [1.0, 1.0])
    # offsets of taps (0,0),(1,0),(0,1),(1,1) with strides (1,2)
    assert list(sp["offset"]) == [0, 1, 2, 3]


# --- test/grid.jl:17-26 : low-level tap / weight KAT ---------------------------------------------
def test_lowlevel_tap_weight_kat(oracle):
    sp = oracle.set_position("quadratic", (4, 6), (1, 4), "nearest", [2, 3.5])
    assert sp["coord1d"].tolist() == [[1, 2, 3], [3, 4, 5]]
    assert sp["coef1d"].tolist() == [[1 / 8, 3 / 4, 1 / 8], [1 / 2, 1 / 2, 0]]
    assert sp["wrap"] is False
    z = np.zeros(24)
    z[sp["offset"] + sp["offset_base"]] = sp["coef"]  # (offset+offset_base+1) is 1-based
    z = z.reshape((4, 6), order="F")
    expect = np.array([[0, 0, 1 / 16, 1 / 16, 0, 0], [0, 0, 3 / 8, 3 / 8, 0, 0], [0, 0, 1 / 16, 1 / 16, 0, 0], [0] * 6])
    assert (z == expect).all()


# --- test/grid.jl:29-48 : centre weight == prod(3/4-(x-2)^2) -------------------------------------
@pytest.mark.parametrize("nd", [1, 2, 3, 4])
def test_centre_weight(oracle, nd):
    rng = np.random.default_rng(nd)
    dims = [3] * nd
    s = [3 ** i for i in range(nd)]
    x = rng.random(nd) + 1.5
    sp = oracle.set_position("quadratic", dims, s, "nearest", x)
    assert sp["wrap"] is False
    mid = int(np.ceil(len(sp["coef"]) / 2)) - 1
    assert abs(sp["coef"][mid] - np.prod(0.75 - (x - 2) ** 2)) < eps
    # index table consistent with coord1d: 1 + sum (c-1)*stride
    index = sp["offset"] + sp["offset_base"] + 1
    it = np.ndindex(*([3] * nd))
    ref = sorted(1 + sum((sp["coord1d"][d][c[d]] - 1) * s[d] for d in range(nd)) for c in it)
    assert sorted(index.tolist()) == ref


# --- test/grid.jl:50-98 : on-grid reproduction, all BCs x orders, 2/3/4-D, scalar + outer ---------
@pytest.mark.parametrize("shape", [(4, 10), (4, 5, 4), (4, 5, 4, 3)])
@pytest.mark.parametrize("bc", BCS)
@pytest.mark.parametrize("order", ORDERS)
def test_on_grid_values(oracle, shape, bc, order):
    if not supported(bc, order):
        pytest.skip("cubic only for nil/nan/na (README.md:146)")
    rng = np.random.default_rng(hash((shape, bc, order)) % 2 ** 32)
    A = rng.standard_normal(shape)
    ig = mk(oracle, A, bc, order)
    for idx in np.ndindex(*shape):
        assert abs(ig[[i + 1 for i in idx]] - A[idx]) < EPS
    v = ig.outer(*[np.arange(1, n + 1) for n in shape])
    assert np.all(np.abs(v - A) < EPS)


# --- test/grid.jl:100-121 : boundary-condition sequences ------------------------------------------
@pytest.mark.parametrize("order", ["nearest", "linear", "quadratic"])
def test_bc_sequences(oracle, order):
    A = np.array([1.0, 2, 3, 4])
    ig = mk(oracle, A, "reflect", order)
    y = ig.outer(np.arange(-3, 9))
    assert np.all(np.abs(y - A[[3, 2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 0]]) < EPS)
    with pytest.raises(oracle.OracleError):
        ig.outer(np.arange(-3, 9), np.arange(2, 5))
    ig = mk(oracle, A, "nearest", order)
    y = ig.outer(np.arange(0, 7))
    assert np.all(np.abs(y - A[[0, 0, 1, 2, 3, 3, 3]]) < EPS)
    ig = mk(oracle, A, "nil", order)
    with pytest.raises(oracle.OracleError):
        ig[-0.8]
    for bc in ("na", "nan"):
        ig = mk(oracle, A, bc, order)
        assert np.isnan(ig[-0.8])


# --- test/grid.jl:123-154 + README.md:44-73 : quadratic exactness, low- and high-level ------------
def test_quadratic_exactness(oracle):
    c, a, o = 2.3, 8.1, 1.6
    qf = lambda x: a * (x - c) ** 2 + o
    y = qf(np.arange(1.0, 6.0))
    yc = oracle.invert(y, "nan", "quadratic")
    x = 1.8
    v, g, h = oracle.lowlevel(yc, "quadratic", "nil", [x])
    assert abs(v - qf(x)) < 100 * eps
    assert abs(g[0] - 2 * a * (x - c)) < 100 * eps
    assert abs(h[0, 0] - 2 * a) < 100 * eps
    yi = mk(oracle, y, "nil", "quadratic")
    assert abs(yi[x] - qf(x)) < 100 * eps
    v, g = yi.valgrad(x)
    assert abs(v - qf(x)) < 100 * eps and abs(g - 2 * a * (x - c)) < 100 * eps
    v2, g2, h2 = yi.valgradhess(x)
    assert abs(v2 - qf(x)) < 100 * eps and abs(g2 - 2 * a * (x - c)) < 100 * eps and abs(h2 - 2 * a) < 100 * eps
    # README.md:44-46, 66-68 -- printed by the reference itself; bit-for-bit
    assert yi[3] == 5.569000000000003
    assert yi.valgradhess(4.25) == (32.40025, 31.59000000000001, 16.200000000000017)
    assert abs(yi[1.9] - 2.8959999999999995) <= 2 * np.spacing(2.896)  # README.md:51 (older release)
    v, g = yi.valgrad(4.25)  # README.md:60 was printed by an older release: 1-2 ulp
    assert abs(v - 32.40025000000001) < 4 * np.spacing(32.4) and abs(g - 31.590000000000003) < 4 * np.spacing(31.6)


def _dnum(f, x, h=None):
    h = h or np.spacing(max(abs(x), 1.0)) ** (1 / 3)
    return (f(x + h) - f(x - h)) / ((x + h) - (x - h))


def _d2num(f, x, h=None):
    h = h or np.spacing(max(abs(x), 1.0)) ** (1 / 3)
    return (f(x + 2 * h) - 2 * f(x) + f(x - 2 * h)) / (4 * h * h)


# --- test/grid.jl:156-177 : derivatives vs centred differences, 1-D, six BCs ----------------------
@pytest.mark.parametrize("bc", ["nan", "na", "nearest", "periodic", "reflect", "fill"])
def test_quadratic_derivatives_1d(oracle, bc):
    c, a, o = 2.3, 8.1, 1.6
    y = a * (np.arange(1.0, 6.0) - c) ** 2 + o
    yi = mk(oracle, y, bc, "quadratic")
    x = 1.8
    with pytest.raises(IndexError):
        yi[[1.1, 2.8]]  # BoundsError: wrong dimensionality on a 1-D grid
    f = lambda t: yi[t]
    v, g = yi.valgrad(x)
    gn = _dnum(f, x)
    assert abs(g - gn) < EPS * (abs(g) + abs(gn))
    v2, g2, h2 = yi.valgradhess(x)
    hn = _d2num(f, x)
    assert abs(g2 - gn) < EPS * (abs(g2) + abs(gn))
    assert abs(h2 - hn) < np.cbrt(eps) * (abs(h2) + abs(hn))


# --- test/grid.jl:179-199 : 2-D BCreflect quadratic derivatives -----------------------------------
def test_quadratic_derivatives_2d(oracle):
    rng = np.random.default_rng(7)
    y = rng.random((7, 8))
    yi = mk(oracle, y, "reflect", "quadratic")
    x = np.array([2.2, 3.1])
    v, g = yi.valgrad(*x)
    v2, g2, h2 = yi.valgradhess(*x)
    assert v == v2 and np.all(g == g2)
    hstep = eps ** (1 / 3) * 2
    for d in range(2):
        e = np.zeros(2); e[d] = hstep
        gn = (yi[x + e] - yi[x - e]) / (2 * hstep)
        assert abs(g[d] - gn) < EPS * (abs(g[d]) + abs(gn)) + 1e-8
        hn = (yi[x + 2 * e] - 2 * yi[x] + yi[x - 2 * e]) / (4 * hstep ** 2)
        assert abs(h2[d, d] - hn) < 1e-5
    e1, e2 = np.array([hstep, 0]), np.array([0, hstep])
    hn = (yi[x + e1 + e2] - yi[x + e1 - e2] - yi[x - e1 + e2] + yi[x - e1 - e2]) / (4 * hstep ** 2)
    assert abs(h2[0, 1] - hn) < 1e-5 and h2[0, 1] == h2[1, 0]


# --- test/grid.jl:201-217 : nearest / linear value and gradient -----------------------------------
def test_nearest_linear(oracle):
    y = np.arange(2.0, 7.0)
    ig = mk(oracle, y, "nan", "nearest")
    v, g = ig.valgrad(1.8)
    assert v == 3 and g == 0
    assert np.all(ig.outer(np.arange(1.8, 5.41, 1.0)) == np.arange(3, 7))
    ig = mk(oracle, y, "nan", "linear")
    v, g = ig.valgrad(1.8)
    assert abs(v - 2.8) < EPS and abs(g - 1.0) < EPS
    assert np.all(np.abs(ig.outer(np.arange(1.8, 5.41, 1.0)) - np.arange(2.8, 5.81, 1.0)) < EPS)


# --- test/grid.jl:260-264 : 5-D linear through the generic path ----------------------------------
def test_5d_linear_generic(oracle):
    B = np.random.default_rng(5).random((4, 4, 4, 4, 4))
    Bi = mk(oracle, B, "nil", "linear")
    assert Bi[[2, 2, 2, 2, 2]] == B[1, 1, 1, 1, 1]


# --- test/cubic_interpolation.jl:1-37 ------------------------------------------------------------
def test_cubic_polynomial(oracle):
    a, b, c, d = 4.3, 1.7, 6.2, 5.9
    cf = lambda x: a * x ** 3 + b * x ** 2 + c * x + d
    cg = lambda x: 3 * a * x ** 2 + 2 * b * x + c
    ch = lambda x: 6 * a * x + 2 * b
    y = cf(np.arange(1.0, 9.0))
    yc = oracle.invert(y, "nan", "cubic")
    x = 3.7
    v, g, h = oracle.lowlevel(yc, "cubic", "nil", [x])
    assert abs(cf(x) - v) < 0.01 * (abs(cf(x)) + abs(v))
    assert abs(cg(x) - g[0]) < 0.1 * (abs(cg(x)) + abs(v))
    assert abs(ch(x) - h[0, 0]) < 0.1 * (abs(ch(x)) + abs(v))
    yi = mk(oracle, y, "nil", "cubic")
    assert abs(yi[x] - cf(x)) < 0.01 * (abs(cf(x)) + abs(yi[x]))
    v, g, h = yi.valgradhess(x)
    assert abs(v - cf(x)) < 0.01 * (abs(cf(x)) + abs(v))
    assert abs(g - cg(x)) < 0.01 * (abs(cg(x)) + abs(v))
    assert abs(h - ch(x)) < 0.01 * (abs(ch(x)) + abs(v))


# --- test/cubic_interpolation.jl:39-62 -----------------------------------------------------------
@pytest.mark.parametrize("bc", ["nan", "na"])
def test_cubic_derivatives_1d(oracle, bc):
    a, b, c, d = 4.3, 1.7, 6.2, 5.9
    y = a * np.arange(1.0, 9.0) ** 3 + b * np.arange(1.0, 9.0) ** 2 + c * np.arange(1.0, 9.0) + d
    yi = mk(oracle, y, bc, "cubic")
    x = 3.7
    f = lambda t: yi[t]
    v, g, h = yi.valgradhess(x)
    gn, hn = _dnum(f, x), _d2num(f, x)
    assert abs(g - gn) < 10 * np.cbrt(eps) * (abs(g) + abs(gn))
    assert abs(h - hn) < 10 * np.cbrt(eps) * (abs(h) + abs(hn))


# --- test/cubic_interpolation.jl:64-89 -----------------------------------------------------------
def test_cubic_derivatives_2d(oracle):
    y = np.random.default_rng(11).random((7, 8))
    yi = mk(oracle, y, "nan", "cubic")
    x = np.array([3.4, 6.1])
    v, g = yi.valgrad(*x)
    v2, g2, h2 = yi.valgradhess(*x)
    assert v == v2 and np.all(g == g2)
    hs = np.cbrt(eps) * 4
    tol4 = 10 * np.sqrt(np.sqrt(eps))
    for d in range(2):
        e = np.zeros(2); e[d] = hs
        gn = (yi[x + e] - yi[x - e]) / (2 * hs)
        assert abs(g[d] - gn) < np.cbrt(eps) * (abs(g[d]) + abs(gn)) * 10
        hn = (yi[x + 2 * e] - 2 * yi[x] + yi[x - 2 * e]) / (4 * hs ** 2)
        assert abs(h2[d, d] - hn) < tol4 * (abs(h2[d, d]) + abs(hn))
    e1, e2 = np.array([hs, 0]), np.array([0, hs])
    hn = (yi[x + e1 + e2] - yi[x + e1 - e2] - yi[x - e1 + e2] + yi[x - e1 - e2]) / (4 * hs ** 2)
    assert abs(h2[0, 1] - hn) < tol4 * (abs(h2[0, 1]) + abs(hn)) and h2[0, 1] == h2[1, 0]


# --- test/coord.jl:1-44 + README.md:76-97 : CoordInterpGrid ---------------------------------------
def test_coord_grid(oracle):
    # -1.0:0.1:1.0 == FloatRange(-10.0, 1.0, 21, 10.0)
    rx = (-10.0, 1.0, 21, 10.0)
    x = np.array([(rx[0] + i * rx[1]) / rx[3] for i in range(21)])
    z = np.sin(x)
    cg = oracle.OracleGrid(z, "nil", "quadratic", ranges=[rx])
    assert abs(cg[0.0] - np.sin(0.0)) < 100 * eps
    assert abs(cg[0.3] - np.sin(0.3)) < 100 * eps
    assert cg[-1.0] == -0.8414709848078965  # README.md:83-84
    for t in (0.0, 0.3):
        v, g = cg.valgrad(t)
        assert abs(g - np.cos(t)) < 0.001
    ry = (-20.0, 1.0, 41, 10.0)  # -2.0:0.1:2.0
    y = np.array([(ry[0] + i * ry[1]) / ry[3] for i in range(41)])
    z2 = np.sin(x[:, None] + y[None, :])
    with pytest.raises(ValueError):
        oracle.OracleGrid(z2, "nil", "quadratic", ranges=[ry, rx])
    cg = oracle.OracleGrid(z2, "nil", "quadratic", ranges=[rx, ry])
    assert abs(cg[[0.0, 0.0]]) < 1e-4 and abs(cg[[0.0, 0.3]] - np.sin(0.3)) < 1e-4 and abs(cg[[0.3, 0.0]] - np.sin(0.3)) < 1e-4
    v, g = cg.valgrad(0.0, 0.0)
    assert abs(g[0] - 1) < 1e-3 and abs(g[1] - 1) < 1e-3
    v, g = cg.valgrad(0.3, 0.0)
    assert abs(g[0] - np.cos(0.3)) < 1e-3 and abs(g[1] - np.cos(0.3)) < 1e-3
    # README.md:88-96: x = -1.0:0.1:1.0, y = 2.0:0.5:10.0 == FloatRange(4.0, 1.0, 17, 2.0)
    ry2 = (4.0, 1.0, 17, 2.0)
    y2 = np.array([(ry2[0] + i) / 2.0 for i in range(17)])
    z3 = np.sin(x[:, None] + y2[None, :])
    cg = oracle.OracleGrid(z3, "nil", "quadratic", ranges=[rx, ry2])
    assert abs(cg[[0.2, 4.1]] - (-0.9156696045493824)) < 1e-14
    # issue #29 (test/coord.jl:38-44): BCfill + Linear keeps size and corner value
    rxg = (-20.0, 1.0, 41, 2.0)    # -10:.5:10
    ryg = (-60.0, 3.0, 44, 20.0)   # -3:.15:3.5
    xg = np.array([(rxg[0] + i * rxg[1]) / rxg[3] for i in range(41)])
    yg = np.array([(ryg[0] + i * ryg[1]) / ryg[3] for i in range(44)])
    zg = np.sin(xg)[:, None] * np.cos(yg)[None, :]
    ig = oracle.OracleGrid(zg, 0, "linear", ranges=[rxg, ryg])
    assert ig.shape == zg.shape
    assert ig[[xg[0], yg[0]]] == zg[0, 0]


# --- test/grid.jl:266-326 : restrict / prolong KATs (exact ==) ------------------------------------
def test_restrict_prolong_kats(oracle):
    O = oracle
    a = np.zeros(5); a[2] = 1
    p = O.prolong_sizes(a, [9])
    assert np.all(p == [0, 0, 0, 0.5, 1, 0.5, 0, 0, 0])
    assert np.all(O.restrict(np.array([0, 0, 0, 0, 1.0, 0, 0, 0, 0]), 0) == [0, 0, 1, 0, 0])
    assert np.all(O.restrict(np.array([0, 1.0, 0, 0, 0, 0, 0, 0, 0]), 0) == [0.5, 0.5, 0, 0, 0])
    assert np.all(O.restrict(p, 0) == [0, 0.25, 1.5, 0.25, 0])
    assert np.all(O.prolong_sizes(a, [8]) == [0, 0, 0.25, 0.75, 0.75, 0.25, 0, 0])
    a[2] = 0; a[0] = 1
    assert np.all(O.prolong_sizes(a, [9]) == [1, 0.5, 0, 0, 0, 0, 0, 0, 0])
    assert np.all(O.prolong_sizes(a, [8]) == [0.75, 0.25, 0, 0, 0, 0, 0, 0])
    a = np.array([0, 0, 1.0, 0])
    assert np.all(O.prolong_sizes(a, [7]) == [0, 0, 0, 0.5, 1, 0.5, 0])
    assert np.all(O.prolong_sizes(a, [6]) == [0, 0, 0.25, 0.75, 0.75, 0.25])
    assert np.all(O.restrict(np.array([0, 1.0, 0, 0, 0, 0, 0, 0]), 0) == [0.25, 0.75, 0, 0, 0])
    assert np.all(O.restrict(np.array([0, 0, 1.0, 0, 0, 0, 0, 0]), 0) == [0, 0.75, 0.25, 0, 0])
    # 2-D
    a = np.zeros((5, 5)); a[2, 2] = 1
    col9 = np.array([0, 0, 0, 0.5, 1, 0.5, 0, 0, 0]); col8 = np.array([0, 0, 0.25, 0.75, 0.75, 0.25, 0, 0])
    e9 = np.zeros((9, 5)); e9[:, 2] = col9
    e8 = np.zeros((8, 5)); e8[:, 2] = col8
    assert np.all(O.prolong(a, 0, 9) == e9) and np.all(O.prolong(a, 1, 9) == e9.T)
    assert np.all(O.prolong(a, 0, 8) == e8) and np.all(O.prolong(a, 1, 8) == e8.T)
    r3 = np.zeros((3, 5)); r3[1, 2] = 1
    assert np.all(O.restrict_flags(a, [True, False]) == r3) and np.all(O.restrict_flags(a, [False, True]) == r3.T)
    a = np.zeros((6, 6)); a[2, 2] = 1
    r4 = np.zeros((4, 6)); r4[:, 2] = [0, 0.75, 0.25, 0]
    assert np.all(O.restrict_flags(a, [True, False]) == r4) and np.all(O.restrict_flags(a, [False, True]) == r4.T)
    with pytest.raises(O.OracleError):
        O.prolong(np.zeros(5), 0, 7)  # src/restrict_prolong.jl:328-330


# --- README.md:236-241 : pyramid schedule ---------------------------------------------------------
def test_restrict_size_schedule(oracle):
    sz = [1920, 1080]
    rows = [[], []]
    for d in range(2):
        n = sz[d]
        for _ in range(4):
            n = oracle.restrict_size(n)
            rows[d].append(n)
    assert rows == [[961, 481, 241, 121], [541, 271, 136, 69]]
    n, chain = 1024, []
    while n > 17:
        n = oracle.restrict_size(n); chain.append(n)
    assert chain == [513, 257, 129, 65, 33, 17]


# --- src/boundaryconditions.jl:22-26 : wrap --------------------------------------------------------
def test_wrap(oracle):
    assert [oracle.wrap("reflect", p, 4) for p in range(-3, 9)] == [4, 3, 2, 1, 1, 2, 3, 4, 4, 3, 2, 1]
    assert [oracle.wrap("periodic", p, 4) for p in range(-3, 9)] == [1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3, 4]
    assert [oracle.wrap("nearest", p, 4) for p in range(-1, 7)] == [1, 1, 1, 2, 3, 4, 4, 4]
    assert oracle.wrap("nan", -5, 4) == -5


# --- prefilter matrices: independent dense check of the five systems (src/interp.jl:955-1014) ------
def _dense(bc, order, n):
    M = np.zeros((n, n))
    if order == "quadratic":
        for i in range(n):
            M[i, i] = 0.75
            if i > 0: M[i, i - 1] = 0.125
            if i < n - 1: M[i, i + 1] = 0.125
        if bc in ("nil", "nan", "na"):
            M[0, :3] = [9 / 8, -1 / 4, 1 / 8]; M[-1, -3:] = [1 / 8, -1 / 4, 9 / 8]
        elif bc == "reflect":
            M[0, 0] = M[-1, -1] = 7 / 8
        elif bc == "periodic":
            M[0, -1] = M[-1, 0] = 1 / 8
        else:
            M[0, 1] = M[-1, -2] = 1 / 4
    else:
        for i in range(n):
            M[i, i] = 4 / 6
            if i > 0: M[i, i - 1] = 1 / 6
            if i < n - 1: M[i, i + 1] = 1 / 6
        M[0, :] = 0; M[0, :3] = [1, -2, 1]
        M[1, :] = 0; M[1, 1] = 1
        M[-2, :] = 0; M[-2, -2] = 1
        M[-1, :] = 0; M[-1, -3:] = [1, -2, 1]
    return M


@pytest.mark.parametrize("bc,order", [("nan", "quadratic"), ("reflect", "quadratic"), ("periodic", "quadratic"),
                                      ("nearest", "quadratic"), ("nan", "cubic")])
@pytest.mark.parametrize("n", [5, 8, 33, 258])
def test_prefilter_vs_dense(oracle, bc, order, n):
    f = np.random.default_rng(n).standard_normal(n)
    c = oracle.invert(f, bc, order)
    M = _dense(bc, order, n)
    ref = np.linalg.solve(M, f)
    assert np.max(np.abs(c - ref)) < 64 * eps * np.max(np.abs(ref))
    assert np.max(np.abs(M @ c - f)) < 64 * eps * np.max(np.abs(f))


@pytest.mark.parametrize("bc,order", [("nan", "quadratic"), ("reflect", "quadratic"), ("periodic", "quadratic"), ("nearest", "quadratic"), ("nan", "cubic")])
def test_prefilter_vs_long_double_witness(oracle, bc, order):
    """Second witness for the prefilter's last bits (parity-unpinned by the reference's own tests): the oracle's
    LU + Woodbury restatement against a dense long-double Gaussian elimination of the same matrix at n = 258 (the C2 line
    length) -- within 4 ulp of max|f| for every system, Float64; within 4 Float32 ulp for Float32."""
    from test_fast_prefilter_host import dense_matrix, solve_longdouble

    n = 258
    for dt, e in ((np.float64, eps), (np.float32, 1.1920929e-07)):
        for seed in range(3):
            f = np.random.default_rng(seed).standard_normal(n).astype(dt)
            c = oracle.invert(f, bc, order)
            truth = solve_longdouble(dense_matrix(bc, order, n, dt), f)
            err = float(np.abs(c.astype(np.longdouble) - truth).max())
            assert err <= 4 * e * float(np.abs(f).max()), (bc, order, dt, err / (e * float(np.abs(f).max())))


def test_prefilter_lu_matches_lapack_gttrf(oracle):
    """The dgttrf transliteration (Julia Base lufact!(::Tridiagonal)) against LAPACK itself."""
    from scipy.linalg import lapack

    for bc, order in [("nan", "quadratic"), ("reflect", "quadratic"), ("nearest", "quadratic"), ("nan", "cubic")]:
        n = 258
        F = oracle.prefilter_factors(bc, order, n)
        M = _dense(bc, order, n)
        # strip the Woodbury part: tridiagonal only
        dl = np.diag(M, -1).copy(); d = np.diag(M).copy(); du = np.diag(M, 1).copy()
        dl2, d2, du2, du22, ipiv, info = lapack.dgttrf(dl, d, du)
        assert info == 0
        assert np.array_equal(ipiv, F["ipiv"])  # LAPACK pivots are 1-based
        for a, b in ((dl2, F["dl"]), (d2, F["d"]), (du2, F["du"]), (du22, F["du2"])):
            assert np.max(np.abs(a - b)) <= 2 * eps * max(1.0, np.max(np.abs(a)))
        assert (F["ipiv"] != np.arange(1, n + 1)).sum() == (1 if order == "cubic" else 0)


def test_prefilter_multidim_is_linewise(oracle):
    """interp_invert! sweeps every 1-D line of each dim in dimlist order (src/interp.jl:916-929)."""
    A = np.random.default_rng(3).standard_normal((6, 7, 5))
    full = oracle.invert(A, "reflect", "quadratic")
    step = A.copy(order="F")
    for d in range(3):
        step = oracle.invert(step, "reflect", "quadratic", dimlist=[d])
    assert np.array_equal(full, step)
    # dim 1 sweep == 1-D solves of each column
    one = oracle.invert(A, "reflect", "quadratic", dimlist=[1])
    for i in range(6):
        for k in range(5):
            assert np.array_equal(one[i, :, k], oracle.invert(A[i, :, k].copy(), "reflect", "quadratic"))


def test_float32_promotion_rules(oracle):
    """Float32 weights go through Float64 where the reference's literals force it (interp.jl:815-847)."""
    rng = np.random.default_rng(0)
    A = rng.random((9, 8)).astype(np.float32)
    for bc, order in [("reflect", "quadratic"), ("nan", "cubic"), ("periodic", "linear")]:
        co32 = oracle.make_coefs(A, bc, order)
        assert co32.dtype == np.float32
        X = (1 + rng.random((200, 2)) * np.array([7.0, 6.0])).astype(np.float32)
        if order == "cubic":
            X = np.maximum(X, 1.0).astype(np.float32)
        v32, g32, _, _ = oracle.evaluate(co32, bc, order, X, 3)
        v64, g64, _, _ = oracle.evaluate(co32.astype(np.float64), bc, order, X.astype(np.float64), 3)
        assert np.max(np.abs(v32 - v64)) < 2e-6 * max(1, np.max(np.abs(v64)))
        assert np.max(np.abs(g32 - g64)) < 1e-5 * max(1, np.max(np.abs(g64)))


def test_divergences_documented(oracle):
    """D1-D4 (see oracle header / DESIGN.md)."""
    A = np.arange(1.0, 9.0)
    ig = mk(oracle, A, "nan", "cubic")
    assert np.isnan(ig[0.5])          # D2: user x in [0,1) -> invalid instead of OOB read
    assert not np.isnan(ig[1.0]) and not np.isnan(ig[8.0]) and np.isnan(ig[8.0001])
    ig = mk(oracle, A, "nan", "nearest")
    assert ig[0.5] == 1.0 and ig[8.5] == 8.0     # D3 clamp (round-half-even gives 0 / 8)
    ig = mk(oracle, A, "reflect", "quadratic")
    assert np.isnan(ig[np.nan]) and np.isnan(ig[np.inf]) and np.isnan(ig[1e300])  # D4
    ig = mk(oracle, A, "nan", "linear")
    v, g = ig.valgrad(8.0)            # D1: tap len+1 under a non-zero gradient weight is skipped
    assert v == 8.0 and g == -8.0


@pytest.mark.parametrize("bc,order,shape", [("nan", "cubic", (12, 13, 11)), ("reflect", "quadratic", (30, 20)), ("periodic", "linear", (6, 7, 5, 6)),
                                            ("fill", "quadratic", (9, 8, 7)), ("nan", "cubic", (40,))])
def test_extended_precision_truth_brackets_the_reference_arithmetic(oracle, bc, order, shape):
    """oracle.evaluate_truth (long double, OURS) vs the reference-order double arithmetic: the two
    agree to a few ulp of the magnitude scale S = sum|c_i a_i| -- the yardstick for GB_FAST."""
    rng = np.random.default_rng(5)
    A = rng.random(shape)
    co = oracle.make_coefs(A, bc, order, -1.25 if bc == "fill" else None)
    X = 1.0 + (np.array(shape) - 1.0) * rng.random((3000, len(shape)))
    v, g, _, _ = oracle.evaluate(co, bc, order, X, 3)
    tv, tg, sv, sg = oracle.evaluate_truth(co, bc, order, X)
    eps = 2.220446049250313e-16
    assert np.all(np.abs(v - tv) <= 8 * eps * sv) and np.all(np.abs(g - tg) <= 8 * eps * sg)
    assert np.all(sv >= np.abs(tv) * (1 - 1e-12)) and np.all(sg >= np.abs(tg) * (1 - 1e-12))


@pytest.mark.parametrize("L", [9, 10, 17, 32])
def test_balanced_restrict_prolong_uses_every_point_equally(oracle, L):
    """src/restrict_prolong.jl:189-195: with the "b" variants sum(P*R, 1) is constant (the plain pair lets
    the edges decay); restrict_extrap reproduces a linear ramp at the edges."""
    n = oracle.restrict_size(L)
    PRb = np.zeros((L, L))
    PR = np.zeros((L, L))
    for k in range(L):
        e = np.zeros(L)
        e[k] = 1.0
        PRb[:, k] = oracle.prolongb(oracle.restrictb(e, 0, 1.0), 0, L)
        PR[:, k] = oracle.prolong(oracle.restrict(e, 0, 1.0), 0, L)
    cs = PRb.sum(axis=0)
    assert np.allclose(cs, cs[0], rtol=0, atol=1e-14)
    cp = PR.sum(axis=0)
    assert abs(cp[0] - cp[L // 2]) > 0.1  # the plain pair does not
    ramp = np.arange(1.0, L + 1.0)
    r = oracle.restrict_extrap(ramp, 0, 0.5)  # scale 1/2: restriction of a ramp is a ramp of half the samples
    assert r.shape == (n,)
    d = np.diff(r)
    assert np.allclose(d, d[0], rtol=0, atol=1e-12)


def test_interp_irregular_reference_semantics(oracle):
    """src/interp.jl:351-377 by hand: knots [1,2,4,8], samples [10,20,40,80]."""
    g, c = [1.0, 2.0, 4.0, 8.0], [10.0, 20.0, 40.0, 80.0]
    x = [1.0, 1.5, 2.0, 3.0, 8.0, 0.5, 9.0]
    lin, _ = oracle.irregular(g, c, "nan", "linear", x)
    assert np.array_equal(lin[:5], [10.0, 15.0, 20.0, 30.0, 80.0]) and np.isnan(lin[5:]).all()
    near, _ = oracle.irregular(g, c, "nearest", "nearest", x)
    assert np.array_equal(near, [10.0, 20.0, 20.0, 40.0, 80.0, 10.0, 80.0])  # ties go to the upper knot (strict <, :373)
    fwd, _ = oracle.irregular(g, c, "fill", "forward", x, fill=-1.0)
    assert np.array_equal(fwd, [20.0, 20.0, 20.0, 40.0, 80.0, -1.0, -1.0])
    bwd, _ = oracle.irregular(g, c, "fill", "backward", x, fill=-1.0)
    assert np.array_equal(bwd, [10.0, 10.0, 10.0, 20.0, 40.0, -1.0, -1.0])
    with pytest.raises(oracle.OracleError):
        oracle.irregular(g, c, "nil", "linear", [0.0])


def test_oracle_reproduces_committed_golden_vectors(oracle):
    """tests/golden/*.npz were written by make_golden.py from this oracle; the CPU suite keeps the two in
    step (a drift of the restatement would otherwise only show up on the GPU box)."""
    import glob
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = sorted(glob.glob(os.path.join(here, "eval_*.npz")))
    assert len(files) >= 9
    for f in files:
        z = np.load(f)
        bc, order, mask = str(z["bc"]), str(z["order"]), int(z["mask"])
        co = oracle.make_coefs(z["A"], bc, order, float(z["fill"]) if bc == "fill" else None)
        assert co.dtype == z["coefs"].dtype and np.array_equal(co, z["coefs"], equal_nan=True)
        v, g, h, _ = oracle.evaluate(co, bc, order, z["X"], mask, raise_invalid=False)
        assert np.array_equal(v, z["val"], equal_nan=True)
        if mask & 2:
            assert np.array_equal(g, z["grad"], equal_nan=True)
        if mask & 4:
            assert np.array_equal(h, z["hess"], equal_nan=True)
    for f in sorted(glob.glob(os.path.join(here, "mg_*.npz"))):
        z = np.load(f)
        assert np.array_equal(oracle.restrict_flags(z["A"], [True] * z["A"].ndim), z["R"])
        assert np.array_equal(oracle.prolong_sizes(z["R"], list(z["A"].shape)), z["P"])
    for f in sorted(glob.glob(os.path.join(here, "mgb_*.npz"))):
        z = np.load(f)
        for d in range(z["A"].ndim):
            assert np.array_equal(oracle.restrictb(z["A"], d, 1.0), z[f"restrictb_{d}"])
            assert np.array_equal(oracle.restrict_extrap(z["A"], d, 1.0), z[f"extrap_{d}"])
    z = np.load(os.path.join(here, "irregular_1d.npz"))
    assert np.array_equal(oracle.irregular(z["knots"], z["samples"], "nan", "linear", z["x"])[0], z["nan_linear"], equal_nan=True)
