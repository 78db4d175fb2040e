"""ctypes front-end of the CPU oracle (oracle/grid_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Arrays follow Julia's convention: shape (n1, n2, ...) with dim 1 fastest in memory, i.e. numpy
Fortran order.  Query batches are (nq, N) C-contiguous arrays (== Julia's N x nq Matrix, one
column per query) or a tuple of N vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC = {"nil": 0, "nan": 1, "na": 2, "reflect": 3, "periodic": 4, "nearest": 5, "fill": 6}
ORDER = {"nearest": 0, "linear": 1, "quadratic": 2, "cubic": 3}
OUT_VAL, OUT_GRAD, OUT_HESS = 1, 2, 4
ERR_OK, ERR_ARG, ERR_INVALID_LOCATION, ERR_UNSUPPORTED = 0, 1, 2, 3


class OracleError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"oracle error {code} {msg}")
        self.code = code


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "grid_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_restrict_size.restype = C.c_int64
        _LIB.oracle_restrict_size.argtypes = [C.c_int64]
    return _LIB


def _dt(a):
    if a.dtype == np.float32:
        return 0
    if a.dtype == np.float64:
        return 1
    raise TypeError(a.dtype)


def _i64(v):
    return (C.c_int64 * len(v))(*[int(x) for x in v])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f(a):
    return np.asfortranarray(a)


def wrap(bc, pos, length):
    out = C.c_int64()
    lib().oracle_wrap(BC[bc], C.c_int64(pos), C.c_int64(length), C.byref(out))
    return out.value


def restrict_size(n):
    return lib().oracle_restrict_size(int(n))


def needs_shift(bc, order):
    return bc == "fill" or (order == "cubic" and bc in ("nil", "nan", "na"))


def pad1(A, f):
    A = _f(A)
    out = np.empty(tuple(s + 2 for s in A.shape), dtype=A.dtype, order="F")
    rc = lib().oracle_pad1(_dt(A), _p(A), A.ndim, _i64(A.shape), C.c_double(f), _p(out))
    if rc:
        raise OracleError(rc)
    return out


def invert(A, bc, order, dimlist=None, nthreads=0):
    """interp_invert!(A, BC, IT, dimlist) on a copy; dimlist is 0-based here."""
    A = np.array(A, order="F", copy=True)
    if dimlist is None:
        dimlist = list(range(A.ndim))
    dl = (C.c_int * len(dimlist))(*dimlist)
    rc = lib().oracle_invert(_dt(A), _p(A), A.ndim, _i64(A.shape), BC[bc], ORDER[order], dl, len(dimlist), nthreads)
    if rc:
        raise OracleError(rc)
    return A


def make_coefs(A, bc, order, fill=None):
    """What InterpGrid's constructors store (src/interp.jl:48-72)."""
    A = _f(np.asarray(A))
    if bc == "fill":
        if fill is None:
            raise ValueError("Construct BCfill InterpGrids by supplying the fill value")
        return invert(pad1(A, fill), "nearest", order)
    if order == "cubic" and bc in ("nil", "nan", "na"):
        return invert(pad1(A, 0.0), bc, order)
    return invert(A, bc, order)


def _queries(X, N):
    """-> (array, xs_q, xs_d, nq)"""
    if isinstance(X, (tuple, list)) and not np.isscalar(X[0]) and np.ndim(X[0]) == 1:
        cols = [np.ascontiguousarray(c) for c in X]
        if len(cols) != N:
            raise OracleError(ERR_ARG, "Incorrect number of dimensions supplied")
        dt = np.result_type(*cols)
        dt = np.float32 if dt == np.float32 else np.float64
        soa = np.ascontiguousarray(np.stack([c.astype(dt) for c in cols], axis=0))
        return soa, 1, soa.shape[1], soa.shape[1]
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    X = np.ascontiguousarray(X.reshape(-1, N) if X.ndim == 1 else X)
    if X.shape[1] != N:
        raise OracleError(ERR_ARG, "Incorrect number of dimensions supplied")
    return X, N, 1, X.shape[0]


def evaluate(coefs, bc, order, X, out_mask=OUT_VAL, affine=None, nthreads=0, raise_invalid=True):
    """Loop of scalar getindex/valgrad/valgradhess calls over the query batch.

    coefs: the STORED (padded, prefiltered) array; X: user coordinates (the +1 shift of setx is
    applied inside).  Returns (val, grad, hess, first_invalid)."""
    coefs = _f(coefs)
    N = coefs.ndim
    Xa, xs_q, xs_d, nq = _queries(X, N)
    T = coefs.dtype
    val = np.empty(nq, dtype=T)
    grad = np.empty((nq, N), dtype=T) if out_mask & (OUT_GRAD | OUT_HESS) else None
    hess = np.empty((nq, N, N), dtype=T) if out_mask & OUT_HESS else None
    aff = None
    if affine is not None:
        aff = np.ascontiguousarray(np.asarray(affine, dtype=np.float64).reshape(N, 4))
    bad = C.c_int64(-1)
    rc = lib().oracle_eval(_dt(coefs), _dt(Xa), _p(coefs), N, _i64(coefs.shape), BC[bc], ORDER[order], C.c_int64(nq),
                           _p(Xa), C.c_int64(xs_q), C.c_int64(xs_d), _p(aff), out_mask, _p(val), _p(grad), _p(hess),
                           C.byref(bad), nthreads)
    if rc == ERR_INVALID_LOCATION:
        if raise_invalid:
            raise OracleError(rc, f"Invalid location (query {bad.value})")
    elif rc:
        raise OracleError(rc)
    return val, grad, hess, bad.value


def evaluate_truth(coefs, bc, order, X, nthreads=0):
    """Extended-precision (x87 long double) value and gradient at the reference's taps, plus the
    magnitude scales S = sum|c_i a_i| per output: (val, grad, sval, sgrad).  float64 grids, no affine
    map.  OURS (ulp accounting for the GB_FAST contraction), not part of the reference."""
    coefs = _f(np.asarray(coefs, dtype=np.float64))
    N = coefs.ndim
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1, N))
    nq = X.shape[0]
    val, sval = np.empty(nq), np.empty(nq)
    grad, sgrad = np.empty((nq, N)), np.empty((nq, N))
    rc = lib().oracle_eval_truth_f64(_p(coefs), N, _i64(coefs.shape), BC[bc], ORDER[order], C.c_int64(nq), _p(X), _p(val), _p(grad),
                                     _p(sval), _p(sgrad), nthreads)
    if rc:
        raise OracleError(rc)
    return val, grad, sval, sgrad


def restrict(A, dim, scale=1.0, nthreads=0):
    """restrict(A, dim, scale) with 0-based dim."""
    A = _f(A)
    shp = list(A.shape)
    shp[dim] = restrict_size(shp[dim])
    R = np.empty(shp, dtype=A.dtype, order="F")
    rc = lib().oracle_restrict(_dt(A), _p(A), A.ndim, _i64(A.shape), dim, C.c_double(scale), _p(R), nthreads)
    if rc:
        raise OracleError(rc)
    return R


def prolong(A, dim, length, nthreads=0):
    A = _f(A)
    shp = list(A.shape)
    shp[dim] = int(length)
    P = np.empty(shp, dtype=A.dtype, order="F")
    rc = lib().oracle_prolong(_dt(A), _p(A), A.ndim, _i64(A.shape), dim, C.c_int64(length), _p(P), nthreads)
    if rc:
        raise OracleError(rc, f"Cannot prolong a dimension of length {A.shape[dim]} to {length}")
    return P


def _edge(variant, A, dim, scale, length):
    A = _f(A)
    shp = list(A.shape)
    shp[dim] = int(length) if variant == 3 else restrict_size(shp[dim])
    out = np.empty(shp, dtype=A.dtype, order="F")
    rc = lib().oracle_edge_variant(variant, _dt(A), _p(A), A.ndim, _i64(A.shape), dim, C.c_double(scale), C.c_int64(length), _p(out), 0)
    if rc:
        raise OracleError(rc)
    return out


def restrictb(A, dim, scale=1.0):
    """restrictb(A, dim, scale), 0-based dim (src/restrict_prolong.jl:196-224)."""
    return _edge(1, A, dim, scale, 0)


def restrict_extrap(A, dim, scale=1.0):
    """restrict_extrap(A, dim, scale), 0-based dim (:287-313)."""
    return _edge(2, A, dim, scale, 0)


def prolongb(A, dim, length):
    """prolongb(A, dim, len), 0-based dim (:231-267)."""
    return _edge(3, A, dim, 1.0, length)


IRR_ORDER = {"nearest": 0, "linear": 1, "forward": 4, "backward": 5}


def irregular(knots, samples, bc, it, x, fill=float("nan"), raise_invalid=True):
    """InterpIrregular(knots, samples, BC, IT)[x] for a vector x (src/interp.jl:318-377).  All arrays of one
    float type.  Returns (values, first_invalid)."""
    g = np.ascontiguousarray(knots)
    dt = g.dtype if g.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    g = np.ascontiguousarray(g, dtype=dt)
    c = np.ascontiguousarray(samples, dtype=dt)
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=dt)
    out = np.empty(x.shape[0], dtype=dt)
    bad = C.c_int64(-1)
    rc = lib().oracle_irregular(_dt(g), _p(g), _p(c), C.c_int64(g.shape[0]), BC[bc], IRR_ORDER[it], C.c_double(fill), _p(x), _p(out),
                                C.c_int64(x.shape[0]), C.byref(bad))
    if rc == ERR_INVALID_LOCATION:
        if raise_invalid:
            raise OracleError(rc, f"BoundsError (query {bad.value})")
    elif rc:
        raise OracleError(rc)
    return out, bad.value


def restrict_flags(A, flags, scale=1.0):
    """restrict(A, flag[, scale]) via mapdim (src/utilities.jl:7-20); flags: (ndims, ncols) bools."""
    flags = np.atleast_2d(np.asarray(flags, dtype=bool).T).T if np.ndim(flags) == 1 else np.asarray(flags, dtype=bool)
    if flags.ndim == 1:
        flags = flags[:, None]
    if flags.shape[0] != np.ndim(A):
        raise OracleError(ERR_ARG, "Boolean flag must have as many rows as dimensions of A")
    for j in range(flags.shape[1]):
        for i in range(flags.shape[0]):
            if flags[i, j]:
                A = restrict(A, i, scale)
    return A


def prolong_sizes(A, sz):
    """prolong(A, sz::Array{Int}) (src/restrict_prolong.jl:87-100)."""
    sz = np.asarray(sz, dtype=np.int64)
    if sz.ndim == 1:
        sz = sz[:, None]
    if sz.shape[0] != np.ndim(A):
        raise OracleError(ERR_ARG, "Array of sizes must have as many rows as dimensions of A")
    for j in range(sz.shape[1]):
        for i in range(sz.shape[0]):
            if sz[i, j] > A.shape[i]:
                A = prolong(A, i, int(sz[i, j]))
    return A


def synth(dtype, n, seed, lo=0.0, hi=1.0, first=0):
    out = np.empty(int(n), dtype=dtype)
    rc = lib().oracle_synth(_dt(out), _p(out), C.c_int64(n), C.c_uint64(seed), C.c_double(lo), C.c_double(hi), C.c_int64(first))
    if rc:
        raise OracleError(rc)
    return out


def max_threads():
    return lib().oracle_max_threads()


def set_position(order, dims, strides, bc, x, calc_grad=True, calc_hess=False):
    """Low-level set_position dump (f64) for the tap/weight KATs."""
    N = len(dims)
    L = ORDER[order] + 1
    n = L ** N
    coord1d = np.zeros((N, 4), dtype=np.int64)
    coef1d = np.zeros((N, 4)); gcoef1d = np.zeros((N, 4)); hcoef1d = np.zeros((N, 4))
    offset = np.zeros(n, dtype=np.int64); index = np.zeros(n, dtype=np.int64); coef = np.zeros(n)
    flags = np.zeros(3, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    rc = lib().oracle_set_position_f64(ORDER[order], N, _i64(dims), _i64(strides), BC[bc], int(calc_grad), int(calc_hess),
                                       _p(x), _p(coord1d), _p(coef1d), _p(gcoef1d), _p(hcoef1d), _p(offset), _p(index),
                                       _p(coef), _p(flags))
    return dict(rc=rc, coord1d=coord1d[:, :L], coef1d=coef1d[:, :L], gcoef1d=gcoef1d[:, :L], hcoef1d=hcoef1d[:, :L],
                offset=offset, index=index, coef=coef, wrap=bool(flags[0]), valid=bool(flags[1]), offset_base=int(flags[2]))


def lowlevel(A, order, bc, x, hess=True):
    """InterpGridCoefs + set_position + interp on raw (already inverted) coefficients, no shift."""
    A = _f(np.asarray(A, dtype=np.float64))
    N = A.ndim
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    val = C.c_double()
    grad = np.zeros(N); h = np.zeros((N, N))
    mask = OUT_VAL | OUT_GRAD | (OUT_HESS if hess else 0)
    rc = lib().oracle_lowlevel_f64(_p(A), ORDER[order], N, _i64(A.shape), BC[bc], _p(x), mask, C.byref(val), _p(grad), _p(h))
    if rc:
        raise OracleError(rc, "Invalid location")
    return val.value, grad, h


def prefilter_factors(bc, order, n):
    dl = np.zeros(n - 1); d = np.zeros(n); du = np.zeros(n - 1); du2 = np.zeros(max(n - 2, 0))
    ipiv = np.zeros(n, dtype=np.int64); Cp = np.zeros(4); vcol = np.zeros(2, dtype=np.int64); wb = C.c_int()
    rc = lib().oracle_prefilter_factors_f64(BC[bc], ORDER[order], C.c_int64(n), _p(dl), _p(d), _p(du), _p(du2), _p(ipiv),
                                            _p(Cp), _p(vcol), C.byref(wb))
    if rc:
        raise OracleError(rc)
    return dict(dl=dl, d=d, du=du, du2=du2, ipiv=ipiv, Cp=Cp.reshape(2, 2, order="F"), vcol=vcol, woodbury=bool(wb.value))


class OracleGrid:
    """Scalar-call convenience mirroring InterpGrid / CoordInterpGrid for the KAT replays."""

    def __init__(self, A, bc, order, fill=None, ranges=None):
        A = np.asarray(A)
        if A.dtype not in (np.float32, np.float64):
            A = A.astype(np.float64)
        if isinstance(bc, (int, float)) and not isinstance(bc, bool):
            fill, bc = float(bc), "fill"
        self.bc, self.order, self.N, self.shape = bc, order, A.ndim, A.shape
        self.coefs = make_coefs(A, bc, order, fill)
        self.affine = None
        if ranges is not None:
            if tuple(len_ for (_, _, len_, _) in ranges) != tuple(A.shape):
                raise ValueError("DimensionMismatch: Coordinate lengths do not match grid size.")
            # FloatRange(start, step, len, divisor)
            self.affine = [[0.0, r[0], r[1], r[3]] for r in ranges]

    def _x(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        if self.N == 1 and x.size == 2:
            if x[1] != 1:
                raise IndexError("BoundsError")
            x = x[:1]
        if x.size != self.N:
            raise OracleError(ERR_ARG, "Incorrect number of dimensions supplied")
        return x.reshape(1, self.N)

    def __getitem__(self, x):
        v, _, _, _ = evaluate(self.coefs, self.bc, self.order, self._x(x), OUT_VAL, self.affine)
        return v[0]

    def valgrad(self, *x):
        v, g, _, _ = evaluate(self.coefs, self.bc, self.order, self._x(x), OUT_VAL | OUT_GRAD, self.affine)
        return v[0], (g[0, 0] if self.N == 1 else g[0])

    def valgradhess(self, *x):
        v, g, h, _ = evaluate(self.coefs, self.bc, self.order, self._x(x), OUT_VAL | OUT_GRAD | OUT_HESS, self.affine)
        if self.N == 1:
            return v[0], g[0, 0], h[0, 0, 0]
        return v[0], g[0], h[0].T  # stored column-major per query

    def outer(self, *vecs):
        """tensor-product getindex (src/interp.jl:273-306)."""
        if len(vecs) != self.N:
            raise OracleError(ERR_ARG, "Dimensionality mismatch")
        arrs = [np.asarray(v) for v in vecs]
        dt = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64  # setx adds its 1 in the query's type
        mesh = np.meshgrid(*[a.astype(dt) for a in arrs], indexing="ij")
        X = np.stack([m.ravel(order="F") for m in mesh], axis=1)
        v, _, _, _ = evaluate(self.coefs, self.bc, self.order, X, OUT_VAL, self.affine)
        return v.reshape([len(v_) for v_ in vecs], order="F")
