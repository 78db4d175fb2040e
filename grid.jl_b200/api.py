"""Host-side mirror of Grid.jl's exported API (src/Grid.jl:28-70) over the C-ABI library.

Names, argument meaning and error behaviour follow the reference so that its tests can be replayed
almost verbatim (tests/test_gpu_reference_kats.py).  Conventions kept from Julia:
  * grid coordinates are continuous and 1-BASED (G[1.0] is the first sample), `dim` arguments and
    `dimlist` entries are 1-based;
  * array shapes are Julia sizes, memory is column-major (numpy arrays of any order are accepted
    and converted; device tensors must already be column-major, see `jl_empty`).
Batched entry points (absent upstream, defined as "a loop of the scalar call"):
  getindex_(out, G, X), valgrad_(v, g, G, X), valgradhess_(v, g, h, G, X) and the allocating
  getindex_batch / valgrad_batch / valgradhess_batch.  X is (nq, N) (one query per row == Julia's
  N x nq Matrix in memory) or a tuple of N vectors; host (numpy) or device (torch.cuda) resident.

All numerical work happens in libgridb200.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from fractions import Fraction  # noqa: F401  (kept for users building ranges)

import numpy as np

from . import _lib as L
from ._lib import BoundsError, DimensionMismatch, ErrorException, MethodError

# ---- type tags (src/boundaryconditions.jl:5-12, src/interpflags.jl:1-7) -------------------------


class BoundaryCondition:
    code = -1


class BCnil(BoundaryCondition):
    code = 0


class BCnan(BoundaryCondition):
    code = 1


class BCna(BoundaryCondition):
    code = 2


class BCreflect(BoundaryCondition):
    code = 3


class BCperiodic(BoundaryCondition):
    code = 4


class BCnearest(BoundaryCondition):
    code = 5


class BCfill(BoundaryCondition):
    code = 6


class InterpType:
    code = -1


class InterpNearest(InterpType):
    code = 0


class InterpLinear(InterpType):
    code = 1


class InterpQuadratic(InterpType):
    code = 2


class InterpCubic(InterpType):
    code = 3


class InterpForward(InterpType):  # InterpIrregular only (src/interp.jl:371)
    code = 4


class InterpBackward(InterpType):  # InterpIrregular only (:372)
    code = 5


def npoints(it):  # src/interp.jl:577-582
    return _it(it) + 1


def _bc(bc):
    if isinstance(bc, type) and issubclass(bc, BoundaryCondition) and bc.code >= 0:
        return bc.code
    raise ErrorException(L.GB_ERR_ARG, f"not a boundary condition: {bc!r}")


def _it(it):
    if isinstance(it, type) and issubclass(it, InterpType) and it.code >= 0:
        return it.code
    raise ErrorException(L.GB_ERR_ARG, f"not an interpolation type: {it!r}")


def wrap(bc, pos, length):  # src/boundaryconditions.jl:22-26 (pure integer rule, host side)
    b = _bc(bc)
    if b == 3:
        r = (pos - 1) % (2 * length)
        return r + 1 if r < length else 2 * length - r
    if b == 4:
        return (pos - 1) % length + 1
    if b in (5, 6):
        return 1 if pos < 1 else (length if pos > length else pos)
    return pos


# ---- array plumbing -------------------------------------------------------------------------------

_NP2GB = {np.dtype(np.float32): L.GB_F32, np.dtype(np.float64): L.GB_F64}


def _torch():
    import torch

    return torch


def _is_tensor(a):
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


def jl_empty(shape, dtype, device="cuda"):
    """A device array with Julia (column-major) layout and the given Julia size."""
    torch = _torch()
    shape = tuple(int(s) for s in shape)
    t = torch.empty(shape[::-1], dtype=dtype, device=device)
    return t.permute(*range(len(shape) - 1, -1, -1))


def jl_from_numpy(a, device="cuda"):
    torch = _torch()
    a = np.asarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)  # .T of F-order data is C-contiguous
    return t.permute(*range(a.ndim - 1, -1, -1))


def jl_to_numpy(t):
    n = t.dim()
    return np.asfortranarray(t.permute(*range(n - 1, -1, -1)).contiguous().cpu().numpy().T)


def _is_colmajor(t):
    n = t.dim()
    return n <= 1 and t.is_contiguous() or t.permute(*range(n - 1, -1, -1)).is_contiguous()


class _Arr:
    """Pointer + metadata of a host (numpy) or device (torch.cuda) array in Julia layout."""

    def __init__(self, a, writable=False, name="array"):
        if _is_tensor(a):
            torch = _torch()
            if not a.is_cuda:
                raise ErrorException(L.GB_ERR_ARG, f"{name}: torch tensors must live on a CUDA device (use numpy for host data)")
            if a.dtype not in (torch.float32, torch.float64):
                raise ErrorException(L.GB_ERR_ARG, f"{name}: only Float32/Float64 are supported")
            if not _is_colmajor(a):
                raise ErrorException(L.GB_ERR_ARG, f"{name}: device arrays must be column-major (see jl_empty / jl_from_numpy)")
            self.mem, self.obj = L.GB_DEVICE, a
            self.dtype = L.GB_F32 if a.dtype == torch.float32 else L.GB_F64
            self.shape = tuple(a.shape)
            self.ptr = a.data_ptr()
            self.device = a.device
        else:
            a0 = a
            a = np.asarray(a)
            if a.dtype not in _NP2GB:
                if writable:
                    raise ErrorException(L.GB_ERR_ARG, f"{name}: only Float32/Float64 are supported")
                a = a.astype(np.float64)
            f = np.asfortranarray(a)
            if writable and f is not a and not (f.ctypes.data == a.ctypes.data):
                raise ErrorException(L.GB_ERR_ARG, f"{name}: in-place operation needs a column-major (Fortran-ordered) array")
            del a0
            self.mem, self.obj = L.GB_HOST, f
            self.dtype = _NP2GB[f.dtype]
            self.shape = tuple(f.shape)
            self.ptr = f.ctypes.data
            self.device = None

    @property
    def ndim(self):
        return len(self.shape)

    def stream(self):
        if self.mem == L.GB_DEVICE:
            return _torch().cuda.current_stream(self.device).cuda_stream
        return None

    def empty_like_shape(self, shape):
        """New array in the same memory space / dtype with Julia size `shape`."""
        if self.mem == L.GB_DEVICE:
            return jl_empty(shape, self.obj.dtype, self.device)
        return np.empty(tuple(int(s) for s in shape), dtype=self.obj.dtype, order="F")


def _i64s(v):
    return (C.c_int64 * len(v))(*[int(x) for x in v])


# ---- ranges for CoordInterpGrid (src/coord.jl) ------------------------------------------------------


class FloatRange:
    """Julia 0.4's FloatRange{Float64}(start, step, len, divisor): element i is (start+i*step)/divisor."""

    kind = 0

    def __init__(self, start, step, len, divisor=1.0):
        self.start, self.step, self.len, self.divisor = float(start), float(step), int(len), float(divisor)

    def __len__(self):
        return self.len

    def __getitem__(self, i):
        return (self.start + i * self.step) / self.divisor

    def values(self):
        return np.array([self[i] for i in range(self.len)])

    def row(self):
        return [0.0, self.start, self.step, self.divisor]


class StepRange:
    """start:step:stop over integers (coordlookup src/coord.jl:31)."""

    kind = 1

    def __init__(self, start, step, stop):
        self.start, self.step = int(start), int(step)
        self.len = max(0, (int(stop) - self.start) // self.step + 1)

    def __len__(self):
        return self.len

    def row(self):
        return [1.0, float(self.start), float(self.step), 1.0]


class UnitRange(StepRange):
    kind = 2

    def __init__(self, start, stop):
        super().__init__(start, 1, stop)

    def row(self):
        return [2.0, float(self.start), 1.0, 1.0]


def _rat(x):
    """Julia Base `rat` (twiceprecision-free 0.4 version): continued-fraction rational of a float."""
    y = x
    a = d = 1
    b = c = 0
    m = 16777216.0  # maxintfloat(Float32)
    while abs(y) <= m:
        f = math.trunc(y)
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > int(m):
            return c, d
        if b != 0 and float(a) / float(b) == x:
            break
        if y == 0:
            break
        y = 1.0 / y
    return a, b


def frange(start, step, stop):
    """`start:step:stop` for Float64 as Julia 0.4 builds it (colon with float-range "lifting")."""
    start, step, stop = float(start), float(step), float(stop)
    if step == 0:
        raise ValueError("range step cannot be zero")
    if start == stop:
        return FloatRange(start, step, 1, 1.0)
    if (0 < step) != (start < stop):
        return FloatRange(start, step, 0, 1.0)
    r = (stop - start) / step
    n = round(r)  # ties to even, like Julia
    lo = np.nextafter((np.nextafter(stop, -np.inf) - np.nextafter(start, np.inf)) / n, -np.inf)
    hi = np.nextafter((np.nextafter(stop, np.inf) - np.nextafter(start, -np.inf)) / n, np.inf)
    if lo <= step <= hi:
        a0, b = _rat(start)
        a = float(a0)
        if b != 0 and a / float(b) == start:
            c0, d = _rat(step)
            c = float(c0)
            if d != 0 and c / float(d) == step:
                e = b * d // math.gcd(b, d)
                a *= e // b
                c *= e // d
                if (a + n * c) / float(e) == stop:
                    return FloatRange(a, c, n + 1, float(e))
    return FloatRange(start, step, math.floor(r) + 1, 1.0)


def _as_range(r):
    if isinstance(r, (FloatRange, StepRange)):
        return r
    if isinstance(r, range):
        return UnitRange(r.start, r.stop - 1) if r.step == 1 else StepRange(r.start, r.step, r.stop - (1 if r.step > 0 else -1))
    raise ErrorException(L.GB_ERR_ARG, "coordinates must be FloatRange / StepRange / UnitRange / range objects")


# ---- query batches -----------------------------------------------------------------------------------


class _Queries:
    def __init__(self, X, N):
        self.cols = None
        if isinstance(X, (tuple, list)) and len(X) > 0 and (np.ndim(X[0]) == 1 or (_is_tensor(X[0]) and X[0].dim() == 1)):
            if len(X) != N:
                raise ErrorException(L.GB_ERR_ARG, "Incorrect number of dimensions supplied")
            if _is_tensor(X[0]):
                torch = _torch()
                dt = torch.float32 if all(c.dtype == torch.float32 for c in X) else torch.float64
                self.cols = [c.to(dt).contiguous() for c in X]
                self.mem, self.device = L.GB_DEVICE, X[0].device
                self.xdtype = L.GB_F32 if dt == torch.float32 else L.GB_F64
                self.nq = int(self.cols[0].shape[0])
                ptrs = [c.data_ptr() for c in self.cols]
            else:
                arrs = [np.asarray(c) for c in X]
                dt = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64
                self.cols = [np.ascontiguousarray(a, dtype=dt) for a in arrs]
                self.mem, self.device = L.GB_HOST, None
                self.xdtype = _NP2GB[np.dtype(dt)]
                self.nq = int(self.cols[0].shape[0])
                ptrs = [c.ctypes.data for c in self.cols]
            if any(int(c.shape[0]) != self.nq for c in self.cols):
                raise ErrorException(L.GB_ERR_ARG, "coordinate vectors must have equal lengths")
            self.ptrs = (C.c_void_p * N)(*ptrs)
            return
        if _is_tensor(X):
            torch = _torch()
            if not X.is_cuda:
                raise ErrorException(L.GB_ERR_ARG, "query tensors must live on a CUDA device")
            if X.dtype not in (torch.float32, torch.float64):
                X = X.to(torch.float64)
            if X.dim() == 1 and N == 1:
                X = X.reshape(-1, 1)
            if X.dim() != 2 or X.shape[1] != N:
                raise ErrorException(L.GB_ERR_ARG, "Incorrect number of dimensions supplied")
            self.X = X.contiguous()
            self.mem, self.device = L.GB_DEVICE, X.device
            self.xdtype = L.GB_F32 if X.dtype == torch.float32 else L.GB_F64
            self.nq, self.ptr = int(X.shape[0]), self.X.data_ptr()
        else:
            X = np.asarray(X)
            if X.dtype not in _NP2GB:
                X = X.astype(np.float64)
            if X.ndim == 1 and N == 1:
                X = X.reshape(-1, 1)
            if X.ndim != 2 or X.shape[1] != N:
                raise ErrorException(L.GB_ERR_ARG, "Incorrect number of dimensions supplied")
            self.X = np.ascontiguousarray(X)
            self.mem, self.device = L.GB_HOST, None
            self.xdtype = _NP2GB[self.X.dtype]
            self.nq, self.ptr = int(X.shape[0]), self.X.ctypes.data

    def stream(self):
        if self.mem == L.GB_DEVICE:
            return _torch().cuda.current_stream(self.device).cuda_stream
        return None


def _out_array(q, shape, dtype_code):
    if q.mem == L.GB_DEVICE:
        torch = _torch()
        return torch.empty(shape, dtype=torch.float32 if dtype_code == L.GB_F32 else torch.float64, device=q.device)
    return np.empty(shape, dtype=np.float32 if dtype_code == L.GB_F32 else np.float64)


def _out_ptr(a, q, shape, dtype_code, what):
    """Validate a caller-provided output buffer; returns its pointer."""
    if a is None:
        return None
    if _is_tensor(a):
        torch = _torch()
        ok = q.mem == L.GB_DEVICE and a.is_cuda and a.is_contiguous() and a.dtype == (torch.float32 if dtype_code == L.GB_F32 else torch.float64)
        if not ok or tuple(a.shape) != tuple(shape):
            raise ErrorException(L.GB_ERR_ARG, f"Wrong number of components for the {what}")
        return a.data_ptr()
    if q.mem != L.GB_HOST or not isinstance(a, np.ndarray) or not a.flags.c_contiguous or a.dtype != (np.float32 if dtype_code == L.GB_F32 else np.float64):
        raise ErrorException(L.GB_ERR_ARG, f"{what}: output must be a C-contiguous array of the grid's element type in the queries' memory space")
    if tuple(a.shape) != tuple(shape):
        raise ErrorException(L.GB_ERR_ARG, f"Wrong number of components for the {what}")
    return a.ctypes.data


# ---- InterpGrid ------------------------------------------------------------------------------------------


class AbstractInterpGrid:
    pass


class InterpGrid(AbstractInterpGrid):
    """InterpGrid(A, BC, IT) or InterpGrid(A, fill::Number, IT) (src/interp.jl:42-72).

    A: numpy array or column-major CUDA tensor.  The padded, prefiltered coefficients live on the
    current CUDA device for the lifetime of the object."""

    def __init__(self, A, bc, it, _from_coefs=False, fast_prefilter=False):
        """fast_prefilter: run interp_invert! with the line-parallel GB_FAST solver where it applies (grids with few long
        lines; a few ulp of max|A| instead of bit-identical coefficients)."""
        lib = L.lib()
        order = _it(it)
        fill = float("nan")
        if isinstance(bc, (int, float, np.integer, np.floating)) and not isinstance(bc, bool):
            fill, bcode, bc = float(bc), 6, BCfill  # the Number constructor selects BCfill (:66-72)
        else:
            bcode = _bc(bc)
            if bcode == 6 and not _from_coefs:
                raise ErrorException(L.GB_ERR_ARG, "Construct BCfill InterpGrids by supplying the fill value")  # :49-51
        a = _Arr(A, name="A")
        self.BC, self.IT = bc, it
        self._h = C.c_void_p()
        if _from_coefs:
            L.check(lib.gb_grid_create_from_coefs(C.byref(self._h), a.dtype, a.ndim, _i64s(a.shape), a.ptr, a.mem, bcode, order, fill, a.stream()))
        else:
            L.check(lib.gb_grid_create_opt(C.byref(self._h), a.dtype, a.ndim, _i64s(a.shape), a.ptr, a.mem, bcode, order, fill,
                                           L.GB_FAST if fast_prefilter else L.GB_EXACT, a.stream()))
        self._lib = lib
        self._read_info(fill)

    def _read_info(self, fill):
        lib = self._lib
        dt, nd = C.c_int(), C.c_int()
        ud, sd = (C.c_int64 * 8)(), (C.c_int64 * 8)()
        L.check(lib.gb_grid_info(self._h, C.byref(dt), C.byref(nd), ud, sd, None, None, None))
        self.ndim = nd.value
        self.dtype_code = dt.value
        self.eltype = np.float32 if dt.value == L.GB_F32 else np.float64
        self._size = tuple(ud[i] for i in range(self.ndim))
        self._stored = tuple(sd[i] for i in range(self.ndim))
        self.fillval = fill
        self.affine = None

    @classmethod
    def _adopt(cls, handle, bc, it, fill=float("nan")):
        """Wrap an existing gb_grid handle (gb_grid_replicate); the object owns it from now on."""
        g = cls.__new__(cls)
        g.BC, g.IT = bc, it
        g._h, g._lib = handle, L.lib()
        g._read_info(fill)
        return g

    @classmethod
    def from_coefs(cls, coefs, bc, it, fill=float("nan")):
        """Adopt padded + prefiltered coefficients (e.g. broadcast from another GPU)."""
        g = cls.__new__(cls)
        if isinstance(bc, type) and bc is BCfill:
            InterpGrid.__init__(g, coefs, BCfill, it, _from_coefs=True)
            g.fillval = fill
        else:
            InterpGrid.__init__(g, coefs, bc, it, _from_coefs=True)
        return g

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.gb_grid_destroy(h)
            except Exception:
                pass
            self._h = None

    # size(G) (src/interp.jl:853-856)
    @property
    def shape(self):
        return self._size

    def size(self, i=None):
        return self._size if i is None else self._size[i - 1]

    @property
    def stored_shape(self):
        return self._stored

    @property
    def coefs(self):
        """G.coefs downloaded to the host (Julia layout)."""
        out = np.empty(self._stored, dtype=self.eltype, order="F")
        L.check(self._lib.gb_grid_coefs_download(self._h, out.ctypes.data))
        return out

    def coefs_device(self):
        """Borrowed column-major CUDA view of the resident coefficients (zero copy)."""
        torch = _torch()
        p = C.c_void_p()
        rc = self._lib.gb_grid_coefs_device(self._h, C.byref(p))
        if rc == L.GB_ERR_UNSUPPORTED:  # resident layout differs from the reference's (blocked): hand out a copy
            out = jl_empty(self._stored, torch.float32 if self.dtype_code == L.GB_F32 else torch.float64)
            L.check(self._lib.gb_grid_coefs_copy(self._h, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            return out
        L.check(rc)
        n = int(np.prod(self._stored))
        esz = 4 if self.dtype_code == L.GB_F32 else 8

        class _Holder:  # __cuda_array_interface__ provider; keeps the grid alive
            pass

        h = _Holder()
        h.owner = self
        h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4" if esz == 4 else "<f8", "data": (p.value, False), "version": 3}
        flat = torch.as_tensor(h, device="cuda")
        return flat.reshape(self._stored[::-1]).permute(*range(self.ndim - 1, -1, -1))

    # -- evaluation core --
    def _eval(self, X, mask, val=None, grad=None, hess=None, fast=False, affine=None, tiled=None, sorted=False):
        N = self.ndim
        q = X if isinstance(X, _Queries) else _Queries(X, N)
        dt = self.dtype_code
        if val is None:
            val = _out_array(q, (q.nq,), dt)
        pv = _out_ptr(val, q, (q.nq,), dt, "value")
        pg = ph = None
        if mask & L.GB_OUT_GRAD:
            if grad is None:
                grad = _out_array(q, (q.nq, N), dt)
            pg = _out_ptr(grad, q, (q.nq, N), dt, "gradient")
        if mask & L.GB_OUT_HESS:
            if hess is None:
                hess = _out_array(q, (q.nq, N, N), dt)
            ph = _out_ptr(hess, q, (q.nq, N, N), dt, "hessian")
        aff = None
        if affine is not None:
            aff = (C.c_double * (4 * N))(*[v for row in affine for v in row])
        bad = C.c_int64(-1)
        pbad = C.byref(bad) if self.BC is BCnil else None
        flags = (L.GB_FAST if fast else L.GB_EXACT) | (0 if tiled is None else (L.GB_TILED if tiled else L.GB_PLAIN)) | (L.GB_SORTED if sorted else 0)
        if q.cols is not None:
            rc = self._lib.gb_eval_soa(self._h, mask, q.nq, q.ptrs, q.xdtype, aff, pv, pg, ph, q.mem, flags, pbad, q.stream())
        else:
            rc = self._lib.gb_eval(self._h, mask, q.nq, q.ptr, q.xdtype, N, 1, aff, pv, pg, ph, q.mem, flags, pbad, q.stream())
        L.check(rc, bad.value)
        return val, grad, hess

    def _eval_packed(self, X, mask, out=None, fast=False, affine=None, tiled=None, sorted=False):
        """GB_OUT_PACKED: one record per query -- value, gradient, Hessian as in `mask` -- in an (nq, W) array (a Julia
        W x nq Matrix); W = 1 [+ N [+ N*N]]."""
        N = self.ndim
        q = X if isinstance(X, _Queries) else _Queries(X, N)
        dt = self.dtype_code
        W = 1 + (N if mask & (L.GB_OUT_GRAD | L.GB_OUT_HESS) else 0) + (N * N if mask & L.GB_OUT_HESS else 0)
        nch = getattr(self, "nchan", 1)
        shape = (q.nq, W) if nch == 1 else (q.nq, nch, W)
        if out is None:
            out = _out_array(q, shape, dt)
        po = _out_ptr(out, q, shape, dt, "packed output")
        if mask & L.GB_OUT_HESS:
            mask |= L.GB_OUT_GRAD
        aff = None
        if affine is not None:
            aff = (C.c_double * (4 * N))(*[v for row in affine for v in row])
        bad = C.c_int64(-1)
        pbad = C.byref(bad) if self.BC is BCnil else None
        flags = (L.GB_FAST if fast else L.GB_EXACT) | (0 if tiled is None else (L.GB_TILED if tiled else L.GB_PLAIN)) | (L.GB_SORTED if sorted else 0)
        m = mask | L.GB_OUT_PACKED
        if q.cols is not None:
            rc = self._lib.gb_eval_soa(self._h, m, q.nq, q.ptrs, q.xdtype, aff, po, None, None, q.mem, flags, pbad, q.stream())
        else:
            rc = self._lib.gb_eval(self._h, m, q.nq, q.ptr, q.xdtype, N, 1, aff, po, None, None, q.mem, flags, pbad, q.stream())
        L.check(rc, bad.value)
        return out

    def _scalar_args(self, x):
        N = self.ndim
        if N == 1 and len(x) == 2:  # src/interp.jl:147-153
            if x[1] != 1:
                raise BoundsError(L.GB_ERR_ARG, "BoundsError")
            x = x[:1]
        if len(x) != N:
            raise ErrorException(L.GB_ERR_ARG, "Incorrect number of dimensions supplied")  # :88-90
        return np.asarray([x], dtype=np.float64)

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        if all(np.ndim(i) == 0 and not _is_tensor(i) for i in idx):
            v, _, _ = self._eval(self._scalar_args(idx), L.GB_OUT_VAL)
            return v[0]
        return self._outer(idx)

    def _outer(self, vecs, fast=False, affine=None):
        """Tensor-product getindex (src/interp.jl:273-306)."""
        N = self.ndim
        if len(vecs) == 1 and N > 1:
            raise ErrorException(L.GB_ERR_ARG, "Linear indexing not supported")  # :281
        if len(vecs) != N:
            raise ErrorException(L.GB_ERR_ARG, "Dimensionality mismatch")  # :293-295
        dev = _is_tensor(vecs[0])
        if dev:
            torch = _torch()
            dt = torch.float32 if all(v.dtype == torch.float32 for v in vecs) else torch.float64
            cols = [v.to(dt).contiguous() for v in vecs]
            ptrs = [c.data_ptr() for c in cols]
            xd = L.GB_F32 if dt == torch.float32 else L.GB_F64
            out = jl_empty([c.shape[0] for c in cols], torch.float32 if self.dtype_code == L.GB_F32 else torch.float64, cols[0].device)
            optr, mem, stream = out.data_ptr(), L.GB_DEVICE, torch.cuda.current_stream(cols[0].device).cuda_stream
        else:
            arrs = [np.atleast_1d(np.asarray(v)) for v in vecs]
            dt = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64
            cols = [np.ascontiguousarray(a, dtype=dt) for a in arrs]
            ptrs = [c.ctypes.data for c in cols]
            xd = _NP2GB[np.dtype(dt)]
            out = np.empty([c.shape[0] for c in cols], dtype=self.eltype, order="F")
            optr, mem, stream = out.ctypes.data, L.GB_HOST, None
        aff = None
        if affine is not None:
            aff = (C.c_double * (4 * N))(*[v for row in affine for v in row])
        lens = _i64s([c.shape[0] for c in cols])
        L.check(self._lib.gb_eval_outer(self._h, lens, (C.c_void_p * N)(*ptrs), xd, aff, optr, mem, L.GB_FAST if fast else L.GB_EXACT, stream))
        return out


class InterpGridChannels(InterpGrid):
    """Multi-valued InterpGrid: A has one trailing channel dimension, A[..., c] is channel c.

    The reference serves multi-valued functions (RGB images, spectra) through its low-level
    interface -- `set_position` once, `interp(ic, A, index)` per channel (README.md:151-209,
    src/interp.jl:390-429,489-508) -- so that the taps and weights of a location are computed once.
    This is the batched form: every evaluation returns one value (gradient, Hessian) per channel,
    `val[q, c]`, `grad[q, c, :]`, `hess[q, c, :, :]`, each identical to what InterpGrid(A[..., c], BC, IT)
    returns.  `InterpGridChannels.from_coefs(coefs, bc, it, raw=True)` adopts coefficients prepared
    with `interp_invert_(A, bc, it, dimlist=1:N)` and takes ARRAY index coordinates like
    `set_position` does (no setx shift)."""

    def __init__(self, A, bc, it, _from_coefs=False, _raw=False):
        lib = L.lib()
        order = _it(it)
        fill = float("nan")
        if isinstance(bc, (int, float, np.integer, np.floating)) and not isinstance(bc, bool):
            fill, bcode, bc = float(bc), 6, BCfill
        else:
            bcode = _bc(bc)
            if bcode == 6 and not _from_coefs:
                raise ErrorException(L.GB_ERR_ARG, "Construct BCfill InterpGrids by supplying the fill value")
        a = _Arr(A, name="A")
        if a.ndim < 2:
            raise ErrorException(L.GB_ERR_ARG, "InterpGridChannels needs at least one interpolated dimension and the channel dimension")
        nd, nchan = a.ndim - 1, int(a.shape[-1])
        self.BC, self.IT = bc, it
        self._h = C.c_void_p()
        if _from_coefs:
            L.check(lib.gb_grid_create_from_coefs_channels(C.byref(self._h), a.dtype, nd, _i64s(a.shape[:-1]), nchan, a.ptr, a.mem, bcode, order,
                                                           fill, 1 if _raw else 0, a.stream()))
        else:
            L.check(lib.gb_grid_create_channels(C.byref(self._h), a.dtype, nd, _i64s(a.shape[:-1]), nchan, a.ptr, a.mem, bcode, order, fill, a.stream()))
        self._lib = lib
        dt, ndc = C.c_int(), C.c_int()
        ud, sd = (C.c_int64 * 8)(), (C.c_int64 * 8)()
        L.check(lib.gb_grid_info(self._h, C.byref(dt), C.byref(ndc), ud, sd, None, None, None))
        self.ndim = ndc.value
        self.nchan = nchan
        self.dtype_code = dt.value
        self.eltype = np.float32 if dt.value == L.GB_F32 else np.float64
        self._size = tuple(ud[i] for i in range(self.ndim))
        self._stored = tuple(sd[i] for i in range(self.ndim))
        self.fillval = fill
        self.affine = None

    @classmethod
    def from_coefs(cls, coefs, bc, it, raw=False):
        g = cls.__new__(cls)
        InterpGridChannels.__init__(g, coefs, bc, it, _from_coefs=True, _raw=raw)
        return g

    @property
    def coefs(self):
        out = np.empty(self._stored + (self.nchan,), dtype=self.eltype, order="F")
        L.check(self._lib.gb_grid_coefs_download(self._h, out.ctypes.data))
        return out

    def coefs_device(self):
        raise MethodError(L.GB_ERR_UNSUPPORTED, "coefs_device is not provided for multi-valued grids")

    def _eval(self, X, mask, val=None, grad=None, hess=None, fast=False, affine=None, tiled=None):
        N, Cn = self.ndim, self.nchan
        q = X if isinstance(X, _Queries) else _Queries(X, N)
        dt = self.dtype_code
        if val is None:
            val = _out_array(q, (q.nq, Cn), dt)
        pv = _out_ptr(val, q, (q.nq, Cn), dt, "value")
        pg = ph = None
        if mask & L.GB_OUT_GRAD:
            if grad is None:
                grad = _out_array(q, (q.nq, Cn, N), dt)
            pg = _out_ptr(grad, q, (q.nq, Cn, N), dt, "gradient")
        if mask & L.GB_OUT_HESS:
            if hess is None:
                hess = _out_array(q, (q.nq, Cn, N, N), dt)
            ph = _out_ptr(hess, q, (q.nq, Cn, N, N), dt, "hessian")
        aff = None
        if affine is not None:
            aff = (C.c_double * (4 * N))(*[v for row in affine for v in row])
        bad = C.c_int64(-1)
        pbad = C.byref(bad) if self.BC is BCnil else None
        flags = L.GB_FAST if fast else L.GB_EXACT
        if q.cols is not None:
            rc = self._lib.gb_eval_soa(self._h, mask, q.nq, q.ptrs, q.xdtype, aff, pv, pg, ph, q.mem, flags, pbad, q.stream())
        else:
            rc = self._lib.gb_eval(self._h, mask, q.nq, q.ptr, q.xdtype, N, 1, aff, pv, pg, ph, q.mem, flags, pbad, q.stream())
        L.check(rc, bad.value)
        return val, grad, hess

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        v, _, _ = self._eval(self._scalar_args(idx), L.GB_OUT_VAL)
        return v[0]


class InterpIrregular(AbstractInterpGrid):
    """InterpIrregular(knots, A, BC, IT) / InterpIrregular(knots, A, fill::Number, IT) (src/interp.jl:318-350):
    1-D interpolation on sorted, non-uniform knots; IT in InterpForward / InterpBackward / InterpNearest /
    InterpLinear; BCreflect and BCperiodic are rejected like in the reference.  `G[x]` for a scalar,
    `G[xs]` / `getindex_batch(G, xs)` for a vector (host array or CUDA tensor)."""

    def __init__(self, knots, A, bc, it):
        if isinstance(knots, (tuple, list)) and len(knots) > 0 and np.ndim(knots[0]) == 1:
            if len(knots) != 1:
                raise ErrorException(L.GB_ERR_UNSUPPORTED, "Sorry, for now only 1d is supported")  # :332-334
            knots = knots[0]
        self.fillval = float("nan")
        if isinstance(bc, (int, float, np.integer, np.floating)) and not isinstance(bc, bool):
            self.fillval, bc = float(bc), BCfill  # :345-349
        code = _bc(bc)
        if code in (3, 4):
            raise ErrorException(L.GB_ERR_UNSUPPORTED, "Sorry, reflecting or periodic boundary conditions not yet supported")  # :335-337
        self.order = it.code if isinstance(it, type) else it().code
        if self.order not in (0, 1, 4, 5):
            raise MethodError(L.GB_ERR_UNSUPPORTED, "InterpIrregular: InterpForward, InterpBackward, InterpNearest or InterpLinear")
        self.BC, self.IT, self.bcode, self.ndim = bc, it, code, 1
        a = np.asarray(A)
        dt = np.float32 if (a.dtype == np.float32 and np.asarray(knots).dtype == np.float32) else np.float64
        self.eltype = dt
        self.grid = np.ascontiguousarray(knots, dtype=dt).copy()
        self.coefs = np.ascontiguousarray(a, dtype=dt).copy()
        if self.grid.ndim != 1 or self.coefs.shape != self.grid.shape:
            raise DimensionMismatch(L.GB_ERR_DIM, "InterpIrregular: knots and samples must be vectors of equal length")
        if np.any(self.grid[1:] < self.grid[:-1]):
            raise ErrorException(L.GB_ERR_ARG, "Coordinates must be supplied in increasing order")  # :340-342
        self._dev = None

    @property
    def shape(self):
        return self.coefs.shape

    def _device_arrays(self, device):
        if self._dev is None or self._dev[0].device != device:
            torch = _torch()
            self._dev = (torch.from_numpy(self.grid).to(device), torch.from_numpy(self.coefs).to(device))
        return self._dev

    def _eval(self, x):
        lib = L.lib()
        dtc = L.GB_F32 if self.eltype == np.float32 else L.GB_F64
        bad = C.c_int64(-1)
        if _is_tensor(x):
            torch = _torch()
            xt = x.to(torch.float32 if self.eltype == np.float32 else torch.float64).contiguous().reshape(-1)
            g, c = self._device_arrays(xt.device)
            out = torch.empty_like(xt)
            rc = lib.gb_interp_irregular(dtc, self.grid.shape[0], g.data_ptr(), c.data_ptr(), self.bcode, self.order, self.fillval, xt.numel(),
                                         xt.data_ptr(), out.data_ptr(), L.GB_DEVICE, C.byref(bad), torch.cuda.current_stream(xt.device).cuda_stream)
        else:
            xa = np.ascontiguousarray(np.atleast_1d(x), dtype=self.eltype)
            out = np.empty_like(xa)
            rc = lib.gb_interp_irregular(dtc, self.grid.shape[0], self.grid.ctypes.data, self.coefs.ctypes.data, self.bcode, self.order, self.fillval,
                                         xa.shape[0], xa.ctypes.data, out.ctypes.data, L.GB_HOST, C.byref(bad), None)
        if rc == L.GB_ERR_INVALID_LOCATION:
            raise BoundsError(rc, f"BoundsError (query {bad.value})")  # :356-360
        L.check(rc)
        return out

    def __getitem__(self, x):
        if np.ndim(x) == 0 and not _is_tensor(x):
            return self._eval([x])[0]
        return self._eval(x)


def _grid_of(G):
    return G.grid if isinstance(G, CoordInterpGrid) else G


# ---- scalar valgrad / valgradhess (src/interp.jl:158-270, src/coord.jl:58-69) -----------------------------


def valgrad(*args):
    """valgrad(G, x...) -> (val, g)   or   valgrad(g, G, x...) -> val  (g filled in place)."""
    if isinstance(args[0], AbstractInterpGrid):
        G, x = args[0], args[1:]
        ig = _grid_of(G)
        v, g, _ = ig._eval(ig._scalar_args(x), L.GB_OUT_VAL | L.GB_OUT_GRAD, affine=G.affine)
        return (v[0], g[0, 0]) if ig.ndim == 1 else (v[0], g[0].copy())
    gbuf, G, x = args[0], args[1], args[2:]
    ig = _grid_of(G)
    if len(gbuf) != ig.ndim:
        raise ErrorException(L.GB_ERR_ARG, "Wrong number of components for the gradient")  # :159-161
    v, g, _ = ig._eval(ig._scalar_args(x), L.GB_OUT_VAL | L.GB_OUT_GRAD, affine=G.affine)
    gbuf[:] = g[0]
    return v[0]


def valgradhess(*args):
    """valgradhess(G, x...) -> (val, g, h)   or   valgradhess(g, h, G, x...) -> val."""
    if isinstance(args[0], AbstractInterpGrid):
        G, x = args[0], args[1:]
        if isinstance(G, CoordInterpGrid):
            raise MethodError(L.GB_ERR_UNSUPPORTED, "valgradhess has no method for CoordInterpGrid")
        v, g, h = G._eval(G._scalar_args(x), L.GB_OUT_VAL | L.GB_OUT_GRAD | L.GB_OUT_HESS)
        if G.ndim == 1:
            return v[0], g[0, 0], h[0, 0, 0]
        return v[0], g[0].copy(), h[0].T.copy()
    gbuf, hbuf, G, x = args[0], args[1], args[2], args[3:]
    if isinstance(G, CoordInterpGrid):
        raise MethodError(L.GB_ERR_UNSUPPORTED, "valgradhess has no method for CoordInterpGrid")
    if len(gbuf) != G.ndim:
        raise ErrorException(L.GB_ERR_ARG, "Wrong number of components for the gradient")  # :206-208
    if np.shape(hbuf) != (G.ndim, G.ndim):
        raise ErrorException(L.GB_ERR_ARG, "Wrong number of components for the hessian")  # :209-211
    v, g, h = G._eval(G._scalar_args(x), L.GB_OUT_VAL | L.GB_OUT_GRAD | L.GB_OUT_HESS)
    gbuf[:] = g[0]
    hbuf[:, :] = h[0].T
    return v[0]


# ---- batched entry points ----------------------------------------------------------------------------------


# `fast` selects the GB_FAST contraction; `tiled` (None = library heuristic, True / False = force)
# only changes scheduling (binned shared-memory bricks vs plain global gathers), never results.


def _irregular_batch(G, X, out=None):
    """InterpIrregular in the batched entry points: 1-D, value only (src/interp.jl:318-377)."""
    v = G._eval(X)
    if out is None:
        return v
    if _is_tensor(out):
        out.copy_(v if _is_tensor(v) else _torch().as_tensor(v))
    else:
        out[...] = v.cpu().numpy() if _is_tensor(v) else v
    return out


def getindex_(out, G, X, fast=False, tiled=None, sorted=False):
    if isinstance(G, InterpIrregular):
        return _irregular_batch(G, X, out)
    ig = _grid_of(G)
    ig._eval(X, L.GB_OUT_VAL, val=out, fast=fast, affine=G.affine, tiled=tiled, sorted=sorted)
    return out


def valgrad_(v, g, G, X, fast=False, tiled=None, sorted=False):
    ig = _grid_of(G)
    ig._eval(X, L.GB_OUT_VAL | L.GB_OUT_GRAD, val=v, grad=g, fast=fast, affine=G.affine, tiled=tiled, sorted=sorted)
    return v


def valgradhess_(v, g, h, G, X, fast=False, tiled=None, sorted=False):
    if isinstance(G, CoordInterpGrid):
        raise MethodError(L.GB_ERR_UNSUPPORTED, "valgradhess has no method for CoordInterpGrid")
    G._eval(X, L.GB_OUT_VAL | L.GB_OUT_GRAD | L.GB_OUT_HESS, val=v, grad=g, hess=h, fast=fast, tiled=tiled, sorted=sorted)
    return v


def valgrad_packed_(out, G, X, fast=False, tiled=None, sorted=False):
    """New batched layout (GB_OUT_PACKED): out[q] = (value, gradient...) of query q, one (N+1)-element record per query --
    a Julia (N+1) x nq Matrix.  One full-sector write per query; the tiled path skips its unpack pass."""
    return _grid_of(G)._eval_packed(X, L.GB_OUT_VAL | L.GB_OUT_GRAD, out=out, fast=fast, affine=G.affine, tiled=tiled, sorted=sorted)


def valgrad_packed(G, X, fast=False, tiled=None, sorted=False):
    return valgrad_packed_(None, G, X, fast=fast, tiled=tiled, sorted=sorted)


def getindex_batch(G, X, fast=False, tiled=None, sorted=False):
    if isinstance(G, InterpIrregular):
        return _irregular_batch(G, X)
    return _grid_of(G)._eval(X, L.GB_OUT_VAL, fast=fast, affine=G.affine, tiled=tiled, sorted=sorted)[0]


def valgrad_batch(G, X, fast=False, tiled=None, sorted=False):
    v, g, _ = _grid_of(G)._eval(X, L.GB_OUT_VAL | L.GB_OUT_GRAD, fast=fast, affine=G.affine, tiled=tiled, sorted=sorted)
    return v, g


def valgradhess_batch(G, X, fast=False, tiled=None, sorted=False):
    """h[q] is the symmetric N x N Hessian of query q."""
    if isinstance(G, CoordInterpGrid):
        raise MethodError(L.GB_ERR_UNSUPPORTED, "valgradhess has no method for CoordInterpGrid")
    return G._eval(X, L.GB_OUT_VAL | L.GB_OUT_GRAD | L.GB_OUT_HESS, fast=fast, tiled=tiled, sorted=sorted)


# ---- CoordInterpGrid (src/coord.jl:3-69) ----------------------------------------------------------------------


class CoordInterpGrid(AbstractInterpGrid):
    """CoordInterpGrid(coord, A, BC, IT) / CoordInterpGrid(coord, grid::InterpGrid)."""

    def __init__(self, coord, A, *args):
        grid = A if isinstance(A, InterpGrid) else InterpGrid(A, *args)
        if isinstance(coord, (FloatRange, StepRange, range)):
            coord = (coord,)
        coord = tuple(_as_range(r) for r in coord)
        if tuple(len(r) for r in coord) != tuple(grid.shape):
            raise DimensionMismatch(L.GB_ERR_DIM, "Coordinate lengths do not match grid size.")  # :7
        self.coord, self.grid = coord, grid
        self.affine = [r.row() for r in coord]
        self.ndim = grid.ndim

    @property
    def shape(self):
        return self.grid.shape

    def size(self, i=None):
        return self.grid.size(i)

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        g = self.grid
        if all(np.ndim(i) == 0 and not _is_tensor(i) for i in idx):
            v, _, _ = g._eval(g._scalar_args(idx), L.GB_OUT_VAL, affine=self.affine)
            return v[0]
        return g._outer(idx, affine=self.affine)


# ---- prefilter / padding ----------------------------------------------------------------------------------------


def filledges_(A, val):
    """filledges!(A, val) (src/utilities.jl:22-33): every element on a face of the array is set to val, in place."""
    a = _Arr(A, writable=True, name="A")
    L.check(L.lib().gb_filledges(a.dtype, a.ndim, _i64s(a.shape), a.ptr, float(val), a.mem, a.stream()))
    if a.mem == L.GB_HOST and isinstance(A, np.ndarray) and a.obj is not A:
        A[...] = a.obj
    return A


def interp_invert_(A, bc, it, dimlist=None, fast=False):
    """interp_invert!(A, BC, IT[, dimlist]) in place (src/interp.jl:909-953); dimlist is 1-based.
    fast=True: the line-parallel GB_FAST solver where it applies (same solution to a few ulp of max|A|)."""
    lib = L.lib()
    a = _Arr(A, writable=True, name="A")
    if isinstance(bc, (int, float)) and not isinstance(bc, bool):
        bcode = 5  # a Number selects the BCnearest system, as the fill constructor does (:68)
    else:
        bcode = _bc(bc)
    if dimlist is None:
        dl, n = None, 0
    else:
        dimlist = [int(d) for d in (dimlist if np.ndim(dimlist) else [dimlist])]
        for d in dimlist:
            if not 1 <= d <= a.ndim:
                raise ErrorException(L.GB_ERR_ARG, "dimlist entry out of range")
        dl, n = (C.c_int * len(dimlist))(*[d - 1 for d in dimlist]), len(dimlist)
    L.check(lib.gb_invert_opt(a.dtype, a.ndim, _i64s(a.shape), a.ptr, a.mem, bcode, _it(it), dl, n, L.GB_FAST if fast else L.GB_EXACT, a.stream()))
    if a.mem == L.GB_HOST and isinstance(A, np.ndarray) and a.obj is not A:
        A[...] = a.obj  # 1-D / already-F-ordered views share memory; this covers harmless copies
    return A


def pad1(A, f, *dimpad):
    """pad1(A, f, 1:N) (src/interp.jl:1305-1310).  Only padding of every dimension is supported."""
    lib = L.lib()
    if isinstance(f, type) and issubclass(f, BoundaryCondition):
        if f is BCfill:
            raise ErrorException(L.GB_ERR_ARG, "For BCfill, specify the fill value instead of the type BCfill")  # :1304
        return A.copy() if isinstance(A, np.ndarray) else A.clone()  # :1303
    a = _Arr(A, name="A")
    out = a.empty_like_shape([s + 2 for s in a.shape])
    o = _Arr(out, name="out")
    L.check(lib.gb_pad1(a.dtype, a.ndim, _i64s(a.shape), a.ptr, float(f), o.ptr, a.mem, a.stream()))
    return out


# ---- restrict / prolong (src/restrict_prolong.jl) ------------------------------------------------------------------


def restrict_size(sz, flag=None):
    """restrict_size(len) or restrict_size(dims, flags) -> (ndims, ncols) schedule (:334-351)."""
    if flag is None:
        n = int(sz)
        return (n + 1) // 2 if n % 2 else n // 2 + 1
    flag = _flags(flag, len(sz))
    out = np.zeros(flag.shape, dtype=np.int64)
    prev = list(sz)
    for j in range(flag.shape[1]):
        for i in range(flag.shape[0]):
            out[i, j] = restrict_size(prev[i]) if flag[i, j] else prev[i]
        prev = list(out[:, j])
    return out


def prolong_size(sz, arg):
    """prolong_size(n, len) -> len (error on mismatch, :322-332) or prolong_size(dims, flags) (:352-355)."""
    if np.ndim(sz) == 0:
        n, ln = int(sz), int(arg)
        newsz = 2 * n - 1 if ln % 2 else 2 * (n - 1)
        if newsz != ln:
            raise ErrorException(L.GB_ERR_PROLONG_SIZE, f"Cannot prolong a dimension of length {n} to {ln}")
        return newsz
    rs = restrict_size(sz, arg)
    return np.concatenate([rs[:, -2::-1], np.asarray(sz, dtype=np.int64)[:, None]], axis=1)


def _flags(flag, ndims):
    flag = np.asarray(flag, dtype=bool)
    if flag.ndim == 1:
        flag = flag[:, None]
    if flag.shape[0] != ndims:
        raise ErrorException(L.GB_ERR_ARG, "Boolean flag must have as many rows as dimensions of A")  # utilities.jl:8-10
    return flag


def _restrict_dim(a, dim, scale):
    lib = L.lib()
    shp = list(a.shape)
    shp[dim] = restrict_size(shp[dim])
    out = a.empty_like_shape(shp)
    o = _Arr(out, name="out")
    L.check(lib.gb_restrict(a.dtype, a.ndim, _i64s(a.shape), a.ptr, dim, float(scale), o.ptr, a.mem, a.stream()))
    return out


def _prolong_dim(a, dim, length):
    lib = L.lib()
    prolong_size(a.shape[dim], length)
    shp = list(a.shape)
    shp[dim] = int(length)
    out = a.empty_like_shape(shp)
    o = _Arr(out, name="out")
    L.check(lib.gb_prolong(a.dtype, a.ndim, _i64s(a.shape), a.ptr, dim, int(length), o.ptr, a.mem, a.stream()))
    return out


def restrict(A, dim_or_flag, scale=1.0, fused=True):
    """restrict(A, dim[, scale]) (1-based dim) or restrict(A, flags[, scale]) via mapdim (:103-185).

    With `fused` a column of flags that is all-true on a 3-D array runs as one pass (gb_restrict3), and one that
    restricts dims 1 and 2 of a 2-D / 3-D array as one pass (gb_restrict12), bit-identical to the staged passes."""
    a = _Arr(A, name="A")
    if np.ndim(dim_or_flag) == 0 and not isinstance(dim_or_flag, (bool, np.bool_)):
        dim = int(dim_or_flag)
        if not 1 <= dim <= a.ndim:
            raise ErrorException(L.GB_ERR_ARG, "restrict: dim out of range")
        return _restrict_dim(a, dim - 1, scale)
    flag = _flags(dim_or_flag, a.ndim)
    cur, out = a, A
    lib = L.lib()
    if fused and a.ndim == 3 and flag.shape[1] > 1 and flag.all():
        # every column restricts every dimension: all levels in ONE library call (gb_restrict3_levels; the levels in between
        # never leave the device)
        shp = list(a.shape)
        for _ in range(flag.shape[1]):
            shp = [restrict_size(s) for s in shp]
        out = a.empty_like_shape(shp)
        o = _Arr(out, name="out")
        L.check(lib.gb_restrict3_levels(a.dtype, _i64s(a.shape), int(flag.shape[1]), a.ptr, float(scale), o.ptr, a.mem, a.stream()))
        return out
    for j in range(flag.shape[1]):
        if fused and cur.ndim == 3 and flag[:, j].all():
            shp = [restrict_size(s) for s in cur.shape]
            out = cur.empty_like_shape(shp)
            o = _Arr(out, name="out")
            L.check(lib.gb_restrict3(cur.dtype, _i64s(cur.shape), cur.ptr, float(scale), o.ptr, cur.mem, cur.stream()))
            cur = o
            continue
        if fused and cur.ndim in (2, 3) and flag[0, j] and flag[1, j] and min(cur.shape[:2]) >= 2:
            # dims 1 and 2 in one pass (gb_restrict12: 2-D arrays, or every plane of a 3-D array)
            shp = [restrict_size(cur.shape[0]), restrict_size(cur.shape[1])] + list(cur.shape[2:])
            out = cur.empty_like_shape(shp)
            o = _Arr(out, name="out")
            L.check(lib.gb_restrict12(cur.dtype, cur.ndim, _i64s(cur.shape), cur.ptr, float(scale), o.ptr, cur.mem, cur.stream()))
            cur = o
            continue
        for i in range(flag.shape[0]):
            if flag[i, j]:
                out = _restrict_dim(cur, i, scale)
                cur = _Arr(out, name="out")
    return out


def prolong(A, *args, fused=True):
    """prolong(A, dim, len) (1-based dim, :10-48) or prolong(A, sizes) (:87-100)."""
    a = _Arr(A, name="A")
    if len(args) == 2:
        dim, length = int(args[0]), int(args[1])
        if not 1 <= dim <= a.ndim:
            raise ErrorException(L.GB_ERR_ARG, "prolong: dim out of range")
        return _prolong_dim(a, dim - 1, length)
    sz = np.asarray(args[0], dtype=np.int64)
    if sz.ndim == 1:
        sz = sz[:, None]
    if sz.shape[0] != a.ndim:
        raise ErrorException(L.GB_ERR_ARG, "Array of sizes must have as many rows as dimensions of A")
    cur, out = a, A
    lib = L.lib()
    if fused and a.ndim == 3 and sz.shape[1] > 1:
        # several columns that each enlarge every dimension: all levels in ONE library call (gb_prolong3_levels)
        shp, ok = list(a.shape), True
        for j in range(sz.shape[1]):
            ok = ok and all(int(sz[i, j]) > shp[i] for i in range(3))
            if not ok:
                break
            shp = [prolong_size(shp[i], int(sz[i, j])) for i in range(3)]
        if ok:
            out = a.empty_like_shape(shp)
            o = _Arr(out, name="out")
            L.check(lib.gb_prolong3_levels(a.dtype, _i64s(a.shape), int(sz.shape[1]), _i64s([int(v) for v in sz.T.reshape(-1)]), a.ptr, o.ptr, a.mem, a.stream()))
            return out
    for j in range(sz.shape[1]):
        grow = [sz[i, j] > cur.shape[i] for i in range(cur.ndim)]
        if fused and cur.ndim == 3 and grow == [True, True, False]:  # every plane in one pass (gb_prolong12)
            m1, m2 = prolong_size(cur.shape[0], int(sz[0, j])), prolong_size(cur.shape[1], int(sz[1, j]))
            out = cur.empty_like_shape([m1, m2, cur.shape[2]])
            o = _Arr(out, name="out")
            L.check(lib.gb_prolong12(cur.dtype, 3, _i64s(cur.shape), cur.ptr, m1, m2, o.ptr, cur.mem, cur.stream()))
            cur = o
            continue
        if fused and cur.ndim == 3 and sum(grow) >= 2:
            for i in range(3):
                if grow[i]:
                    prolong_size(cur.shape[i], int(sz[i, j]))
            shp = [int(sz[i, j]) if grow[i] else cur.shape[i] for i in range(3)]
            out = cur.empty_like_shape(shp)
            o = _Arr(out, name="out")
            L.check(lib.gb_prolong3(cur.dtype, _i64s(cur.shape), cur.ptr, _i64s(shp), o.ptr, cur.mem, cur.stream()))
            cur = o
            continue
        if fused and cur.ndim == 2 and all(grow):  # both dimensions of a 2-D array in one pass (gb_prolong12)
            m1, m2 = prolong_size(cur.shape[0], int(sz[0, j])), prolong_size(cur.shape[1], int(sz[1, j]))
            out = cur.empty_like_shape([m1, m2])
            o = _Arr(out, name="out")
            L.check(lib.gb_prolong12(cur.dtype, 2, _i64s(cur.shape), cur.ptr, m1, m2, o.ptr, cur.mem, cur.stream()))
            cur = o
            continue
        for i in range(cur.ndim):
            if sz[i, j] > cur.shape[i]:
                out = _prolong_dim(cur, i, int(sz[i, j]))
                cur = _Arr(out, name="out")
    return out


# ---- misc --------------------------------------------------------------------------------------------------------


# ---- restrictb / prolongb / restrict_extrap (src/restrict_prolong.jl:187-320) ----------------------------


def _edge_dim(fn_name, a, dim, arg):
    lib = L.lib()
    shp = list(a.shape)
    if fn_name == "gb_prolongb":
        prolong_size(a.shape[dim], arg)
        shp[dim] = int(arg)
    else:
        shp[dim] = restrict_size(shp[dim])
    out = a.empty_like_shape(shp)
    o = _Arr(out, name="out")
    val = int(arg) if fn_name == "gb_prolongb" else float(arg)
    L.check(getattr(lib, fn_name)(a.dtype, a.ndim, _i64s(a.shape), a.ptr, dim, val, o.ptr, a.mem, a.stream()))
    return out


def _edge_restrict(fn_name, A, dim_or_flag, scale):
    a = _Arr(A, name="A")
    if np.ndim(dim_or_flag) == 0 and not isinstance(dim_or_flag, (bool, np.bool_)):
        dim = int(dim_or_flag)
        if not 1 <= dim <= a.ndim:
            raise ErrorException(L.GB_ERR_ARG, "dim out of range")
        return _edge_dim(fn_name, a, dim - 1, scale)
    flag = _flags(dim_or_flag, a.ndim)
    cur, out = a, A
    for j in range(flag.shape[1]):  # mapdim, src/utilities.jl:7-20
        for i in range(flag.shape[0]):
            if flag[i, j]:
                out = _edge_dim(fn_name, cur, i, scale)
                cur = _Arr(out, name="out")
    return out


def restrictb(A, dim_or_flag, scale=1.0):
    """restrictb(A, dim[, scale]) / restrictb(A, flags[, scale]): "balanced" restriction, edge planes
    weighted (2,1)*scale/sqrt(3) (odd length) or (3,1)*scale/(2 sqrt(2)) (even) (:196-229)."""
    return _edge_restrict("gb_restrictb", A, dim_or_flag, scale)


def restrict_extrap(A, dim_or_flag, scale=1.0):
    """restrict_extrap(A, dim[, scale]) / (A, flags[, scale]): restrict, then linear extrapolation at the
    edges (:287-320)."""
    return _edge_restrict("gb_restrict_extrap", A, dim_or_flag, scale)


def prolongb(A, *args):
    """prolongb(A, dim, len) (1-based dim) or prolongb(A, sizes) (:231-281)."""
    a = _Arr(A, name="A")
    if len(args) == 2:
        dim, length = int(args[0]), int(args[1])
        if not 1 <= dim <= a.ndim:
            raise ErrorException(L.GB_ERR_ARG, "prolongb: dim out of range")
        return _edge_dim("gb_prolongb", a, dim - 1, length)
    sz = np.asarray(args[0], dtype=np.int64)
    if sz.ndim == 1:
        sz = sz[:, None]
    if sz.shape[0] != a.ndim:
        raise ErrorException(L.GB_ERR_ARG, "Array of sizes must have as many rows as dimensions of A")
    cur, out = a, A
    for j in range(sz.shape[1]):
        for i in range(sz.shape[0]):
            if sz[i, j] > cur.shape[i]:
                out = _edge_dim("gb_prolongb", cur, i, int(sz[i, j]))
                cur = _Arr(out, name="out")
    return out


def launch_count():
    return int(L.lib().gb_launch_count())


PROFILE_STAGES = ("tile_hist", "tile_offsets", "tile_scatter", "eval_brick", "eval_deferred", "unpermute")


def profile_enable(on=True):
    """Record CUDA events around the stages of every tiled evaluation (gb_profile_enable)."""
    L.check(L.lib().gb_profile_enable(1 if on else 0))


def profile_read():
    """Stage durations (ms) of the last tiled evaluation as {stage: ms}, {} if none ran."""
    ms = (C.c_double * 8)()
    n = C.c_int(0)
    L.check(L.lib().gb_profile_read(ms, 8, C.byref(n)))
    out = {PROFILE_STAGES[i]: float(ms[i]) for i in range(n.value)}
    if n.value:
        out["coherent"] = float(ms[6])  # 1.0: the batch arrived spatially coherent and was evaluated in place by the plain kernel
    return out


def synth_uniform(out, seed, lo=0.0, hi=1.0, first=0):
    """Fill a CUDA tensor with the replayable synthetic stream."""
    torch = _torch()
    assert out.is_cuda and out.dtype in (torch.float32, torch.float64)
    flat_ok = out.is_contiguous() or _is_colmajor(out)
    assert flat_ok
    dt = L.GB_F32 if out.dtype == torch.float32 else L.GB_F64
    L.check(L.lib().gb_synth_uniform(dt, out.data_ptr(), out.numel(), int(seed), float(lo), float(hi), int(first), torch.cuda.current_stream(out.device).cuda_stream))
    return out
