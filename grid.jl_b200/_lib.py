"""ctypes binding of libgridb200.so (include/gridb200.h).

There is no fallback: if the CUDA library is missing or a symbol is absent, importing the compute
entry points raises.  Loading the library does not need a GPU (the symbol table can be inspected on
a CPU-only box); calling a compute entry point does.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgridb200.so")

GB_F32, GB_F64 = 0, 1
GB_HOST, GB_DEVICE = 0, 1
GB_OUT_VAL, GB_OUT_GRAD, GB_OUT_HESS, GB_OUT_PACKED = 1, 2, 4, 8
GB_EXACT, GB_FAST, GB_TILED, GB_PLAIN, GB_SORTED = 0, 1, 2, 4, 8
(GB_OK, GB_ERR_ARG, GB_ERR_INVALID_LOCATION, GB_ERR_UNSUPPORTED, GB_ERR_CUDA, GB_ERR_DIM, GB_ERR_PROLONG_SIZE) = range(7)

vp, i64, i32, f64, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64
p_i64, p_i32, p_f64, p_vp = C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/gridb200.h one to one (tests/test_abi.py checks it
# against the header text).
SIGNATURES = {
    "gb_version": (i32, []),
    "gb_last_error": (C.c_char_p, []),
    "gb_device_count": (i32, [p_i32]),
    "gb_set_device": (i32, [i32]),
    "gb_launch_count": (i64, []),
    "gb_profile_enable": (i32, [i32]),
    "gb_profile_read": (i32, [C.POINTER(C.c_double), i32, p_i32]),
    "gb_grid_create": (i32, [p_vp, i32, i32, p_i64, vp, i32, i32, i32, f64, vp]),
    "gb_grid_create_from_coefs": (i32, [p_vp, i32, i32, p_i64, vp, i32, i32, i32, f64, vp]),
    "gb_grid_create_channels": (i32, [p_vp, i32, i32, p_i64, i64, vp, i32, i32, i32, f64, vp]),
    "gb_grid_create_from_coefs_channels": (i32, [p_vp, i32, i32, p_i64, i64, vp, i32, i32, i32, f64, i32, vp]),
    "gb_grid_channels": (i32, [vp, p_i64]),
    "gb_grid_info": (i32, [vp, p_i32, p_i32, p_i64, p_i64, p_i32, p_i32, p_f64]),
    "gb_grid_coefs_device": (i32, [vp, p_vp]),
    "gb_grid_coefs_download": (i32, [vp, vp]),
    "gb_grid_coefs_copy": (i32, [vp, vp, vp]),
    "gb_grid_destroy": (i32, [vp]),
    "gb_invert": (i32, [i32, i32, p_i64, vp, i32, i32, i32, p_i32, i32, vp]),
    "gb_invert_opt": (i32, [i32, i32, p_i64, vp, i32, i32, i32, p_i32, i32, i32, vp]),
    "gb_grid_create_opt": (i32, [p_vp, i32, i32, p_i64, vp, i32, i32, i32, f64, i32, vp]),
    "gb_filledges": (i32, [i32, i32, p_i64, vp, f64, i32, vp]),
    "gb_comm_unique_id": (i32, [vp, C.c_size_t]),
    "gb_comm_init_rank": (i32, [p_vp, i32, i32, vp, C.c_size_t]),
    "gb_comm_init_all": (i32, [p_vp, i32, p_i32]),
    "gb_comm_init_loopback": (i32, [p_vp, i32]),
    "gb_comm_size": (i32, [vp, p_i32, p_i32, p_i32]),
    "gb_comm_destroy": (i32, [vp]),
    "gb_shard_range": (i32, [i64, i32, i32, p_i64, p_i64]),
    "gb_slab_plan": (i32, [i32, i64, i64, i32, i32, p_i64, p_i64, p_i64, p_i64, p_i64, i32, p_i32]),
    "gb_grid_replicate": (i32, [vp, i32, p_vp, i32, i32, p_i64, i32, i32, f64]),
    "gb_eval_sharded": (i32, [vp, p_vp, i32, i64, vp, i32, p_f64, vp, vp, vp, i32, p_i64]),
    "gb_restrict3_sharded": (i32, [vp, i32, p_i64, p_vp, f64, p_vp, p_vp]),
    "gb_prolong3_sharded": (i32, [vp, i32, p_i64, p_vp, p_i64, p_vp, p_vp]),
    "gb_restrict3_sharded_levels": (i32, [vp, i32, p_i64, i32, p_vp, f64, p_vp, p_vp]),
    "gb_prolong3_sharded_levels": (i32, [vp, i32, p_i64, i32, p_i64, p_vp, p_vp, p_vp]),
    "gb_comm_stats": (i32, [vp, p_f64, i32, p_i32]),
    "gb_comm_peer_halo": (i32, [vp, i32, p_i32]),
    "gb_comm_last_path": (i32, [vp, p_i32]),
    "gb_pad1": (i32, [i32, i32, p_i64, vp, f64, vp, i32, vp]),
    "gb_eval": (i32, [vp, i32, i64, vp, i32, i64, i64, p_f64, vp, vp, vp, i32, i32, p_i64, vp]),
    "gb_eval_soa": (i32, [vp, i32, i64, p_vp, i32, p_f64, vp, vp, vp, i32, i32, p_i64, vp]),
    "gb_eval_outer": (i32, [vp, p_i64, p_vp, i32, p_f64, vp, i32, i32, vp]),
    "gb_restrict_size": (i64, [i64]),
    "gb_prolong_size": (i32, [i64, i64, p_i64]),
    "gb_restrict": (i32, [i32, i32, p_i64, vp, i32, f64, vp, i32, vp]),
    "gb_prolong": (i32, [i32, i32, p_i64, vp, i32, i64, vp, i32, vp]),
    "gb_interp_irregular": (i32, [i32, i64, vp, vp, i32, i32, f64, i64, vp, vp, i32, p_i64, vp]),
    "gb_restrictb": (i32, [i32, i32, p_i64, vp, i32, f64, vp, i32, vp]),
    "gb_restrict_extrap": (i32, [i32, i32, p_i64, vp, i32, f64, vp, i32, vp]),
    "gb_prolongb": (i32, [i32, i32, p_i64, vp, i32, i64, vp, i32, vp]),
    "gb_restrict_slab": (i32, [i32, i32, p_i64, vp, i32, f64, i64, i64, i64, i64, vp, i32, vp]),
    "gb_prolong_slab": (i32, [i32, i32, p_i64, vp, i32, i64, i64, i64, i64, i64, vp, i32, vp]),
    "gb_restrict12": (i32, [i32, i32, p_i64, vp, f64, vp, i32, vp]),
    "gb_prolong12": (i32, [i32, i32, p_i64, vp, i64, i64, vp, i32, vp]),
    "gb_restrict3_slab": (i32, [i32, p_i64, vp, f64, i64, i64, i64, i64, vp, i32, vp]),
    "gb_prolong3_slab": (i32, [i32, p_i64, vp, p_i64, i64, i64, i64, i64, vp, i32, vp]),
    "gb_restrict3": (i32, [i32, p_i64, vp, f64, vp, i32, vp]),
    "gb_prolong3": (i32, [i32, p_i64, vp, p_i64, vp, i32, vp]),
    "gb_restrict3_levels": (i32, [i32, p_i64, i32, vp, f64, vp, i32, vp]),
    "gb_prolong3_levels": (i32, [i32, p_i64, i32, p_i64, vp, vp, i32, vp]),
    "gb_device_alloc": (i32, [p_vp, C.c_size_t]),
    "gb_device_free": (i32, [vp]),
    "gb_memcpy_h2d": (i32, [vp, vp, C.c_size_t, vp]),
    "gb_memcpy_d2h": (i32, [vp, vp, C.c_size_t, vp]),
    "gb_host_alloc_pinned": (i32, [p_vp, C.c_size_t]),
    "gb_host_free_pinned": (i32, [vp]),
    "gb_host_register": (i32, [vp, C.c_size_t]),
    "gb_host_unregister": (i32, [vp]),
    "gb_stream_synchronize": (i32, [vp]),
    "gb_synth_uniform": (i32, [i32, vp, i64, u64, f64, f64, i64, vp]),
}

_lib = None


class GridB200Error(RuntimeError):
    """Base class; `.code` is the C status."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class ErrorException(GridB200Error):
    """Julia's ErrorException (error(...)): arity / shape problems, "Invalid location"."""


class MethodError(GridB200Error):
    """Combination the reference has no method for (Cubic with reflect/..., Hessian of Linear)."""


class DimensionMismatch(GridB200Error):
    pass


class BoundsError(GridB200Error, IndexError):
    pass


class CudaError(GridB200Error):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or make -C grid.jl_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().gb_last_error().decode("utf-8", "replace")


def check(rc, first_invalid=None):
    if rc == GB_OK:
        return
    msg = last_error()
    if rc == GB_ERR_INVALID_LOCATION:
        where = f" (query {first_invalid})" if first_invalid is not None and first_invalid >= 0 else ""
        raise ErrorException(rc, "Invalid location" + where)
    if rc == GB_ERR_ARG or rc == GB_ERR_PROLONG_SIZE:
        raise ErrorException(rc, msg)
    if rc == GB_ERR_UNSUPPORTED:
        raise MethodError(rc, msg)
    if rc == GB_ERR_DIM:
        raise DimensionMismatch(rc, msg)
    if rc == GB_ERR_CUDA:
        raise CudaError(rc, msg)
    raise GridB200Error(rc, msg)
