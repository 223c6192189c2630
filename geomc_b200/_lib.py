"""ctypes binding of libgeomc_b200.so (the C ABI declared in include/geomc_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, this raises.
Nothing here imports, links or calls anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GMB_LIB_PATH: an alternative build of the SAME library (tuning experiments with different compile-time constants)
LIB_PATH = os.environ.get("GMB_LIB_PATH") or os.path.join(_HERE, "libgeomc_b200.so")

LAYOUT_AOS = 0
LAYOUT_SOA = 1
MAX_N = 16
MAX_M = 4


class GmbError(RuntimeError):
    pass


_lib = None

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> argtypes (per dtype suffix); every function returns int
_TYPED = {
    "gmb_aos_to_soa": [_i, _i64, _vp, _vp, _vp],
    "gmb_soa_to_aos": [_i, _i64, _vp, _vp, _vp],
    "gmb_plu_decomp": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_lup_decomp": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_lu_backsolve": [_i, _i, _i64, _vp, _vp, _i, _i, _i, _vp],
    "gmb_plu_backsolve": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "gmb_apply_permutation": [_i, _i, _i64, _vp, _vp, _i, _i, _vp],
    "gmb_plu_solve": [_i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "gmb_linear_solve_vec": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_cholesky": [_i, _i64, _vp, _vp, _i, _i, _vp],
    "gmb_cholesky_solve": [_i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_inv": [_i, _i64, _vp, _vp, _vp, _i, _i, _i, _vp],
    "gmb_solve_residual": [_i, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_orthogonal": [_i, _i64, _vp, _vp, _i, _vp],
    "gmb_nullspace": [_i, _i, _i64, _vp, _vp, _vp, _i, _vp],
    "gmb_simplex_contains": [_i, _i64, _vp, _i64, _vp, _vp, _vp, _i, _vp],
    "gmb_trace_simplex": [_i, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i, _vp],
    "gmb_simplex_intersect": [_i, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i, _vp],
    "gmb_mul": [_i, _i, _i, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gmb_xf_from_matrix": [_i, _i64, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_xf_translation": [_i, _i64, _vp, _vp, _i, _vp],
    "gmb_xf_scale": [_i, _i64, _vp, _vp, _i, _vp],
    "gmb_xf_compose": [_i, _i64, _vp, _vp, _vp, _i, _i, _vp],
    "gmb_xf_apply": [_i, _i, _i64, _vp, _i64, _vp, _vp, _i, _vp],
    "gmb_host_plu_solve": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i, _i],
    "gmb_host_inv": [_i, _i64, _vp, _vp, _vp, _i, _i, _i],
    "gmb_host_cholesky_solve": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i],
    "gmb_host_plu_decomp": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i],
    "gmb_host_plu_solve_multi": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp, _i],
    "gmb_host_inv_multi": [_i, _i64, _vp, _vp, _vp, _i, _i, _vp, _i],
    "gmb_host_cholesky_solve_multi": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _vp, _i],
    "gmb_host_plu_decomp_multi": [_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _vp, _i],
}

# every symbol include/geomc_b200.h declares (tests/test_abi.py checks the header against this)
UNTYPED = ["gmb_version", "gmb_error_string", "gmb_device_count", "gmb_soa_tile_width", "gmb_soa_elems",
           "gmb_checksum_bytes", "gmb_host_alloc", "gmb_host_free", "gmb_launch_count", "gmb_set_option", "gmb_set_device",
           "gmb_device_alloc", "gmb_device_free", "gmb_memcpy_h2d", "gmb_memcpy_d2h", "gmb_memset_device",
           "gmb_stream_synchronize", "gmb_get_device", "gmb_host_register", "gmb_host_unregister", "gmb_shard_range",
           "gmb_host_alloc_wc", "gmb_bind_thread_near_device", "gmb_host_alloc_near"]
EXPORTS = UNTYPED + [f"{k}_{s}" for k in _TYPED for s in ("f32", "f64")]


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GmbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C geomc_b200/csrc`. geomc_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, argtypes in _TYPED.items():
        for suf in ("f32", "f64"):
            f = getattr(L, f"{name}_{suf}")
            f.argtypes = argtypes
            f.restype = C.c_int
    L.gmb_version.restype = C.c_char_p
    L.gmb_error_string.restype = C.c_char_p
    L.gmb_error_string.argtypes = [_i]
    L.gmb_device_count.restype = _i
    L.gmb_soa_tile_width.restype = _i
    L.gmb_soa_tile_width.argtypes = [_i]
    L.gmb_soa_elems.restype = _i64
    L.gmb_soa_elems.argtypes = [_i64, _i, _i]
    L.gmb_checksum_bytes.restype = _i
    L.gmb_checksum_bytes.argtypes = [_vp, _i64, _i64, _vp, _vp]
    L.gmb_host_alloc.restype = _vp
    L.gmb_host_alloc.argtypes = [C.c_size_t]
    L.gmb_bind_thread_near_device.restype = _i
    L.gmb_bind_thread_near_device.argtypes = [_i]
    L.gmb_host_alloc_near.restype = _vp
    L.gmb_host_alloc_near.argtypes = [C.c_size_t, _i]
    L.gmb_host_alloc_wc.restype = _vp
    L.gmb_host_alloc_wc.argtypes = [C.c_size_t]
    L.gmb_host_free.restype = None
    L.gmb_host_free.argtypes = [_vp]
    L.gmb_launch_count.restype = _i64
    L.gmb_set_option.restype = _i
    L.gmb_set_option.argtypes = [C.c_char_p, _i]
    L.gmb_set_device.restype = _i
    L.gmb_set_device.argtypes = [_i]
    L.gmb_get_device.restype = _i
    L.gmb_get_device.argtypes = [C.POINTER(_i)]
    L.gmb_host_register.restype = _i
    L.gmb_host_register.argtypes = [_vp, C.c_size_t]
    L.gmb_host_unregister.restype = _i
    L.gmb_host_unregister.argtypes = [_vp]
    L.gmb_shard_range.restype = _i
    L.gmb_shard_range.argtypes = [_i64, _i, _i, _i64, C.POINTER(_i64), C.POINTER(_i64)]
    L.gmb_device_alloc.restype = _vp
    L.gmb_device_alloc.argtypes = [C.c_size_t]
    L.gmb_device_free.restype = None
    L.gmb_device_free.argtypes = [_vp]
    for name in ("gmb_memcpy_h2d", "gmb_memcpy_d2h"):
        getattr(L, name).restype = _i
        getattr(L, name).argtypes = [_vp, _vp, C.c_size_t, _vp]
    L.gmb_memset_device.restype = _i
    L.gmb_memset_device.argtypes = [_vp, _i, C.c_size_t, _vp]
    L.gmb_stream_synchronize.restype = _i
    L.gmb_stream_synchronize.argtypes = [_vp]
    # test / tuning harness: GMB_OPT_<NAME>=<int> in the environment of THIS python process is forwarded to
    # gmb_set_option (the shared library itself never reads the environment)
    for key, val in os.environ.items():
        if key.startswith("GMB_OPT_"):
            rc = L.gmb_set_option(key[len("GMB_OPT_"):].lower().encode(), int(val))
            if rc != 0:
                raise GmbError(f"unknown option in the environment: {key}")
    _lib = L
    return L


def set_option(name: str, value: int) -> None:
    """Force a kernel family (tests, tuning); -1 restores the measured dispatch table."""
    check(lib().gmb_set_option(name.encode(), int(value)), f"gmb_set_option({name})")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise GmbError(f"{what} failed: {lib().gmb_error_string(rc).decode()} (code {rc})")


def call(name: str, suffix: str, *args) -> None:
    check(getattr(lib(), f"{name}_{suffix}")(*args), f"{name}_{suffix}")
