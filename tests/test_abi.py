"""CPU suite: the C-ABI library loads without a GPU and exports every symbol include/geomc_b200.h
declares; argument validation works; compute calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "geomc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from geomc_b200 import _lib
    L = _lib.lib()
    declared = header_symbols()
    assert len(declared) > 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/geomc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, "python binding table and header disagree"


def test_container_helpers_without_gpu():
    from geomc_b200 import _lib
    L = _lib.lib()
    assert L.gmb_soa_tile_width(4) == 128 and L.gmb_soa_tile_width(8) == 64
    assert L.gmb_soa_elems(1, 16, 4) == 128 * 16
    assert L.gmb_soa_elems(129, 16, 4) == 256 * 16
    assert L.gmb_soa_elems(0, 16, 4) == 0
    assert b"sm_100a" in L.gmb_version()
    assert b"no CPU fallback" in L.gmb_error_string(-3)


def test_argument_validation_and_no_cpu_fallback():
    from geomc_b200 import _lib
    L = _lib.lib()
    # invalid arguments are rejected before anything touches a device
    assert L.gmb_plu_solve_f32(4, 1, -1, None, None, None, None, None, 0, 0, 1, None) == -1
    assert L.gmb_plu_solve_f32(4, 1, 8, None, None, None, None, None, 0, 0, 1, None) == -1
    assert L.gmb_plu_solve_f32(4, 1, 8, None, None, None, None, None, 0, 7, 1, None) == -1
    assert L.gmb_plu_solve_f32(17, 1, 0, None, None, None, None, None, 0, 0, 1, None) == -2
    # rows (f)3 / (f)4: shapes, dimensions, modes, transform counts and aliasing are checked the same way
    assert L.gmb_mul_f32(4, 4, 4, 8, None, None, None, 0, 0, 1, 1, 1, None) == -1
    assert L.gmb_mul_f64(17, 4, 4, 0, None, None, None, 0, 0, 1, 1, 1, None) == -2
    assert L.gmb_mul_f64(0, 4, 4, 0, None, None, None, 0, 0, 1, 1, 1, None) == -1
    assert L.gmb_xf_from_matrix_f32(8, 0, None, None, None, 0, 1, None) == -2        # dimensions 2..7
    assert L.gmb_xf_from_matrix_f32(3, 4, None, None, None, 0, 1, None) == -1
    assert L.gmb_xf_apply_f64(3, 6, 0, None, 0, None, None, 0, None) == -1           # mode 0..5
    assert L.gmb_xf_apply_f64(3, 0, 5, 16, 2, 16, 16, 0, None) == -1                 # one transform, or one per point
    assert L.gmb_xf_compose_f32(3, 5, 16, 32, 16, 0, 0, None) == -1                  # out must not alias an input
    assert L.gmb_xf_translation_f32(1, 0, None, None, 0, None) == -2
    if L.gmb_device_count() == 0:
        a = np.zeros((8, 16), np.float32)
        b = np.zeros((8, 4), np.float32)
        ok = np.zeros(8, np.uint8)
        rc = L.gmb_host_plu_solve_f32(4, 1, 8, a.ctypes.data, b.ctypes.data, b.ctypes.data, ok.ctypes.data, 0, 1, 0)
        assert rc == -3, "without a GPU the host entry point must fail loudly, not compute on the CPU"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "geomc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "libgeomc_ref" not in text, f


def test_c_and_python_shard_ranges_agree():
    """gmb_shard_range (used by the multi-device host entry points) == geomc_b200.sharding.shard_range (used by bench.py and the
    gloo tests): every system belongs to exactly one shard, boundaries on multiples of `align`."""
    import ctypes
    from geomc_b200 import _lib
    from geomc_b200.sharding import shard_range
    L = _lib.lib()
    for B in (0, 1, 7, 1023, 1024, 150001, 1 << 24):
        for world in (1, 2, 3, 8):
            for align in (1, 64, 128, 1024):
                covered = 0
                for r in range(world):
                    lo, hi = ctypes.c_int64(), ctypes.c_int64()
                    assert L.gmb_shard_range(B, r, world, align, ctypes.byref(lo), ctypes.byref(hi)) == 0
                    assert (lo.value, hi.value) == shard_range(B, r, world, align)
                    assert lo.value == covered and (lo.value % align == 0 or lo.value == B)
                    covered = hi.value
                assert covered == B
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    assert L.gmb_shard_range(10, 2, 2, 1, ctypes.byref(lo), ctypes.byref(hi)) != 0      # rank out of range
