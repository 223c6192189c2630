"""The C++ host side (include/geomc_b200/batch.hpp): compiles against the C ABI with a plain C++
compiler -- with stand-in types and, where /root/reference exists, with the reference's REAL
SimpleMatrix / Vec types -- fails loudly without a GPU, and passes its checks on one."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_batch_hpp.cpp")
LIBDIR = os.path.join(ROOT, "geomc_b200")


def build(tmp_path, with_geomc):
    out = str(tmp_path / ("t_geomc" if with_geomc else "t_plain"))
    cmd = ["g++", "-O1", "-g", "-rdynamic", "-w", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR, "-lgeomc_b200",
           f"-Wl,-rpath,{LIBDIR}", "-o", out]
    if with_geomc:
        cmd[1:1] = ["-std=c++23", "-DWITH_GEOMC", "-include", os.path.join(ROOT, "oracle", "shim", "prelude.h"),
                    "-I", os.path.join(ROOT, "oracle", "shim"), "-I", "/root/reference"]
        cmd.insert(cmd.index(SRC) + 1, "/root/reference/geomc/GeomException.cpp")
    else:
        cmd[1:1] = ["-std=c++17"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ missing")
@pytest.mark.parametrize("with_geomc", [False, True])
def test_header_compiles_and_fails_loudly_without_gpu(tmp_path, with_geomc):
    if with_geomc and not os.path.isdir("/root/reference/geomc"):
        pytest.skip("/root/reference absent")
    exe = build(tmp_path, with_geomc)
    from geomc_b200 import _lib
    if _lib.lib().gmb_device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_api_on_gpu(tmp_path):
    exe = build(tmp_path, False)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "batch.hpp OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


REAL = os.path.join(ROOT, "oracle", "_ref", "test_batch_hpp_geomc")
REAL_BENCH = os.path.join(ROOT, "oracle", "_ref", "bench_object_inv_geomc")


@pytest.mark.gpu
def test_cpp_host_api_with_the_references_real_types_on_gpu():
    """The object forms on geom::SimpleMatrix / Vec / Simplex / Ray / AffineTransform themselves: the binary is built by
    oracle/Makefile on the box that has /root/reference (-DWITH_GEOMC) and travels with the snapshot, like _ref/*.so."""
    if not os.path.exists(REAL):
        if os.path.isdir("/root/reference/geomc"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "hosttests"], check=True, capture_output=True)
        else:
            pytest.skip("oracle/_ref/test_batch_hpp_geomc was not built (no /root/reference on the build box)")
    r = subprocess.run([REAL], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "batch.hpp OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_object_form_inverse_is_zero_copy_and_matches_the_raw_entry_point(tmp_path):
    """geom::batch::inv on an array of SimpleMatrix<float,4,4> objects hands the array itself to gmb_host_inv_f32
    (no repacking): identical bits, and -- on registered memory -- the same end-to-end rate as raw pinned pointers."""
    import json
    exe = REAL_BENCH
    if not os.path.exists(exe):
        exe = str(tmp_path / "bench_object_inv")
        cmd = ["g++", "-std=c++17", "-O2", "-w", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "bench_object_inv.cpp"),
               "-L", LIBDIR, "-lgeomc_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "20", "0", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["identical_to_raw"] is True and d["singular"] >= 1, d
    assert d["sizeof_object"] == 64
