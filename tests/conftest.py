import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def bits_equal(a, b):
    """Bit-exact comparison of float arrays; NaNs compare equal to NaNs regardless of payload/sign
    (x86 and the GPU canonicalise NaNs differently; the reference pins no NaN payload)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind != "f":
        return bool(np.array_equal(a, b))
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    u = f"u{a.itemsize}"
    return bool(np.array_equal(a[~na].view(u), b[~nb].view(u)))


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b)))) if a.dtype.kind == "f" else np.argwhere(a != b)
    if len(bad) == 0 and a.dtype.kind == "f":
        u = f"u{a.itemsize}"
        ok = np.isnan(a) & np.isnan(b)
        bad = np.argwhere((a.view(u) != b.view(u)) & ~ok)
        if len(bad):
            i = tuple(bad[0])
            return (f"{len(bad)} bit-level diffs (signed zero), first at {i}: {a[i]!r} ({a.view(u)[i]:#x}) vs "
                    f"{b[i]!r} ({b.view(u)[i]:#x}); rows {sorted(set(int(x[0]) for x in bad))[:8]}")
    if len(bad) == 0:
        return "no value difference (signed zero?)"
    i = tuple(bad[0])
    return f"{len(bad)} diffs, first at {i}: {a[i]!r} vs {b[i]!r}"


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle as po
    return po.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle as po
    if not po.ref_available() and not os.path.isdir("/root/reference/geomc"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return po.ref()


@pytest.fixture(scope="session")
def orc():
    """The checker of the GPU parity tests: oracle/_ref (the unmodified reference, compiled; it travels to the GPU box)
    when present, else the C port -- no indirection through the port when the reference itself can answer."""
    from oracle import pyoracle as po
    return po.checker()


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_SIZES = (2, 3, 4, 5, 7, 8, 11, 16)


def load_golden(tag, n):
    return np.load(os.path.join(GOLDEN_DIR, f"hotpath_{tag}_n{n}.npz"))
