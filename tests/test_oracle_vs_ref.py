"""CPU suite: oracle.c pinned bit-for-bit against the reference itself (oracle/_ref) on fresh
random batches -- every hot-path function, N = 2..16, float and double, both layouts.
Skipped where the compiled reference is not present."""
import numpy as np
import pytest

from conftest import bits_equal, first_diff
from oracle import pyoracle as po


def check(got, want, what):
    for i, (u, v) in enumerate(zip(got if isinstance(got, tuple) else (got,), want if isinstance(want, tuple) else (want,))):
        assert bits_equal(u, v), f"{what}[{i}]: {first_diff(u, v)}"


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 8, 9, 13, 16])
def test_port_equals_reference(port, ref, dt, n):
    rng = np.random.default_rng(77 * n + np.dtype(dt).itemsize)
    B = 600
    A = rng.standard_normal((B, n * n)).astype(dt)
    A[::50, :] = 0
    A[1::50, 0::n] = 0
    x1 = rng.standard_normal((B, n)).astype(dt)
    x3 = rng.standard_normal((B, n * 3)).astype(dt)
    S = po.spd_exact(rng, B, n, dt)
    S[::40, 0] = -1
    with np.errstate(all="ignore"):
        for rm in (True, False):
            check(port.decomp_plu(A, n, n, rm), ref.decomp_plu(A, n, n, rm), "decomp_plu")
            check(port.decomp_lup(A, n, n, rm), ref.decomp_lup(A, n, n, rm), "decomp_lup")
            check(port.linear_solve(A, x1, n, 1, 0, rm), ref.linear_solve(A, x1, n, 1, 0, rm), "linear_solve")
            check(port.linear_solve(A, x3, n, 3, 1, rm), ref.linear_solve(A, x3, n, 3, 1, rm), "linear_solve m3 s1")
            lu, p, _, _ = ref.decomp_plu(A, n, n, rm)
            check(port.backsolve_lu(lu, x3, n, 3, 1, rm), ref.backsolve_lu(lu, x3, n, 3, 1, rm), "backsolve_lu")
            check(port.backsolve_plu(lu, p, x3, n, 3, 0, rm), ref.backsolve_plu(lu, p, x3, n, 3, 0, rm), "backsolve_plu")
            check(port.apply_perm(p, x3, n, 3, rm), ref.apply_perm(p, x3, n, 3, rm), "apply_perm")
            check(port.cholesky(S, n, rm), ref.cholesky(S, n, rm), "cholesky")
            for drm in (True, False):
                init = rng.standard_normal((B, n * n)).astype(dt)
                check(port.inv(A, n, rm, drm, init), ref.inv(A, n, rm, drm, init), "inv")
        check(port.cholesky(S, n, True), ref.cholesky_mtx(S, n), "cholesky(SimpleMatrix*)")
        check(port.inv_raw(A, n), ref.inv_raw(A, n), "inv_raw")
        for m, skip in ((1, 0), (2, 1)):
            xv = rng.standard_normal((B, n * m)).astype(dt)
            check(port.linear_solve_vec(A, xv, n, m, skip), ref.linear_solve_vec(A, xv, n, m, skip), "linear_solve(Vec)")
        check(port.decomp_plu(A[:, : (n - 1) * n], n - 1, n, True), ref.decomp_plu(A[:, : (n - 1) * n], n - 1, n, True), "rect plu")
