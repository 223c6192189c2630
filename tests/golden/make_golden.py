"""Generate tests/golden/*.npz from the reference itself.

Runs ONLY in the build container (needs oracle/_ref/libgeomc_ref.so, which oracle/Makefile
compiles from the unmodified headers under /root/reference).  The reference ships no golden
vectors or known-answer tests for this path (property tests with fixed
seeds only), so these fixtures are outputs of the reference's own functions on seeded inputs:

    python tests/golden/make_golden.py

Every fixture holds the inputs and every output of one reference function for a handful of
systems, including deliberately singular / non-SPD / non-finite ones.  tests/test_oracle_golden.py
replays them through oracle.c (CPU suite) and tests/test_gpu_parity.py through the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SIZES = (2, 3, 4, 5, 7, 8, 11, 16)
B = 32


def inputs(rng, n, dt):
    """48 systems: exact-grid random, plus the singular/edge cases of SURVEY.md 8a (a6 probe)."""
    A = po.exact_grid(rng, (B, n * n), dt)
    A[0] = 0                                   # all-zero
    A[1].reshape(n, n)[:, 0] = 0               # zero first column  -> degenerate at step 0
    A[2].reshape(n, n)[n - 1] = A[2].reshape(n, n)[0]  # duplicate row -> zero LAST pivot (ok stays true)
    A[3].reshape(n, n)[:, 0] = [1.0 if i % 2 == 0 else -1.0 for i in range(n)]  # |ties|: lowest index wins
    A[4].reshape(n, n)[1:, 1] = 0              # second column zero below the diagonal
    A[4].reshape(n, n)[0, 1] = 1
    A[5].reshape(n, n)[0, 0] = np.nan          # NaN on the diagonal is never displaced
    A[6].reshape(n, n)[n - 1, 0] = np.inf
    A[7] = np.eye(n, dtype=dt).reshape(-1)
    A[8] = np.eye(n, dtype=dt)[::-1].reshape(-1)   # anti-diagonal permutation
    A[9] *= dt(2.0 ** -120) if dt == np.float32 else dt(2.0 ** -1000)   # denormal territory
    A[10] *= dt(2.0 ** 60)
    return A


def main():
    R = po.ref()
    R.set_threads(1)
    total = 0
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for n in SIZES:
            rng = np.random.default_rng(1000 * n + (32 if dt == np.float32 else 64))
            A = inputs(rng, n, dt)
            b1 = po.exact_grid(rng, (B, n), dt)
            b3 = po.exact_grid(rng, (B, n * 3), dt)
            S = po.spd_exact(rng, B, n, dt)
            S[0].reshape(n, n)[0, 0] = -1          # not SPD at the first pivot
            S[1].reshape(n, n)[n - 1, n - 1] = -50  # not SPD at the last pivot
            S[2] = 0
            d = {"A": A, "b1": b1, "b3": b3, "S": S}
            with np.errstate(all="ignore"):
                for rm, r in ((True, "rm"), (False, "cm")):
                    lu, p, par, dg = R.decomp_plu(A, n, n, rm)
                    d[f"plu_{r}_lu"], d[f"plu_{r}_p"], d[f"plu_{r}_parity"], d[f"plu_{r}_degen"] = lu, p, par, dg
                    lu2, p2, par2, dg2 = R.decomp_lup(A, n, n, rm)
                    d[f"lup_{r}_lu"], d[f"lup_{r}_p"], d[f"lup_{r}_parity"], d[f"lup_{r}_degen"] = lu2, p2, par2, dg2
                    a1, x1, ok1 = R.linear_solve(A, b1, n, 1, 0, rm)
                    d[f"solve_{r}_a"], d[f"solve_{r}_x"], d[f"solve_{r}_ok"] = a1, x1, ok1
                    a3, x3, ok3 = R.linear_solve(A, b3, n, 3, 1, rm)
                    d[f"solve3s1_{r}_x"], d[f"solve3s1_{r}_ok"] = x3, ok3
                    d[f"bslu_{r}_x"] = R.backsolve_lu(lu, b3, n, 3, 0, rm)
                    d[f"bsplu_{r}_x"] = R.backsolve_plu(lu, p, b3, n, 3, 0, rm)
                    pa, _ = R.apply_perm(p, b3, n, 3, rm)
                    d[f"perm_{r}_a"] = pa
                    L, okc = R.cholesky(S, n, rm)
                    d[f"chol_{r}_l"], d[f"chol_{r}_ok"] = L, okc
                    for drm, q in ((True, "rm"), (False, "cm")):
                        init = np.full_like(A, 7)
                        inv, oki = R.inv(A, n, rm, drm, init)
                        d[f"inv_{r}_{q}"], d[f"inv_{r}_{q}_ok"] = inv, oki
                bv, xv, okv = R.linear_solve_vec(A, b1, n, 1, 0)
                d["solvevec_bases"], d["solvevec_x"], d["solvevec_ok"] = bv, xv, okv
                # rectangular decomp_plu as used by nullspace / orthogonal (Orthogonal.h:83,161)
                rows = n - 1
                Rm = po.exact_grid(rng, (B, rows * n), dt)
                lur, pr, parr, dgr = R.decomp_plu(Rm, rows, n, True)
                d["rect_m"], d["rect_lu"], d["rect_p"], d["rect_parity"], d["rect_degen"] = Rm, lur, pr, parr, dgr
            path = os.path.join(OUT, f"hotpath_{tag}_n{n}.npz")
            np.savez_compressed(path, **d)
            total += os.path.getsize(path)
    print(f"wrote {2 * len(SIZES)} fixtures, {total / 1024:.0f} KiB; reference = {R.build_info()}")


if __name__ == "__main__":
    main()
