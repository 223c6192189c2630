"""CPU suite: oracle.c (the plain-C restatement) against the committed golden fixtures.

The fixtures are outputs of the reference's own functions (tests/golden/make_golden.py ran
oracle/_ref, i.e. the unmodified geomc headers).  Bit-exact, NaN == NaN.
"""
import numpy as np
import pytest

from conftest import GOLDEN_SIZES, bits_equal, first_diff, load_golden

DT = {"f32": np.float32, "f64": np.float64}


def check(got, want, what):
    assert bits_equal(got, want), f"{what}: {first_diff(got, want)}"


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("n", GOLDEN_SIZES)
def test_port_matches_golden(port, tag, n):
    g = load_golden(tag, n)
    A, b1, b3, S = g["A"], g["b1"], g["b3"], g["S"]
    with np.errstate(all="ignore"):
        for rm, r in ((True, "rm"), (False, "cm")):
            lu, p, par, dg = port.decomp_plu(A, n, n, rm)
            check(lu, g[f"plu_{r}_lu"], "plu lu")
            check(p, g[f"plu_{r}_p"], "plu reorder")
            check(par, g[f"plu_{r}_parity"], "plu parity")
            check(dg, g[f"plu_{r}_degen"], "plu degenerate_ct")
            lu2, p2, par2, dg2 = port.decomp_lup(A, n, n, rm)
            check(lu2, g[f"lup_{r}_lu"], "lup lu")
            check(p2, g[f"lup_{r}_p"], "lup reorder")
            check(par2, g[f"lup_{r}_parity"], "lup parity")
            check(dg2, g[f"lup_{r}_degen"], "lup degenerate_ct")
            a1, x1, ok1 = port.linear_solve(A, b1, n, 1, 0, rm)
            check(a1, g[f"solve_{r}_a"], "linear_solve a")
            check(x1, g[f"solve_{r}_x"], "linear_solve x")
            check(ok1, g[f"solve_{r}_ok"], "linear_solve ok")
            _, x3, ok3 = port.linear_solve(A, b3, n, 3, 1, rm)
            check(x3, g[f"solve3s1_{r}_x"], "linear_solve m=3 skip=1 x")
            check(ok3, g[f"solve3s1_{r}_ok"], "linear_solve m=3 skip=1 ok")
            check(port.backsolve_lu(lu, b3, n, 3, 0, rm), g[f"bslu_{r}_x"], "backsolve_lu")
            check(port.backsolve_plu(lu, p, b3, n, 3, 0, rm), g[f"bsplu_{r}_x"], "backsolve_plu")
            check(port.apply_perm(p, b3, n, 3, rm)[0], g[f"perm_{r}_a"], "apply_permutation")
            L, okc = port.cholesky(S, n, rm)
            check(L, g[f"chol_{r}_l"], "cholesky L")
            check(okc, g[f"chol_{r}_ok"], "cholesky ok")
            for drm, q in ((True, "rm"), (False, "cm")):
                inv, oki = port.inv(A, n, rm, drm, np.full_like(A, 7))
                check(inv, g[f"inv_{r}_{q}"], f"inv {r}->{q}")
                check(oki, g[f"inv_{r}_{q}_ok"], f"inv ok {r}->{q}")
        bv, xv, okv = port.linear_solve_vec(A, b1, n, 1, 0)
        check(bv, g["solvevec_bases"], "linear_solve(Vec) bases")
        check(xv, g["solvevec_x"], "linear_solve(Vec) x")
        check(okv, g["solvevec_ok"], "linear_solve(Vec) ok")
        lur, pr, parr, dgr = port.decomp_plu(g["rect_m"], n - 1, n, True)
        check(lur, g["rect_lu"], "rect plu")
        check(pr, g["rect_p"], "rect reorder")
        check(parr, g["rect_parity"], "rect parity")
        check(dgr, g["rect_degen"], "rect degen")


def test_golden_covers_edge_cases():
    """The fixtures must actually contain the edge cases the reference's semantics hinge on."""
    g = load_golden("f32", 4)
    ok = g["solve_rm_ok"]
    assert ok[0] == 0 and ok[1] == 0          # all-zero / zero first column -> false
    assert ok[2] == 1                         # zero LAST pivot is never tested (LUDecomp.h:203)
    assert not np.isfinite(g["solve_rm_x"][2]).all()
    assert np.array_equal(g["solve_rm_x"][0], g["b1"][0])   # x untouched on failure (:391-393)
    assert g["plu_rm_p"][3][0] == 0           # ties: lowest index wins (:209 strict >)
    assert g["inv_rm_rm_ok"][0] == 0 and (g["inv_rm_rm"][0] == 7).all()   # out untouched on det == 0
    assert g["chol_rm_ok"][0] == 0 and g["chol_rm_ok"][1] == 0 and g["chol_rm_ok"][3] == 1
