"""GPU suite, large differential batches: every kernel family on >= 2^18 .. 2^20 systems per case, generic data with
hostile systems mixed in (exact ties, zero columns, duplicate rows, NaN / Inf, magnitudes at both ends of the fast
division window, small integers), against the compiled reference (oracle/_ref) when it is on the box, else the C port.

Why batches this large: the fast paths (one shared reciprocal per pivot, the speculative solve of gmb_rowfast.cuh and
its hand-back list) take their slow branches roughly once per 10^3 .. 10^7 systems on generic data; batches of a few
hundred never reach them.
"""
import numpy as np
import pytest

from conftest import bits_equal, first_diff

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gm():
    import geomc_b200 as gm
    gm.lib()
    assert torch.cuda.is_available()
    return gm


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle as po
    return po.best()          # the unmodified reference when oracle/_ref is present (it travels with the repo)


def hostile_batch(rng, B, n, dt, every=61):
    """Exact-grid random systems; every `every`-th system gets one of 12 structures that leave the generic path."""
    from oracle import pyoracle as po
    A = po.exact_grid(rng, (B, n * n), dt)
    b = po.exact_grid(rng, (B, n), dt)
    idx = np.arange(5, B, every)
    kind = (idx // every) % 12
    big, tiny = (1e30, 1e-30) if dt == np.float32 else (1e200, 1e-200)
    for k in range(12):
        s = idx[kind == k]
        if len(s) == 0:
            continue
        M = A[s].reshape(len(s), n, n)
        if k == 0:
            M[:] = 0
        elif k == 1:
            M[:] = np.where(np.arange(n * n).reshape(n, n) % 3 == 0, 1, -1)          # ties everywhere
        elif k == 2:
            M[:, :, 0] = 0                                                             # zero first column
        elif k == 3:
            M[:, 1, :] = M[:, 0, :]                                                    # duplicate row
        elif k == 4:
            M[:, 3 % n, 2 % n] = np.nan
        elif k == 5:
            M[:, 0, 0] = np.inf
        elif k == 6:
            M *= dt(big)                                                               # beyond the division window
        elif k == 7:
            M *= dt(tiny)
        elif k == 8:
            M[:] = np.trunc(M)                                                         # small integers: ties and zeros
        elif k == 9:
            M[:, n - 1, n - 1] = np.nan
            b[s, 1 % n] = 0
        elif k == 10:
            M[:, :, n // 2] *= dt(tiny)                                                # one tiny column: tiny pivots
        elif k == 11:
            M[:, n // 2, :] = -M[:, n // 2 - 1, :]                                     # equal magnitudes, opposite signs
        A[s] = M.reshape(len(s), n * n)
    return A, b


def up(gm, a, soa):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return gm.BatchSoA.from_aos(t) if soa else t


def down(x):
    if hasattr(x, "to_aos"):
        x = x.to_aos()
    return x.cpu().numpy()


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [9, 11, 12, 13, 16])
def test_speculative_solve_is_the_exact_solve(gm, orc, n, dt, soa):
    """gmb_rowfast.cuh: the speculative kernel + the exact re-solve of the systems it hands back must produce the bits
    of the exact kernel alone (option rowfast = 0) and of the reference, for every system of a hostile batch, in both
    matrix orders and with skip > 0."""
    rng = np.random.default_rng(1000 * n + np.dtype(dt).itemsize + int(soa))
    B = 1 << 18
    A, b = hostile_batch(rng, B, n, dt)
    dA, db = up(gm, A, soa), up(gm, b, soa)
    try:
        for rm, skip in ((True, 0), (False, 0), (True, 2)):
            gm.set_option("rowfast", 1)
            x1, ok1 = gm.linear_solve(dA, db, n, 1, skip, rm)
            gm.set_option("rowfast", 0)
            x0, ok0 = gm.linear_solve(dA, db, n, 1, skip, rm)
            x1, ok1, x0, ok0 = down(x1), down(ok1), down(x0), down(ok0)
            assert bits_equal(ok1, ok0), f"ok n={n} rm={rm} skip={skip}: {first_diff(ok1, ok0)}"
            assert bits_equal(x1, x0), f"x n={n} rm={rm} skip={skip}: {first_diff(x1, x0)}"
            assert 0 < int((ok0 == 0).sum()) < B // 8              # the hostile systems are there, the rest is regular
            # and against the reference itself on a slice that contains ~500 hostile systems
            S = 1 << 15
            with np.errstate(all="ignore"):
                _, xe, oke = orc.linear_solve(A[:S], b[:S], n, 1, skip, rm)
            assert bits_equal(ok1[:S], oke), f"ok vs {orc.kind} n={n} rm={rm}: {first_diff(ok1[:S], oke)}"
            assert bits_equal(x1[:S], xe), f"x vs {orc.kind} n={n} rm={rm} skip={skip}: {first_diff(x1[:S], xe)}"
    finally:
        gm.set_option("rowfast", -1)


def test_speculative_solve_in_place(gm):
    """x may alias b: the systems handed back must still see their right-hand side."""
    n, B = 16, 1 << 16
    rng = np.random.default_rng(77)
    A, b = hostile_batch(rng, B, n, np.float32)
    dA, db = up(gm, A, False), up(gm, b, False)
    try:
        gm.set_option("rowfast", 0)
        x0, _ = gm.linear_solve(dA, db, n)
        gm.set_option("rowfast", 1)
        xin = db.clone()
        gm.linear_solve(dA, xin, n, x_out=xin)
        assert bits_equal(down(xin), down(x0)), first_diff(down(xin), down(x0))
    finally:
        gm.set_option("rowfast", -1)
