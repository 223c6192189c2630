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


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, 6, 7, 8])
def test_lane_cholesky_is_the_row_cholesky(gm, orc, n, dt, soa):
    """gmb_lane.cuh k_lane_chol (staged ingest, one thread per system) against the row kernel (option lanechol = 0) and the
    reference: factor, ok flags and the fused solve, SPD and non-SPD systems, extreme scalings."""
    if dt == np.float64 and n > 6:
        pytest.skip("double: n = 5, 6 only (registers)")
    from oracle import pyoracle as po
    rng = np.random.default_rng(3000 * n + np.dtype(dt).itemsize + int(soa))
    B = (1 << 16) + 77
    S = po.spd_exact(rng, B, n, dt)
    S[::53, 0] = -1                                              # not positive definite
    S[7::211] = 0
    kmax = 30 if dt == np.float32 else 200
    k = rng.integers(-kmax, kmax + 1, size=(B, n))
    k[::3] = 0
    with np.errstate(all="ignore"):
        S = np.ldexp(S.astype(np.float64).reshape(B, n, n), k[:, :, None] + k[:, None, :]).astype(dt).reshape(B, n * n)
    b1 = po.exact_grid(rng, (B, n), dt)
    b2 = po.exact_grid(rng, (B, 2 * n), dt)
    dS = up(gm, S, soa)
    try:
        for rm in (True, False):
            for m, b in ((1, b1), (2, b2)):
                db = up(gm, b, soa)
                gm.set_option("lanechol", 1)
                x1, ok1, L1 = gm.cholesky_solve(dS, db, n, m, rm, want_l=True)
                Lf1, okf1 = gm.cholesky(dS, n, rm)
                gm.set_option("lanechol", 0)
                if m == 1:
                    x0, ok0, L0 = gm.cholesky_solve(dS, db, n, m, rm, want_l=True)
                    assert bits_equal(down(x1), down(x0)), f"x n={n} rm={rm}: {first_diff(down(x1), down(x0))}"
                    assert bits_equal(down(L1), down(L0)), f"L n={n} rm={rm}: {first_diff(down(L1), down(L0))}"
                    assert bits_equal(down(ok1), down(ok0))
                with np.errstate(all="ignore"):
                    Le, oke = orc.cholesky(S, n, rm)
                    xe = po.port().cholesky_solve(Le, b, n, m, rm)       # the reference has no Cholesky solve: the port's definition
                assert bits_equal(down(L1), Le), f"L vs {orc.kind} n={n} rm={rm}: {first_diff(down(L1), Le)}"
                assert bits_equal(down(Lf1), Le), f"cholesky() vs {orc.kind} n={n} rm={rm}: {first_diff(down(Lf1), Le)}"
                assert bits_equal(down(ok1), oke) and bits_equal(down(okf1), oke)
                assert bits_equal(down(x1), xe), f"x vs port n={n} m={m} rm={rm}: {first_diff(down(x1), xe)}"
                assert 0 < int((oke == 0).sum()) < B // 8
    finally:
        gm.set_option("lanechol", -1)


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, 7, 9, 12, 16])
def test_multi_rhs_solve_takes_the_subwarp_kernel(gm, orc, n, dt, soa):
    """m = 2..4 right-hand sides for n >= 5 ride through k_sub_plu as extra columns (no local-memory generic kernel)."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(4000 * n + np.dtype(dt).itemsize + int(soa))
    B = (1 << 13) + 5
    A, _ = hostile_batch(rng, B, n, dt)
    dA = up(gm, A, soa)
    launches = gm.lib().gmb_launch_count
    for m in (2, 3, 4):
        b = po.exact_grid(rng, (B, m * n), dt)
        db = up(gm, b, soa)
        for rm, skip in ((True, 0), (False, 1)):
            x, ok, lu = gm.linear_solve(dA, db, n, m, skip, rm, want_lu=True)
            with np.errstate(all="ignore"):
                lue, xe, oke = orc.linear_solve(A, b, n, m, skip, rm)
            tag = f"n={n} m={m} rm={rm} skip={skip}"
            assert bits_equal(down(ok), oke), f"ok {tag}: {first_diff(down(ok), oke)}"
            assert bits_equal(down(x), xe), f"x {tag}: {first_diff(down(x), xe)}"
            assert bits_equal(down(lu), lue), f"lu {tag}: {first_diff(down(lu), lue)}"


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, 6, 7, 8, 9, 10, 11, 12, 13, 14])
def test_deferred_rider_inverse_is_the_reference_inverse(gm, orc, n, dt, soa):
    """k_sub_plu MODE_INVD (one lane per system, the identity is NOT carried through the elimination; the columns of the
    inverse are produced in i-space and stored at their permuted positions) against the kernels that carry the riders
    (option invd = 0) and against the reference, on a hostile 2^17 batch: singular systems (destination untouched), ties,
    NaN / Inf (dense forward sweep inside the kernel), huge / tiny scalings (IEEE-division redo), all four layout pairs."""
    # one lane per system (k_sub_plu MODE_INVD) up to float n = 14 / double n = 10 (the wide 255-register instantiations from float
    # n = 12 / double n = 8 on); where no deferred-rider kernel exists both arms run the same kernel and the comparison with the
    # reference is what counts
    rng = np.random.default_rng(5000 * n + np.dtype(dt).itemsize + int(soa))
    B = ((1 << 17) if n <= 10 else (1 << 15)) + 33
    A, _ = hostile_batch(rng, B, n, dt, every=37)
    dA = up(gm, A, soa)
    try:
        for srm in (True, False):
            for drm in (True, False):
                init = np.full_like(A, 7)
                gm.set_option("invd", 1)
                o1, k1 = gm.inv(dA, n, srm, drm, out=up(gm, init, soa))
                gm.set_option("invd", 0)
                o0, k0 = gm.inv(dA, n, srm, drm, out=up(gm, init, soa))
                o1, k1, o0, k0 = down(o1), down(k1), down(o0), down(k0)
                tag = f"n={n} {srm}->{drm}"
                assert bits_equal(k1, k0), f"ok {tag}: {first_diff(k1, k0)}"
                assert bits_equal(o1, o0), f"inverse {tag}: {first_diff(o1, o0)}"
                assert 0 < int((k0 == 0).sum()) < B // 4
                S = 1 << 14
                with np.errstate(all="ignore"):
                    oe, ke = orc.inv(A[:S], n, srm, drm, init[:S])
                assert bits_equal(k1[:S], ke), f"ok vs {orc.kind} {tag}: {first_diff(k1[:S], ke)}"
                assert bits_equal(o1[:S], oe), f"inverse vs {orc.kind} {tag}: {first_diff(o1[:S], oe)}"
    finally:
        gm.set_option("invd", -1)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 3, 4])
def test_million_system_differential_register_kernels(gm, orc, n, dt):
    """2^20 systems through the register kernels (n <= 4: the metric kernel's family) against the compiled reference, with the
    hostile structures AND row / element scalings far outside the normal exponent range -- the once-per-10^7 cases of the IEEE
    division slow path: fused solve, decomposition, inverse, SoA container."""
    rng = np.random.default_rng(7000 + 10 * n + np.dtype(dt).itemsize)
    B = 1 << 20
    A, b = hostile_batch(rng, B, n, dt, every=53)
    kmax = 60 if dt == np.float32 else 480
    with np.errstate(all="ignore"):
        scale = np.ldexp(1.0, rng.integers(-kmax, kmax + 1, size=(B, n, 1)))
        scale[::2] = 1.0
        A = (A.reshape(B, n, n).astype(np.float64) * scale).astype(dt).reshape(B, n * n)
        sub = rng.integers(0, B, size=B // 64)
        A[sub, 0] = np.ldexp(A[sub, 0].astype(np.float64), -(140 if dt == np.float32 else 1040)).astype(dt)   # subnormal / underflowing entries
    dA, db = up(gm, A, True), up(gm, b, True)
    with np.errstate(all="ignore"):
        x, ok = gm.linear_solve(dA, db, n)
        _, xe, oke = orc.linear_solve(A, b, n)
        assert bits_equal(down(ok), oke), f"ok n={n}: {first_diff(down(ok), oke)}"
        assert bits_equal(down(x), xe), f"x n={n}: {first_diff(down(x), xe)}"
        lu, p, par, dg = gm.decomp_plu(dA, n)
        e = orc.decomp_plu(A, n)
        assert bits_equal(down(lu), e[0]), f"lu n={n}: {first_diff(down(lu), e[0])}"
        assert bits_equal(down(p).astype(np.int64), e[1]) and bits_equal(down(par), e[2])
        assert bits_equal(down(dg).astype(np.int64), e[3])
        init = np.full_like(A, 7)
        out, oki = gm.inv(dA, n, out=up(gm, init, True))
        ie, ike = orc.inv(A, n, True, True, init)
        assert bits_equal(down(oki), ike), f"inv ok n={n}"
        assert bits_equal(down(out), ie), f"inv n={n}: {first_diff(down(out), ie)}"


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [8, 9, 11, 12, 13, 14, 16])
def test_one_lane_cholesky_large_n(gm, orc, n, dt, soa):
    """k_row_chol with ONE lane per system beyond n = 8 (chol_group: SoA every float size and double n <= 12, AoS float n <= 14 /
    double n <= 9) on 2^14 systems: SPD, non-SPD (ok = 0 at different columns), all-zero and badly scaled systems; factor, ok flags
    and the fused solve against the reference / the port's definition of the solve, both matrix orders."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(8000 * n + np.dtype(dt).itemsize + int(soa))
    B = (1 << 14) + 45
    S = po.spd_exact(rng, B, n, dt)
    S[::53, 0] = -1                                              # fails at the first column
    S[11::97, n * (n - 1) + (n - 1)] = -3                        # fails at the last diagonal entry
    S[7::211] = 0
    kmax = 30 if dt == np.float32 else 200
    k = rng.integers(-kmax, kmax + 1, size=(B, n))
    k[::3] = 0
    with np.errstate(all="ignore"):
        S = np.ldexp(S.astype(np.float64).reshape(B, n, n), k[:, :, None] + k[:, None, :]).astype(dt).reshape(B, n * n)
    b = po.exact_grid(rng, (B, n), dt)
    dS, db = up(gm, S, soa), up(gm, b, soa)
    for rm in (True, False):
        x, ok, L = gm.cholesky_solve(dS, db, n, 1, rm, want_l=True)
        Lf, okf = gm.cholesky(dS, n, rm)
        with np.errstate(all="ignore"):
            Le, oke = orc.cholesky(S, n, rm)
            xe = po.port().cholesky_solve(Le, b, n, 1, rm)       # the reference has no Cholesky solve: the port's definition
        assert bits_equal(down(ok), oke) and bits_equal(down(okf), oke), f"ok n={n} rm={rm}"
        assert bits_equal(down(L), Le), f"L n={n} rm={rm}: {first_diff(down(L), Le)}"
        assert bits_equal(down(Lf), Le), f"cholesky() n={n} rm={rm}: {first_diff(down(Lf), Le)}"
        assert bits_equal(down(x), xe), f"x n={n} rm={rm}: {first_diff(down(x), xe)}"
        assert 0 < int((oke == 0).sum()) < B // 8


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, 8, 9, 11, 13, 14])
def test_one_lane_backsolve_hostile(gm, orc, n, dt, soa):
    """k_lane_backsolve (gmb_lanebs.cuh) on the LU factors of a hostile 2^14 batch -- zero and NaN pivots, Inf, huge and tiny
    scalings: the sweeps divide by whatever sits on the diagonal, exactly as the reference does -- with 1..4 right-hand sides,
    skip = 0 / 2, both orders, against the reference; and against the row kernel (option lanebs = 0) for the same bits."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(9000 * n + np.dtype(dt).itemsize + int(soa))
    B = (1 << 14) + 21
    A, _ = hostile_batch(rng, B, n, dt, every=29)
    try:
        for rm in (True, False):
            with np.errstate(all="ignore"):
                lu_np, p_np, _, _ = orc.decomp_plu(A, n, n, rm)
            lu, p = up(gm, lu_np, soa), torch.from_numpy(p_np.astype(np.uint8)).cuda()
            for m in (1, 2, 3, 4):
                b = po.exact_grid(rng, (B, n * m), dt)
                for skip in (0, 2):
                    gm.set_option("lanebs", -1)
                    x1 = down(gm.backsolve_plu(lu, p, up(gm, b, soa), n, m, skip, rm))
                    y1 = down(gm.backsolve_lu(lu, up(gm, b, soa), n, m, skip, rm))
                    gm.set_option("lanebs", 0)
                    x0 = down(gm.backsolve_plu(lu, p, up(gm, b, soa), n, m, skip, rm))
                    y0 = down(gm.backsolve_lu(lu, up(gm, b, soa), n, m, skip, rm))
                    with np.errstate(all="ignore"):
                        xe = orc.backsolve_plu(lu_np, p_np, b, n, m, skip, rm)
                        ye = orc.backsolve_lu(lu_np, b, n, m, skip, rm)
                    tag = f"n={n} m={m} rm={rm} skip={skip}"
                    assert bits_equal(x1, xe), f"backsolve_plu vs {orc.kind} {tag}: {first_diff(x1, xe)}"
                    assert bits_equal(y1, ye), f"backsolve_lu vs {orc.kind} {tag}: {first_diff(y1, ye)}"
                    assert bits_equal(x1, x0) and bits_equal(y1, y0), f"lane kernel vs row kernel {tag}"
    finally:
        gm.set_option("lanebs", -1)
