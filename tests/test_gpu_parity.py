"""GPU suite: the CUDA path, called through the C ABI, against the CPU oracle.

* the committed golden fixtures (outputs of the reference itself, tests/golden/),
* seeded random batches against oracle/_ref = the compiled reference itself (conftest.orc; the C port only when
  _ref is absent -- test_oracle_vs_ref.py pins the two to each other),
for every hot-path function, float and double, AoS and SoA-interleaved layouts, row- and
column-major, N = 2..16, including singular / non-SPD / non-finite systems.

Bar (BASELINE.json north_star): pivot permutation, parity, degenerate counts and ok flags are
bit-exact.  Factors and solutions are ALSO required bit-exact here (tolerance 0 ulp; NaN == NaN),
which is stricter than the stated "float <= 4 ulp / double <= 1e-12": the kernels spell the same
IEEE operation sequence as the reference.
"""
import numpy as np
import pytest

from conftest import GOLDEN_SIZES, bits_equal, first_diff, load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TDT = {np.dtype(np.float32): "float32", np.dtype(np.float64): "float64"}


@pytest.fixture(scope="module")
def gm():
    import geomc_b200 as gm
    gm.lib()   # raises if the CUDA library is missing -- no fallback
    assert torch.cuda.is_available()
    return gm


def up(gm, a, soa):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return gm.BatchSoA.from_aos(t) if soa else t


def down(x):
    if hasattr(x, "to_aos"):
        x = x.to_aos()
    return x.cpu().numpy()


def check(got, want, what):
    assert bits_equal(got, want), f"{what}: {first_diff(got, want)}"


def run_all(gm, orc, n, A, b1, b3, S, soa, expect=None):
    """Run every entry point on one input set and compare with `expect` (golden dict) or the checker (conftest.orc)."""
    with np.errstate(all="ignore"):
        for rm, r in ((True, "rm"), (False, "cm")):
            tag = f"n={n} soa={soa} {r}"
            dA, db1, db3, dS = up(gm, A, soa), up(gm, b1, soa), up(gm, b3, soa), up(gm, S, soa)
            # a1 decomp_plu
            lu, p, par, dg = gm.decomp_plu(dA, n, n, rm)
            e = (expect[f"plu_{r}_lu"], expect[f"plu_{r}_p"], expect[f"plu_{r}_parity"], expect[f"plu_{r}_degen"]) if expect is not None else orc.decomp_plu(A, n, n, rm)
            check(down(lu), e[0], f"decomp_plu lu {tag}")
            check(down(p).astype(np.int64), e[1], f"decomp_plu reorder {tag}")
            check(down(par), e[2], f"decomp_plu parity {tag}")
            check(down(dg).astype(np.int64), e[3], f"decomp_plu degenerate {tag}")
            # a2 decomp_lup
            lu2, p2, par2, dg2 = gm.decomp_lup(dA, n, n, rm)
            e = (expect[f"lup_{r}_lu"], expect[f"lup_{r}_p"], expect[f"lup_{r}_parity"], expect[f"lup_{r}_degen"]) if expect is not None else orc.decomp_lup(A, n, n, rm)
            check(down(lu2), e[0], f"decomp_lup lu {tag}")
            check(down(p2).astype(np.int64), e[1], f"decomp_lup reorder {tag}")
            check(down(par2), e[2], f"decomp_lup parity {tag}")
            check(down(dg2).astype(np.int64), e[3], f"decomp_lup degenerate {tag}")
            # a6 linear_solve (fused), m = 1 and m = 3 / skip = 1, plus the LU it leaves in `a`
            x, ok, lua = gm.linear_solve(dA, db1, n, 1, 0, rm, want_lu=True)
            e = (expect[f"solve_{r}_a"], expect[f"solve_{r}_x"], expect[f"solve_{r}_ok"]) if expect is not None else orc.linear_solve(A, b1, n, 1, 0, rm)
            check(down(x), e[1], f"linear_solve x {tag}")
            check(down(ok), e[2], f"linear_solve ok {tag}")
            check(down(lua), e[0], f"linear_solve a (LU) {tag}")
            # the same call without the LU output takes its own kernels for n >= 9 (only the live part of a row is exchanged)
            x_nolu, ok_nolu = gm.linear_solve(dA, db1, n, 1, 0, rm)
            check(down(x_nolu), e[1], f"linear_solve (no LU output) x {tag}")
            check(down(ok_nolu), e[2], f"linear_solve (no LU output) ok {tag}")
            x3, ok3 = gm.linear_solve(dA, db3, n, 3, 1, rm)
            e = (None, expect[f"solve3s1_{r}_x"], expect[f"solve3s1_{r}_ok"]) if expect is not None else orc.linear_solve(A, b3, n, 3, 1, rm)
            check(down(x3), e[1], f"linear_solve m=3 skip=1 x {tag}")
            check(down(ok3), e[2], f"linear_solve m=3 skip=1 ok {tag}")
            # a3 / a4 / a5 on the reference LU
            lu_np, p_np = down(lu), down(p)
            e = expect[f"bslu_{r}_x"] if expect is not None else orc.backsolve_lu(lu_np, b3, n, 3, 0, rm)
            check(down(gm.backsolve_lu(lu, db3, n, 3, 0, rm)), e, f"backsolve_lu {tag}")
            e = expect[f"bsplu_{r}_x"] if expect is not None else orc.backsolve_plu(lu_np, p_np.astype(np.int64), b3, n, 3, 0, rm)
            check(down(gm.backsolve_plu(lu, p, db3, n, 3, 0, rm)), e, f"backsolve_plu {tag}")
            e = expect[f"perm_{r}_a"] if expect is not None else orc.apply_perm(p_np.astype(np.int64), b3, n, 3, rm)[0]
            check(down(gm.apply_permutation(p, db3, n, 3, rm)), e, f"apply_permutation {tag}")
            # a8 cholesky
            L, okc = gm.cholesky(dS, n, rm)
            e = (expect[f"chol_{r}_l"], expect[f"chol_{r}_ok"]) if expect is not None else orc.cholesky(S, n, rm)
            check(down(L), e[0], f"cholesky L {tag}")
            check(down(okc), e[1], f"cholesky ok {tag}")
            # new: cholesky_solve against the port's definition
            Lp, okp = orc.cholesky(S, n, rm)
            xs, oks, Ls = gm.cholesky_solve(dS, db1, n, 1, rm, want_l=True)
            check(down(Ls), Lp, f"cholesky_solve L {tag}")
            check(down(oks), okp, f"cholesky_solve ok {tag}")
            check(down(xs), orc.cholesky_solve(Lp, b1, n, 1, rm), f"cholesky_solve x {tag}")
            # a13 inv, all four layout combinations, destination pre-filled with 7
            for drm, q in ((True, "rm"), (False, "cm")):
                init = np.full_like(A, 7)
                out, oki = gm.inv(dA, n, rm, drm, out=up(gm, init, soa))
                e = (expect[f"inv_{r}_{q}"], expect[f"inv_{r}_{q}_ok"]) if expect is not None else orc.inv(A, n, rm, drm, init)
                check(down(out), e[0], f"inv {r}->{q} {tag}")
                check(down(oki), e[1], f"inv ok {r}->{q} {tag}")
        # a7 Vec-form linear_solve
        xv, okv = gm.linear_solve_vec(up(gm, A, soa), up(gm, b1, soa), n, 1, 0)
        e = (None, expect["solvevec_x"], expect["solvevec_ok"]) if expect is not None else orc.linear_solve_vec(A, b1, n, 1, 0)
        check(down(xv), e[1], f"linear_solve(Vec) x n={n} soa={soa}")
        check(down(okv), e[2], f"linear_solve(Vec) ok n={n} soa={soa}")


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("n", GOLDEN_SIZES)
def test_golden(gm, orc, tag, n, soa):
    g = load_golden(tag, n)
    run_all(gm, orc, n, g["A"], g["b1"], g["b3"], g["S"], soa, expect=g)
    # rectangular decomp_plu ((n-1) x n), as used by nullspace / orthogonal
    lur, pr, parr, dgr = gm.decomp_plu(up(gm, g["rect_m"], soa), n - 1, n, True)
    check(down(lur), g["rect_lu"], "rect lu")
    check(down(pr).astype(np.int64).reshape(g["rect_p"].shape), g["rect_p"], "rect reorder")
    check(down(parr), g["rect_parity"], "rect parity")
    check(down(dgr).astype(np.int64), g["rect_degen"], "rect degen")


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", list(range(2, 17)))
def test_random_vs_oracle(gm, orc, dt, n, soa):
    """Seeded random batches whose size is NOT a multiple of the tile width (ragged last tile)."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(4242 + 100 * n + np.dtype(dt).itemsize)
    B = 1000 + 37 if n <= 8 else 300 + 5
    A = rng.standard_normal((B, n * n)).astype(dt)
    A[::97] = 0
    A[5::101, 0::n] = 0
    b1 = rng.standard_normal((B, n)).astype(dt)
    b3 = rng.standard_normal((B, n * 3)).astype(dt)
    S = po.spd_exact(rng, B, n, dt)
    S[::53, 0] = -1
    run_all(gm, orc, n, A, b1, b3, S, soa)


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [3, 4, 5, 6, 8, 9, 12, 13, 16])
def test_extreme_magnitudes_vs_oracle(gm, orc, dt, n, soa):
    """Rows scaled by powers of two far outside the fast domain of the shared-reciprocal division (RcpOf, gmb_math.cuh:
    float operands beyond 2^+-57, double quotients near the ends of the exponent range), subnormal and zero entries,
    overflowing multipliers: every quotient must still be the correctly rounded one (= the oracle's)."""
    from oracle import pyoracle as po
    f32 = np.dtype(dt) == np.float32
    rng = np.random.default_rng(777 + 10 * n + np.dtype(dt).itemsize)
    B = 257
    kmax, ksub = (70, 140) if f32 else (520, 1040)
    A = rng.standard_normal((B, n, n))
    A = np.ldexp(A, rng.integers(-kmax, kmax + 1, size=(B, n, 1)))            # row scaling
    A[3::7] = np.ldexp(A[3::7], rng.integers(-8, 9, size=A[3::7].shape))       # element scaling
    A[5::11, :, 0] = np.ldexp(A[5::11, :, 0], -ksub)                          # subnormal / underflowing first column
    A[2::13, 1, :] = 0.0
    A[4::17, :, 1] = 0.0                                                      # a zero column: degenerate
    with np.errstate(all="ignore"):
        A = A.astype(dt).reshape(B, n * n)
        b1 = np.ldexp(rng.standard_normal((B, n)), rng.integers(-kmax // 2, kmax // 2 + 1, size=(B, n))).astype(dt)
        b3 = np.ldexp(rng.standard_normal((B, n * 3)), rng.integers(-kmax // 2, kmax // 2 + 1, size=(B, n * 3))).astype(dt)
        S = po.spd_exact(rng, B, n, np.float64).reshape(B, n, n)
        k = rng.integers(-kmax // 3, kmax // 3 + 1, size=(B, n))
        S = np.ldexp(S, k[:, :, None] + k[:, None, :]).astype(dt).reshape(B, n * n)   # D S D stays symmetric positive definite
    run_all(gm, orc, n, A, b1, b3, S, soa)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_empty_and_tiny_batches(gm, dt):
    for B in (0, 1, 3):
        for soa in (False, True):
            A = np.tile(np.eye(4, dtype=dt).reshape(1, 16), (B, 1)) * 2
            b = np.ones((B, 4), dt)
            x, ok = gm.linear_solve(up(gm, A, soa) if B else (gm.BatchSoA(0, 16, getattr(torch, TDT[np.dtype(dt)])) if soa else torch.empty((0, 16), dtype=getattr(torch, TDT[np.dtype(dt)]), device="cuda")),
                                    up(gm, b, soa) if B else (gm.BatchSoA(0, 4, getattr(torch, TDT[np.dtype(dt)])) if soa else torch.empty((0, 4), dtype=getattr(torch, TDT[np.dtype(dt)]), device="cuda")),
                                    4)
            assert down(ok).shape == (B,)
            if B:
                assert np.array_equal(down(x), np.full((B, 4), 0.5, dt))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("E", [4, 9, 16, 3, 25, 64, 256])
def test_aos_soa_round_trip(gm, dt, E):
    rng = np.random.default_rng(E)
    for B in (1, 127, 128, 129, 1000):
        a = rng.standard_normal((B, E)).astype(dt)
        soa = up(gm, a, True)
        assert np.array_equal(down(soa), a)
        # container layout: element e of system s at ((s // W) * E + e) * W + s % W
        W = soa.W
        flat = soa.data.cpu().numpy()
        s = B - 1
        for e in (0, E - 1):
            assert flat[((s // W) * E + e) * W + s % W] == a[s, e]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 3, 4, 6, 8, 12, 16])
def test_unaligned_aos_batches_take_the_elementwise_kernels(gm, orc, n, dt):
    """The reference's objects are only 4/8-byte aligned: an AoS batch at an odd offset must still be served -- by
    element-wise kernels, bit-identically -- by EVERY entry point (include/geomc_b200.h, "Alignment")."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(n)
    B = 777
    tdt = torch.float32 if dt == np.float32 else torch.float64
    A = rng.standard_normal((B, n * n)).astype(dt)
    A[::59] = 0
    b = rng.standard_normal((B, n)).astype(dt)
    S = po.spd_exact(rng, B, n, dt)
    S[::41, 0] = -1

    def odd(h):
        buf = torch.zeros(h.size + 1, dtype=tdt, device="cuda")
        v = buf[1:].view(*h.shape)
        v.copy_(torch.from_numpy(h))
        assert v.data_ptr() % 16 != 0
        return v
    dA, db, dS = odd(A), odd(b), odd(S)
    with np.errstate(all="ignore"):
        x, ok = gm.linear_solve(dA, db, n)
        _, xe, oke = orc.linear_solve(A, b, n)
        check(down(x), xe, f"unaligned linear_solve n={n}")
        check(down(ok), oke, "unaligned ok")
        lu, p, par, dg = gm.decomp_plu(dA, n)
        e = orc.decomp_plu(A, n)
        check(down(lu), e[0], "unaligned decomp lu")
        check(down(p).astype(np.int64), e[1], "unaligned decomp perm")
        # inverse: adjugate for n <= 4 (element-wise kernel), PLU above; all four layout combinations
        for srm in (True, False):
            for drm in (True, False):
                init = np.full_like(A, 7)
                out, oki = gm.inv(dA, n, srm, drm, out=odd(init))
                e = orc.inv(A, n, srm, drm, init)
                check(down(out), e[0], f"unaligned inv n={n} {srm}->{drm}")
                check(down(oki), e[1], f"unaligned inv ok n={n}")
        # Vec form (inverse-then-multiply below 5 unknowns)
        xv, okv = gm.linear_solve_vec(dA, db, n)
        e = orc.linear_solve_vec(A, b, n, 1, 0)
        check(down(xv), e[1], f"unaligned linear_solve(Vec) n={n}")
        check(down(okv), e[2], f"unaligned linear_solve(Vec) ok n={n}")
        # Cholesky and the fused Cholesky solve
        L, okc = gm.cholesky(dS, n)
        Le, okce = orc.cholesky(S, n, True)
        check(down(L), Le, f"unaligned cholesky n={n}")
        check(down(okc), okce, "unaligned cholesky ok")
        xs, oks = gm.cholesky_solve(dS, db, n)
        check(down(oks), okce, "unaligned cholesky_solve ok")
        check(down(xs), orc.cholesky_solve(Le, b, n, 1, True), f"unaligned cholesky_solve n={n}")
    # container ingest / egress with an element-aligned AoS side
    soa = gm.BatchSoA.from_aos(dA)
    check(down(soa), A, "aos_to_soa from an unaligned AoS batch")
    back = odd(np.zeros_like(A))
    gm.lib()  # (the C ABI directly: to_aos() would allocate an aligned destination)
    from geomc_b200 import _lib
    _lib.call("gmb_soa_to_aos", "f32" if dt == np.float32 else "f64", n * n, B, soa.data.data_ptr(), back.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    check(down(back), A, "soa_to_aos into an unaligned AoS batch")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [8, 12, 16])
def test_batches_aligned_to_16_but_not_32_bytes(gm, orc, dt, n):
    """The row kernels read AoS rows with 256-bit loads when the batch is 32-byte aligned and with 128-bit loads when
    it is only 16-byte aligned (ldg_row, gmb_io.cuh): same results either way."""
    rng = np.random.default_rng(31 * n)
    B = 333
    tdt = torch.float32 if dt == np.float32 else torch.float64
    off = 16 // np.dtype(dt).itemsize
    A = rng.standard_normal((B, n * n)).astype(dt)
    b = rng.standard_normal((B, n)).astype(dt)
    from oracle import pyoracle as po
    S = po.spd_exact(rng, B, n, dt)
    bufs = []
    for h in (A, b, S):
        buf = torch.zeros(h.size + off, dtype=tdt, device="cuda")
        v = buf[off:].view(*h.shape)
        v.copy_(torch.from_numpy(h))
        assert v.data_ptr() % 32 == 16
        bufs.append(v)
    dA, db, dS = bufs
    x, ok = gm.linear_solve(dA, db, n)
    _, xe, oke = orc.linear_solve(A, b, n)
    check(down(x), xe, f"16-byte aligned linear_solve n={n}")
    check(down(ok), oke, "16-byte aligned ok")
    out, oki = gm.inv(dA, n)
    e = orc.inv(A, n, True, True, np.zeros_like(A))
    check(down(out)[down(oki) != 0], e[0][e[1] != 0], f"16-byte aligned inv n={n}")
    Lp, okp = orc.cholesky(S, n, True)
    xs, oks = gm.cholesky_solve(dS, db, n)
    check(down(oks), okp, "16-byte aligned cholesky ok")
    check(down(xs), orc.cholesky_solve(Lp, b, n, 1, True), f"16-byte aligned cholesky_solve n={n}")


def test_host_api_matches_device_api(gm, orc):
    rng = np.random.default_rng(9)
    B, n = 70001, 4
    A = rng.standard_normal((B, n * n)).astype(np.float32)
    A[::1000] = 0
    b = rng.standard_normal((B, n)).astype(np.float32)
    x = np.zeros_like(b)
    ok = np.zeros(B, np.uint8)
    gm.host_linear_solve(A, b, x, ok, n)
    _, xe, oke = orc.linear_solve(A, b, n)
    check(x, xe, "host linear_solve x")
    check(ok, oke, "host linear_solve ok")
    out = np.full_like(A, 3)
    gm.host_inv(A, out, ok, n)
    oe, oke = orc.inv(A, n, True, True, np.full_like(A, 3))
    check(out, oe, "host inv")
    check(ok, oke, "host inv ok")
    from oracle import pyoracle as po
    S = po.spd_exact(rng, B, 3, np.float64)
    b3 = po.exact_grid(rng, (B, 3), np.float64)
    x3 = np.zeros_like(b3)
    gm.host_cholesky_solve(S, b3, x3, ok, 3)
    L, okc = orc.cholesky(S, 3)
    check(x3, orc.cholesky_solve(L, b3, 3), "host cholesky_solve")
    check(ok, okc, "host cholesky ok")
    A8 = rng.standard_normal((B // 7, 64))
    lu = A8.copy()
    p = np.zeros((B // 7, 8), np.uint8)
    par = np.zeros(B // 7, np.uint8)
    dg = np.zeros(B // 7, np.uint8)
    gm.host_decomp_plu(lu, p, par, dg, 8)
    e = orc.decomp_plu(A8, 8)
    check(lu, e[0], "host decomp lu")
    check(p.astype(np.int64), e[1], "host decomp perm")
    check(par, e[2], "host decomp parity")


def test_multi_device_host_entry_points(gm, orc):
    """gmb_host_*_multi_*: the batch is sharded contiguously (gmb_shard_range, 1024-system granules) over the listed devices,
    one pipeline and one host thread per device.  Runs on every device the box has (one is enough: the sharding, the
    thread-per-device driver and the offset arithmetic are the same); results are those of the single-device call."""
    import ctypes
    ndev = gm.lib().gmb_device_count()
    devs = list(range(ndev))
    rng = np.random.default_rng(19)
    B, n = 150001, 4
    A = rng.standard_normal((B, n * n)).astype(np.float32)
    A[::997] = 0
    b = rng.standard_normal((B, n)).astype(np.float32)
    with np.errstate(all="ignore"):
        _, xe, oke = orc.linear_solve(A, b, n)
        ie, ike = orc.inv(A, n, True, True, np.full_like(A, 3))
    for use in ([0], devs, devs[::-1]):
        x, ok = np.zeros_like(b), np.zeros(B, np.uint8)
        gm.host_linear_solve(A, b, x, ok, n, device=list(use))
        check(x, xe, f"multi-device linear_solve x on {use}")
        check(ok, oke, f"multi-device linear_solve ok on {use}")
        out = np.full_like(A, 3)
        gm.host_inv(A, out, ok, n, device=list(use))
        check(out, ie, f"multi-device inv on {use}")
        check(ok, ike, f"multi-device inv ok on {use}")
    A6 = rng.standard_normal((B // 5, 36))
    e = orc.decomp_plu(A6, 6)
    lu, p = A6.copy(), np.zeros((B // 5, 6), np.uint8)
    par, dg = np.zeros(B // 5, np.uint8), np.zeros(B // 5, np.uint8)
    gm.host_decomp_plu(lu, p, par, dg, 6, device=devs)
    check(lu, e[0], "multi-device decomp lu")
    check(p.astype(np.int64), e[1], "multi-device decomp perm")
    check(par, e[2], "multi-device decomp parity")
    # the C and Python shard helpers agree (bench.py and the gloo tests use the Python one)
    from geomc_b200.sharding import shard_range
    for Bq, world, align in ((150001, 3, 1024), (7, 8, 1), (1 << 20, 8, 128), (0, 2, 64)):
        for r in range(world):
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            assert gm.lib().gmb_shard_range(Bq, r, world, align, ctypes.byref(lo), ctypes.byref(hi)) == 0
            assert (lo.value, hi.value) == shard_range(Bq, r, world, align), (Bq, world, align, r)
    # argument errors: duplicate / out-of-range devices
    with pytest.raises(Exception):
        gm.host_linear_solve(A, b, x, ok, n, device=[0, 0])
    with pytest.raises(Exception):
        gm.host_linear_solve(A, b, x, ok, n, device=[ndev])


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("n", [5, 6, 8, 9, 12, 13, 16])
def test_double_pivots_near_dbl_max(gm, orc, n, soa):
    """RcpOf<double> (gmb_math.cuh): the reciprocal of a divisor beyond 2^1022 is subnormal and rcp.approx.ftz flushes it -- such
    pivots must take the IEEE division (nvcc's own in-line __ddiv_rn sends |b| >= 2^1017 to its subroutine too).  Matrices whose
    entries sit in [2^1015, DBL_MAX), so that pivots and diagonal entries do, with right-hand sides of every magnitude: every
    kernel family that shares a reciprocal per pivot, against the reference."""
    rng = np.random.default_rng(900 + n)
    B = 2049
    with np.errstate(all="ignore"):
        A = np.ldexp(rng.uniform(0.5, 1.0, (B, n * n)) * rng.choice([-1.0, 1.0], (B, n * n)), rng.integers(1016, 1025, (B, n * n)))
        A[::3] = np.ldexp(A[::3], -rng.integers(0, 40, (A[::3].shape[0], 1)))          # some systems a little lower
        A = A.astype(np.float64)
        b = np.ldexp(rng.standard_normal((B, n)), rng.integers(-300, 1000, (B, n))).astype(np.float64)
        assert np.isfinite(A).all() and np.isfinite(b).all()
        dA, db = up(gm, A, soa), up(gm, b, soa)
        x, ok = gm.linear_solve(dA, db, n)
        _, xe, oke = orc.linear_solve(A, b, n)
        check(down(ok), oke, f"ok n={n}")
        check(down(x), xe, f"linear_solve with pivots near DBL_MAX n={n} soa={soa}")
        lu, p, par, dg = gm.decomp_plu(dA, n)
        e = orc.decomp_plu(A, n)
        check(down(lu), e[0], f"decomp_plu near DBL_MAX n={n}")
        check(down(p).astype(np.int64), e[1], "perm")
        init = np.full_like(A, 7)
        out, oki = gm.inv(dA, n, out=up(gm, init, soa))
        ie, ike = orc.inv(A, n, True, True, init)
        check(down(oki), ike, f"inv ok n={n}")
        check(down(out), ie, f"inv near DBL_MAX n={n}")
        lu2, p2, _, _ = gm.decomp_lup(dA, n)
        e2 = orc.decomp_lup(A, n)
        check(down(lu2), e2[0], f"decomp_lup near DBL_MAX n={n}")
        xb = gm.backsolve_plu(lu, p, db, n)
        check(down(xb), orc.backsolve_plu(e[0], e[1], b, n), f"backsolve_plu near DBL_MAX n={n}")


@pytest.mark.parametrize("n", [3, 6, 8, 13])
def test_cholesky_solve_against_a_long_double_host_solve(gm, n):
    """cholesky_solve has no counterpart in the reference (PARITY UNPINNED: the definition is this framework's, oracle.c
    restates it): besides the bit comparison with the port's definition elsewhere, the solution is held to the north star's
    double-precision bar -- relative residual <= 1e-12 -- against a host computation in long double."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(600 + n)
    B = 4096
    S = po.spd_exact(rng, B, n, np.float64)
    b = po.exact_grid(rng, (B, n), np.float64)
    x, ok = gm.cholesky_solve(torch.from_numpy(S).cuda(), torch.from_numpy(b).cuda(), n)
    x, ok = down(x), down(ok)
    assert ok.all()
    ld = np.longdouble
    Al, xl, bl = S.reshape(B, n, n).astype(ld), x.astype(ld), b.astype(ld)
    res = np.abs(np.einsum("sij,sj->si", Al, xl) - bl).max(axis=1)
    scale = np.abs(Al).sum(axis=2).max(axis=1) * np.abs(xl).max(axis=1) + np.abs(bl).max(axis=1)
    rel = (res / scale).max()
    assert rel <= 1e-12, f"relative residual {rel} (n={n})"
    # and the long-double solution itself: forward / backward substitution on the long-double Cholesky factor
    y = np.linalg.solve(S.reshape(B, n, n), b[..., None])[..., 0]
    cond = np.linalg.cond(S.reshape(B, n, n))
    err = np.abs(y - x).max(axis=1) / np.abs(y).max(axis=1)
    assert (err <= 1e-13 * cond + 1e-15).all(), f"solution error beyond cond * 1e-13 (n={n}): {err.max()}"


def test_full_size_properties(gm):
    """BASELINE-size batch (2^24 4x4 float systems): size-independent properties instead of the oracle.
    (1) residual of A x = b stays small on every system, (2) inverse round trip inv(inv(A)) ~ A,
    (3) the SoA and AoS paths produce the same bits, (4) the checksum of checksums over two halves
    equals the checksum of the whole."""
    B, n = 1 << 24, 4
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randint(-(1 << 19), 1 << 19, (B, 16), device="cuda", generator=g).float() * 2.0 ** -17)
    b = (torch.randint(-(1 << 19), 1 << 19, (B, 4), device="cuda", generator=g).float() * 2.0 ** -17)
    x, ok = gm.linear_solve(A, b, n)
    stats = gm.solve_residual(A, x, b, ok, n).cpu().numpy()
    n_ok = B - stats[2]
    assert stats[2] < 10 and stats[3] < 10
    assert stats[0] / n_ok < 1e-4          # mean max-residual (float, |entries| < 4)
    xs, oks = gm.linear_solve(gm.BatchSoA.from_aos(A), gm.BatchSoA.from_aos(b), n)
    assert torch.equal(xs.to_aos().view(torch.int32), x.view(torch.int32))
    assert torch.equal(oks, ok)
    half = B // 2
    c = gm.checksum(ok).item()
    acc = gm.checksum(ok[:half], 0)
    gm.checksum(ok[half:], half // 8, out=acc)               # checksum of checksums
    assert acc.item() == c
    okf = ok.clone()
    okf[12345] ^= 1
    assert gm.checksum(okf).item() != c
    small = ok[:100003].contiguous()
    from geomc_b200.linalg import checksum_host
    assert gm.checksum(small).item() % (1 << 64) == checksum_host(small.cpu().numpy()) % (1 << 64)
    inv1, ok1 = gm.inv(A, n)
    inv2, ok2 = gm.inv(inv1, n)
    good = (ok1 == 1) & (ok2 == 1)
    rel = ((inv2 - A).abs().amax(dim=1) / A.abs().amax(dim=1))[good]
    assert rel.median().item() < 1e-4


@pytest.mark.parametrize("force", ["9", "17"])
def test_both_large_n_solve_kernels(force):
    """The dispatch table picks one of two fused-solve kernels per (type, n); force each of them for
    every n = 9..16 (row-distributed with option rowplu_min_n = 9, column-distributed with 17; speculative kernel off) in a fresh
    process and rerun the oracle comparison."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GMB_OPT_ROWPLU_MIN_N=force, GMB_OPT_ROWFAST="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sel = " or ".join(f"{n}-" for n in range(9, 17))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                        "-k", f"(test_random_vs_oracle or test_golden or test_extreme_magnitudes) and ({sel}) and not both_large"],
                       env=env, capture_output=True, text=True, cwd=root, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


def test_row_backsolve_kernel_without_the_lane_kernel():
    """backsolve_lu / backsolve_plu with 1..4 right-hand sides take the one-lane kernel (gmb_lanebs.cuh) for float n <= 13 /
    double n <= 9; option lanebs = 0 sends every call to the row kernel (run-time m) instead: rerun the oracle comparison for
    n = 5..13 in a fresh process so that both kernels stay covered."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GMB_OPT_LANEBS="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sel = " or ".join(f"{n}-" for n in range(5, 14))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                        "-k", f"(test_random_vs_oracle or test_backsolve_single_rhs) and ({sel})"],
                       env=env, capture_output=True, text=True, cwd=root, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("force", ["0", "1"])
def test_both_large_n_inverse_kernels(force):
    """inv for n >= 5 has a row-distributed (gmb_rowinv.cuh) and a column-distributed kernel; the dispatcher picks per
    (type, n).  option rowinv = 0 / 1 forces one for every n: rerun the oracle comparison for n = 5..16 in a fresh process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GMB_OPT_ROWINV=force)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sel = " or ".join(f"{n}-" for n in range(5, 17))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                        "-k", f"(test_random_vs_oracle or test_extreme_magnitudes) and ({sel})"],
                       env=env, capture_output=True, text=True, cwd=root, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 4, 6, 7, 8, 9, 12, 15, 16])
def test_skip_rows_single_rhs(gm, orc, dt, n, soa):
    """`skip` > 0 with ONE right-hand side goes through the register / column / row kernels (the m = 3
    cases above take the generic path): rows < skip keep their forward-substituted values
    (LUDecomp.h:356), bit for bit, in both layouts and both matrix orders."""
    rng = np.random.default_rng(31 * n + np.dtype(dt).itemsize)
    B = 513
    A = rng.standard_normal((B, n * n)).astype(dt)
    A[::101] = 0
    b = rng.standard_normal((B, n)).astype(dt)
    for skip in sorted({1, 2, n - 1, n} - {0}):
        if skip > n:
            continue
        for rm in (True, False):
            x, ok = gm.linear_solve(up(gm, A, soa), up(gm, b, soa), n, 1, skip, rm)
            _, xe, oke = orc.linear_solve(A, b, n, 1, skip, rm)
            check(down(x), xe, f"skip={skip} n={n} rm={rm} soa={soa}")
            check(down(ok), oke, f"skip={skip} ok")
    # and through the Vec-form entry point for n >= 5 (column-major PLU, LUDecomp.h:428-439)
    if n >= 5:
        xv, okv = gm.linear_solve_vec(up(gm, A, soa), up(gm, b, soa), n, 1, 1)
        _, xe, oke = orc.linear_solve_vec(A, b, n, 1, 1)
        check(down(xv), xe, f"Vec-form skip=1 n={n}")
        check(down(okv), oke, "Vec-form ok")


@pytest.mark.parametrize("force", ["7", "17"])
def test_both_large_n_decomp_kernels(force):
    """decomp_plu has the same two kernel families; force each for every n it serves and recheck
    factors, permutation, parity and degenerate counts against the oracle."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GMB_OPT_ROWDECOMP_MIN_N=force)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sel = " or ".join(f"{n}-" for n in range(7, 17))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                        "-k", f"(test_random_vs_oracle or test_golden or test_extreme_magnitudes) and ({sel}) and not both_large"],
                       env=env, capture_output=True, text=True, cwd=root, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 3, 4, 6, 9, 12, 16])
def test_no_out_of_bounds_writes(gm, dt, n, soa):
    """(compute-sanitizer is closed on this pool.)  Every output buffer sits inside a larger allocation
    filled with a sentinel; after running each kernel family on a ragged batch the guard bands on both
    sides must be untouched."""
    tdt = getattr(torch, TDT[np.dtype(dt)])
    B = 257 + n
    GUARD = 4096                       # elements / bytes on each side, keeps 16-byte alignment
    SENT = -12345.0

    def guarded(nelem, dtype):
        buf = torch.full((nelem + 2 * GUARD,), SENT if dtype != torch.uint8 else 0xAB, dtype=dtype, device="cuda")
        return buf, buf[GUARD:GUARD + nelem]

    def intact(buf, nelem):
        s = SENT if buf.dtype != torch.uint8 else 0xAB
        return bool((buf[:GUARD] == s).all()) and bool((buf[GUARD + nelem:] == s).all())

    def batch(E, fill=None):
        nelem = gm.lib().gmb_soa_elems(B, E, np.dtype(dt).itemsize) if soa else B * E
        buf, view = guarded(nelem, tdt)
        if fill is not None:
            src = torch.from_numpy(fill).cuda()
            if soa:
                tmp = gm.BatchSoA.from_aos(src)
                view.copy_(tmp.data)
            else:
                view.copy_(src.reshape(-1))
        obj = gm.BatchSoA(B, E, tdt, "cuda", view) if soa else view.view(B, E)
        return buf, nelem, obj

    rng = np.random.default_rng(n)
    A = rng.standard_normal((B, n * n)).astype(dt)
    b = rng.standard_normal((B, n)).astype(dt)
    bufA, neA, dA = batch(n * n, A)
    bufb, neb, db = batch(n, b)
    bufx, nex, dx = batch(n)
    bufo, neo, dout = batch(n * n)
    bufok, ok = guarded(B, torch.uint8)
    bufp, perm = guarded(B * n, torch.uint8)
    bufpar, par = guarded(B, torch.uint8)
    bufdg, dg = guarded(B, torch.uint8)
    gm.linear_solve(dA, db, n, x_out=dx, ok_out=ok)
    gm.inv(dA, n, out=dout, ok_out=ok)
    gm.cholesky_solve(dA, db, n, x_out=dx, ok_out=ok)
    gm.decomp_plu(dA, n, inplace=True, out=(perm.view(B, n), par, dg))
    gm.cholesky(dA, n, inplace=True)
    torch.cuda.synchronize()
    for name, buf, ne in (("A", bufA, neA), ("b", bufb, neb), ("x", bufx, nex), ("inv", bufo, neo), ("ok", bufok, B),
                          ("perm", bufp, B * n), ("parity", bufpar, B), ("degenerate", bufdg, B)):
        assert intact(buf, ne), f"guard band of {name} was overwritten (n={n}, soa={soa})"


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [3, 5, 6, 9, 12, 16])
def test_backsolve_single_rhs(gm, orc, dt, n, soa):
    """backsolve_lu / backsolve_plu with ONE right-hand side (n >= 5: the row-distributed kernel; the
    m = 3 cases in run_all go through the generic path), skip = 0 and skip > 0, both matrix orders."""
    rng = np.random.default_rng(77 * n + np.dtype(dt).itemsize)
    B = 300 + n
    A = rng.standard_normal((B, n * n)).astype(dt)
    b = rng.standard_normal((B, n)).astype(dt)
    for rm in (True, False):
        lu_np, p_np, _, _ = orc.decomp_plu(A, n, n, rm)
        lu, p = up(gm, lu_np, soa), torch.from_numpy(p_np.astype(np.uint8)).cuda()
        for skip in (0, 1, n - 1):
            with np.errstate(all="ignore"):
                check(down(gm.backsolve_lu(lu, up(gm, b, soa), n, 1, skip, rm)), orc.backsolve_lu(lu_np, b, n, 1, skip, rm),
                      f"backsolve_lu n={n} rm={rm} skip={skip} soa={soa}")
                check(down(gm.backsolve_plu(lu, p, up(gm, b, soa), n, 1, skip, rm)),
                      orc.backsolve_plu(lu_np, p_np, b, n, 1, skip, rm), f"backsolve_plu n={n} rm={rm} skip={skip} soa={soa}")


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [5, 8, 11, 13])
def test_backsolve_two_and_four_rhs(gm, orc, dt, n, soa):
    """backsolve_lu / backsolve_plu with m = 2 and m = 4 right-hand sides (the one-lane kernel's other instantiations; m = 1
    and m = 3 are covered above and in run_all), skip = 0 and 1, both orders, a batch with a ragged last warp."""
    rng = np.random.default_rng(991 * n + np.dtype(dt).itemsize)
    B = 32 * 7 + 19
    A = rng.standard_normal((B, n * n)).astype(dt)
    for rm in (True, False):
        lu_np, p_np, _, _ = orc.decomp_plu(A, n, n, rm)
        lu, p = up(gm, lu_np, soa), torch.from_numpy(p_np.astype(np.uint8)).cuda()
        for m in (2, 4):
            b = rng.standard_normal((B, n * m)).astype(dt)
            for skip in (0, 1):
                with np.errstate(all="ignore"):
                    check(down(gm.backsolve_lu(lu, up(gm, b, soa), n, m, skip, rm)), orc.backsolve_lu(lu_np, b, n, m, skip, rm),
                          f"backsolve_lu n={n} m={m} rm={rm} skip={skip} soa={soa}")
                    check(down(gm.backsolve_plu(lu, p, up(gm, b, soa), n, m, skip, rm)),
                          orc.backsolve_plu(lu_np, p_np, b, n, m, skip, rm), f"backsolve_plu n={n} m={m} rm={rm} skip={skip} soa={soa}")


@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("n,dt", [(9, np.float64), (10, np.float64), (13, np.float32), (14, np.float32), (12, np.float32)])
def test_tiny_batches_through_the_one_lane_kernels(gm, orc, n, dt, soa):
    """Batches smaller than / not a multiple of a warp (B = 1, 31, 33) and the empty batch through the wide one-lane kernels, the
    one-lane backsolve and the one-lane Cholesky: the staged (full-warp) and the per-lane (ragged warp) I/O paths of the same
    kernel must agree with the reference."""
    from oracle import pyoracle as po
    for B in (1, 31, 33):
        rng = np.random.default_rng(31337 + 7 * n + B)
        A = rng.standard_normal((B, n * n)).astype(dt)
        A[0] = 0
        b1 = rng.standard_normal((B, n)).astype(dt)
        b3 = rng.standard_normal((B, n * 3)).astype(dt)
        S = po.spd_exact(rng, B, n, dt)
        run_all(gm, orc, n, A, b1, b3, S, soa)
    tdt = getattr(torch, TDT[np.dtype(dt)])
    A0 = gm.BatchSoA(0, n * n, tdt) if soa else torch.empty((0, n * n), dtype=tdt, device="cuda")
    b0 = gm.BatchSoA(0, n, tdt) if soa else torch.empty((0, n), dtype=tdt, device="cuda")
    p0 = torch.empty((0, n), dtype=torch.uint8, device="cuda")
    gm.decomp_plu(A0, n)
    gm.decomp_lup(A0, n)
    gm.linear_solve(A0, b0, n)
    gm.inv(A0, n)
    gm.cholesky_solve(A0, b0, n)
    gm.backsolve_plu(A0, p0, b0, n, 1)
    gm.backsolve_lu(A0, b0, n, 1)
    torch.cuda.synchronize()
