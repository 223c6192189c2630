"""Rows (f)3 and (f)4 of the scope table: the steps either side of the inverse.

(f)4  batched mul / mul_acc (linalg/Matrix.h:386 / :433 -> mtxdetail/MatrixMult.h:54-80) and the
      matrix * Vec / Vec * matrix products (:181-190 / :239-248);
(f)3  AffineTransform (linalg/AffineTransform.h): transformation(mat) :862, translation :814, scale :838,
      composition `*` / `/` :142-159, and the six apply functions :224-293 on point arrays.

CPU: oracle.c against the reference-generated fixtures (tests/golden/xform_*.npz) and against the compiled
reference on fresh random batches, plus the properties regression/matrix_tests.cpp (M * M^-1 = I) and
regression/similarity.cpp (composition == successive application) rely on.  GPU: the CUDA kernels against
both, bit for bit, AoS and SoA."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, bits_equal, first_diff
from oracle import pyoracle as po

MUL_SHAPES = [(2, 2, 2), (3, 3, 3), (4, 4, 4), (5, 5, 5), (8, 8, 8), (2, 3, 4), (4, 3, 2), (3, 5, 2), (5, 2, 7), (16, 16, 16),
              (16, 3, 9), (7, 16, 5)]
XF_DIMS = range(2, 8)
XF_MODES = po.CpuLinalg.XF_MODES


def golden(tag):
    return np.load(os.path.join(GOLDEN_DIR, f"xform_{tag}.npz"))


def eq(got, want, what):
    assert bits_equal(got, want), f"{what}: {first_diff(got, want)}"


def check_mul_golden(impl, g, what):
    for (r, k, c) in MUL_SHAPES:
        key = f"{r}_{k}_{c}"
        a, b, d0 = g[f"mul_a_{key}"], g[f"mul_b_{key}"], g[f"mul_d0_{key}"]
        eq(impl.mul(a, b, r, k, c), g[f"mul_d_{key}"], f"{what} mul {key}")
        eq(impl.mul(a, b, r, k, c, d0, acc=True), g[f"mul_acc_{key}"], f"{what} mul_acc {key}")
        eq(impl.mul(a, b, r, k, c, a_rm=False, b_rm=False, d_rm=False), g[f"mul_cm_{key}"], f"{what} mul col-major {key}")


def check_xf_golden(impl, g, N, what):
    m, t, p = g[f"xf_m_{N}"], g[f"xf_t_{N}"], g[f"xf_p_{N}"]
    xm = impl.xf_from_matrix(m, N, True)[0]
    eq(xm, g[f"xf_from_rm_{N}"], f"{what} transformation N={N}")
    eq(impl.xf_from_matrix(m, N, False)[0], g[f"xf_from_cm_{N}"], f"{what} transformation col-major N={N}")
    xt = impl.xf_translation(t, N)
    eq(xt, g[f"xf_transl_{N}"], f"{what} translation N={N}")
    eq(impl.xf_translation(t, N, scale=True), g[f"xf_scale_{N}"], f"{what} scale N={N}")
    comp = impl.xf_compose(xt, xm, N)
    eq(comp, g[f"xf_comp_{N}"], f"{what} compose N={N}")
    eq(impl.xf_compose(comp, xm, N, inverse=True), g[f"xf_div_{N}"], f"{what} compose-inverse N={N}")
    for mode in XF_MODES:
        eq(impl.xf_apply(comp, p, N, mode), g[f"xf_{mode}_{N}"], f"{what} {mode} N={N}")
        eq(impl.xf_apply(comp[7:8], p, N, mode), g[f"xf_{mode}_bcast_{N}"], f"{what} {mode} broadcast N={N}")


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_matches_golden(port, tag):
    g = golden(tag)
    with np.errstate(all="ignore"):
        check_mul_golden(port, g, "port")
        for N in XF_DIMS:
            check_xf_golden(port, g, N, "port")
    # what the fixtures pin: a singular matrix leaves `inv` as the identity (inv()'s result is ignored,
    # AffineTransform.h:873) and a product that starts from `sum = 0` turns -0 into +0
    K = 4
    for row in (0, 1):
        assert np.array_equal(g["xf_from_rm_3"][row, K * K:].reshape(K, K), np.eye(K))
    assert not np.signbit(g["mul_d_3_3_3"][1]).all()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_port_equals_reference(port, ref, dt):
    rng = np.random.default_rng(5)
    B = 300
    with np.errstate(all="ignore"):
        for (r, k, c) in MUL_SHAPES + [(11, 11, 11), (13, 13, 13)]:
            a = rng.standard_normal((B, r * k)).astype(dt)
            b = rng.standard_normal((B, k * c)).astype(dt)
            d0 = rng.standard_normal((B, r * c)).astype(dt)
            for lay in range(8):
                arm, brm, drm = bool(lay & 4), bool(lay & 2), bool(lay & 1)
                for acc in (False, True):
                    try:
                        want = ref.mul(a, b, r, k, c, d0, acc, arm, brm, drm)
                    except NotImplementedError:
                        continue
                    eq(port.mul(a, b, r, k, c, d0, acc, arm, brm, drm), want, f"mul {r, k, c} layouts {lay} acc {acc}")
        for n in range(2, 9):
            a = rng.standard_normal((B, n * n)).astype(dt)
            v = rng.standard_normal((B, n)).astype(dt)
            for vf in (False, True):
                for rm in (True, False):
                    eq(port.mul_vec(a, v, n, vf, rm), ref.mul_vec(a, v, n, vf, rm), f"mul_vec n={n} vec_first={vf} rm={rm}")
        for N in XF_DIMS:
            m = rng.standard_normal((B, N * N)).astype(dt)
            m[5] = 0
            m[6, :N] = 0
            t = rng.standard_normal((B, N)).astype(dt)
            p = rng.standard_normal((B, N)).astype(dt)
            for rm in (True, False):
                eq(port.xf_from_matrix(m, N, rm)[0], ref.xf_from_matrix(m, N, rm)[0], f"transformation N={N} rm={rm}")
            for sc in (False, True):
                eq(port.xf_translation(t, N, sc), ref.xf_translation(t, N, sc), f"translation/scale N={N}")
            xm, xt = port.xf_from_matrix(m, N)[0], port.xf_translation(t, N)
            comp = port.xf_compose(xt, xm, N)
            for inv in (False, True):
                eq(port.xf_compose(comp, xm, N, inv), ref.xf_compose(comp, xm, N, inv), f"compose N={N} inverse={inv}")
            for mode in XF_MODES:
                for xf in (comp, comp[3:4]):
                    eq(port.xf_apply(xf, p, N, mode), ref.xf_apply(xf, p, N, mode), f"{mode} N={N} count={len(xf)}")


@pytest.mark.parametrize("dt,tol", [(np.float32, 2e-3), (np.float64, 1e-9)])
def test_properties(port, dt, tol):
    """M * M^-1 = I (regression/matrix_tests.cpp:196-250); (xf1 * xf2)(p) = xf1(xf2(p)) and
    xf^-1(xf(p)) = p (regression/similarity.cpp:86-132)."""
    rng = np.random.default_rng(11)
    B = 200
    for n in (2, 3, 4, 6, 9):
        A = (rng.standard_normal((B, n * n)) + 3 * np.eye(n).ravel()).astype(dt)
        Ainv, ok = port.inv(A, n)
        assert ok.all()
        P = port.mul(A, Ainv, n, n, n).reshape(B, n, n)
        assert np.abs(P - np.eye(n)).max() < tol * 50
    for N in XF_DIMS:
        m1 = (rng.standard_normal((B, N * N)) + 3 * np.eye(N).ravel()).astype(dt)
        m2 = (rng.standard_normal((B, N * N)) + 3 * np.eye(N).ravel()).astype(dt)
        t = rng.standard_normal((B, N)).astype(dt)
        p = rng.standard_normal((B, N)).astype(dt)
        x1 = port.xf_compose(port.xf_translation(t, N), port.xf_from_matrix(m1, N)[0], N)
        x2 = port.xf_from_matrix(m2, N)[0]
        both = port.xf_compose(x1, x2, N)
        a = port.xf_apply(both, p, N, "apply")
        b = port.xf_apply(x1, port.xf_apply(x2, p, N, "apply"), N, "apply")
        assert np.abs(a - b).max() < tol * 100
        back = port.xf_apply(both, a, N, "apply_inverse")
        assert np.abs(back - p).max() < tol * 1000
        # x2 * (x1 / x2) is x1 again (x1 / x2 = "x1, then the inverse of x2", AffineTransform.h:150-152)
        q = port.xf_compose(x2, port.xf_compose(x1, x2, N, inverse=True), N)
        assert np.abs(port.xf_apply(q, p, N, "apply") - port.xf_apply(x1, p, N, "apply")).max() < tol * 1000
        # normals stay perpendicular to transformed tangents: n . d = (M^-T n) . (M d)
        d = rng.standard_normal((B, N)).astype(dt)
        nn = port.xf_apply(x1, p, N, "apply_normal")
        dd = port.xf_apply(x1, d, N, "apply_direction")
        assert np.abs((nn * dd).sum(1) - (p * d).sum(1)).max() < tol * 1000


# ====================================================================================== GPU
def _gm():
    import torch
    import geomc_b200 as gm
    return torch, gm


def _dev(torch, gm, a, soa):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return gm.BatchSoA.from_aos(t) if soa else t


def _host(x):
    return (x.to_aos() if hasattr(x, "to_aos") else x).cpu().numpy()


class GpuImpl:
    """The CUDA path behind the oracle's numpy interface, so the golden checks run on it unchanged."""

    XF_MODES = XF_MODES

    def __init__(self, soa):
        self.torch, self.gm = _gm()
        self.soa = soa

    def _d(self, a):
        return _dev(self.torch, self.gm, a, self.soa)

    def mul(self, a, b, r, k, c, d_init=None, acc=False, a_rm=True, b_rm=True, d_rm=True):
        d = None if d_init is None else self._d(d_init)
        return _host(self.gm.mul(self._d(a), self._d(b), r, k, c, out=d, acc=acc, a_row_major=a_rm, b_row_major=b_rm,
                                 d_row_major=d_rm))

    def xf_from_matrix(self, m, N, row_major=True):
        xf, ok = self.gm.xf_from_matrix(self._d(m), N, row_major)
        return _host(xf), ok.cpu().numpy()

    def xf_translation(self, t, N, scale=False):
        return _host(self.gm.xf_scale(self._d(t), N) if scale else self.gm.xf_translation(self._d(t), N))

    def xf_compose(self, x1, x2, N, inverse=False):
        return _host(self.gm.xf_compose(self._d(x1), self._d(x2), N, inverse))

    def xf_apply(self, xf, p, N, mode="apply"):
        bcast = xf.shape[0] == 1 and p.shape[0] != 1
        dxf = self.torch.from_numpy(np.ascontiguousarray(xf)).cuda() if bcast else self._d(xf)
        return _host(self.gm.xf_apply(dxf, self._d(p), N, mode))


@pytest.mark.gpu
@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_gpu_matches_golden(tag, soa):
    g = golden(tag)
    impl = GpuImpl(soa)
    check_mul_golden(impl, g, f"gpu soa={soa}")
    for N in XF_DIMS:
        check_xf_golden(impl, g, N, f"gpu soa={soa}")


@pytest.mark.gpu
@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_mul_equals_oracle(dt, soa):
    """Random batches with ragged tails; every layout combination; aliasing (d = a, d = b) for the register shapes."""
    orc = po.best() if po.ref_available() else po.port()
    port = po.port()
    impl = GpuImpl(soa)
    torch, gm = _gm()
    rng = np.random.default_rng(3)
    for (r, k, c) in [(2, 2, 2), (3, 3, 3), (4, 4, 4), (4, 4, 1), (1, 3, 3), (3, 4, 2), (2, 2, 4), (5, 5, 5), (6, 6, 1), (7, 3, 5), (9, 9, 9),
                      (16, 16, 16), (1, 16, 16), (12, 7, 1)]:
        B = 1000 + 37 * r + c
        a = rng.standard_normal((B, r * k)).astype(dt)
        b = rng.standard_normal((B, k * c)).astype(dt)
        d0 = rng.standard_normal((B, r * c)).astype(dt)
        for lay in range(8):
            arm, brm, drm = bool(lay & 4), bool(lay & 2), bool(lay & 1)
            for acc in (False, True):
                want = port.mul(a, b, r, k, c, d0, acc, arm, brm, drm)
                eq(impl.mul(a, b, r, k, c, d0, acc, arm, brm, drm), want, f"mul {r, k, c} layouts {lay} acc={acc} soa={soa}")
        try:
            eq(impl.mul(a, b, r, k, c), orc.mul(a, b, r, k, c), f"mul vs reference {r, k, c}")
        except NotImplementedError:
            pass
    for n in (2, 3, 4):                                   # destination aliases an operand (Matrix.h:409-417)
        B = 777
        a = rng.standard_normal((B, n * n)).astype(dt)
        b = rng.standard_normal((B, n * n)).astype(dt)
        want = port.mul(a, b, n, n, n)
        da, db = impl._d(a), impl._d(b)
        eq(_host(gm.mul(da, db, n, n, n, out=da)), want, f"mul into a, n={n}")
        da = impl._d(a)
        eq(_host(gm.mul(da, db, n, n, n, out=db)), want, f"mul into b, n={n}")


@pytest.mark.gpu
@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_xf_equals_oracle(dt, soa):
    orc = po.best()
    impl = GpuImpl(soa)
    rng = np.random.default_rng(4)
    with np.errstate(all="ignore"):
        for N in XF_DIMS:
            B = 2000 + 129 * N
            m = rng.standard_normal((B, N * N)).astype(dt)
            m[5] = 0
            m[6, :N] = 0
            m[::97] = 0
            t = rng.standard_normal((B, N)).astype(dt)
            p = rng.standard_normal((B, N)).astype(dt)
            for rm in (True, False):
                got, ok = impl.xf_from_matrix(m, N, rm)
                eq(got, orc.xf_from_matrix(m, N, rm)[0], f"transformation N={N} rm={rm}")
                assert np.array_equal(ok, po.port().xf_from_matrix(m, N, rm)[1])
            xm = orc.xf_from_matrix(m, N)[0]
            for sc in (False, True):
                eq(impl.xf_translation(t, N, sc), orc.xf_translation(t, N, sc), f"translation/scale N={N}")
            xt = orc.xf_translation(t, N)
            comp = orc.xf_compose(xt, xm, N)
            eq(impl.xf_compose(xt, xm, N), comp, f"compose N={N}")
            eq(impl.xf_compose(comp, xm, N, True), orc.xf_compose(comp, xm, N, True), f"compose-inverse N={N}")
            for mode in XF_MODES:
                eq(impl.xf_apply(comp, p, N, mode), orc.xf_apply(comp, p, N, mode), f"{mode} N={N}")
                eq(impl.xf_apply(comp[11:12], p, N, mode), orc.xf_apply(comp[11:12], p, N, mode), f"{mode} broadcast N={N}")


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_empty_and_tiny_batches(dt):
    """B = 0 launches nothing and returns empty results; B = 1 and a ragged B = 3 match the oracle in both layouts."""
    torch, gm = _gm()
    tdt = torch.float32 if dt == np.float32 else torch.float64
    port = po.port()
    N = 3
    E = 2 * (N + 1) ** 2
    for soa in (False, True):
        empty = (lambda e: gm.BatchSoA(0, e, tdt)) if soa else (lambda e: torch.empty((0, e), dtype=tdt, device="cuda"))
        assert _host(gm.mul(empty(16), empty(16), 4, 4, 4)).shape == (0, 16)
        assert _host(gm.mul(empty(81), empty(81), 9, 9, 9)).shape == (0, 81)
        xf, ok = gm.xf_from_matrix(empty(N * N), N)
        assert _host(xf).shape == (0, E) and ok.shape == (0,)
        assert _host(gm.xf_compose(empty(E), empty(E), N)).shape == (0, E)
        assert _host(gm.xf_apply(empty(E), empty(N), N)).shape == (0, N)
        impl = GpuImpl(soa)
        rng = np.random.default_rng(8)
        for B in (1, 3):
            m = rng.standard_normal((B, N * N)).astype(dt)
            p = rng.standard_normal((B, N)).astype(dt)
            xm = port.xf_from_matrix(m, N)[0]
            eq(impl.xf_from_matrix(m, N)[0], xm, f"B={B} transformation")
            eq(impl.xf_apply(xm, p, N, "apply_inverse"), port.xf_apply(xm, p, N, "apply_inverse"), f"B={B} apply_inverse")
            a = rng.standard_normal((B, 35)).astype(dt)
            b = rng.standard_normal((B, 15)).astype(dt)
            eq(impl.mul(a, b, 7, 5, 3), port.mul(a, b, 7, 5, 3), f"B={B} mul 7x5x3")


@pytest.mark.gpu
def test_gpu_inverse_round_trip_full_size():
    """Size-independent property at 2^22 systems: inv on the device, then M * M^-1 on the device, is the identity."""
    torch, gm = _gm()
    B, n = 1 << 22, 4
    gen = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((B, n * n), device="cuda", generator=gen) + 3 * torch.eye(n, device="cuda").reshape(1, -1)
    Ainv, ok = gm.inv(A, n)
    P = gm.mul(A, Ainv, n, n, n)
    err = (P - torch.eye(n, device="cuda").reshape(1, -1)).abs().amax(dim=1)
    good = ok.bool()
    assert good.float().mean() > 0.999 and float(err[good].median()) < 1e-5 and float(err[good].quantile(0.999)) < 5e-2
