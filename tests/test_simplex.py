"""Row (f)2 of the scope table: batched simplex queries, the in-library consumers of linear_solve
(shape/Simplex.h: Simplex::contains :450, trace_simplex :647, Simplex::intersect(Ray) :251).
CPU: oracle.c against the reference-generated fixtures and against the compiled reference, plus the
geometric properties regression/simplex.cpp relies on.  GPU: the CUDA kernels against both, bit for bit.

trace_simplex's ray parameter: the reference stores x[N], one past the end of its Vec<T,N> (:675, undefined
behaviour); oracle.c and the CUDA path store x[N-1], the coefficient of -direction it means.  It is checked
against the geometry (origin + s * direction lies on the simplex) instead of against the reference."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, bits_equal, first_diff

DIMS = range(2, 8)


def golden(tag):
    return np.load(os.path.join(GOLDEN_DIR, f"simplex_{tag}.npz"))


def check_against_golden(impl, g, N, what):
    nv = g[f"c_nv_{N}"]
    inside = impl.simplex_contains(g[f"c_verts_{N}"], g[f"c_p_{N}"], N, nv)
    assert np.array_equal(inside, g[f"c_in_{N}"]), f"{what} contains N={N}: {first_diff(inside, g[f'c_in_{N}'])}"
    lohi = impl.simplex_intersect(g[f"c_verts_{N}"], g[f"i_o_{N}"], g[f"i_d_{N}"], N, nv)
    assert np.array_equal((lohi[:, 0] <= lohi[:, 1]).astype(np.uint8), g[f"i_hit_{N}"]), f"{what} intersect hit N={N}"
    m = nv != N
    assert bits_equal(lohi[m], g[f"i_lohi_{N}"][m]), f"{what} intersect N={N}: {first_diff(lohi[m], g[f'i_lohi_{N}'][m])}"
    for wuv in (1, 0):
        hit, s, uv = impl.trace_simplex(g[f"t_verts_{N}"], g[f"t_o_{N}"], g[f"t_d_{N}"], N, bool(wuv))
        assert np.array_equal(hit, g[f"t_hit{wuv}_{N}"]), f"{what} trace hit N={N} uv={wuv}"
        if wuv:
            assert bits_equal(uv, g[f"t_uv_{N}"]), f"{what} trace uv N={N}: {first_diff(uv, g[f't_uv_{N}'])}"


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_matches_golden(port, tag):
    g = golden(tag)
    with np.errstate(all="ignore"):
        for N in DIMS:
            check_against_golden(port, g, N, "port")
    # the fixtures hold the interesting cases: vertex and edge-midpoint queries, degenerate simplexes (never
    # contain anything, Simplex.h:466), sub-dimensional simplexes (:451)
    assert g["c_in_3"][3] == 0 and g["c_in_3"][4] == 0 and (g["c_in_3"][g["c_nv_3"] < 4] == 0).all()
    # with a null `uv` the reference solves with skip = N-1 and then tests the un-substituted rows (:664-672):
    # for N >= 5 (the PLU path) that rejects everything; kept, because it is what the reference returns
    assert g["t_hit1_5"].sum() > 0 and g["t_hit0_5"].sum() == 0 and np.array_equal(g["t_hit1_3"], g["t_hit0_3"])


def random_scene(rng, N, K, B, dt):
    verts = rng.standard_normal((B, K * N)).astype(dt)
    V = verts.reshape(B, K, N)
    w = rng.dirichlet(np.ones(K), B)
    inside = np.einsum("bk,bkd->bd", w, V.astype(np.float64)).astype(dt)
    aim = (np.arange(B) % 2 == 0)[:, None]
    p = np.where(aim, inside, rng.standard_normal((B, N))).astype(dt)
    o = (3 * rng.standard_normal((B, N))).astype(dt)
    d = np.where(aim, inside - o, rng.standard_normal((B, N))).astype(dt)
    verts[::41] = 0
    verts[5::43, N:2 * N] = verts[5::43, :N]
    return verts, p, o, d


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_port_equals_reference(port, ref, dt):
    rng = np.random.default_rng(23)
    B = 600
    with np.errstate(all="ignore"):
        for N in DIMS:
            verts, p, o, d = random_scene(rng, N, N + 1, B, dt)
            nv = np.full(B, N + 1, np.uint8)
            nv[3::7] = N
            nv[4::11] = N - 1
            assert np.array_equal(port.simplex_contains(verts, p, N, nv), ref.simplex_contains(verts, p, N, nv)), N
            a, b = port.simplex_intersect(verts, o, d, N, nv), ref.simplex_intersect(verts, o, d, N, nv)
            m = nv != N
            assert bits_equal(a[m], b[m]), (N, first_diff(a[m], b[m]))
            assert np.array_equal(a[~m, 0] <= a[~m, 1], b[~m, 0] <= b[~m, 1]), N
            tv, _, to, td = random_scene(rng, N, N, B, dt)
            for wuv in (True, False):
                ha, sa, ua = port.trace_simplex(tv, to, td, N, wuv)
                hb, sb, ub = ref.trace_simplex(tv, to, td, N, wuv)
                assert np.array_equal(ha, hb) and bits_equal(ua, ub), (N, wuv)


@pytest.mark.parametrize("N", [2, 3, 4, 5, 7])
def test_geometry(port, N):
    """What regression/simplex.cpp and shape.cpp rely on: convex combinations are inside, far points are not;
    a ray aimed at a facet point hits it at origin + s * direction with barycentric coordinates uv."""
    rng = np.random.default_rng(100 + N)
    B = 400
    verts = rng.standard_normal((B, (N + 1) * N))
    V = verts.reshape(B, N + 1, N)
    w = rng.dirichlet(2 * np.ones(N + 1), B)
    pin = np.einsum("bk,bkd->bd", w, V)
    assert port.simplex_contains(verts, pin, N).all()
    assert not port.simplex_contains(verts, pin + 100.0, N).any()
    # ray through an interior point: the parameter interval contains the parameter of that point (= 1)
    o = 5 * rng.standard_normal((B, N))
    lohi = port.simplex_intersect(verts, o, pin - o, N)
    assert (lohi[:, 0] <= 1 + 1e-9).all() and (lohi[:, 1] >= 1 - 1e-9).all()
    # facet hit
    tv = verts[:, :N * N].copy()
    wf = rng.dirichlet(2 * np.ones(N), B)
    target = np.einsum("bk,bkd->bd", wf, tv.reshape(B, N, N))
    hit, s, uv = port.trace_simplex(tv, o, target - o, N, True)
    assert hit.all()
    assert np.abs(s - 1).max() < 1e-8 and np.abs(uv - wf[:, 1:]).max() < 1e-8
    assert np.abs(o + s[:, None] * (target - o) - target).max() < 1e-7


# ------------------------------------------------------------------------------------------ GPU
def _gm():
    pytest.importorskip("torch")
    import geomc_b200 as gm
    gm.lib()
    return gm


class GpuImpl:
    """The CUDA path behind the oracle's numpy interface."""

    def __init__(self, gm, soa):
        import torch
        self.gm, self.soa, self.torch = gm, soa, torch

    def up(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return self.gm.BatchSoA.from_aos(t) if self.soa else t

    @staticmethod
    def down(x):
        return (x.to_aos() if hasattr(x, "to_aos") else x).cpu().numpy()

    def _nv(self, nv):
        return None if nv is None else self.torch.from_numpy(np.ascontiguousarray(nv, np.uint8)).cuda()

    def simplex_contains(self, verts, p, N, nverts=None):
        return self.down(self.gm.simplex_contains(self.up(verts), self.up(p), N, self._nv(nverts)))

    def simplex_intersect(self, verts, o, d, N, nverts=None):
        return self.down(self.gm.simplex_intersect(self.up(verts), self.up(o), self.up(d), N, self._nv(nverts)))

    def trace_simplex(self, verts, o, d, N, want_uv=True):
        hit, s, uv = self.gm.trace_simplex(self.up(verts), self.up(o), self.up(d), N, want_uv)
        return self.down(hit), self.down(s), (self.down(uv) if uv is not None else np.zeros((len(verts), N - 1), verts.dtype))


@pytest.mark.gpu
@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_gpu_matches_golden(tag, soa):
    gpu = GpuImpl(_gm(), soa)
    g = golden(tag)
    for N in DIMS:
        check_against_golden(gpu, g, N, f"gpu soa={soa}")


@pytest.mark.gpu
@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_random_vs_oracle(port, dt, soa):
    gpu = GpuImpl(_gm(), soa)
    rng = np.random.default_rng(77)
    B = 5000 + 37                                             # ragged: not a multiple of the tile or the block
    with np.errstate(all="ignore"):
        for N in DIMS:
            verts, p, o, d = random_scene(rng, N, N + 1, B, dt)
            nv = np.full(B, N + 1, np.uint8)
            nv[3::7] = N
            nv[4::11] = N - 1
            for nvv in (nv, None):
                a, b = gpu.simplex_contains(verts, p, N, nvv), port.simplex_contains(verts, p, N, nvv)
                assert np.array_equal(a, b), f"contains N={N}: {first_diff(a, b)}"
                a, b = gpu.simplex_intersect(verts, o, d, N, nvv), port.simplex_intersect(verts, o, d, N, nvv)
                assert bits_equal(a, b), f"intersect N={N}: {first_diff(a, b)}"
            tv, _, to, td = random_scene(rng, N, N, B, dt)
            for wuv in (True, False):
                ha, sa, ua = gpu.trace_simplex(tv, to, td, N, wuv)
                hb, sb, ub = port.trace_simplex(tv, to, td, N, wuv)
                assert np.array_equal(ha, hb), f"trace hit N={N} uv={wuv}"
                assert bits_equal(sa, sb), f"trace s N={N} uv={wuv}: {first_diff(sa, sb)}"
                assert bits_equal(ua, ub), f"trace uv N={N}: {first_diff(ua, ub)}"


@pytest.mark.gpu
def test_gpu_object_array_strides_and_empty(port):
    """An array of Simplex<float,3> objects (13 floats + index_t n, padded: 56 bytes = 14 floats) and an array of
    Ray<float,3> (origin, direction interleaved) go in as they are, through the strides of the C ABI."""
    gm = _gm()
    import torch
    from geomc_b200 import _lib
    rng = np.random.default_rng(3)
    N, B = 3, 1000
    verts, p, o, d = random_scene(rng, N, N + 1, B, np.float32)
    objs = np.zeros((B, 14), np.float32)
    objs[:, :12] = verts
    rays = np.concatenate([o, d], axis=1)
    t_objs, t_rays, t_p = (torch.from_numpy(x).cuda() for x in (objs, rays, p))
    inside = torch.empty(B, dtype=torch.uint8, device="cuda")
    _lib.call("gmb_simplex_contains", "f32", N, B, t_objs.data_ptr(), 14, None, t_p.data_ptr(), inside.data_ptr(), 0, None)
    lohi = torch.empty((B, 2), dtype=torch.float32, device="cuda")
    _lib.call("gmb_simplex_intersect", "f32", N, B, t_objs.data_ptr(), 14, None, t_rays.data_ptr(), t_rays.data_ptr() + 12,
              6, lohi.data_ptr(), 0, None)
    torch.cuda.synchronize()
    with np.errstate(all="ignore"):
        assert np.array_equal(inside.cpu().numpy(), port.simplex_contains(verts, p, N))
        assert bits_equal(lohi.cpu().numpy(), port.simplex_intersect(verts, o, d, N))
    # empty batch and bad arguments
    e = torch.empty((0, 12), dtype=torch.float32, device="cuda")
    assert gm.simplex_contains(e, torch.empty((0, 3), dtype=torch.float32, device="cuda"), 3).numel() == 0
    with pytest.raises(gm.GmbError):
        _lib.call("gmb_simplex_contains", "f32", 8, B, t_objs.data_ptr(), 0, None, t_p.data_ptr(), inside.data_ptr(), 0, None)
    with pytest.raises(gm.GmbError):
        _lib.call("gmb_simplex_contains", "f32", 3, B, t_objs.data_ptr(), 5, None, t_p.data_ptr(), inside.data_ptr(), 0, None)
