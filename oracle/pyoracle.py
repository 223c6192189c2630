"""numpy/ctypes front end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``port()``  -> oracle/liboracle.so      (oracle.c, the plain-C restatement; always buildable)
* ``ref()``   -> oracle/_ref/libgeomc_ref.so  (the reference's own headers compiled by oracle/Makefile;
                 present wherever it was built from /root/reference and shipped with the snapshot)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``geomc_b200`` never does.

All batches are AoS: ``a`` has shape (B, n*n) (flat matrix per system, interpreted row- or
column-major by the ``row_major`` flag, MatrixGlue.h:28-58), ``x`` has shape (B, n*m).
Functions never modify their arguments; they return fresh arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build(verbose: bool = False) -> None:
    """(Re)build liboracle.so and, when /root/reference exists, _ref/libgeomc_ref.so."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


class CpuLinalg:
    """One of the two checker libraries, behind one numpy interface."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.path, self.prefix, self.kind = path, prefix, kind
        self.lib = C.CDLL(path)

    # ------------------------------------------------------------------ helpers
    def _fn(self, name, dt):
        f = getattr(self.lib, f"{self.prefix}_{name}_{_SUF[np.dtype(dt)]}")
        f.restype = C.c_int
        return f

    @staticmethod
    def _prep(a, cols=None):
        a = np.ascontiguousarray(a).copy()
        if a.dtype not in _SUF:
            raise TypeError(a.dtype)
        if a.ndim != 2:
            raise ValueError("expected (B, elems)")
        return a

    def max_threads(self) -> int:
        if self.kind != "reference":
            return 1
        self.lib.ref_max_threads.restype = C.c_int
        return int(self.lib.ref_max_threads())

    def set_threads(self, n: int) -> None:
        if self.kind == "reference":
            self.lib.ref_set_threads(C.c_int(n))

    def build_info(self) -> str:
        f = self.lib.ref_build_info if self.kind == "reference" else self.lib.orc_build_info
        f.restype = C.c_char_p
        return f().decode()

    # ------------------------------------------------------------------ factorizations
    def _decomp(self, which, m, rows, cols, row_major, n_reorder):
        m = self._prep(m)
        B = m.shape[0]
        assert m.shape[1] == rows * cols
        reorder = np.zeros((B, n_reorder), np.int64)
        parity = np.zeros(B, np.uint8)
        degen = np.zeros(B, np.int64)
        rc = self._fn(which, m.dtype)(C.c_int(rows), C.c_int(cols), C.c_int64(B), _ptr(m), _ptr(reorder),
                                      _ptr(parity), _ptr(degen), C.c_int(int(row_major)))
        assert rc == 0
        return m, reorder, parity, degen

    def decomp_plu(self, m, rows, cols=None, row_major=True):
        """geom::decomp_plu (LUDecomp.h:188) -> (lu, reorder[B,rows], parity[B], degenerate_ct[B])."""
        cols = rows if cols is None else cols
        return self._decomp("decomp_plu", m, rows, cols, row_major, rows)

    def decomp_lup(self, m, rows, cols=None, row_major=True):
        """geom::decomp_lup (LUDecomp.h:102) -> (lu, reorder[B,cols], parity[B], degenerate_ct[B])."""
        cols = rows if cols is None else cols
        return self._decomp("decomp_lup", m, rows, cols, row_major, cols)

    def backsolve_lu(self, lu, x, n, m=1, skip=0, row_major=True):
        """geom::backsolve_lu (LUDecomp.h:332) -> x."""
        lu = self._prep(lu)
        x = self._prep(x)
        B = lu.shape[0]
        rc = self._fn("backsolve_lu", lu.dtype)(C.c_int(n), C.c_int(m), C.c_int64(B), _ptr(lu), _ptr(x),
                                                C.c_int(skip), C.c_int(int(row_major)))
        assert rc == 0
        return x

    def backsolve_plu(self, plu, p, b, n, m=1, skip=0, row_major=True):
        """geom::backsolve_plu (LUDecomp.h:271) -> x."""
        plu = self._prep(plu)
        b = self._prep(b)
        p = np.ascontiguousarray(p, dtype=np.int64)
        x = np.zeros_like(b)
        B = plu.shape[0]
        rc = self._fn("backsolve_plu", plu.dtype)(C.c_int(n), C.c_int(m), C.c_int64(B), _ptr(plu), _ptr(p), _ptr(x),
                                                  _ptr(b), C.c_int(skip), C.c_int(int(row_major)))
        assert rc == 0
        return x

    def apply_perm(self, p, a, n, m, row_major=True):
        """geom::detail::destructive_apply_permutation (LUDecomp.h:35) -> (P a, clobbered p)."""
        a = self._prep(a)
        p = np.ascontiguousarray(p, dtype=np.int64).copy()
        B = a.shape[0]
        name = "apply_perm"
        rc = self._fn(name, a.dtype)(C.c_int(n), C.c_int(m), C.c_int64(B), _ptr(p), _ptr(a), C.c_int(int(row_major)))
        assert rc == 0
        return a, p

    def linear_solve(self, a, x, n, m=1, skip=0, row_major=True):
        """geom::linear_solve pointer form (LUDecomp.h:388) -> (a_after (LU), x_after, ok[B])."""
        a = self._prep(a)
        x = self._prep(x)
        B = a.shape[0]
        ok = np.zeros(B, np.uint8)
        rc = self._fn("linear_solve", a.dtype)(C.c_int(n), C.c_int(m), C.c_int64(B), _ptr(a), _ptr(x), C.c_int(skip),
                                               _ptr(ok), C.c_int(int(row_major)))
        assert rc == 0
        return a, x, ok

    def linear_solve_vec(self, bases, x, n, m=1, skip=0):
        """geom::linear_solve Vec form (LUDecomp.h:417) -> (bases_after, x_after, ok[B])."""
        bases = self._prep(bases)
        x = self._prep(x)
        B = bases.shape[0]
        ok = np.zeros(B, np.uint8)
        rc = self._fn("linear_solve_vec", bases.dtype)(C.c_int(n), C.c_int(m), C.c_int64(B), _ptr(bases), _ptr(x),
                                                       C.c_int(skip), _ptr(ok))
        assert rc == 0
        return bases, x, ok

    def cholesky(self, a, n, row_major=True):
        """geom::cholesky<T,RowMajor>(T*, n) (Cholesky.h:37) -> (L (upper zeroed), ok[B])."""
        a = self._prep(a)
        B = a.shape[0]
        ok = np.zeros(B, np.uint8)
        rc = self._fn("cholesky", a.dtype)(C.c_int(n), C.c_int64(B), _ptr(a), _ptr(ok), C.c_int(int(row_major)))
        assert rc == 0
        return a, ok

    def cholesky_mtx(self, a, n):
        """geom::cholesky(SimpleMatrix*) (Cholesky.h:83); reference library only."""
        a = self._prep(a)
        B = a.shape[0]
        ok = np.zeros(B, np.uint8)
        rc = self._fn("cholesky_mtx", a.dtype)(C.c_int(n), C.c_int64(B), _ptr(a), _ptr(ok))
        assert rc == 0
        return a, ok

    def cholesky_solve(self, l, x, n, m=1, row_major=True):
        """Framework-defined L L^T x = b solve (no reference counterpart; port only)."""
        if self.kind == "reference":
            raise NotImplementedError("the reference has no Cholesky solve")
        l = self._prep(l)
        x = self._prep(x)
        B = l.shape[0]
        rc = self._fn("cholesky_solve", l.dtype)(C.c_int(n), C.c_int(m), C.c_int64(B), _ptr(l), _ptr(x),
                                                 C.c_int(int(row_major)))
        assert rc == 0
        return x

    def inv(self, src, n, src_row_major=True, dst_row_major=True, dst_init=None):
        """geom::inv(Md*, const Mx&) (Matrix.h:717) -> (dst, ok[B]); dst keeps dst_init where ok == 0."""
        src = self._prep(src)
        B = src.shape[0]
        dst = np.zeros_like(src) if dst_init is None else self._prep(dst_init)
        ok = np.zeros(B, np.uint8)
        rc = self._fn("inv", src.dtype)(C.c_int(n), C.c_int64(B), _ptr(src), _ptr(dst), _ptr(ok),
                                        C.c_int(int(src_row_major)), C.c_int(int(dst_row_major)))
        assert rc == 0
        return dst, ok

    def inv_raw(self, src, n, dst_init=None):
        """inv2x2 / inv3x3 / inv4x4 / invNxN(T*,T*,n) on flat arrays (MatrixInv.h:49/75/117/186)."""
        src = self._prep(src)
        B = src.shape[0]
        dst = np.zeros_like(src) if dst_init is None else self._prep(dst_init)
        ok = np.zeros(B, np.uint8)
        rc = self._fn("inv_raw", src.dtype)(C.c_int(n), C.c_int64(B), _ptr(src), _ptr(dst), _ptr(ok))
        assert rc == 0
        return dst, ok

    def orthogonal(self, v, n):
        """geom::orthogonal<T,N> (Orthogonal.h:73): v = (B, (n-1)*n) -> (B, n)."""
        v = self._prep(v)
        B = v.shape[0]
        assert v.shape[1] == (n - 1) * n
        out = np.zeros((B, n), v.dtype)
        rc = self._fn("orthogonal", v.dtype)(C.c_int(n), C.c_int64(B), _ptr(v), _ptr(out))
        assert rc == 0
        return out

    def nullspace(self, bases, N, nb):
        """geom::nullspace<T,N>(bases, nb, null_basis) (Orthogonal.h:136): (B, nb*N) -> ((B, (N-nb)*N), ok[B])."""
        bases = self._prep(bases)
        B = bases.shape[0]
        assert bases.shape[1] == nb * N and nb < N
        null = np.zeros((B, (N - nb) * N), bases.dtype)
        ok = np.zeros(B, np.uint8)
        rc = self._fn("nullspace", bases.dtype)(C.c_int(N), C.c_int(nb), C.c_int64(B), _ptr(bases), _ptr(null), _ptr(ok))
        assert rc == 0
        return null, ok

    # ---- simplex queries (shape/Simplex.h) ------------------------------------------
    def simplex_contains(self, verts, p, N, nverts=None):
        """Simplex<T,N>::contains (Simplex.h:450): verts (B, (N+1)*N), p (B, N) -> inside[B]."""
        verts, p = self._prep(verts), self._prep(p)
        B = verts.shape[0]
        assert verts.shape[1] == (N + 1) * N and p.shape == (B, N)
        nv = None if nverts is None else np.ascontiguousarray(nverts, np.uint8)
        inside = np.zeros(B, np.uint8)
        rc = self._fn("simplex_contains", verts.dtype)(C.c_int(N), C.c_int64(B), _ptr(verts), _ptr(nv), _ptr(p), _ptr(inside))
        assert rc == 0
        return inside

    def trace_simplex(self, verts, origin, direction, N, want_uv=True):
        """trace_simplex<T,N> (Simplex.h:647): verts (B, N*N), rays (B, N) -> (hit[B], s[B], uv (B, N-1)).
        s and uv are only written where hit; the reference library's s is x[N] (out of bounds, :669) and is
        not comparable -- the port returns x[N-1]."""
        verts, origin, direction = self._prep(verts), self._prep(origin), self._prep(direction)
        B = verts.shape[0]
        assert verts.shape[1] == N * N and origin.shape == (B, N) and direction.shape == (B, N)
        hit = np.zeros(B, np.uint8)
        s = np.zeros(B, verts.dtype)
        uv = np.zeros((B, N - 1), verts.dtype)
        rc = self._fn("trace_simplex", verts.dtype)(C.c_int(N), C.c_int64(B), _ptr(verts), _ptr(origin), _ptr(direction),
                                                    C.c_int(1 if want_uv else 0), _ptr(hit), _ptr(s), _ptr(uv))
        assert rc == 0
        return hit, s, uv

    def simplex_intersect(self, verts, origin, direction, N, nverts=None):
        """Simplex<T,N>::intersect(Ray) (Simplex.h:251): verts (B, (N+1)*N) -> Rect<T,1> as (B, 2) = lo, hi."""
        verts, origin, direction = self._prep(verts), self._prep(origin), self._prep(direction)
        B = verts.shape[0]
        assert verts.shape[1] == (N + 1) * N and origin.shape == (B, N) and direction.shape == (B, N)
        nv = None if nverts is None else np.ascontiguousarray(nverts, np.uint8)
        lohi = np.zeros((B, 2), verts.dtype)
        rc = self._fn("simplex_intersect", verts.dtype)(C.c_int(N), C.c_int64(B), _ptr(verts), _ptr(nv), _ptr(origin),
                                                        _ptr(direction), _ptr(lohi))
        assert rc == 0
        return lohi

    # ---- mul / mul_acc (Matrix.h:386 / :433, MatrixMult.h:54-80) ----------------------
    def mul(self, a, b, r, k, c, d_init=None, acc=False, a_rm=True, b_rm=True, d_rm=True):
        """geom::mul(&d, a, b) / mul_acc: a (B, r*k), b (B, k*c) -> d (B, r*c).  The reference library only
        instantiates a table of shapes (oracle/ref_driver_xform.cpp); others raise NotImplementedError."""
        a, b = self._prep(a), self._prep(b)
        B = a.shape[0]
        assert a.shape[1] == r * k and b.shape == (B, k * c)
        d = np.zeros((B, r * c), a.dtype) if d_init is None else self._prep(d_init)
        rc = self._fn("mul", a.dtype)(C.c_int(r), C.c_int(k), C.c_int(c), C.c_int64(B), _ptr(a), _ptr(b), _ptr(d),
                                      C.c_int(int(acc)), C.c_int(int(a_rm)), C.c_int(int(b_rm)), C.c_int(int(d_rm)))
        if rc != 0:
            raise NotImplementedError(f"mul {r}x{k}x{c} layouts {a_rm, b_rm, d_rm} not instantiated in {self.kind}")
        return d

    def mul_vec(self, a, v, n, vec_first=False, row_major=True):
        """SimpleMatrix<T,n,n> * Vec<T,n> (MatrixMult.h:181) or Vec * SimpleMatrix (:239): -> (B, n)."""
        a, v = self._prep(a), self._prep(v)
        B = a.shape[0]
        assert a.shape[1] == n * n and v.shape == (B, n)
        out = np.zeros((B, n), a.dtype)
        rc = self._fn("mul_vec", a.dtype)(C.c_int(n), C.c_int(int(vec_first)), C.c_int64(B), _ptr(a), _ptr(v), _ptr(out),
                                          C.c_int(int(row_major)))
        assert rc == 0
        return out

    # ---- AffineTransform (linalg/AffineTransform.h); records = mat (N+1)^2 | inv (N+1)^2 ----
    def xf_from_matrix(self, m, N, row_major=True):
        """geom::transformation(mat) (AffineTransform.h:862): m (B, N*N) -> (B, 2*(N+1)^2).
        The port also returns inv()'s success flag, which the reference ignores."""
        m = self._prep(m)
        B = m.shape[0]
        assert m.shape[1] == N * N
        out = np.zeros((B, 2 * (N + 1) ** 2), m.dtype)
        if self.kind == "reference":
            rc = self._fn("xf_from_matrix", m.dtype)(C.c_int(N), C.c_int64(B), _ptr(m), _ptr(out), C.c_int(int(row_major)))
            assert rc == 0
            return out, None
        ok = np.zeros(B, np.uint8)
        rc = self._fn("xf_from_matrix", m.dtype)(C.c_int(N), C.c_int64(B), _ptr(m), _ptr(out), _ptr(ok), C.c_int(int(row_major)))
        assert rc == 0
        return out, ok

    def xf_translation(self, t, N, scale=False):
        """geom::translation(Vec) (AffineTransform.h:814) / geom::scale(Vec) (:838): t (B, N) -> records."""
        t = self._prep(t)
        B = t.shape[0]
        assert t.shape[1] == N
        out = np.zeros((B, 2 * (N + 1) ** 2), t.dtype)
        rc = self._fn("xf_translation", t.dtype)(C.c_int(N), C.c_int64(B), _ptr(t), _ptr(out), C.c_int(int(scale)))
        assert rc == 0
        return out

    def xf_compose(self, xf1, xf2, N, inverse=False):
        """xf1 * xf2 (AffineTransform.h:142) or xf1 / xf2 (:154)."""
        xf1, xf2 = self._prep(xf1), self._prep(xf2)
        B = xf1.shape[0]
        assert xf1.shape == xf2.shape == (B, 2 * (N + 1) ** 2)
        out = np.zeros_like(xf1)
        rc = self._fn("xf_compose", xf1.dtype)(C.c_int(N), C.c_int64(B), _ptr(xf1), _ptr(xf2), _ptr(out), C.c_int(int(inverse)))
        assert rc == 0
        return out

    XF_MODES = ("apply", "apply_inverse", "apply_direction", "apply_inverse_direction", "apply_normal",
                "apply_inverse_normal")

    def xf_apply(self, xf, p, N, mode="apply"):
        """AffineTransform::apply & co. (AffineTransform.h:224-293): xf (B or 1, 2*(N+1)^2), p (B, N) -> (B, N)."""
        xf, p = self._prep(xf), self._prep(p)
        B = p.shape[0]
        assert xf.shape[0] in (1, B) and xf.shape[1] == 2 * (N + 1) ** 2 and p.shape[1] == N
        out = np.zeros_like(p)
        rc = self._fn("xf_apply", p.dtype)(C.c_int(N), C.c_int(self.XF_MODES.index(mode)), C.c_int64(B), _ptr(xf),
                                           C.c_int64(xf.shape[0]), _ptr(p), _ptr(out))
        assert rc == 0
        return out

    # in-place, zero-copy variants used only by bench.py's CPU baseline timing loop
    def linear_solve_inplace(self, a, x, ok, n, m=1, skip=0, row_major=True):
        rc = self._fn("linear_solve", a.dtype)(C.c_int(n), C.c_int(m), C.c_int64(a.shape[0]), _ptr(a), _ptr(x),
                                               C.c_int(skip), _ptr(ok), C.c_int(int(row_major)))
        assert rc == 0


_cache: dict = {}


def port() -> CpuLinalg:
    if "port" not in _cache:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _cache["port"] = CpuLinalg(path, "orc", "port")
    return _cache["port"]


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libgeomc_ref.so"))


def ref() -> CpuLinalg:
    if "ref" not in _cache:
        path = os.path.join(_HERE, "_ref", "libgeomc_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/geomc"):
                build()
            if not os.path.exists(path):
                raise FileNotFoundError("oracle/_ref/libgeomc_ref.so not built (needs /root/reference)")
        _cache["ref"] = CpuLinalg(path, "ref", "reference")
    return _cache["ref"]


def best() -> CpuLinalg:
    """The reference itself when its compiled library is present, else the port."""
    return ref() if ref_available() else port()


class Checker:
    """What the GPU parity tests compare with: every function of best() -- the compiled reference when oracle/_ref is
    on the box -- except the one the reference does not have (cholesky_solve, defined by the port: parity unpinned)."""

    def __init__(self):
        self._best, self._port = best(), port()
        self.kind = self._best.kind

    def __getattr__(self, name):
        return getattr(self._port if name == "cholesky_solve" else self._best, name)


def checker() -> Checker:
    if "checker" not in _cache:
        _cache["checker"] = Checker()
    return _cache["checker"]


# ---------------------------------------------------------------------- shared input generators
def exact_grid(rng: np.random.Generator, shape, dtype, bits: int = 20, scale_log2: int = -17):
    """Contraction-proof inputs: `bits`-bit signed integers times 2**scale_log2 (exact in f32 and f64).

    Default: uniform on a grid of step 2^-17 in [-4, 4).  SURVEY.md section 7 item 2.
    """
    half = 1 << (bits - 1)
    v = rng.integers(-half, half, size=shape, dtype=np.int64)
    return (v.astype(np.float64) * (2.0 ** scale_log2)).astype(dtype)


def spd_exact(rng: np.random.Generator, B: int, n: int, dtype):
    """A = R R^T + I with R on a coarse exact grid, so that every product and sum is exact in f64
    (and in f32 for the coarse grid used): 8-bit integers times 2^-5."""
    R = exact_grid(rng, (B, n, n), np.float64, bits=8, scale_log2=-5)
    A = R @ np.swapaxes(R, 1, 2) + np.eye(n)[None]
    return A.reshape(B, n * n).astype(dtype)
