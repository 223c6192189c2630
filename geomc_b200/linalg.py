"""Batched geomc/linalg on B200: Python mirror of the reference's function names, one call per BATCH.

Every function takes either AoS CUDA tensors of shape (B, E) -- the reference's own per-object
storage, E = n*n for matrices and n*m for right-hand sides, flat in the order the ``row_major``
flag says (MatrixGlue.h:28-58) -- or :class:`BatchSoA` containers (the SoA-interleaved batch
layout), and goes straight through the C ABI (include/geomc_b200.h) to the CUDA kernels.
torch is used for device memory and streams only.  No CPU path exists.

Reference function                                   -> here
    decomp_plu            LUDecomp.h:188             -> decomp_plu
    decomp_lup            LUDecomp.h:102             -> decomp_lup
    backsolve_lu          LUDecomp.h:332             -> backsolve_lu
    backsolve_plu         LUDecomp.h:271             -> backsolve_plu
    destructive_apply_permutation  LUDecomp.h:35     -> apply_permutation
    linear_solve(T*,...)  LUDecomp.h:388             -> linear_solve
    linear_solve(Vec[],…) LUDecomp.h:417             -> linear_solve_vec
    cholesky              Cholesky.h:37              -> cholesky
    (none)                                           -> cholesky_solve   (new: fused factor + solve)
    inv                   Matrix.h:717               -> inv
    mul / mul_acc         Matrix.h:386 / :433        -> mul
    transformation / translation / scale   AffineTransform.h:862 / :814 / :838   -> xf_from_matrix / xf_translation / xf_scale
    AffineTransform `*`, `/`               AffineTransform.h:142 / :154          -> xf_compose
    AffineTransform::apply & co.           AffineTransform.h:224-293             -> xf_apply
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import torch

from . import _lib
from ._lib import LAYOUT_AOS, LAYOUT_SOA

_SUF = {torch.float32: "f32", torch.float64: "f64"}


class _Stream(int):
    """The raw cudaStream_t of torch's current stream on `device`, remembering the device it belongs to."""
    device = None


def _stream_ptr(device) -> "_Stream":
    st = _Stream(torch.cuda.current_stream(device).cuda_stream)
    st.device = device
    return st


def _call(name: str, suf: str, *args) -> None:
    """C ABI call whose last argument is a _Stream: the stream's device is made current for the duration of the call
    (kernel launches and cudaFuncSetAttribute go to the CURRENT device, which need not be the tensors' device)."""
    st = args[-1]
    with torch.cuda.device(st.device):
        _lib.call(name, suf, *args[:-1], int(st))


def _same(a: "Batch", *others, what: str = "operands") -> None:
    """dtype, device and layout of every operand must equal those of `a` (the C ABI takes raw pointers)."""
    la, _, _, dta, deva, _ = _info(a)
    for o in others:
        if o is None:
            continue
        lo, _, _, dto, devo, _ = _info(o)
        if (lo, dto, devo) != (la, dta, deva):
            raise ValueError(f"{what}: layout / dtype / device differ ({lo}, {dto}, {devo}) vs ({la}, {dta}, {deva})")


def _shape(x: "Batch", B: int, E: int, what: str) -> None:
    _, Bx, Ex, _, _, _ = _info(x)
    if (Bx, Ex) != (B, E):
        raise ValueError(f"{what}: expected ({B}, {E}), got ({Bx}, {Ex})")


def _flags(t: torch.Tensor, B: int, device, what: str) -> None:
    if t.dtype != torch.uint8 or t.device != device or t.numel() < B or not t.is_contiguous():
        raise ValueError(f"{what}: a contiguous uint8 tensor of {B} entries on {device} is required")


class BatchSoA:
    """SoA-interleaved batch container (replaces an array of SimpleMatrix / Vec objects).

    B systems of E elements each; tiles of W = 512/sizeof(T) systems; element e of system s lives
    at ``data[(s // W * E + e) * W + s % W]``.  Storage is whole tiles.
    """

    def __init__(self, B: int, E: int, dtype=torch.float32, device="cuda", data: Optional[torch.Tensor] = None):
        self.B, self.E, self.dtype = int(B), int(E), dtype
        self.W = _lib.lib().gmb_soa_tile_width(torch.empty(0, dtype=dtype).element_size())
        n = _lib.lib().gmb_soa_elems(self.B, self.E, torch.empty(0, dtype=dtype).element_size())
        if data is None:
            data = torch.zeros(n, dtype=dtype, device=device)
        assert data.numel() == n and data.dtype == dtype and data.is_cuda and data.is_contiguous()
        self.data = data

    @property
    def device(self):
        return self.data.device

    @property
    def tiles(self) -> int:
        return (self.B + self.W - 1) // self.W

    @classmethod
    def from_aos(cls, aos: torch.Tensor) -> "BatchSoA":
        """Coalesced AoS -> SoA ingest (gmb_aos_to_soa_*)."""
        assert aos.is_cuda and aos.dim() == 2 and aos.is_contiguous()
        B, E = aos.shape
        out = cls(B, E, aos.dtype, aos.device)
        _call("gmb_aos_to_soa", _SUF[aos.dtype], E, B, aos.data_ptr(), out.data.data_ptr(), _stream_ptr(aos.device))
        return out

    def to_aos(self) -> torch.Tensor:
        out = torch.empty((self.B, self.E), dtype=self.dtype, device=self.device)
        _call("gmb_soa_to_aos", _SUF[self.dtype], self.E, self.B, self.data.data_ptr(), out.data_ptr(),
                  _stream_ptr(self.device))
        return out

    def clone(self) -> "BatchSoA":
        return BatchSoA(self.B, self.E, self.dtype, self.device, self.data.clone())

    def empty_like(self, E: Optional[int] = None) -> "BatchSoA":
        return BatchSoA(self.B, self.E if E is None else E, self.dtype, self.device)


Batch = Union[torch.Tensor, BatchSoA]


def _info(a: Batch):
    if isinstance(a, BatchSoA):
        return LAYOUT_SOA, a.B, a.E, a.dtype, a.device, a.data.data_ptr()
    assert a.is_cuda and a.dim() == 2 and a.is_contiguous(), "AoS batches are contiguous CUDA tensors of shape (B, E)"
    return LAYOUT_AOS, a.shape[0], a.shape[1], a.dtype, a.device, a.data_ptr()


def _clone(a: Batch) -> Batch:
    return a.clone()


def _new_like(a: Batch, E: Optional[int] = None) -> Batch:
    if isinstance(a, BatchSoA):
        return a.empty_like(E)
    return torch.empty((a.shape[0], a.shape[1] if E is None else E), dtype=a.dtype, device=a.device)


def _ptr(a: Optional[Batch]) -> Optional[int]:
    if a is None:
        return None
    return a.data.data_ptr() if isinstance(a, BatchSoA) else a.data_ptr()


def _u8(B: int, k: int, device) -> torch.Tensor:
    # every entry is written by the kernels, so no fill kernel is launched
    return torch.empty((B, k) if k > 1 else (B,), dtype=torch.uint8, device=device)


def decomp_plu(m: Batch, rows: int, cols: Optional[int] = None, row_major: bool = True, inplace: bool = False,
               out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None):
    """LU = P M with row pivoting.  -> (lu, reorder uint8[B,rows], parity uint8[B], degenerate_ct uint8[B]).

    `out` = preallocated (reorder, parity, degenerate_ct) uint8 tensors to write into."""
    cols = rows if cols is None else cols
    layout, B, E, dt, dev, _ = _info(m)
    assert E == rows * cols
    lu = m if inplace else _clone(m)
    if out is not None:
        perm, parity, degen = out
    else:
        perm, parity, degen = _u8(B, rows, dev).reshape(B, rows), _u8(B, 1, dev), _u8(B, 1, dev)
    _call("gmb_plu_decomp", _SUF[dt], rows, cols, B, _ptr(lu), perm.data_ptr(), parity.data_ptr(), degen.data_ptr(),
              layout, int(row_major), _stream_ptr(dev))
    return lu, perm, parity, degen


def decomp_lup(m: Batch, rows: int, cols: Optional[int] = None, row_major: bool = True, inplace: bool = False):
    """LU = M P with column pivoting.  -> (lu, reorder uint8[B,cols], parity, degenerate_ct)."""
    cols = rows if cols is None else cols
    layout, B, E, dt, dev, _ = _info(m)
    assert E == rows * cols
    lu = m if inplace else _clone(m)
    perm, parity, degen = _u8(B, cols, dev).reshape(B, cols), _u8(B, 1, dev), _u8(B, 1, dev)
    _call("gmb_lup_decomp", _SUF[dt], rows, cols, B, _ptr(lu), perm.data_ptr(), parity.data_ptr(), degen.data_ptr(),
              layout, int(row_major), _stream_ptr(dev))
    return lu, perm, parity, degen


def backsolve_lu(lu: Batch, x: Batch, n: int, m: int = 1, skip: int = 0, row_major: bool = True, inplace: bool = False):
    """Solve L U X = X_in (X already permuted).  -> X."""
    layout, B, E, dt, dev, _ = _info(lu)
    assert E == n * n
    _same(lu, x, what="backsolve_lu")
    _shape(x, B, n * m, "backsolve_lu: x")
    out = x if inplace else _clone(x)
    _call("gmb_lu_backsolve", _SUF[dt], n, m, B, _ptr(lu), _ptr(out), skip, layout, int(row_major), _stream_ptr(dev))
    return out


def backsolve_plu(plu: Batch, p: torch.Tensor, b: Batch, n: int, m: int = 1, skip: int = 0, row_major: bool = True):
    """Solve L U X = P B given the row-source array p (uint8[B,n]).  -> X."""
    layout, B, E, dt, dev, _ = _info(plu)
    assert E == n * n and p.dtype == torch.uint8 and p.is_contiguous()
    _same(plu, b, what="backsolve_plu")
    _shape(b, B, n * m, "backsolve_plu: b")
    x = _new_like(b)
    _call("gmb_plu_backsolve", _SUF[dt], n, m, B, _ptr(plu), p.data_ptr(), _ptr(x), _ptr(b), skip, layout,
              int(row_major), _stream_ptr(dev))
    return x


def apply_permutation(p: torch.Tensor, a: Batch, n: int, m: int, row_major: bool = True, inplace: bool = False):
    """A <- P A (row i <- row p[i]).  p is left intact."""
    layout, B, E, dt, dev, _ = _info(a)
    assert E == n * m and p.dtype == torch.uint8
    out = a if inplace else _clone(a)
    _call("gmb_apply_permutation", _SUF[dt], n, m, B, p.data_ptr(), _ptr(out), layout, int(row_major), _stream_ptr(dev))
    return out


def linear_solve(a: Batch, b: Batch, n: int, m: int = 1, skip: int = 0, row_major: bool = True,
                 x_out: Optional[Batch] = None, ok_out: Optional[torch.Tensor] = None, want_lu: bool = False):
    """Fused PLU solve of A X = B (pointer-form linear_solve).  -> (x, ok uint8[B][, lu]).

    ok[s] == 0 iff the reference returns false; then x[s] == b[s] ("x untouched").
    """
    layout, B, E, dt, dev, _ = _info(a)
    assert E == n * n
    _same(a, b, x_out, what="linear_solve")
    _shape(b, B, n * m, "linear_solve: b")
    x = _new_like(b) if x_out is None else x_out
    _shape(x, B, n * m, "linear_solve: x_out")
    ok = _u8(B, 1, dev) if ok_out is None else ok_out
    _flags(ok, B, dev, "linear_solve: ok_out")
    lu = _new_like(a) if want_lu else None
    _call("gmb_plu_solve", _SUF[dt], n, m, B, _ptr(a), _ptr(b), _ptr(x), ok.data_ptr(), _ptr(lu), skip, layout,
              int(row_major), _stream_ptr(dev))
    return (x, ok, lu) if want_lu else (x, ok)


def linear_solve_vec(bases: Batch, b: Batch, n: int, m: int = 1, skip: int = 0):
    """Vec-form linear_solve: bases = n column vectors; n < 5 is inverse-then-multiply.  -> (x, ok)."""
    layout, B, E, dt, dev, _ = _info(bases)
    assert E == n * n
    _same(bases, b, what="linear_solve_vec")
    _shape(b, B, n * m, "linear_solve_vec: b")
    x = _new_like(b)
    ok = _u8(B, 1, dev)
    _call("gmb_linear_solve_vec", _SUF[dt], n, m, B, _ptr(bases), _ptr(b), _ptr(x), ok.data_ptr(), skip, layout,
              _stream_ptr(dev))
    return x, ok


def cholesky(a: Batch, n: int, row_major: bool = True, inplace: bool = False):
    """In-place lower Cholesky factor.  -> (L, ok)."""
    layout, B, E, dt, dev, _ = _info(a)
    assert E == n * n
    l = a if inplace else _clone(a)
    ok = _u8(B, 1, dev)
    _call("gmb_cholesky", _SUF[dt], n, B, _ptr(l), ok.data_ptr(), layout, int(row_major), _stream_ptr(dev))
    return l, ok


def cholesky_solve(a: Batch, b: Batch, n: int, m: int = 1, row_major: bool = True, x_out: Optional[Batch] = None,
                   ok_out: Optional[torch.Tensor] = None, want_l: bool = False):
    """Fused Cholesky factor + solve (framework-defined; the reference has no Cholesky solve).  -> (x, ok[, L])."""
    layout, B, E, dt, dev, _ = _info(a)
    assert E == n * n
    _same(a, b, x_out, what="cholesky_solve")
    _shape(b, B, n * m, "cholesky_solve: b")
    x = _new_like(b) if x_out is None else x_out
    _shape(x, B, n * m, "cholesky_solve: x_out")
    ok = _u8(B, 1, dev) if ok_out is None else ok_out
    _flags(ok, B, dev, "cholesky_solve: ok_out")
    l = _new_like(a) if want_l else None
    _call("gmb_cholesky_solve", _SUF[dt], n, m, B, _ptr(a), _ptr(b), _ptr(x), ok.data_ptr(), _ptr(l), layout,
              int(row_major), _stream_ptr(dev))
    return (x, ok, l) if want_l else (x, ok)


def inv(a: Batch, n: int, src_row_major: bool = True, dst_row_major: bool = True, out: Optional[Batch] = None,
        ok_out: Optional[torch.Tensor] = None):
    """Matrix inverse (adjugate for n <= 4, PLU for n >= 5).  -> (inverse, ok); out[s] untouched where ok == 0."""
    layout, B, E, dt, dev, _ = _info(a)
    assert E == n * n
    if out is None:
        out = _new_like(a)
        (out.data if isinstance(out, BatchSoA) else out).zero_()
    _same(a, out, what="inv")
    _shape(out, B, n * n, "inv: out")
    ok = _u8(B, 1, dev) if ok_out is None else ok_out
    _flags(ok, B, dev, "inv: ok_out")
    _call("gmb_inv", _SUF[dt], n, B, _ptr(a), _ptr(out), ok.data_ptr(), layout, int(src_row_major),
              int(dst_row_major), _stream_ptr(dev))
    return out, ok


def orthogonal(v: Batch, n: int) -> Batch:
    """geom::orthogonal: v = n-1 vectors of n per system, shape (B, (n-1)*n) -> (B, n)."""
    layout, B, E, dt, dev, _ = _info(v)
    assert E == (n - 1) * n
    out = _new_like(v, n)
    _call("gmb_orthogonal", _SUF[dt], n, B, _ptr(v), _ptr(out), layout, _stream_ptr(dev))
    return out


def nullspace(bases: Batch, N: int, nb: int):
    """geom::nullspace: bases = nb vectors of N per system -> (null basis (B, (N-nb)*N), ok uint8[B])."""
    layout, B, E, dt, dev, _ = _info(bases)
    assert E == nb * N and 1 <= nb < N
    null = _new_like(bases, (N - nb) * N)
    ok = _u8(B, 1, dev)
    _call("gmb_nullspace", _SUF[dt], N, nb, B, _ptr(bases), _ptr(null), ok.data_ptr(), layout, _stream_ptr(dev))
    return null, ok


# ---- batched simplex queries (geomc/shape/Simplex.h) ---------------------------------
def simplex_contains(verts: Batch, p: Batch, dim: int, nverts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Simplex<T,N>::contains (Simplex.h:450): verts (B, (N+1)*N) vertex-major, p (B, N) -> inside uint8[B].
    nverts (uint8[B], optional) = Simplex::n; simplexes with fewer than N+1 vertices contain nothing."""
    layout, B, E, dt, dev, _ = _info(verts)
    assert E == (dim + 1) * dim and _info(p)[2] == dim
    inside = _u8(B, 1, dev)
    _call("gmb_simplex_contains", _SUF[dt], dim, B, _ptr(verts), 0, None if nverts is None else nverts.data_ptr(),
              _ptr(p), inside.data_ptr(), layout, _stream_ptr(dev))
    return inside


def trace_simplex(verts: Batch, origin: Batch, direction: Batch, dim: int, want_uv: bool = True):
    """trace_simplex<T,N> (Simplex.h:647): verts (B, N*N) = N vertices of N, rays (B, N) each ->
    (hit uint8[B], s (B,), uv (B, N-1) or None).  s and uv are written only where hit (zero elsewhere);
    want_uv=False solves with skip = N-1, as the reference does for a null `uv`."""
    layout, B, E, dt, dev, _ = _info(verts)
    assert E == dim * dim and _info(origin)[2] == dim and _info(direction)[2] == dim
    hit = _u8(B, 1, dev)
    s = torch.zeros(B, dtype=dt, device=dev)
    uv = None
    if want_uv:
        uv = _new_like(verts, dim - 1)
        (uv.data if isinstance(uv, BatchSoA) else uv).zero_()
    _call("gmb_trace_simplex", _SUF[dt], dim, B, _ptr(verts), 0, _ptr(origin), _ptr(direction), 0, hit.data_ptr(),
              s.data_ptr(), _ptr(uv), layout, _stream_ptr(dev))
    return hit, s, uv


def simplex_intersect(verts: Batch, origin: Batch, direction: Batch, dim: int, nverts: Optional[torch.Tensor] = None) -> Batch:
    """Simplex<T,N>::intersect(Ray) (Simplex.h:251): verts (B, (N+1)*N) -> the Rect<T,1> as (B, 2) = lo, hi."""
    layout, B, E, dt, dev, _ = _info(verts)
    assert E == (dim + 1) * dim and _info(origin)[2] == dim and _info(direction)[2] == dim
    interval = _new_like(verts, 2)
    _call("gmb_simplex_intersect", _SUF[dt], dim, B, _ptr(verts), 0, None if nverts is None else nverts.data_ptr(),
              _ptr(origin), _ptr(direction), 0, _ptr(interval), layout, _stream_ptr(dev))
    return interval


def mul(a: Batch, b: Batch, r: int, k: int, c: int, out: Optional[Batch] = None, acc: bool = False,
        a_row_major: bool = True, b_row_major: bool = True, d_row_major: bool = True) -> Batch:
    """geom::mul(&d, a, b) / mul_acc (Matrix.h:386 / :433): a (B, r*k) * b (B, k*c) -> d (B, r*c).
    c == 1: matrix * Vec; r == 1: Vec * matrix.  `out` may be `a` or `b` for shapes up to 4 x 4 x 4."""
    layout, B, E, dt, dev, _ = _info(a)
    assert E == r * k and _info(b)[2] == k * c and _info(b)[0] == layout
    if out is None:
        assert not acc, "mul_acc needs the destination to accumulate into"
        out = _new_like(a, r * c)
    assert _info(out)[2] == r * c and _info(out)[0] == layout
    _call("gmb_mul", _SUF[dt], r, k, c, B, _ptr(a), _ptr(b), _ptr(out), int(acc), layout, int(a_row_major),
              int(b_row_major), int(d_row_major), _stream_ptr(dev))
    return out


XF_MODES = ("apply", "apply_inverse", "apply_direction", "apply_inverse_direction", "apply_normal", "apply_inverse_normal")


def xf_from_matrix(m: Batch, dim: int, row_major: bool = True) -> Tuple[Batch, torch.Tensor]:
    """geom::transformation(mat) (AffineTransform.h:862): m (B, N*N) -> (records (B, 2*(N+1)^2), ok uint8[B]).
    A record is mat | inv, both (N+1) x (N+1) row-major; a singular m gives inv = identity (and ok = 0)."""
    layout, B, E, dt, dev, _ = _info(m)
    assert E == dim * dim
    xf = _new_like(m, 2 * (dim + 1) ** 2)
    ok = _u8(B, 1, dev)
    _call("gmb_xf_from_matrix", _SUF[dt], dim, B, _ptr(m), _ptr(xf), ok.data_ptr(), layout, int(row_major),
              _stream_ptr(dev))
    return xf, ok


def _xf_from_vec(name: str, t: Batch, dim: int) -> Batch:
    layout, B, E, dt, dev, _ = _info(t)
    assert E == dim
    xf = _new_like(t, 2 * (dim + 1) ** 2)
    _call(name, _SUF[dt], dim, B, _ptr(t), _ptr(xf), layout, _stream_ptr(dev))
    return xf


def xf_translation(t: Batch, dim: int) -> Batch:
    """geom::translation(Vec<T,N>) (AffineTransform.h:814): t (B, N) -> records."""
    return _xf_from_vec("gmb_xf_translation", t, dim)


def xf_scale(sx: Batch, dim: int) -> Batch:
    """geom::scale(Vec<T,N>) (AffineTransform.h:838): sx (B, N) -> records (inv diagonal = 1 / sx)."""
    return _xf_from_vec("gmb_xf_scale", sx, dim)


def xf_compose(xf1: Batch, xf2: Batch, dim: int, inverse: bool = False) -> Batch:
    """xf1 * xf2 (AffineTransform.h:142: apply xf2, then xf1) or, with inverse=True, xf1 / xf2 (:154)."""
    layout, B, E, dt, dev, _ = _info(xf1)
    assert E == 2 * (dim + 1) ** 2 and _info(xf2)[2] == E and _info(xf2)[0] == layout
    out = _new_like(xf1)
    _call("gmb_xf_compose", _SUF[dt], dim, B, _ptr(xf1), _ptr(xf2), _ptr(out), int(inverse), layout, _stream_ptr(dev))
    return out


def xf_apply(xf, p: Batch, dim: int, mode: str = "apply") -> Batch:
    """AffineTransform::apply & co. (AffineTransform.h:224-293) on a point array p (B, N).
    xf: a batch of B records (same layout as p), or ONE record as a tensor of shape (1, 2*(N+1)^2) / (2*(N+1)^2,)
    applied to every point."""
    layout, B, E, dt, dev, _ = _info(p)
    assert E == dim
    E_xf = 2 * (dim + 1) ** 2
    if isinstance(xf, BatchSoA) and xf.B == 1:
        xf = xf.to_aos()                                   # a one-record container is not a packed record
    if isinstance(xf, torch.Tensor) and xf.numel() == E_xf:
        assert xf.is_cuda and xf.is_contiguous() and xf.dtype == dt
        count, pxf = 1, xf.data_ptr()
    else:
        assert _info(xf)[0] == layout and _info(xf)[1] == B and _info(xf)[2] == E_xf
        count, pxf = B, _ptr(xf)
    out = _new_like(p)
    _call("gmb_xf_apply", _SUF[dt], dim, XF_MODES.index(mode), B, pxf, count, _ptr(p), _ptr(out), layout,
              _stream_ptr(dev))
    return out


def solve_residual(a: Batch, x: Batch, b: Batch, ok: Optional[torch.Tensor], n: int, row_major: bool = True) -> torch.Tensor:
    """Device-side batch statistics: [sum of per-system max|Ax-b|, max, #ok==0, #non-finite] (float64[4])."""
    layout, B, E, dt, dev, _ = _info(a)
    stats = torch.zeros(4, dtype=torch.float64, device=dev)
    _call("gmb_solve_residual", _SUF[dt], n, B, _ptr(a), _ptr(x), _ptr(b), None if ok is None else ok.data_ptr(),
              stats.data_ptr(), layout, int(row_major), _stream_ptr(dev))
    return stats


def checksum(t: torch.Tensor, word_offset: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Additive 64-bit checksum of a device buffer's bytes (int64[1], wraps mod 2^64).

    Shards hashed with their global 8-byte-word offsets sum to the checksum of the whole buffer."""
    assert t.is_cuda and t.is_contiguous()
    if out is None:
        out = torch.zeros(1, dtype=torch.int64, device=t.device)
    _lib.check(_lib.lib().gmb_checksum_bytes(t.data_ptr(), t.numel() * t.element_size(), word_offset, out.data_ptr(),
                                             _stream_ptr(t.device)), "gmb_checksum_bytes")
    return out


def checksum_host(buf) -> int:
    """The same checksum computed with numpy on the host (for tests / cross-checks of small buffers)."""
    import numpy as np
    b = np.frombuffer(np.ascontiguousarray(buf).tobytes(), dtype=np.uint8)
    pad = (-len(b)) % 8
    w = np.concatenate([b, np.zeros(pad, np.uint8)]).view(np.uint64)
    with np.errstate(over="ignore"):
        z = w + np.uint64(0x9E3779B97F4A7C15) * (np.arange(len(w), dtype=np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(z.sum(dtype=np.uint64))


# ------------------------------------------------------------------------- host-buffer entry points
def _host_ptr(t):
    """(address, dtype name) of a contiguous numpy array or CPU torch tensor."""
    import numpy as np
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data, t.dtype.name
    assert not t.is_cuda and t.is_contiguous()
    return t.data_ptr(), str(t.dtype).replace("torch.", "")


def _host_suf(*arrays) -> str:
    """f32 / f64 from the DTYPE of the floating-point operands (they must all agree; anything else is refused)."""
    names = {_host_ptr(a)[1] for a in arrays}
    if names == {"float32"}:
        return "f32"
    if names == {"float64"}:
        return "f64"
    raise TypeError(f"host batches must be all float32 or all float64, got {sorted(names)}")


def _host_u8(t) -> int:
    p, name = _host_ptr(t)
    if name != "uint8":
        raise TypeError(f"flag / index outputs are uint8 arrays, got {name}")
    return p


def host_linear_solve(a, b, x, ok, n: int, m: int = 1, skip: int = 0, row_major: bool = True, device=0) -> None:
    """Host AoS buffers in, host buffers out (pipelined H2D / kernel / D2H).  numpy or CPU torch tensors.
    `device` is one device index, or a sequence of them: the batch is then sharded contiguously over those GPUs
    (gmb_host_plu_solve_multi_*: one pipeline and one host thread per device, no collective)."""
    suf = _host_suf(a, b, x)
    B = a.shape[0]
    if isinstance(device, (list, tuple)):
        devs = (C.c_int * len(device))(*device)
        _lib.call("gmb_host_plu_solve_multi", suf, n, m, B, _host_ptr(a)[0], _host_ptr(b)[0], _host_ptr(x)[0], _host_u8(ok), skip,
                  int(row_major), C.cast(devs, C.c_void_p), len(device))
        return
    _lib.call("gmb_host_plu_solve", suf, n, m, B, _host_ptr(a)[0], _host_ptr(b)[0], _host_ptr(x)[0], _host_u8(ok), skip,
              int(row_major), device)


def host_inv(a, out, ok, n: int, src_row_major: bool = True, dst_row_major: bool = True, device=0) -> None:
    suf = _host_suf(a, out)
    if isinstance(device, (list, tuple)):
        devs = (C.c_int * len(device))(*device)
        _lib.call("gmb_host_inv_multi", suf, n, a.shape[0], _host_ptr(a)[0], _host_ptr(out)[0], _host_u8(ok), int(src_row_major),
                  int(dst_row_major), C.cast(devs, C.c_void_p), len(device))
        return
    _lib.call("gmb_host_inv", suf, n, a.shape[0], _host_ptr(a)[0], _host_ptr(out)[0], _host_u8(ok), int(src_row_major),
              int(dst_row_major), device)


def host_cholesky_solve(a, b, x, ok, n: int, m: int = 1, row_major: bool = True, device=0) -> None:
    suf = _host_suf(a, b, x)
    if isinstance(device, (list, tuple)):
        devs = (C.c_int * len(device))(*device)
        _lib.call("gmb_host_cholesky_solve_multi", suf, n, m, a.shape[0], _host_ptr(a)[0], _host_ptr(b)[0], _host_ptr(x)[0],
                  _host_u8(ok), int(row_major), C.cast(devs, C.c_void_p), len(device))
        return
    _lib.call("gmb_host_cholesky_solve", suf, n, m, a.shape[0], _host_ptr(a)[0], _host_ptr(b)[0], _host_ptr(x)[0],
              _host_u8(ok), int(row_major), device)


def host_decomp_plu(a, perm, parity, degen, rows: int, cols: Optional[int] = None, row_major: bool = True, device=0) -> None:
    cols = rows if cols is None else cols
    suf = _host_suf(a)
    if isinstance(device, (list, tuple)):
        devs = (C.c_int * len(device))(*device)
        _lib.call("gmb_host_plu_decomp_multi", suf, rows, cols, a.shape[0], _host_ptr(a)[0], _host_u8(perm), _host_u8(parity),
                  _host_u8(degen), int(row_major), C.cast(devs, C.c_void_p), len(device))
        return
    _lib.call("gmb_host_plu_decomp", suf, rows, cols, a.shape[0], _host_ptr(a)[0], _host_u8(perm), _host_u8(parity),
              _host_u8(degen), int(row_major), device)
