"""Generate tests/golden/simplex_*.npz (row (f)2: Simplex::contains / trace_simplex / Simplex::intersect)
from the reference itself.  Same rules as make_golden.py: runs only where oracle/_ref was built from
/root/reference.

    python tests/golden/make_golden_simplex.py

trace_simplex's ray parameter is NOT stored: the reference reads x[N], one past the end of its Vec<T,N>
(shape/Simplex.h:675), so that value is whatever lies on its stack.  For the same reason the n == N rows of
`intersect` store only whether the interval is non-empty.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B = 48


def scene(rng, N, K, dt, B=B):
    """K vertices of N per simplex on the exact grid; points / rays: a third aimed inside, a third random,
    plus vertices, centroids, degenerate simplexes and axis-parallel rays."""
    verts = po.exact_grid(rng, (B, K * N), dt)
    V = verts.reshape(B, K, N)
    w = rng.dirichlet(np.ones(K), B)
    inside = np.einsum("bk,bkd->bd", w, V.astype(np.float64)).astype(dt)
    p = np.where((np.arange(B) % 3 == 0)[:, None], inside, po.exact_grid(rng, (B, N), dt))
    p[1] = V[1, 0]                                             # a vertex
    p[2] = ((V[2, 0].astype(np.float64) + V[2, 1]) / 2).astype(dt)   # an edge midpoint
    verts[3, N:2 * N] = verts[3, :N]                           # two coincident vertices
    verts[4] = 0                                               # all vertices at the origin
    o = po.exact_grid(rng, (B, N), dt) * 4
    d = np.where((np.arange(B) % 3 != 1)[:, None], inside - o, po.exact_grid(rng, (B, N), dt)).astype(dt)
    d[5] = 0
    d[5, 0] = 1                                                # axis-parallel ray
    d[6] = 0                                                   # null direction
    return verts, p.astype(dt), o.astype(dt), d


def main():
    R = po.ref()
    R.set_threads(1)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        g = {}
        for N in range(2, 8):
            rng = np.random.default_rng(9100 + N + (0 if dt == np.float32 else 100))
            verts, p, o, d = scene(rng, N, N + 1, dt)
            nv = np.full(B, N + 1, np.uint8)
            nv[7::8] = N
            nv[9::16] = N - 1
            g[f"c_verts_{N}"], g[f"c_p_{N}"], g[f"c_nv_{N}"] = verts, p, nv
            g[f"c_in_{N}"] = R.simplex_contains(verts, p, N, nv)
            g[f"i_o_{N}"], g[f"i_d_{N}"] = o, d
            lohi = R.simplex_intersect(verts, o, d, N, nv)
            g[f"i_hit_{N}"] = (lohi[:, 0] <= lohi[:, 1]).astype(np.uint8)
            lohi[nv == N] = 0                                  # see the module docstring
            g[f"i_lohi_{N}"] = lohi
            tverts, _, to, td = scene(rng, N, N, dt)
            g[f"t_verts_{N}"], g[f"t_o_{N}"], g[f"t_d_{N}"] = tverts, to, td
            for wuv in (1, 0):
                hit, _, uv = R.trace_simplex(tverts, to, td, N, bool(wuv))
                g[f"t_hit{wuv}_{N}"] = hit
                if wuv:
                    g[f"t_uv_{N}"] = uv
        np.savez_compressed(os.path.join(OUT, f"simplex_{tag}.npz"), **g)
    print("wrote simplex fixtures")


if __name__ == "__main__":
    main()
