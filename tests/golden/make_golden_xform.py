"""Generate tests/golden/xform_*.npz (rows (f)3 and (f)4: AffineTransform construction / composition /
application, and batched mul) from the reference itself.  Same rules as make_golden.py: runs only where
oracle/_ref was built from /root/reference.

    python tests/golden/make_golden_xform.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B = 24
MUL_SHAPES = [(2, 2, 2), (3, 3, 3), (4, 4, 4), (5, 5, 5), (8, 8, 8), (2, 3, 4), (4, 3, 2), (3, 5, 2), (5, 2, 7), (16, 16, 16),
              (16, 3, 9), (7, 16, 5)]
XF_DIMS = range(2, 8)


def matrices(rng, N, dt):
    """N x N on the exact grid; rows: 0 = zero matrix, 1 = duplicate row (singular), 2 = identity,
    3 = signed zeros in the products, 4 = huge / tiny scales."""
    m = po.exact_grid(rng, (B, N * N), dt)
    m[0] = 0
    m[1, N:2 * N] = m[1, :N]
    m[2] = np.eye(N, dtype=dt).ravel()
    m[3, ::2] = -0.0
    m[4] *= dt(2.0) ** 40
    m[5] *= dt(2.0) ** -40
    return m


def main():
    R = po.ref()
    R.set_threads(1)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        g = {}
        with np.errstate(all="ignore"):
            for (r, k, c) in MUL_SHAPES:
                rng = np.random.default_rng(7000 + 100 * r + 10 * k + c + (0 if dt == np.float32 else 5))
                a = po.exact_grid(rng, (B, r * k), dt)
                b = po.exact_grid(rng, (B, k * c), dt)
                d0 = po.exact_grid(rng, (B, r * c), dt)
                a[0] = 0
                a[1, ::2] = -0.0
                b[2] = np.inf
                b[3, 0] = np.nan
                key = f"{r}_{k}_{c}"
                g[f"mul_a_{key}"], g[f"mul_b_{key}"], g[f"mul_d0_{key}"] = a, b, d0
                g[f"mul_d_{key}"] = R.mul(a, b, r, k, c)
                g[f"mul_acc_{key}"] = R.mul(a, b, r, k, c, d0, acc=True)
                g[f"mul_cm_{key}"] = R.mul(a, b, r, k, c, a_rm=False, b_rm=False, d_rm=False)
            for N in XF_DIMS:
                rng = np.random.default_rng(7900 + N + (0 if dt == np.float32 else 50))
                m = matrices(rng, N, dt)
                t = po.exact_grid(rng, (B, N), dt)
                p = po.exact_grid(rng, (B, N), dt)
                g[f"xf_m_{N}"], g[f"xf_t_{N}"], g[f"xf_p_{N}"] = m, t, p
                xm, _ = R.xf_from_matrix(m, N, True)
                g[f"xf_from_rm_{N}"] = xm
                g[f"xf_from_cm_{N}"] = R.xf_from_matrix(m, N, False)[0]
                xt = R.xf_translation(t, N)
                g[f"xf_transl_{N}"] = xt
                g[f"xf_scale_{N}"] = R.xf_translation(t, N, scale=True)
                comp = R.xf_compose(xt, xm, N)                     # translation(t) * transformation(m), shape_generation.h:301
                g[f"xf_comp_{N}"] = comp
                g[f"xf_div_{N}"] = R.xf_compose(comp, xm, N, inverse=True)
                for mode in R.XF_MODES:
                    g[f"xf_{mode}_{N}"] = R.xf_apply(comp, p, N, mode)
                    g[f"xf_{mode}_bcast_{N}"] = R.xf_apply(comp[7:8], p, N, mode)
        np.savez_compressed(os.path.join(OUT, f"xform_{tag}.npz"), **g)
    print("wrote xform fixtures")


if __name__ == "__main__":
    main()
