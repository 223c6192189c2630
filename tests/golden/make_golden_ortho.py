"""Generate tests/golden/ortho_*.npz (row (f)1: orthogonal / nullspace) from the reference itself.
Same rules as make_golden.py: runs only where oracle/_ref was built from /root/reference.

    python tests/golden/make_golden_ortho.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B = 24


def main():
    R = po.ref()
    R.set_threads(1)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        d = {}
        for n in (2, 3, 4, 5, 8, 16):
            rng = np.random.default_rng(7000 + n + (0 if dt == np.float32 else 100))
            v = po.exact_grid(rng, (B, (n - 1) * n), dt)
            v[0] = 0                                           # degenerate
            if n >= 3:
                v[1, n:2 * n] = v[1, :n]                       # two identical vectors
                v[2, 0::n] = 0                                 # zero first column
            d[f"orth_v_{n}"] = v
            d[f"orth_o_{n}"] = R.orthogonal(v, n)
            if n >= 3:
                for nb in sorted({1, n // 2, n - 2, n - 1} - {0}):
                    bs = po.exact_grid(rng, (B, nb * n), dt)
                    bs[0] = 0
                    if nb >= 2:
                        bs[1, n:2 * n] = 2 * bs[1, :n]        # linearly dependent
                    null, ok = R.nullspace(bs, n, nb)
                    d[f"ns_b_{n}_{nb}"], d[f"ns_x_{n}_{nb}"], d[f"ns_ok_{n}_{nb}"] = bs, null, ok
        np.savez_compressed(os.path.join(OUT, f"ortho_{tag}.npz"), **d)
    print("wrote ortho fixtures")


if __name__ == "__main__":
    main()
