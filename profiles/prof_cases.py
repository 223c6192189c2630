"""Small driver for ncu captures of the secondary kernels (one launch of each after a warm-up).

    python profiles/prof_cases.py [case ...]      cases: c2 c3aos c3soa c4n4 c4n8 c5 c1aos
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import geomc_b200 as gm

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)


def grid(shape, dt):
    return (torch.randint(-(1 << 19), 1 << 19, shape, device=dev, generator=g).to(dt) * 2.0 ** -17)


def run(case):
    f32, f64 = torch.float32, torch.float64
    if case == "c2":
        B = 1 << 23
        R = torch.randint(-128, 128, (B, 3, 3), device=dev, generator=g).double() * 2.0 ** -5
        S = (R @ R.transpose(1, 2) + torch.eye(3, device=dev, dtype=f64)).reshape(B, 9).contiguous()
        Ss, bs = gm.BatchSoA.from_aos(S), gm.BatchSoA.from_aos(grid((B, 3), f64))
        xs, ok = bs.empty_like(), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.cholesky_solve(Ss, bs, 3, x_out=xs, ok_out=ok)
    if case in ("c3aos", "c3soa", "c1aos"):
        B = 1 << 23
        A, b = grid((B, 16), f32), grid((B, 4), f32)
        ok = torch.zeros(B, dtype=torch.uint8, device=dev)
        if case == "c1aos":
            x = torch.empty_like(b)
            return lambda: gm.linear_solve(A, b, 4, x_out=x, ok_out=ok)
        if case == "c3soa":
            A = gm.BatchSoA.from_aos(A)
        out = A.empty_like() if case == "c3soa" else torch.empty_like(A)
        return lambda: gm.inv(A, 4, out=out, ok_out=ok)
    if case.startswith("c4n"):
        n = int(case[3:])
        B = 1 << 21
        M = gm.BatchSoA(B, n * n, f64, dev)
        M.data.copy_(grid(M.data.shape, f64))
        return lambda: gm.decomp_plu(M, n, inplace=True)
    if case == "c5":
        B = 1 << 19
        A, b = grid((B, 256), f32), grid((B, 16), f32)
        x, ok = torch.empty_like(b), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.linear_solve(A, b, 16, x_out=x, ok_out=ok)
    if case.startswith("inv"):            # inv<f|d><n>: matrix inverse, AoS
        dt = f32 if case[3] == "f" else f64
        n = int(case[4:])
        B = 1 << 20
        A = grid((B, n * n), dt)
        out, ok = torch.empty_like(A), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.inv(A, n, out=out, ok_out=ok)
    if case.startswith("bs"):             # bs<f|d><n>: backsolve_plu with three right-hand sides, AoS records
        dt = f32 if case[2] == "f" else f64
        n = int(case[3:])
        B = 1 << 20
        A, b3 = grid((B, n * n), dt), grid((B, 3 * n), dt)
        lu, perm, _, _ = gm.decomp_plu(A, n)
        return lambda: gm.backsolve_plu(lu, perm, b3, n, 3)
    if case.startswith("chol"):           # chol<f|d><n>: fused Cholesky + solve, SoA, SPD inputs
        dt = f32 if case[4] == "f" else f64
        n = int(case[5:])
        B = 1 << 20
        R = torch.randint(-128, 128, (B, n, n), device=dev, generator=g).to(dt) * 2.0 ** -5
        S = gm.BatchSoA.from_aos((R @ R.transpose(1, 2) + torch.eye(n, device=dev, dtype=dt)).reshape(B, n * n).contiguous())
        b = gm.BatchSoA.from_aos(grid((B, n), dt))
        x, ok = b.empty_like(), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.cholesky_solve(S, b, n, x_out=x, ok_out=ok)
    if case.startswith("dec"):            # dec<f|d><n>: PLU decomposition in place, SoA
        dt = f32 if case[3] == "f" else f64
        n = int(case[4:])
        B = 1 << 21
        M = gm.BatchSoA(B, n * n, dt, dev)
        M.data.copy_(grid(M.data.shape, dt))
        outs = (torch.empty((B, n), dtype=torch.uint8, device=dev), torch.empty(B, dtype=torch.uint8, device=dev),
                torch.empty(B, dtype=torch.uint8, device=dev))
        return lambda: gm.decomp_plu(M, n, inplace=True, out=outs)
    if case.startswith("solve"):          # solve<f|d><n>: fused PLU solve, AoS, n x n
        dt = f32 if case[5] == "f" else f64
        n = int(case[6:])
        B = int(os.environ.get("PROF_B", 1 << 20))
        A, b = grid((B, n * n), dt), grid((B, n), dt)
        x, ok = torch.empty_like(b), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.linear_solve(A, b, n, x_out=x, ok_out=ok)
    if case == "c1soa":                   # the metric kernel at the bench size: 2^24 4x4 float systems, SoA container
        B = 1 << 24
        A, b = gm.BatchSoA(B, 16, f32, dev), gm.BatchSoA(B, 4, f32, dev)
        A.data.copy_(grid(A.data.shape, f32))
        b.data.copy_(grid(b.data.shape, f32))
        x, ok = b.empty_like(), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.linear_solve(A, b, 4, 1, 0, True, x_out=x, ok_out=ok)
    if case == "c5big":                   # C5 at its stated size: 2^23 16x16 float systems, AoS
        B = 1 << 23
        A, b = grid((B, 256), f32), grid((B, 16), f32)
        x, ok = torch.empty_like(b), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.linear_solve(A, b, 16, x_out=x, ok_out=ok)
    if case.startswith("ssolve"):         # ssolve<f|d><n>: fused PLU solve, SoA container, n x n
        dt = f32 if case[6] == "f" else f64
        n = int(case[7:])
        B = int(os.environ.get("PROF_B", 1 << 20))
        A, b = gm.BatchSoA.from_aos(grid((B, n * n), dt)), gm.BatchSoA.from_aos(grid((B, n), dt))
        x, ok = b.empty_like(), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.linear_solve(A, b, n, x_out=x, ok_out=ok)
    if case == "c5soa":
        B = 1 << 19
        A, b = gm.BatchSoA.from_aos(grid((B, 256), f32)), gm.BatchSoA.from_aos(grid((B, 16), f32))
        x, ok = b.empty_like(), torch.zeros(B, dtype=torch.uint8, device=dev)
        return lambda: gm.linear_solve(A, b, 16, x_out=x, ok_out=ok)
    if case in ("mul4", "mul16"):         # batched mul, SoA 4x4x4 float / AoS 16x16x16 float
        n = 4 if case == "mul4" else 16
        B = 1 << (24 if n == 4 else 19)
        a, b = grid((B, n * n), f32), grid((B, n * n), f32)
        if n == 4:
            a, b = gm.BatchSoA.from_aos(a), gm.BatchSoA.from_aos(b)
        d = a.empty_like() if n == 4 else torch.empty_like(a)
        return lambda: gm.mul(a, b, n, n, n, out=d)
    if case in ("xf3", "apply3", "apply3one", "compose3"):   # AffineTransform<float,3>, SoA containers, 2^23 transforms / points
        B = 1 << 23
        m = gm.BatchSoA.from_aos(grid((B, 9), f32) + 3 * torch.eye(3, device=dev).reshape(1, 9))
        if case == "xf3":
            return lambda: gm.xf_from_matrix(m, 3)
        xm, _ = gm.xf_from_matrix(m, 3)
        p = gm.BatchSoA.from_aos(grid((B, 3), f32))
        if case == "apply3":
            return lambda: gm.xf_apply(xm, p, 3, "apply")
        if case == "compose3":
            xt = gm.xf_translation(p, 3)
            return lambda: gm.xf_compose(xt, xm, 3)
        one = xm.to_aos()[5:6].contiguous()
        return lambda: gm.xf_apply(one, p, 3, "apply")
    raise SystemExit(f"unknown case {case}")


ONCE = "--once" in sys.argv           # under ncu: a single launch per case
for case in ([a for a in sys.argv[1:] if not a.startswith("--")] or ["c2", "c3aos", "c3soa", "c4n4", "c4n8", "c5"]):
    fn = run(case)
    for _ in range(0 if ONCE else 3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    print(f"{case}: {e0.elapsed_time(e1):.4f} ms")
