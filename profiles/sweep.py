"""Sweep every fast-path kernel family over n = 2..16, float and double (AoS resident in HBM), and print
systems/s and the fraction of the measured HBM peak.   python profiles/sweep.py [nmin nmax] > gpurun_out/sweep.md

Algorithmic bytes per system: solve  E*n*n + 2*E*n (+1);  decomp (in place) 2*E*n*n + n + 2;
inverse 2*E*n*n (+1);  cholesky+solve E*n*(n+1)/2 (lower triangle) + 2*E*n (+1).   E = sizeof(T).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import geomc_b200 as gm

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def grid(shape, dt):
    return torch.randint(-(1 << 19), 1 << 19, shape, device=dev, generator=g).to(dt) * 2.0 ** -17


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("| n | type | solve sys/s (of peak) | decomp_plu | inv | cholesky+solve |")
print("|---|---|---|---|---|---|")
NMIN, NMAX = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 16)
for n in range(NMIN, NMAX + 1):
    for dt, name, E in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        B = 1 << (22 if n <= 4 else 20 if n <= 10 else 19)
        A, b = grid((B, n * n), dt), grid((B, n), dt)
        x, ok = torch.empty_like(b), torch.empty(B, dtype=torch.uint8, device=dev)
        out = torch.empty_like(A)
        R = torch.randint(-128, 128, (B, n, n), device=dev, generator=g).to(dt) * 2.0 ** -5
        S = (R @ R.transpose(1, 2) + torch.eye(n, device=dev, dtype=dt)).reshape(B, n * n).contiguous()
        del R
        outs = (torch.empty((B, n), dtype=torch.uint8, device=dev), torch.empty(B, dtype=torch.uint8, device=dev),
                torch.empty(B, dtype=torch.uint8, device=dev))
        M = A.clone()
        cells = []
        for fn, nbytes in ((lambda: gm.linear_solve(A, b, n, x_out=x, ok_out=ok), E * n * n + 2 * E * n + 1),
                           (lambda: gm.decomp_plu(M, n, inplace=True, out=outs), 2 * E * n * n + n + 2),
                           (lambda: gm.inv(A, n, out=out, ok_out=ok), 2 * E * n * n + 1),
                           (lambda: gm.cholesky_solve(S, b, n, x_out=x, ok_out=ok), E * n * (n + 1) // 2 + 2 * E * n + 1)):
            ms = timeit(fn)
            rate = B / (ms * 1e-3)
            cells.append(f"{rate:.3g} ({rate * nbytes / (peak * 1e9):.2f})")
        print(f"| {n} | {name} | " + " | ".join(cells) + " |", flush=True)
        del A, b, x, ok, out, S, M, outs
