"""Sweep every fast-path kernel family over n = 2..16, float and double, AoS objects AND the SoA container (both resident
in HBM, working set of every case > 2 x L2), and print systems/s and the fraction of the measured HBM peak.
    python profiles/sweep.py [nmin nmax] > gpurun_out/sweep.md

Algorithmic bytes per system: solve  E*n*n + 2*E*n*m (+1);  decomp (in place) 2*E*n*n + n + 2;
inverse 2*E*n*n (+1);  cholesky+solve: SoA E*n*(n+1)/2 (only the lower-triangle planes are read) + 2*E*n (+1), AoS E*n*n +
2*E*n (+1) (records are contiguous: the upper triangle travels too);  backsolve_plu (m = 3) E*n*n + n + 2*E*n*3.   E = sizeof(T).
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


def timeit(fn, reps=3, before=None):
    if before:
        before()
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        if before:
            before()                                   # e.g. restore the input of an in-place kernel (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)                      # ~0.2 ms of device spin ahead of e0: the host launch path of fn() hides behind it
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def restore(dst, src):
    (dst.data if hasattr(dst, "data") and not torch.is_tensor(dst) else dst).copy_(src.data if not torch.is_tensor(src) else src)


print("| n | type | layout | solve sys/s (of peak) | decomp_plu | inv | cholesky+solve | solve m=3 | decomp_lup | backsolve_plu m=3 |")
print("|---|---|---|---|---|---|---|---|---|---|")
NMIN, NMAX = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 16)
WS_MB = int(os.environ.get("GMB_SWEEP_MB", "1024"))      # matrix bytes per case (round 2 until run r2q: 256)
for n in range(NMIN, NMAX + 1):
    for dt, name, E in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        # the matrix batch alone is >= 1 GB (8 x the 126 MB L2), so a best-of-3 launch never runs out of L2
        B = max(1 << 18, 1 << ((WS_MB << 20) // (E * n * n) - 1).bit_length())
        A, b = grid((B, n * n), dt), grid((B, n), dt)
        R = torch.randint(-128, 128, (B, n, n), device=dev, generator=g).to(dt) * 2.0 ** -5
        S = (R @ R.transpose(1, 2) + torch.eye(n, device=dev, dtype=dt)).reshape(B, n * n).contiguous()
        del R
        ok = torch.empty(B, dtype=torch.uint8, device=dev)
        outs = (torch.empty((B, n), dtype=torch.uint8, device=dev), torch.empty(B, dtype=torch.uint8, device=dev),
                torch.empty(B, dtype=torch.uint8, device=dev))
        for lay in ("AoS", "SoA"):
            if lay == "SoA":
                A, b, S = gm.BatchSoA.from_aos(A), gm.BatchSoA.from_aos(b), gm.BatchSoA.from_aos(S)
            x = b.clone()
            out = A.clone()
            M = A.clone()
            b3 = grid((B, 3 * n), dt)
            if lay == "SoA":
                b3 = gm.BatchSoA.from_aos(b3)
            x3 = b3.clone()
            lu, perm, _, _ = gm.decomp_plu(A, n)
            chol_bytes = (E * n * (n + 1) // 2 if lay == "SoA" else E * n * n) + 2 * E * n + 1
            cells = []
            for fn, nbytes, before in ((lambda: gm.linear_solve(A, b, n, x_out=x, ok_out=ok), E * n * n + 2 * E * n + 1, None),
                                       (lambda: gm.decomp_plu(M, n, inplace=True, out=outs), 2 * E * n * n + n + 2, lambda: restore(M, A)),
                                       (lambda: gm.inv(A, n, out=out, ok_out=ok), 2 * E * n * n + 1, None),
                                       (lambda: gm.cholesky_solve(S, b, n, x_out=x, ok_out=ok), chol_bytes, None),
                                       (lambda: gm.linear_solve(A, b3, n, 3, x_out=x3, ok_out=ok), E * n * n + 6 * E * n + 1, None),
                                       (lambda: gm.decomp_lup(M, n, inplace=True), 2 * E * n * n + n + 2, lambda: restore(M, A)),
                                       (lambda: gm.backsolve_plu(lu, perm, b3, n, 3), E * n * n + n + 6 * E * n, None)):
                ms = timeit(fn, before=before)
                rate = B / (ms * 1e-3)
                cells.append(f"{rate:.3g} ({rate * nbytes / (peak * 1e9):.2f})")
            print(f"| {n} | {name} | {lay} | " + " | ".join(cells) + " |", flush=True)
            del x, out, M, b3, x3, lu, perm
        del A, b, ok, S, outs
