"""A/B of the one-thread-per-system kernels (gmb_lane.cuh) against the kernels they would replace, n = 5..8, float and
double, AoS objects and the SoA container: systems/s and the fraction of the measured HBM peak (algorithmic bytes as in
profiles/sweep.py).     python profiles/lane_ab.py > gpurun_out/lane_ab.md"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomc_b200 as gm  # noqa: E402

peak = 6548.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)


def grid(shape, dt):
    return torch.randint(-(1 << 19), 1 << 19, shape, device=dev, generator=g).to(dt) * 2.0 ** -17


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("| n | type | layout | m | solve: exact kernels | solve: lane (speculative) | chol+solve: row kernel | chol+solve: lane |")
print("|---|---|---|---|---|---|---|---|")
for n in range(5, 9):
    for dt, name, E in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        if E == 8 and n > 6:
            continue
        B = max(1 << 18, 1 << ((256 << 20) // (E * n * n) - 1).bit_length())
        A0 = grid((B, n * n), dt)
        R = torch.randint(-128, 128, (B, n, n), device=dev, generator=g).to(dt) * 2.0 ** -5
        S0 = (R @ R.transpose(1, 2) + torch.eye(n, device=dev, dtype=dt)).reshape(B, n * n).contiguous()
        del R
        ok = torch.empty(B, dtype=torch.uint8, device=dev)
        for m in (1, 3):
            b0 = grid((B, n * m), dt)
            for lay in ("AoS", "SoA"):
                A, b, S = (A0, b0, S0) if lay == "AoS" else (gm.BatchSoA.from_aos(A0), gm.BatchSoA.from_aos(b0), gm.BatchSoA.from_aos(S0))
                x = b.clone()
                cells = []
                nb_solve = E * n * n + 2 * E * n * m + 1
                nb_chol = E * n * (n + 1) // 2 + 2 * E * n * m + 1
                for opt, fn, nbytes in (("lanesolve", lambda: gm.linear_solve(A, b, n, m, x_out=x, ok_out=ok), nb_solve),
                                        ("lanechol", lambda: gm.cholesky_solve(S, b, n, m, x_out=x, ok_out=ok), nb_chol)):
                    res = []
                    for v in (0, 1):
                        gm.set_option(opt, v)
                        try:
                            ms = timeit(fn)
                            rate = B / (ms * 1e-3)
                            res.append((f"{rate:.3g} ({rate * nbytes / (peak * 1e9):.2f})", x.to_aos().clone() if lay == "SoA" else x.clone()))
                        except Exception as ex:
                            res.append((f"error {str(ex)[:40]}", None))
                    gm.set_option(opt, -1)
                    same = ""
                    if res[0][1] is not None and res[1][1] is not None:
                        it = torch.int32 if E == 4 else torch.int64
                        same = " =" if torch.equal(res[0][1].view(it), res[1][1].view(it)) else " DIFF"
                    cells += [res[0][0], res[1][0] + same]
                print(f"| {n} | {name} | {lay} | {m} | " + " | ".join(cells) + " |", flush=True)
                del x
        del A0, S0, ok
