"""Throughput of the (f)3 / (f)4 kernels (batched mul, AffineTransform construct / compose / apply), batches
resident in HBM, outputs preallocated, straight through the C ABI.
    python profiles/xform_bench.py > gpurun_out/xform.md
Algorithmic bytes: mul (r*k + k*c + r*c) * E;  transformation N*N*E + 2*K*K*E + 1;  compose 3 * 2*K*K*E;
apply with one transform per point 2*N*E + N*K*E (the N rows of the one matrix the mode reads);
apply with one shared transform 2*N*E.   E = sizeof(T), K = N + 1."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import geomc_b200 as gm
from geomc_b200 import _lib

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6548.8
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def ptr(x):
    return x.data.data_ptr() if isinstance(x, gm.BatchSoA) else x.data_ptr()


def row(label, shape, name, soa, B, ms, nbytes):
    rate = B / (ms * 1e-3)
    print(f"| {label} | {shape} | {name} | {'SoA' if soa else 'AoS'} | {rate:.3g} | {rate * nbytes / (peak * 1e9):.2f} |", flush=True)


print("| op | shape | type | layout | systems/s | of HBM peak |")
print("|---|---|---|---|---|---|")
for dt, name, E in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
    for (r, k, c) in ((4, 4, 4), (3, 3, 3), (4, 4, 1), (2, 2, 2), (8, 8, 8), (16, 16, 16)):
        B = 1 << (24 if max(r, k, c) <= 4 else 21 if max(r, k, c) <= 8 else 19)
        a = torch.randn((B, r * k), device=dev, dtype=dt, generator=g)
        b = torch.randn((B, k * c), device=dev, dtype=dt, generator=g)
        for soa in (False, True):
            cv = (lambda t: gm.BatchSoA.from_aos(t)) if soa else (lambda t: t)
            A, Bm = cv(a), cv(b)
            D = gm.BatchSoA(B, r * c, dt, dev) if soa else torch.empty((B, r * c), device=dev, dtype=dt)
            ms = timeit(lambda: _lib.call("gmb_mul", name, r, k, c, B, ptr(A), ptr(Bm), ptr(D), 0, int(soa), 1, 1, 1, st))
            row("mul", f"{r}x{k}x{c}", name, soa, B, ms, (r * k + k * c + r * c) * E)
            del A, Bm, D
        del a, b
    for N in (2, 3, 4, 7):
        K = N + 1
        EX = 2 * K * K
        B = 1 << (23 if N <= 4 else 21)
        m = torch.randn((B, N * N), device=dev, dtype=dt, generator=g)
        t = torch.randn((B, N), device=dev, dtype=dt, generator=g)
        p = torch.randn((B, N), device=dev, dtype=dt, generator=g)
        for soa in (False, True):
            cv = (lambda x: gm.BatchSoA.from_aos(x)) if soa else (lambda x: x)
            new = (lambda e: gm.BatchSoA(B, e, dt, dev)) if soa else (lambda e: torch.empty((B, e), device=dev, dtype=dt))
            M, T_, P = cv(m), cv(t), cv(p)
            X1, X2, X3, O = new(EX), new(EX), new(EX), new(N)
            ok = torch.empty(B, dtype=torch.uint8, device=dev)
            ms = timeit(lambda: _lib.call("gmb_xf_from_matrix", name, N, B, ptr(M), ptr(X1), ok.data_ptr(), int(soa), 1, st))
            row("transformation", f"N={N}", name, soa, B, ms, N * N * E + EX * E + 1)
            _lib.call("gmb_xf_translation", name, N, B, ptr(T_), ptr(X2), int(soa), st)
            ms = timeit(lambda: _lib.call("gmb_xf_compose", name, N, B, ptr(X2), ptr(X1), ptr(X3), 0, int(soa), st))
            row("compose", f"N={N}", name, soa, B, ms, 3 * EX * E)
            ms = timeit(lambda: _lib.call("gmb_xf_apply", name, N, 0, B, ptr(X3), B, ptr(P), ptr(O), int(soa), st))
            row("apply (per-point xf)", f"N={N}", name, soa, B, ms, 2 * N * E + N * K * E)
            ms = timeit(lambda: _lib.call("gmb_xf_apply", name, N, 4, B, ptr(X3), B, ptr(P), ptr(O), int(soa), st))
            row("apply_normal (per-point xf)", f"N={N}", name, soa, B, ms, 2 * N * E + N * K * E)
            one = (X3.to_aos() if soa else X3)[5:6].contiguous()
            ms = timeit(lambda: _lib.call("gmb_xf_apply", name, N, 0, B, one.data_ptr(), 1, ptr(P), ptr(O), int(soa), st))
            row("apply (one xf)", f"N={N}", name, soa, B, ms, 2 * N * E)
            del M, T_, P, X1, X2, X3, O
        del m, t, p
