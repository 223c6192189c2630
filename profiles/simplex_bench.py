"""Throughput of the fused simplex queries, batches resident in HBM.
    python profiles/simplex_bench.py > gpurun_out/simplex.md
Algorithmic bytes per query: contains (N+1)*N*E + N*E + 1;  trace N*N*E + 2*N*E + 1 + E + (N-1)*E;
intersect (N+1)*N*E + 2*N*E + 2*E.   E = sizeof(T)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import geomc_b200 as gm

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6548.8


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("| query | N | type | layout | queries/s | of HBM peak |")
print("|---|---|---|---|---|---|")
for N in (2, 3, 4, 7):
    for dt, name, E in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        B = 1 << (24 if N <= 4 else 22)
        verts = torch.randn((B, (N + 1) * N), device=dev, dtype=dt, generator=g)
        w = torch.rand((B, N + 1), device=dev, dtype=dt, generator=g)
        w = w / w.sum(1, keepdim=True)
        p = torch.einsum("bk,bkd->bd", w, verts.view(B, N + 1, N)).contiguous()
        o = 3 * torch.randn((B, N), device=dev, dtype=dt, generator=g)
        d = (p - o).contiguous()
        tv = verts[:, :N * N].contiguous()
        for soa in (False, True):
            cv = (lambda t: gm.BatchSoA.from_aos(t)) if soa else (lambda t: t)
            V, P, O, D, TV = cv(verts), cv(p), cv(o), cv(d), cv(tv)
            for label, fn, nbytes in (("contains", lambda: gm.simplex_contains(V, P, N), (N + 1) * N * E + N * E + 1),
                                      ("trace", lambda: gm.trace_simplex(TV, O, D, N), N * N * E + 2 * N * E + 1 + E + (N - 1) * E),
                                      ("intersect", lambda: gm.simplex_intersect(V, O, D, N), (N + 1) * N * E + 2 * N * E + 2 * E)):
                ms = timeit(fn)
                rate = B / (ms * 1e-3)
                print(f"| {label} | {N} | {name} | {'SoA' if soa else 'AoS'} | {rate:.3g} | {rate * nbytes / (peak * 1e9):.2f} |", flush=True)
            del V, P, O, D, TV
        del verts, w, p, o, d, tv
