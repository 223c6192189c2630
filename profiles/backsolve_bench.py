"""backsolve_plu (a4) for n = 8, 12, 16, AoS, one right-hand side: systems/s and fraction of the measured HBM peak
(bytes per system: LU + perm + b + x).   python profiles/backsolve_bench.py"""
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
import geomc_b200 as gm
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
def grid(shape, dt): return torch.randint(-(1 << 19), 1 << 19, shape, device=dev, generator=g).to(dt) * 2.0 ** -17
for n in (8, 12, 16):
    for dt in (torch.float32, torch.float64):
        B = 1 << 19
        A = grid((B, n * n), dt); b = grid((B, n), dt)
        lu, p, par, dg = gm.decomp_plu(A, n)
        x = torch.empty_like(b)
        def f(): return gm.backsolve_plu(lu, p, b, n, 1, 0, True)
        f(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
        E = 4 if dt == torch.float32 else 8
        nbytes = E * n * n + n + 2 * E * n
        print(f"backsolve_plu n={n} {dt}: {best:.3f} ms  {B/best*1e3:.3g} sys/s  of peak {B/best*1e3*nbytes/6548.8e9:.2f}")
