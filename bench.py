#!/usr/bin/env python
"""bench.py -- systems/sec of the batched 4x4 float PLU solve (geom::linear_solve) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is ONE pass of the hot path over one batch of synthetic input: B = 2^24 independent
4x4 float systems per GPU (A: 1 GiB, b: 256 MiB, x: 256 MiB -- far larger than the 126 MB L2, so
no L2 flush is needed between steps), resident in HBM in the SoA-interleaved layout, solved by
ONE launch of the fused PLU-solve kernel through the C ABI (gmb_plu_solve_f32).  Systems are
sharded contiguously across GPUs (weak scaling: every rank owns its own 2^24 systems); the only
collective is the final all-reduce of {residual sum, max, fail count, checksum} scalars.

Prints one JSON line (see the driver contract): value, roofline, cpu_baseline, e2e, clocks, ...
`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified
geomc headers, OpenMP over all host threads; falls back to the C port if _ref is absent).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DIM = 4
LOG2_B = 24                 # systems per GPU per step
ALG_BYTES = 96              # 64 (A) + 16 (b) read, 16 (x) written per system
METRIC = "systems/sec (4x4 f32 PLU solve)"
WORKLOAD = "C1-shape: random 4x4 float systems, fused PLU decompose + linear_solve, 2^24 systems per GPU per step"


def read_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def read_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("plu_solve_f32_soa_bytes_per_system")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads(lib):
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 for its workers; the
    CPU reference arm is meant to use every core, so the OpenMP team size is set explicitly."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if lib.kind == "reference":
        lib.set_threads(n)
        return n
    return 1


def cpu_reference_rate(seconds_budget=12.0, sample_log2=21):
    """The reference's own linear_solve<float,true> on an AoS sample, OpenMP over all host threads."""
    import numpy as np
    from oracle import pyoracle as po
    lib = po.best()
    threads = host_threads(lib)
    Bs = 1 << sample_log2
    rng = np.random.default_rng(123)
    A0 = po.exact_grid(rng, (Bs, 16), np.float32)
    b0 = po.exact_grid(rng, (Bs, 4), np.float32)
    A, b, ok = A0.copy(), b0.copy(), np.zeros(Bs, np.uint8)
    lib.linear_solve_inplace(A, b, ok, N_DIM)          # warm-up: page touch + thread spin-up
    best, reps, t_start = None, 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t_start < seconds_budget and reps < 50):
        np.copyto(A, A0)
        np.copyto(b, b0)
        t0 = time.perf_counter()
        lib.linear_solve_inplace(A, b, ok, N_DIM)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
    return {"value": Bs / best, "unit": "systems/s", "cores": threads, "kind": lib.kind,
            "sample": f"2^{sample_log2} systems per pass (AoS, inputs re-copied before each pass), best of {reps} passes; "
                      f"{lib.build_info()}"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU loop; rank 0 only."""
    if rank != 0:
        return
    import numpy as np
    from oracle import pyoracle as po
    lib = po.best()
    threads = host_threads(lib)
    Bs = 1 << 21
    rng = np.random.default_rng(123)
    A0 = po.exact_grid(rng, (Bs, 16), np.float32)
    b0 = po.exact_grid(rng, (Bs, 4), np.float32)
    A, b, ok = A0.copy(), b0.copy(), np.zeros(Bs, np.uint8)
    for _ in range(max(1, args.warmup)):
        np.copyto(A, A0); np.copyto(b, b0)
        lib.linear_solve_inplace(A, b, ok, N_DIM)
    total = 0.0
    for _ in range(args.steps):
        np.copyto(A, A0); np.copyto(b, b0)
        t0 = time.perf_counter()
        lib.linear_solve_inplace(A, b, ok, N_DIM)
        total += time.perf_counter() - t0
    value = Bs * args.steps / total
    sample = (f"each step = 2^21 of the step's 2^24 systems (bounded sample), AoS, {threads} OpenMP threads; "
              f"{lib.build_info()}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "systems/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": N_DIM, "rhs": 1, "systems_per_gpu_per_step": 1 << LOG2_B,
                       "layout": "AoS (reference objects)"},
            "cpu_baseline": {"value": value, "unit": "systems/s", "cores": threads, "kind": lib.kind, "sample": sample},
            "e2e": {"value": value, "unit": "systems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def side_config(gm, torch, name, fn, nbytes_per_system, B, peak, reps=5):
    """Quick secondary measurement (other BASELINE configs): best-of-`reps` CUDA-event time of one launch."""
    try:
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        rate = B / (best * 1e-3)
        return {"workload": name, "systems_per_s": rate, "ms": best, "alg_bytes_per_system": nbytes_per_system,
                "hbm_frac": rate * nbytes_per_system / (peak * 1e9)}
    except Exception as ex:     # a side config must never take the headline down
        return {"workload": name, "error": repr(ex)[:200]}


_REAL_STDOUT = None


def claim_stdout():
    """Libraries (NCCL's version banner, torchrun children) write to fd 1.  The contract is ONE JSON line on
    stdout, so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-batch", type=int, default=LOG2_B)
    ap.add_argument("--no-sides", action="store_true", help="skip the secondary configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import geomc_b200 as gm
    L = gm.lib()                      # fails loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: geomc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)

    B = 1 << args.log2_batch
    n = N_DIM
    peak, peak_src = read_peak()

    # ---- synthetic inputs, generated on the device directly in the SoA-interleaved container ----
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    A = gm.BatchSoA(B, n * n, torch.float32, dev)
    b = gm.BatchSoA(B, n, torch.float32, dev)
    A.data.copy_(torch.randint(-(1 << 19), 1 << 19, A.data.shape, device=dev, generator=g, dtype=torch.int32).float() * 2.0 ** -17)
    b.data.copy_(torch.randint(-(1 << 19), 1 << 19, b.data.shape, device=dev, generator=g, dtype=torch.int32).float() * 2.0 ** -17)
    x = gm.BatchSoA(B, n, torch.float32, dev)
    ok = torch.zeros(B, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        gm.linear_solve(A, b, n, 1, 0, True, x_out=x, ok_out=ok)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.gmb_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        evs[i + 1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = L.gmb_launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())

    # ---- end-to-end: host AoS buffers (pinned) -> C ABI host call -> host results -------------------
    e2e = None
    try:
        Be = B
        hA = torch.empty((Be, n * n), dtype=torch.float32).pin_memory()
        hb = torch.empty((Be, n), dtype=torch.float32).pin_memory()
        hx = torch.empty((Be, n), dtype=torch.float32).pin_memory()
        hok = torch.empty(Be, dtype=torch.uint8).pin_memory()
        hA.copy_(A.to_aos())
        hb.copy_(b.to_aos())
        torch.cuda.synchronize()
        gm.host_linear_solve(hA, hb, hx, hok, n, device=local_rank)      # warm-up (allocates staging buffers)
        gm.host_linear_solve(hA, hb, hx, hok, n, device=local_rank)
        if world > 1:
            dist.barrier()
        e2e_steps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gm.host_linear_solve(hA, hb, hx, hok, n, device=local_rank)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": Be * world * e2e_steps / float(te.item()), "unit": "systems/s",
               "h2d_bytes_per_step": Be * (64 + 16), "d2h_bytes_per_step": Be * (16 + 1),
               "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
               "api": "gmb_host_plu_solve_f32 (pinned host AoS in, host x/ok out; chunked H2D/kernel/D2H on 3 streams)"}
        # the host path and the resident SoA path must agree bit for bit
        same = bool(torch.equal(hx.cuda(dev).view(torch.int32), x.to_aos().view(torch.int32)))
        e2e["matches_resident_path"] = same
        del hA, hb, hx, hok
    except Exception as ex:
        e2e = {"value": None, "unit": "systems/s", "error": repr(ex)[:300]}

    # ---- verification of the timed run's outputs -----------------------------------------------------
    stats = gm.solve_residual(A, x, b, ok, n)                       # [sum, max, fails, nonfinite]
    csum = gm.checksum(ok, word_offset=rank * (B // 8))
    red = torch.cat([stats[[0, 2, 3]], csum.double() * 0])         # sums
    mx = stats[1:2].clone()
    cs = csum.clone()
    if world > 1:                                                   # the ONLY collective on the path
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    verify = {"mean_residual": float(red[0].item()) / max(1.0, B * world - float(red[1].item())),
              "max_residual": float(mx.item()), "failed_systems": int(red[1].item()),
              "nonfinite_systems": int(red[2].item()), "ok_checksum": int(cs.item()) & ((1 << 64) - 1)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clocks = sampler.stop()
    # oracle check on a sample of the timed run's output (first 2 tiles + last tile)
    try:
        from oracle import pyoracle as po
        W = A.W
        tiles = [0, 1, A.tiles - 1]
        okc = True
        orc = po.best()
        for tl in tiles:
            sub = lambda c, E: gm.BatchSoA(W, E, torch.float32, dev, c.data[tl * E * W:(tl + 1) * E * W]).to_aos().cpu().numpy()
            _, xe, oke = orc.linear_solve(sub(A, 16), sub(b, 4), n)
            xg = sub(x, 4)
            okc &= bool(np.array_equal(xg.view(np.uint32), xe.view(np.uint32)))
            okc &= bool(np.array_equal(ok[tl * W:(tl + 1) * W].cpu().numpy(), oke))
        verify["oracle_sample_bit_exact"] = okc
        verify["oracle"] = orc.kind
    except Exception as ex:
        verify["oracle_sample_bit_exact"] = None
        verify["oracle_error"] = repr(ex)[:200]

    value = B * world * args.steps / (total_ms_max * 1e-3)
    avg_launch_ms = statistics.mean(per_launch_ms)
    achieved = ALG_BYTES * B / (avg_launch_ms * 1e-3) / 1e9
    traffic = read_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if traffic is None else traffic * B, "peak_source": peak_src,
                "kernel": "gmb::k_plu_solve<float,4,1,row_major,SoA>", "alg_bytes_per_system": ALG_BYTES,
                "avg_launch_ms": avg_launch_ms, "min_launch_ms": min(per_launch_ms),
                "frac_of_nominal_8TBs": achieved / 8000.0}

    sides = []
    if not args.no_sides and world == 1:
        f32, f64 = torch.float32, torch.float64
        # C1 from AoS (ingest included)
        Aa, ba = A.to_aos(), b.to_aos()
        xa = torch.empty_like(ba)
        sides.append(side_config(gm, torch, "4x4 f32 PLU solve, AoS in/out (reference objects)",
                                 lambda: gm.linear_solve(Aa, ba, 4, x_out=xa, ok_out=ok), 96, B, peak))
        # C3: 4x4 f32 inverse (AoS SimpleMatrix<float,4,4> and SoA)
        inv_out = torch.empty_like(Aa)
        sides.append(side_config(gm, torch, "C3: 4x4 f32 inverse, AoS", lambda: gm.inv(Aa, 4, out=inv_out, ok_out=ok), 128, B, peak))
        inv_soa = A.empty_like()
        sides.append(side_config(gm, torch, "C3: 4x4 f32 inverse, SoA", lambda: gm.inv(A, 4, out=inv_soa, ok_out=ok), 128, B, peak))
        del Aa, ba, xa, inv_out, inv_soa
        # C2: 3x3 f64 SPD cholesky + solve, 2^24 systems
        R = torch.randint(-128, 128, (B, 3, 3), device=dev, generator=g).double() * 2.0 ** -5
        S = (R @ R.transpose(1, 2) + torch.eye(3, device=dev, dtype=f64)).reshape(B, 9).contiguous()
        bs = torch.randint(-(1 << 19), 1 << 19, (B, 3), device=dev, generator=g).double() * 2.0 ** -17
        Ss, bss = gm.BatchSoA.from_aos(S), gm.BatchSoA.from_aos(bs)
        xs = bss.empty_like()
        # algorithmic bytes: the factorization reads only the LOWER triangle (Cholesky.h:42-44), so the SoA
        # kernel never touches the 3 upper-triangle planes: 48 + 24 read, 24 written = 96 B (the survey's
        # 120 B counts the full 72-byte matrix, which the AoS path does pull in sector-wise).
        sides.append(side_config(gm, torch, "C2: 3x3 f64 SPD cholesky + solve, SoA (96 B/system: lower triangle only)",
                                 lambda: gm.cholesky_solve(Ss, bss, 3, x_out=xs, ok_out=ok), 96, B, peak))
        xs_a = torch.empty_like(bs)
        sides.append(side_config(gm, torch, "C2: 3x3 f64 SPD cholesky + solve, AoS (120 B/system: whole records)",
                                 lambda: gm.cholesky_solve(S, bs, 3, x_out=xs_a, ok_out=ok), 120, B, peak))
        del R, S, bs, Ss, bss, xs, xs_a
        # C4: f64 PLU decompose N = 2..8 (one homogeneous sub-batch per N; working set >> L2 for every N)
        for nn in range(2, 9):
            Bc = 1 << (23 if nn <= 4 else 22)
            M = gm.BatchSoA(Bc, nn * nn, f64, dev)
            M.data.copy_(torch.randint(-(1 << 19), 1 << 19, M.data.shape, device=dev, generator=g).double() * 2.0 ** -17)
            outs = (torch.empty((Bc, nn), dtype=torch.uint8, device=dev), torch.empty(Bc, dtype=torch.uint8, device=dev),
                    torch.empty(Bc, dtype=torch.uint8, device=dev))
            sides.append(side_config(gm, torch, f"C4: {nn}x{nn} f64 PLU decompose (in place), SoA, 2^{23 if nn <= 4 else 22} systems",
                                     lambda: gm.decomp_plu(M, nn, inplace=True, out=outs), 16 * nn * nn + nn + 2, Bc, peak, reps=3))
            del M, outs
        # row (f)1: orthogonal<float,3> (cross product) and <float,4> (3x4 PLU + back-substitution), AoS Vec arrays
        for nn in (3, 4):
            Vn = torch.randint(-(1 << 19), 1 << 19, (B, (nn - 1) * nn), device=dev, generator=g).float() * 2.0 ** -17
            sides.append(side_config(gm, torch, f"(f)1: orthogonal<float,{nn}>, AoS Vec arrays", lambda: gm.orthogonal(Vn, nn),
                                     4 * ((nn - 1) * nn + nn), B, peak, reps=3))
            del Vn
        # C5: 16x16 f32 PLU solve, 2^21 systems
        B5 = 1 << 21
        A16 = torch.randint(-(1 << 19), 1 << 19, (B5, 256), device=dev, generator=g).float() * 2.0 ** -17
        b16 = torch.randint(-(1 << 19), 1 << 19, (B5, 16), device=dev, generator=g).float() * 2.0 ** -17
        x16 = torch.empty_like(b16)
        ok16 = torch.zeros(B5, dtype=torch.uint8, device=dev)
        sides.append(side_config(gm, torch, "C5: 16x16 f32 PLU solve, AoS",
                                 lambda: gm.linear_solve(A16, b16, 16, x_out=x16, ok_out=ok16), 1152, B5, peak, reps=3))
        del A16, b16, x16

    cpu = None
    if not args.no_cpu and world == 1:
        try:
            cpu = cpu_reference_rate()
        except Exception as ex:
            cpu = {"value": None, "unit": "systems/s", "cores": None, "kind": None, "sample": "failed: " + repr(ex)[:200]}

    line = {"metric": METRIC, "value": value, "unit": "systems/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": n, "rhs": 1, "systems_per_gpu_per_step": B, "layout": "SoA-interleaved (resident in HBM)",
                       "l2": "inputs+outputs per step = 1.5 GiB per GPU >> 126 MB L2 (no flush needed)",
                       "sharding": f"contiguous, {world} rank(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "verify": verify, "other_configs": sides}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
