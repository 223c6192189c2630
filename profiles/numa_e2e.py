"""Where does the 8-GPU end-to-end number go?  Host-side bandwidth experiment, one rank per GPU (torchrun):
pinned-copy rates (H2D only, D2H only, both at once) and the host entry point gmb_host_plu_solve_f32, all ranks at once,
(1) with the default CPU affinity and (2) with every rank bound to the CPUs next to its GPU before it allocates its pinned
buffers (gmb_bind_thread_near_device: NUMA-local buffers).     torchrun --nproc-per-node N profiles/numa_e2e.py"""
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomc_b200 as gm  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L = gm.lib()


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            keep = [ln for ln in out.splitlines() if cmd[0] != "lscpu" or any(k in ln for k in ("NUMA", "Socket", "Model name", "CPU(s):"))]
            print("\n".join(keep), flush=True)
        except Exception as ex:
            print("(no", cmd[0], ex, ")")
n, B = 4, 1 << 24
g = torch.Generator().manual_seed(1 + rank)
src = torch.randint(-(1 << 19), 1 << 19, (1 << 20, 16), generator=g).float() * 2.0 ** -17
for bound in (False, True):
    if bound:
        rc = L.gmb_bind_thread_near_device(local)
        cpus = sorted(os.sched_getaffinity(0))
        print(f"rank {rank}: bound={rc} cpus {cpus[0]}..{cpus[-1]} ({len(cpus)})", flush=True)
    hA = torch.empty((B, 16), dtype=torch.float32).pin_memory()
    hb = torch.empty((B, 4), dtype=torch.float32).pin_memory()
    hx = torch.empty((B, 4), dtype=torch.float32).pin_memory()
    hok = torch.empty(B, dtype=torch.uint8).pin_memory()
    for i in range(0, B, 1 << 20):
        hA[i:i + (1 << 20)] = src
        hb[i:i + (1 << 20)] = src[:, :4]
    dA = torch.empty((B, 16), dtype=torch.float32, device=dev)
    dX = torch.empty((B, 4), dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        dA.copy_(hA, non_blocking=True)
    torch.cuda.synchronize()
    res["h2d_GBps_per_gpu"] = 3 * hA.numel() * 4 / allmax(time.perf_counter() - t0) / 1e9
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        hA.copy_(dA, non_blocking=True)
    torch.cuda.synchronize()
    res["d2h_GBps_per_gpu"] = 3 * hA.numel() * 4 / allmax(time.perf_counter() - t0) / 1e9
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            dA.copy_(hA, non_blocking=True)
        with torch.cuda.stream(s2):
            hx.copy_(dX, non_blocking=True)
            hb.copy_(dX, non_blocking=True)
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0)
    res["both_h2d_GBps_per_gpu"] = 3 * hA.numel() * 4 / dt / 1e9
    res["both_d2h_GBps_per_gpu"] = 3 * 2 * hx.numel() * 4 / dt / 1e9
    gm.host_linear_solve(hA, hb, hx, hok, n, device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        gm.host_linear_solve(hA, hb, hx, hok, n, device=local)
    dt = allmax(time.perf_counter() - t0)
    res["e2e_systems_per_s"] = B * world * 5 / dt
    res["e2e_ms_per_step"] = dt / 5 * 1e3
    if rank == 0:
        print(f"{world} rank(s), NUMA-local pinned buffers: {bound} -> " + ", ".join(f"{k} {v:.4g}" for k, v in res.items()), flush=True)
    del hA, hb, hx, hok, dA, dX
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()
    except Exception:
        pass
if world > 1:
    dist.destroy_process_group()
