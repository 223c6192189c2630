"""Multi-GPU plumbing: contiguous sharding of independent systems and the final scalar reduction.

Systems are independent, so the data path has no collective: rank r of G owns the contiguous
range shard_range(B, r, G) in its own HBM.  After the kernels each rank holds a handful of
scalars (residual sum / max, failure counts, an additive checksum); reduce_stats() is the only
collective in the system (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

MASK64 = (1 << 64) - 1


def shard_range(B: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous shard of B systems.

    Shard boundaries are multiples of `align` (use the SoA tile width so that a shard is a whole
    number of tiles and therefore a contiguous byte range of the container); the last rank takes
    the ragged remainder.  Every system belongs to exactly one rank.
    """
    if world < 1 or not (0 <= rank < world) or B < 0 or align < 1:
        raise ValueError("bad shard request")
    units = (B + align - 1) // align
    per, extra = divmod(units, world)
    b_units = rank * per + min(rank, extra)
    e_units = b_units + per + (1 if rank < extra else 0)
    return min(b_units * align, B), min(e_units * align, B)


def checksum_word_offset(begin: int, bytes_per_system: int) -> int:
    """Global 8-byte-word offset of a shard that starts at system `begin` (must be 8-byte aligned)."""
    off = begin * bytes_per_system
    if off % 8:
        raise ValueError("shard must start on an 8-byte boundary of the whole buffer")
    return off // 8


def reduce_stats(local: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """All-reduce the per-shard scalars: keys ending in `_max` use MAX, `checksum` wraps mod 2^64
    (carried as two 32-bit halves so that no float rounding is involved), everything else is summed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = dict(local)
        if "checksum" in out:
            out["checksum"] = int(out["checksum"]) & MASK64
        return out
    keys = sorted(k for k in local if k != "checksum")
    sums = torch.tensor([float(local[k]) for k in keys if not k.endswith("_max")], dtype=torch.float64, device=device)
    maxs = torch.tensor([float(local[k]) for k in keys if k.endswith("_max")], dtype=torch.float64, device=device)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.endswith("_max"):
            out[k] = float(maxs[mi]); mi += 1
        else:
            out[k] = float(sums[si]); si += 1
    if "checksum" in local:
        c = int(local["checksum"]) & MASK64
        halves = torch.tensor([c & 0xFFFFFFFF, c >> 32], dtype=torch.int64, device=device)
        dist.all_reduce(halves, op=dist.ReduceOp.SUM, group=group)
        out["checksum"] = (int(halves[0]) + (int(halves[1]) << 32)) & MASK64
    return out
