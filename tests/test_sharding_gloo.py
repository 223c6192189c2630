"""CPU suite: the N > 1 host logic (contiguous sharding + the final scalar reduction) under
torch.distributed with the gloo backend, world_size 2 -- the same code path bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from geomc_b200.sharding import MASK64, checksum_word_offset, reduce_stats, shard_range


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 127, 128, 129, 1000, 1 << 20, (1 << 20) + 77):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 64, 128):
                ranges = [shard_range(B, r, world, align) for r in range(world)]
                assert ranges[0][0] == 0 and ranges[-1][1] == B
                for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
                    assert e0 == b1 and b0 <= e0
                for b, e in ranges[:-1]:
                    assert (b % align == 0 or b == B) and (e % align == 0 or e == B)
                sizes = [e - b for b, e in ranges]
                assert max(sizes) - min(sizes) <= 2 * align  # one unit of imbalance + the ragged tail
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geomc_b200.linalg import checksum_host
        from oracle import pyoracle as po
        rng = np.random.default_rng(5)                    # every rank generates the same whole batch
        A = po.exact_grid(rng, (B, 16), np.float32)
        A[::97] = 0
        b = po.exact_grid(rng, (B, 4), np.float32)
        lo, hi = shard_range(B, rank, world, align=128)
        _, x, ok = po.port().linear_solve(A[lo:hi], b[lo:hi], 4)      # stand-in for the GPU shard
        res = np.abs(np.einsum("sij,sj->si", A[lo:hi].reshape(-1, 4, 4).astype(np.float64), x.astype(np.float64)) - b[lo:hi])
        res = res.max(axis=1)[ok == 1]
        # shard checksum with GLOBAL word offsets == what gmb_checksum_bytes(word_offset=...) computes
        full = np.zeros(B, np.uint8)
        full[lo:hi] = ok
        woff = checksum_word_offset(lo, 1)
        words = np.frombuffer(full[lo:hi].tobytes() + b"\0" * ((-len(full[lo:hi])) % 8), dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = words + np.uint64(0x9E3779B97F4A7C15) * (np.arange(len(words), dtype=np.uint64) + np.uint64(woff + 1))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            local_sum = int(z.sum(dtype=np.uint64))
        out = reduce_stats({"residual_sum": float(res.sum()), "residual_max": float(res.max()),
                            "failed": float((ok == 0).sum()), "checksum": local_sum})
        if rank == 0:
            _, xw, okw = po.port().linear_solve(A, b, 4)
            resw = np.abs(np.einsum("sij,sj->si", A.reshape(-1, 4, 4).astype(np.float64), xw.astype(np.float64)) - b).max(axis=1)[okw == 1]
            q.put({"got": out, "want": {"residual_sum": float(resw.sum()), "residual_max": float(resw.max()),
                                        "failed": float((okw == 0).sum()), "checksum": checksum_host(okw) & MASK64}})
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_matches_single_rank():
    world, B = 2, 128 * 37            # shard boundary on a tile boundary, 8-byte aligned for the checksum
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got, want = res["got"], res["want"]
    assert got["failed"] == want["failed"] and got["failed"] > 0
    assert got["residual_max"] == want["residual_max"]
    assert abs(got["residual_sum"] - want["residual_sum"]) <= 1e-9 * max(1.0, want["residual_sum"])
    assert got["checksum"] == want["checksum"], "checksum of checksums must equal the checksum of the whole"
