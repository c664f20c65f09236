"""world_size-2 gloo tests of the viewpoint sharding + gather logic (runs on CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quad3dr_b200 import dist as qdist


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, 10001):
        for world in (1, 2, 3, 8):
            parts = [qdist.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
                assert b == c
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _fake_result(lo, hi, seed=1234):
    """Deterministic per-viewpoint lists as a function of the global viewpoint index."""
    offs, leaf, info, tot = [0], [], [], []
    for v in range(lo, hi):
        rng = np.random.default_rng(seed + v)
        k = int(rng.integers(0, 40))
        l = np.sort(rng.choice(5000, size=k, replace=False)).astype(np.int32)
        i = rng.random(k).astype(np.float32)
        leaf.append(l)
        info.append(i)
        offs.append(offs[-1] + k)
        tot.append(np.float32(i.astype(np.float64).sum()))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs and sum(len(x) for x in xs) else np.empty(0, dt)
    return (np.asarray(offs, np.int64), np.asarray(tot, np.float32), cat(leaf, np.int32), cat(info, np.float32))


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = qdist.shard_range(n, rank, world)
        n_max = max(qdist.shard_range(n, r, world)[1] - qdist.shard_range(n, r, world)[0] for r in range(world))
        o, t, l, i = [torch.from_numpy(x) for x in _fake_result(lo, hi)]
        O_, T_, L_, I_ = qdist.gather_results(o, t, l, i, n_max)
        eo, et, el, ei = _fake_result(0, n)
        ok = (np.array_equal(O_.numpy(), eo) and np.array_equal(T_.numpy().view(np.uint32), et.view(np.uint32))
              and np.array_equal(L_.numpy(), el) and np.array_equal(I_.numpy().view(np.uint32), ei.view(np.uint32)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n", [11, 2, 1])
def test_gather_results_world2_gloo(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(world))
    assert got == {0: True, 1: True}


# ---- the torch.distributed plumbing around the C ABI's exchange (quad3dr_b200/dist.py PeerExchange) at world_size 2 ----

def _plumbing_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    try:
        # IPC handles travel as opaque 128-byte blobs, rank order preserved
        blob = bytes([rank]) * 128
        got = qdist.all_gather_bytes(blob)
        ok &= got == [bytes([r]) * 128 for r in range(world)]
        # a rank that failed locally ships an empty blob; everybody sees it and everybody agrees on the outcome
        got = qdist.all_gather_bytes(b"" if rank == 1 else blob)
        ok &= [len(g) for g in got] == [128 if r != 1 else 0 for r in range(world)]
        ok &= qdist.agree(True) is True
        ok &= qdist.agree(rank != 1) is False
        # gather_results refuses a slice larger than n_local_max on ALL ranks (no hang, no corruption)
        lo, hi = qdist.shard_range(7, rank, world)
        o, t, l, i = [torch.from_numpy(x) for x in _fake_result(lo, hi)]
        try:
            qdist.gather_results(o, t, l, i, 3)  # rank 0 holds 4 viewpoints
            ok = False
        except ValueError:
            pass
        try:
            qdist.gather_results(o, t, l, i, 4 + rank)  # n_local_max differs between ranks
            ok = False
        except ValueError:
            pass
        O_, _, _, _ = qdist.gather_results(o, t, l, i, 4)  # and works right after
        ok &= int(O_.numel()) == 8
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_exchange_plumbing_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plumbing_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(world))
    assert got == {0: True, 1: True}
