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


# ---- PeerGather (the overlapped exchange) at world_size 2 on gloo: same logic, shared memory instead of CUDA IPC ----

class ShmTransport:
    """Stand-in for quad3dr_b200.dist.CudaIpcTransport on a machine without GPUs: receive buffers are POSIX shared
    memory segments ("IPC handle" = the segment name), a copy is a memmove, ordering is immediate, the barrier is
    gloo's.  Everything above the transport -- slot layout, per-chunk pushes from the helper thread, header push, double
    buffering, views -- is PeerGather's own code."""

    def __init__(self):
        import ctypes
        from multiprocessing import shared_memory

        self.ctypes = ctypes
        self.shared_memory = shared_memory
        self.segs = {}

    def _map(self, shm):
        anchor = self.ctypes.c_char.from_buffer(shm.buf)
        ptr = self.ctypes.addressof(anchor)
        self.segs[ptr] = [shm, anchor]
        return ptr

    def alloc(self, nbytes):
        shm = self.shared_memory.SharedMemory(create=True, size=nbytes)
        return self._map(shm), shm.name.encode()

    def open(self, handle):
        return self._map(self.shared_memory.SharedMemory(name=handle.decode()))

    def close(self, ptr):
        entry = self.segs.pop(ptr)
        shm = entry[0]
        entry.clear()  # drops the anchor: a segment with an exported buffer cannot be closed
        shm.close()

    def free(self, ptr):
        entry = self.segs.pop(ptr)
        shm = entry[0]
        entry.clear()
        shm.close()
        shm.unlink()

    def agree(self, ok):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def thread_init(self):
        pass

    def mark(self, stream):
        return None

    def after(self, mark):
        pass

    def copy(self, dst, src, nbytes):
        self.ctypes.memmove(dst, src, nbytes)

    def barrier(self, stream):
        dist.barrier()

    def view(self, ptr, n, typestr):
        dt = np.dtype(typestr)
        if n == 0:
            return torch.empty(0, dtype=torch.from_numpy(np.empty(0, dt)).dtype)
        raw = (self.ctypes.c_char * (n * dt.itemsize)).from_address(ptr)
        return torch.from_numpy(np.frombuffer(raw, dtype=dt, count=n).copy())

    def sync(self):
        pass


def _peer_worker(rank, world, port, n, q):
    from types import SimpleNamespace

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    try:
        lo, hi = qdist.shard_range(n, rank, world)
        n_per = [qdist.shard_range(n, r, world)[1] - qdist.shard_range(n, r, world)[0] for r in range(world)]
        pg = qdist.PeerGather(None, max(n_per), 40 * max(n_per) + 8, transport=ShmTransport())
        for step in range(4):
            seed = 1000 * step
            drop = 1 if step == 2 and hi - lo > 0 else 0            # a smaller result in the middle
            o, t, l, i = _fake_result(lo, hi - drop, seed)
            res = SimpleNamespace(offsets=o.ctypes.data, total=t.ctypes.data if t.size else 0,
                                  leaf=l.ctypes.data if l.size else 0, information=i.ctypes.data if i.size else 0)
            pg.begin_step()
            n_local = hi - lo - drop
            for v0 in range(0, n_local, 3):                          # the engine reports chunk after chunk
                nv = min(3, n_local - v0)
                pg.on_chunk(v0, nv, int(o[v0]), int(o[v0 + nv] - o[v0]), res, None)
            pg.end_step(n_local, None)
            n_use = [n_per[r] - (1 if step == 2 and n_per[r] > 0 else 0) for r in range(world)]
            got = pg.views(n_use)
            for r in range(world):
                rlo, rhi = qdist.shard_range(n, r, world)
                eo, et, el, ei = _fake_result(rlo, rlo + n_use[r], seed)
                go, gt, gl, gi = got[r]
                ok &= bool(np.array_equal(go.numpy(), eo) and np.array_equal(gt.numpy().view(np.uint32), et.view(np.uint32))
                           and np.array_equal(gl.numpy(), el) and np.array_equal(gi.numpy().view(np.uint32), ei.view(np.uint32)))
        pg.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 2, 1])
def test_peer_gather_protocol_world2_gloo(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(world))
    assert got == {0: True, 1: True}
