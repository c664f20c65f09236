"""Viewpoint sharding across ranks and the result gather.

The path shards by viewpoint: the map is replicated, rank g scores the contiguous slice
[g*n/G, (g+1)*n/G) and nothing is exchanged during compute.  Afterwards the per-viewpoint records are
gathered with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests): one all_gather of the
fixed-size per-viewpoint header (offsets, totals, entry count) and one all_gather of the hit lists padded to the
largest rank.  Tensors stay on the device they were given on.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of n viewpoints owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_results(offsets: torch.Tensor, total: torch.Tensor, leaf: torch.Tensor, info: torch.Tensor,
                   n_local_max: int, group=None):
    """All-gathers the CSR results of every rank.

    offsets: (n_local + 1,) int64, total: (n_local,) float32, leaf: (e,) int32, info: (e,) float32, all on one device.
    n_local_max: the largest per-rank viewpoint count (same on every rank; headers are padded to it).
    Returns (offsets_all int64 (N+1,), total_all float32 (N,), leaf_all int32 (E,), info_all float32 (E,)) in
    global viewpoint order, on the same device.
    """
    world = dist.get_world_size(group)
    dev = offsets.device
    n_local = total.numel()
    e_local = int(leaf.numel())
    # header: [n_local, e_local, counts (n_local_max), totals bits (n_local_max)] as int64
    head = torch.zeros(2 + 2 * n_local_max, dtype=torch.int64, device=dev)
    head[0] = n_local
    head[1] = e_local
    head[2:2 + n_local] = offsets[1:] - offsets[:-1]
    head[2 + n_local_max:2 + n_local_max + n_local] = total.view(torch.int32).to(torch.int64)
    heads = torch.empty(world * head.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(heads, head, group=group)
    heads = heads.view(world, -1)
    heads_cpu = heads.cpu()
    n_per = heads_cpu[:, 0].tolist()
    e_per = heads_cpu[:, 1].tolist()
    e_max = max(max(e_per), 1)
    send = torch.zeros(2 * e_max, dtype=torch.int32, device=dev)
    send[:e_local] = leaf
    send[e_max:e_max + e_local] = info.view(torch.int32)
    recv = torch.empty(world * 2 * e_max, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, 2, e_max)
    counts = torch.cat([heads[r, 2:2 + n_per[r]] for r in range(world)])
    totals = torch.cat([heads[r, 2 + n_local_max:2 + n_local_max + n_per[r]].to(torch.int32).view(torch.float32)
                        for r in range(world)])
    offsets_all = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=dev)
    offsets_all[1:] = torch.cumsum(counts, 0)
    leaf_all = torch.cat([recv[r, 0, :e_per[r]] for r in range(world)])
    info_all = torch.cat([recv[r, 1, :e_per[r]].view(torch.float32) for r in range(world)])
    return offsets_all, totals, leaf_all, info_all


class CudaArrayView:
    """Zero-copy torch view of a raw device pointer owned by the C library (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_result_tensors(res, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """torch views (no copy) of a quad3dr_b200.engine.DeviceResult."""
    nv, ne = res.n_viewpoints, res.n_entries
    offsets = torch.as_tensor(CudaArrayView(res.offsets, nv + 1, "<i8"), device=device)
    total = torch.as_tensor(CudaArrayView(res.total, max(nv, 1), "<f4"), device=device)[:nv]
    if ne:
        leaf = torch.as_tensor(CudaArrayView(res.leaf, ne, "<i4"), device=device)
        info = torch.as_tensor(CudaArrayView(res.information, ne, "<f4"), device=device)
    else:
        leaf = torch.empty(0, dtype=torch.int32, device=device)
        info = torch.empty(0, dtype=torch.float32, device=device)
    return offsets, total, leaf, info


class CudaIpcTransport:
    """What PeerGather needs from the machine: buffers other ranks can write into, asynchronous copies into them and
    ordering.  This one is the product's: cudaMalloc + CUDA IPC handles, cudaMemcpyAsync on a side stream (copy engines
    over NVLink / NVSwitch), CUDA events, one NCCL all_reduce as the barrier.  (tests/test_dist_cpu.py drives the same
    PeerGather logic at world_size 2 on gloo with a shared-memory stand-in.)"""

    def __init__(self, device: torch.device, group=None):
        from cuda.bindings import runtime as rt

        self.rt = rt
        self.device = torch.device(device)
        self.group = group
        torch.cuda.set_device(self.device)
        self.side = torch.cuda.Stream(device=self.device)
        self.token = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _ok(self, ret):
        if int(ret[0]) != 0:
            raise RuntimeError(f"CUDA runtime error {ret[0]} in PeerGather")
        return ret[1] if len(ret) == 2 else ret[1:]

    def alloc(self, nbytes: int) -> Tuple[int, bytes]:
        ptr = self._ok(self.rt.cudaMalloc(nbytes))
        try:
            h = self._ok(self.rt.cudaIpcGetMemHandle(ptr))
        except Exception:
            self.rt.cudaFree(ptr)
            raise
        return int(ptr), bytes(h.reserved)

    def open(self, handle: bytes) -> int:
        h = self.rt.cudaIpcMemHandle_t()
        h.reserved = handle
        return int(self._ok(self.rt.cudaIpcOpenMemHandle(h, self.rt.cudaIpcMemLazyEnablePeerAccess)))

    def close(self, ptr: int):
        self.rt.cudaIpcCloseMemHandle(ptr)

    def free(self, ptr: int):
        self.rt.cudaFree(ptr)

    def agree(self, ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    def thread_init(self):
        self.rt.cudaSetDevice(self.device.index)

    def mark(self, stream):
        """a point in `stream` (the scoring call's stream) that later copies can be ordered after"""
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def after(self, mark):
        self.side.wait_event(mark)

    def copy(self, dst: int, src: int, nbytes: int):
        (err,) = self.rt.cudaMemcpyAsync(dst, src, nbytes, self.rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice,
                                         self.side.cuda_stream)
        if int(err) != 0:
            raise RuntimeError(f"cudaMemcpyAsync: {err}")

    def barrier(self, stream):
        """every rank's copies issued so far have landed everywhere; `stream` continues after that"""
        with torch.cuda.stream(self.side):
            dist.all_reduce(self.token, group=self.group)
        stream.wait_stream(self.side)

    def view(self, ptr: int, n: int, typestr: str) -> torch.Tensor:
        return torch.as_tensor(CudaArrayView(ptr, max(n, 1), typestr), device=self.device)[:n]

    def sync(self):
        torch.cuda.synchronize(self.device)


class PeerGather:
    """The result exchange of SURVEY.md section 8e, overlapped with the compute.

    Instead of one all_gather after the last kernel, every finished chunk of a scoring call (q3dr_map_set_chunk_callback)
    is PUSHED into all ranks' HBM by the copy engines -- cudaMemcpyAsync into CUDA-IPC-mapped peer buffers over
    NVLink / NVSwitch, on a side stream, issued by a helper thread -- while the SMs cast the next chunk.  The pushes use
    no SM, so they do not compete with the issue-bound raycast kernel; only the last chunk's push and one tiny NCCL
    all_reduce (the barrier that tells every rank that all pushes into its memory have landed) stay exposed.

    Every rank ends a step with the results of ALL ranks in its own receive buffer, one slot per source rank:
        offsets int64 [n_max + 1] | total f32 [n_max] | leaf i32 [cap] | information f32 [cap]
    Two receive buffers alternate between steps, so a fast rank's pushes of step s+1 never touch what a slow rank
    still reads from step s.  Memory: 2 * world * (8 * cap + 12 * n_max) bytes per GPU, sized for the worst case.
    """

    def __init__(self, device, n_local_max: int, cap_entries: int, group=None, transport=None):
        import queue
        import threading

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_max = int(n_local_max)
        self.cap = int(cap_entries)
        self.tr = transport if transport is not None else CudaIpcTransport(device, group)

        def up(x, a=256):
            return (x + a - 1) // a * a

        self.tot_off = up(8 * (self.n_max + 1))
        self.leaf_off = up(self.tot_off + 4 * self.n_max)
        self.info_off = up(self.leaf_off + 4 * self.cap)
        self.slot_bytes = up(self.info_off + 4 * self.cap)
        self.base: List[int] = []
        self.peer = [[0] * self.world for _ in range(2)]
        handles = []
        failure = None
        try:
            for _ in range(2):
                ptr, handle = self.tr.alloc(self.world * self.slot_bytes)
                self.base.append(ptr)
                handles.append(handle)
        except Exception as e:
            failure = e
            handles = [b"", b""]
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        if failure is None and all(len(g[0]) and len(g[1]) for g in gathered):
            try:
                for b in range(2):
                    for p in range(self.world):
                        self.peer[b][p] = self.base[b] if p == self.rank else self.tr.open(gathered[p][b])
            except Exception as e:
                failure = e
        elif failure is None:
            failure = RuntimeError("another rank could not export its receive buffer")
        # all ranks agree on success before anybody relies on it (a one-sided failure must not desynchronise collectives)
        if not self.tr.agree(failure is None):
            for ptr in self.base:
                self.tr.free(ptr)
            self.base = []
            raise RuntimeError(f"PeerGather unavailable: {failure or 'failed on another rank'}")
        self.cur = 1
        self.error = None
        self._last = None
        self._q: "queue.Queue" = queue.Queue()
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()
        dist.barrier(group=group)  # every rank has mapped every buffer before anybody pushes

    def _push(self, dst_off: int, src_ptr: int, nbytes: int):
        """the same bytes into this rank's slot on every rank"""
        if nbytes <= 0:
            return
        for k in range(self.world):
            p = (self.rank + k) % self.world  # start with different targets on different ranks
            self.tr.copy(self.peer[self.cur][p] + self.rank * self.slot_bytes + dst_off, src_ptr, nbytes)

    def _run(self):
        self.tr.thread_init()
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                mark, first_entry, n_entries, leaf_ptr, info_ptr = job
                self.tr.after(mark)
                self._push(self.leaf_off + 4 * first_entry, leaf_ptr + 4 * first_entry, 4 * n_entries)
                self._push(self.info_off + 4 * first_entry, info_ptr + 4 * first_entry, 4 * n_entries)
            except Exception as e:  # surfaced by end_step
                self.error = e
            finally:
                self._q.task_done()

    def begin_step(self):
        self.cur ^= 1
        self._last = None

    def on_chunk(self, first_vp: int, n_vp: int, first_entry: int, n_entries: int, res, stream):
        """chunk callback: order the push after the work issued so far, hand it to the helper thread, return at once
        (the calling thread goes on to launch the next chunk's raycast)"""
        if first_entry + n_entries > self.cap or first_vp + n_vp > self.n_max:
            self.error = RuntimeError("PeerGather: result larger than the receive slot")
            return
        self._last = res
        self._q.put((self.tr.mark(stream), first_entry, n_entries, int(res.leaf or 0), int(res.information or 0)))

    def end_step(self, n_local: int, stream):
        """headers of all local viewpoints, then the barrier; `stream` continues when every rank's pushes have landed"""
        self._q.join()
        if self.error is not None:
            raise self.error
        if self._last is not None and n_local > 0:
            self.tr.after(self.tr.mark(stream))
            self._push(0, int(self._last.offsets), 8 * (n_local + 1))
            self._push(self.tot_off, int(self._last.total), 4 * n_local)
        self.tr.barrier(stream)

    def views(self, n_per_rank: List[int]):
        """per source rank: (offsets int64 (n+1,), total f32 (n,), leaf i32 (e,), info f32 (e,)) -- zero-copy views of
        this rank's receive buffer of the step that just ended (valid until the step after the next one begins)"""
        out = []
        for r in range(self.world):
            slot = self.base[self.cur] + r * self.slot_bytes
            n = int(n_per_rank[r])
            off = self.tr.view(slot, n + 1, "<i8")
            tot = self.tr.view(slot + self.tot_off, n, "<f4")
            e = int(off[n].item()) if n else 0
            leaf = self.tr.view(slot + self.leaf_off, e, "<i4")
            info = self.tr.view(slot + self.info_off, e, "<f4")
            out.append((off, tot, leaf, info))
        return out

    def close(self):
        self._q.put(None)
        self._worker.join(timeout=5)
        self.tr.sync()
        dist.barrier(group=self.group)  # nobody unmaps while somebody may still push
        for b in range(2):
            for p in range(self.world):
                if p != self.rank and self.peer[b][p]:
                    self.tr.close(self.peer[b][p])
        dist.barrier(group=self.group)
        for ptr in self.base:
            self.tr.free(ptr)
        self.base = []
