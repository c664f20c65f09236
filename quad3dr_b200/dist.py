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
    # n_local_max must be the same on every rank and cover every rank's slice; checked collectively BEFORE the first
    # sized collective, so that a bad argument raises on all ranks instead of hanging or corrupting some
    chk = torch.tensor([n_local_max, -n_local_max, n_local], dtype=torch.int64, device=dev)
    dist.all_reduce(chk, op=dist.ReduceOp.MAX, group=group)
    if int(chk[0]) != -int(chk[1]):
        raise ValueError("gather_results: n_local_max differs between ranks")
    if int(chk[2]) > n_local_max:
        raise ValueError(f"gather_results: a rank holds {int(chk[2])} viewpoints, more than n_local_max = {n_local_max}")
    if offsets.numel() != n_local + 1 or info.numel() != e_local:
        raise ValueError("gather_results: inconsistent CSR arrays")
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


def all_gather_bytes(blob: bytes, group=None) -> List[bytes]:
    """Every rank's blob on every rank, in rank order (object collective: works on gloo and nccl groups alike)."""
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, bytes(blob), group=group)
    return [bytes(b) for b in out]


def agree(ok: bool, device=None, group=None) -> bool:
    """True iff every rank passes True; all ranks get the same answer (no one-sided failure may desynchronise the
    collectives that follow)."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


class PeerExchange:
    """The result exchange of SURVEY.md section 8e for one process per GPU: torch.distributed plumbing around the C
    ABI's q3dr_exchange (quad3dr_b200/csrc/q3dr_exchange.cu), which does the work.

    Instead of one all_gather after the last kernel, every finished chunk of a scoring call is pushed into all ranks'
    HBM by the copy engines -- cudaMemcpyAsync into CUDA-IPC-mapped peer buffers over NVLink / NVSwitch, on a side
    stream, issued by the scoring call itself -- while the SMs cast the next chunk; completion travels as flags behind
    the data, and end_step waits for them on the device.  torch.distributed is used ONCE, to ship the IPC handles (and
    for the agreement / teardown barriers); there is no collective in a step.

    Every rank ends a step holding the results of ALL ranks: views() gives zero-copy tensors per source rank."""

    def __init__(self, gmap, device, n_local_max: int, cap_entries: int, group=None):
        from . import engine as E

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.ex = None
        failure = None
        handles = b""
        try:
            self.ex = E.Exchange(gmap, self.world, self.rank, int(n_local_max), int(cap_entries))
            handles = self.ex.export_handles()
        except Exception as e:  # e.g. CUDA IPC not permitted in this container
            failure = e
        gathered = all_gather_bytes(handles, group)
        if failure is None:
            if all(len(h) == 2 * E.IPC_HANDLE_BYTES for h in gathered):
                try:
                    self.ex.import_handles(gathered)
                except Exception as e:
                    failure = e
            else:
                failure = RuntimeError("another rank could not export its receive buffer")
        if not agree(failure is None, self.device, group):
            if self.ex is not None:
                self.ex.close()
                self.ex = None
            raise RuntimeError(f"PeerExchange unavailable: {failure or 'failed on another rank'}")
        dist.barrier(group=group)  # every rank has mapped every buffer before anybody pushes

    def begin_step(self):
        self.ex.begin_step()

    def end_step(self, local_status: int = 0, stream: int = 0):
        """raises on ALL ranks if any rank failed in this step (the failure flag travels with the completion flags)"""
        self.ex.end_step(local_status, stream)

    def views(self):
        """per source rank: (offsets int64 (n+1,), total f32 (n,), leaf i32 (e,), info f32 (e,)) -- zero-copy views of
        this rank's receive buffer of the step that just ended (valid until this rank's next begin_step)"""
        out = []
        for r in range(self.world):
            v = self.ex.view(r)
            n, e = v.n_viewpoints, v.n_entries

            def t(ptr, cnt, typestr):
                if cnt == 0:
                    return torch.empty(0, dtype={"<i8": torch.int64, "<f4": torch.float32, "<i4": torch.int32}[typestr],
                                       device=self.device)
                return torch.as_tensor(CudaArrayView(ptr, cnt, typestr), device=self.device)

            off = t(v.offsets, n + 1, "<i8") if n else torch.zeros(1, dtype=torch.int64, device=self.device)
            out.append((off, t(v.total, n, "<f4"), t(v.leaf, e, "<i4"), t(v.information, e, "<f4")))
        return out

    def close(self):
        if self.ex is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody unmaps while somebody may still push
        self.ex.disconnect()
        dist.barrier(group=self.group)  # nobody frees while somebody still has it mapped
        self.ex.close()
        self.ex = None
