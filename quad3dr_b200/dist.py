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
