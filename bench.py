#!/usr/bin/env python
"""bench.py -- rays/sec of the viewpoint raycast + scoring path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a engine through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on host cores

A "step" is one pass of the hot path over one batch of synthetic viewpoints: every pixel ray of every
viewpoint of the workload is cast through the occupancy map, hit voxels are de-duplicated per viewpoint and
scored.  At N = 1 the workload is BASELINE.json configs[1] ("C2": ~1 M-leaf box-building map, 1000 viewpoints at
640x480).  At N > 1 every rank scores its own 1000-viewpoint slice of an N x 1000 candidate set on a replicated
map (weak scaling) and the per-viewpoint scores / hit lists of all ranks end up on every rank inside the timed step
(q3dr_exchange: copy-engine pushes per finished chunk; Q3DR_GATHER=nccl: one NCCL all_gather after the last kernel);
after the timed loop the gathered bytes are compared with every rank's own result (`gather_verified`).  The line also
carries a `north_star` block: BASELINE.json configs[2] ("C3": ~10 M-leaf map, 1250 viewpoints per rank = 10 000 on 8
GPUs) timed the same way, with its own parity sample.

One JSON line is printed by rank 0.  `value` = device-resident throughput (poses already in HBM, results left
in HBM, CUDA events on the launch stream, max over ranks).  `e2e` = the same metric through the host-buffer
C-ABI call (pinned host poses -> H2D -> kernels -> D2H of the CSR result inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "rays_per_sec"
UNIT = "rays/s"
WORKLOADS = {
    # name: (synth config, viewpoints per rank, width, height)
    "C1": ("C1", 100, 320, 240),
    "C2": ("C2", 1000, 640, 480),
    "C3": ("C3", 1250, 640, 480),
    "C4": ("C4", 500, 1280, 960),
    "small": ("small", 16, 160, 120),
}


def algorithmic_bytes_per_ray(n_leaves: int) -> int:
    """(2 D + 1) * 32 B + 4 B with D = ceil(log2 N_leaf)."""
    d = int(math.ceil(math.log2(max(n_leaves, 2))))
    return (2 * d + 1) * 32 + 4


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_ncu(workload: str) -> dict:
    """Figures of the committed ncu --set full capture of the raycast kernel (profiles/traffic.json), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f).get(workload)
            return t if isinstance(t, dict) else {}
    except Exception:
        return {}


def committed_traffic(workload: str):
    """dram bytes per raycast launch from the committed capture, scaled to nothing: per launch of that capture."""
    return committed_ncu(workload).get("dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.lines.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """All host threads this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def build_inputs(workload: str, rank: int, world: int, use_product: bool = True, nvp=None):
    """(leaves in the reference's leaf order, leaf depths, camera, this rank's poses).  use_product=False computes the
    leaf order with the oracle's own Tree::build restatement, so that the reference arm never maps the product library
    (both give the same permutation: tests/test_host_cpu.py)."""
    from quad3dr_b200 import synth

    cfg, nvp_default, w, h = WORKLOADS[workload]
    nvp = nvp or nvp_default
    leaves = synth.CONFIGS[cfg][0]()
    if use_product:
        from quad3dr_b200 import engine as E

        order, depth = E.splitmedian_order(leaves.boxes)  # tie-break order of the reference's Tree::build
    else:
        from oracle import oracle_py as O

        order = O.OracleTree(leaves.boxes).dfs_order()
        depth = None
    leaves = leaves.permuted(order)
    cam = synth.make_camera(w, h)
    # one global candidate set of world*nvp viewpoints, rank r scores the contiguous slice [r*nvp, (r+1)*nvp)
    all_vps = synth.make_viewpoints(leaves, nvp * world, config_id=2)
    vps = np.ascontiguousarray(all_vps[rank * nvp:(rank + 1) * nvp])
    return leaves, depth, cam, vps


def cpu_baseline(leaves, cam, vps, budget_s: float, threads=None):
    """The oracle (CPU restatement of the reference algorithm) on host cores, bounded sample."""
    from oracle import oracle_py as O

    tree = O.OracleTree(leaves.boxes)
    threads = threads or host_threads()
    # calibrate on two viewpoints, single thread
    t = time.perf_counter()
    tree.score_batch(leaves.weight, leaves.normal, cam.K, cam.width, cam.height, vps[:2], threads=1)
    per_vp = (time.perf_counter() - t) / 2
    n = int(max(threads, min(vps.shape[0], math.ceil(budget_s * threads / max(per_vp, 1e-6)))))
    n = min(n, vps.shape[0])
    t = time.perf_counter()
    r = tree.score_batch(leaves.weight, leaves.normal, cam.K, cam.width, cam.height, vps[:n], threads=threads)
    dt = time.perf_counter() - t
    rays = n * cam.width * cam.height
    return {
        "value": rays / dt, "unit": UNIT, "cores": int(threads), "kind": "port",
        "sample": f"first {n} of {vps.shape[0]} viewpoints at {cam.width}x{cam.height} ({rays} rays), "
                  f"OpenMP over viewpoints, {dt:.2f} s; single-thread rate {cam.width * cam.height / per_vp:.3e} rays/s",
        "_n": n, "_dt": dt, "_result": r, "_tree": tree,
    }


def reference_cuda_baseline(leaves, cam, vps, tree, n_vp: int = 16):
    """The reference's own recursive CUDA kernel (viewpoint_planner/src/bvh/bvh.cu, built with its own
    -use_fast_math into oracle/_ref by oracle/ref_cuda/Makefile) recompiled for sm_100a, on the same map and the first
    n_vp viewpoints of the same workload.  Context only: it is neither the parity oracle nor the product."""
    try:
        from oracle import refcuda_py as R

        if not R.available("fast"):
            return {"unavailable": "oracle/_ref/libq3dr_refcuda_fast.so not built (needs the reference checkout at build time)"}
        rt = R.RefCudaTree(tree, "fast")
        n_vp = min(n_vp, vps.shape[0])
        rt.time_raycasts(cam.K, vps[:2], cam.width, cam.height, 60.0)  # warm-up
        r = rt.time_raycasts(cam.K, vps[:n_vp], cam.width, cam.height, 60.0)
        return {
            "kernel_rays_per_sec": r["rays"] / (r["kernel_ms"] * 1e-3), "api_rays_per_sec": r["rays"] / (r["api_ms"] * 1e-3),
            "unit": UNIT, "viewpoints": int(n_vp), "kernel_ms": r["kernel_ms"], "api_ms": r["api_ms"],
            "what": "bvh::CudaTree<float>::raycastWithScreenCoordinatesRecursive: kernel = its __global__ alone (CUDA events); "
                    "api = the public call incl. cudaDeviceSynchronize + D2H of W*H 56-byte records per viewpoint; "
                    "traversal only (no dedup, no scoring), 32 KiB device stack per thread as the reference sets it",
        }
    except Exception as e:  # context only: never fail the bench on it
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself needs Eigen/Boost, absent
    here) on ALL host threads of the box; each step a bounded sample of the same workload.  Never loads the product
    library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as O

    world = max(int(args.gpus), 1)
    leaves, _, cam, vps = build_inputs(args.workload, 0, 1, use_product=False)
    tree = O.OracleTree(leaves.boxes)
    threads = host_threads()
    t = time.perf_counter()
    tree.score_batch(leaves.weight, leaves.normal, cam.K, cam.width, cam.height, vps[:2], threads=1)
    per_vp = (time.perf_counter() - t) / 2
    total_steps = args.steps + args.warmup
    budget = 150.0 / max(total_steps, 1)  # whole run within a few minutes
    n = int(min(vps.shape[0], max(threads, math.floor(budget * threads / max(per_vp, 1e-6)))))
    n = max(1, min(n, vps.shape[0]))
    times = []
    for s in range(total_steps):
        t = time.perf_counter()
        tree.score_batch(leaves.weight, leaves.normal, cam.K, cam.width, cam.height, vps[:n], threads=threads)
        dt = time.perf_counter() - t
        if s >= args.warmup:
            times.append(dt)
    rays = n * cam.width * cam.height
    ms = 1e3 * sum(times) / len(times)
    value = rays / (ms * 1e-3)
    sample = (f"each step = first {n} of {vps.shape[0]} viewpoints at {cam.width}x{cam.height} ({rays} rays), "
              f"OpenMP over viewpoints on {threads} threads (all host threads; OMP_NUM_THREADS ignored)")
    cfg = workload_config(args.workload, leaves, cam, vps.shape[0], world)
    cfg["gather"] = gather_description(world)
    cfg["normals"] = "per voxel (enable_opengl = false)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(threads), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def gather_description(world: int) -> str:
    if world <= 1:
        return "single GPU: nothing to exchange"
    if os.environ.get("Q3DR_GATHER", "peer") == "peer":
        return ("q3dr_exchange: per finished chunk copy-engine pushes (cudaMemcpyAsync into CUDA-IPC-mapped peer HBM over "
                "NVLink) on a side stream while the next chunk is cast; completion flags behind the data, waited for on "
                "the device; no collective inside a step")
    return "NCCL all_gather of headers + padded hit lists after the last kernel"


def workload_config(workload, leaves, cam, nvp_per_rank, world):
    return {
        "workload": f"{workload}: synthetic box-building occupancy map, {leaves.n} BVH leaves (r={leaves.resolution} m), "
                    f"{nvp_per_rank * world} viewpoints at {cam.width}x{cam.height}, raycast_max_range 60 m, "
                    f"dedup + distance/incidence score, ignore_zero=1",
        "n_leaves": int(leaves.n), "viewpoints": int(nvp_per_rank * world), "viewpoints_per_gpu": int(nvp_per_rank),
        "width": cam.width, "height": cam.height,
        "parallelism": f"viewpoint-sharded x{world}, map replicated",
        "l2": "flushed between timed steps (256 MiB device write); node array "
              + ("fits in the 126 MB L2" if leaves.n < 2_000_000 else "is larger than the 126 MB L2"),
        "timing": "CUDA events on the launch stream around each step (flush excluded), summed over steps, max over ranks",
    }


class Workload:
    """One workload resident on this rank's GPU: map, poses, (at N > 1) the exchange; knows how to run and time steps."""

    def __init__(self, name, rank, world, local_rank, args, nvp=None):
        import torch

        from quad3dr_b200 import engine as E

        self.torch = torch
        self.E = E
        self.name, self.rank, self.world, self.local_rank, self.args = name, rank, world, local_rank, args
        self.dev = torch.device("cuda", local_rank)
        self.leaves, self.depth, self.cam, self.vps = build_inputs(name, rank, world, nvp=nvp)
        self.nvp, self.w, self.h = self.vps.shape[0], self.cam.width, self.cam.height
        self.gmap = E.OccupancyMap(self.leaves.boxes, self.leaves.weight, self.leaves.normal, self.depth, device=local_rank)
        self.mesh = None
        if args.normals == "mesh":
            # the reference's enable_opengl = true branch: normals from a surface mesh at the kept pixel
            from quad3dr_b200 import synth

            tv, tn = synth.make_surface_mesh(self.leaves, cell=args.mesh_cell)
            self.mesh = E.NormalMesh(tv, tn, device=local_rank)
        self.opts = E.default_opts()
        self.window = (0, self.w, 0, self.h)
        self.stream = torch.cuda.current_stream()
        self.d_rt = torch.from_numpy(self.vps).to(self.dev)
        self.h_rt = torch.from_numpy(self.vps).pin_memory()
        self.rays_rank = self.nvp * self.w * self.h
        self.px = None
        self.gather_note = gather_description(world)
        self.gather_mode = "none" if world == 1 else os.environ.get("Q3DR_GATHER", "peer")

    # -- one pass of the path, poses and results resident in HBM
    def score_device(self):
        if self.mesh is not None:
            return self.gmap.score_viewpoints_mesh_normals_device(self.mesh, self.cam.K, self.w, self.h, self.window,
                                                                  self.d_rt.data_ptr(), self.nvp, self.opts,
                                                                  stream=self.stream.cuda_stream)
        return self.gmap.score_viewpoints_device(self.cam.K, self.w, self.window, self.d_rt.data_ptr(), self.nvp, self.opts,
                                                 stream=self.stream.cuda_stream)

    # -- the same through the host-buffer call a user makes (H2D poses, D2H CSR result inside)
    def score_host(self):
        if self.mesh is not None:
            return self.gmap.score_viewpoints_mesh_normals(self.mesh, self.cam.K, self.w, self.h, self.window, self.h_rt.numpy(),
                                                           self.opts, copy=False)
        return self.gmap.score_viewpoints(self.cam.K, self.w, self.window, self.h_rt.numpy(), self.opts, copy=False)

    def setup_exchange(self):
        """N > 1: size the receive slots from a probe pass (the result of fixed inputs is deterministic)."""
        import torch.distributed as dist

        if self.world == 1 or self.gather_mode != "peer":
            return
        from quad3dr_b200 import dist as qdist

        probe = self.score_device()
        e_max = self.torch.tensor([probe.n_entries], dtype=self.torch.int64, device=self.dev)
        dist.all_reduce(e_max, op=dist.ReduceOp.MAX)
        try:
            self.px = qdist.PeerExchange(self.gmap, self.dev, self.nvp, int(int(e_max.item()) * 1.25) + 4096)
        except Exception as e:  # e.g. CUDA IPC not permitted in this container: every rank lands here together
            self.px = None
            self.gather_mode = "nccl"
            self.gather_note = ("NCCL all_gather of headers + padded hit lists after the last kernel "
                                f"(peer push unavailable: {type(e).__name__}: {e})")

    def step_device(self):
        if self.px is not None:
            self.px.begin_step()
            res = self.score_device()
            self.px.end_step(0, self.stream.cuda_stream)
            return res, None
        res = self.score_device()
        gathered = None
        if self.world > 1:
            from quad3dr_b200 import dist as qdist

            off, tot, leaf, info = qdist.device_result_tensors(res, self.dev)
            gathered = qdist.gather_results(off, tot, leaf, info, self.nvp)
        return res, gathered

    def verify_gather(self, res, gathered) -> bool:
        """Every rank compares, for every source rank, what it RECEIVED with what that rank itself computed (checksums
        of the source's own device result, all-gathered).  True only if every slot on every rank matches."""
        import torch.distributed as dist

        from quad3dr_b200 import dist as qdist

        torch = self.torch

        def checksum(off, tot, leaf, info):
            # position-sensitive 63-bit sums: any dropped, shifted or corrupted element changes them
            def wsum(x):
                x = x.to(torch.int64)
                k = torch.arange(1, x.numel() + 1, device=x.device, dtype=torch.int64)
                return (x * (k % 1000003)).sum() if x.numel() else torch.zeros((), dtype=torch.int64, device=x.device)

            return torch.stack([torch.tensor(float(tot.numel()), device=off.device).to(torch.int64),
                                torch.tensor(float(leaf.numel()), device=off.device).to(torch.int64),
                                wsum(off), wsum(tot.view(torch.int32)), wsum(leaf), wsum(info.view(torch.int32))])

        off, tot, leaf, info = qdist.device_result_tensors(res, self.dev)
        mine = checksum(off - off[0], tot, leaf, info)
        allsums = torch.empty(self.world * mine.numel(), dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(allsums, mine)
        allsums = allsums.view(self.world, -1)
        ok = True
        if self.px is not None:
            views = self.px.views()
        else:
            o_all, t_all, l_all, i_all = gathered
            views = []
            v0 = 0
            for r in range(self.world):
                n = int(allsums[r, 0].item())
                e0, e1 = int(o_all[v0].item()), int(o_all[v0 + n].item())
                views.append((o_all[v0:v0 + n + 1] - o_all[v0], t_all[v0:v0 + n], l_all[e0:e1], i_all[e0:e1]))
                v0 += n
        for r in range(self.world):
            o, t, l, i = views[r]
            ok = ok and bool(torch.equal(checksum(o - o[0], t, l, i), allsums[r]))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag.item()))

    def close(self):
        if self.px is not None:
            self.px.close()
            self.px = None
        if self.mesh is not None:
            self.mesh.close()
        self.gmap.close()


def time_workload(wl: "Workload", steps: int, warmup: int, flush, barrier, sample_clocks: bool):
    """W untimed + K timed device-resident steps (CUDA events on the launch stream, L2 flushed between steps), then the
    end-to-end arm with host buffers.  Returns a dict of per-rank numbers."""
    import torch

    wl.gmap.set_result_fields(first_pixel=True, dist_sq=True)
    wl.setup_exchange()
    for _ in range(warmup):
        flush.fill_(1)
        wl.step_device()
    launches = 0
    raycast_ms = 0.0
    raycast_launches = 0
    sampler = ClockSampler(wl.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()
    t_wall0 = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    res = gathered = None
    for s in range(steps):
        flush.fill_(s & 0xFF)
        ev[s][0].record(wl.stream)
        res, gathered = wl.step_device()
        ev[s][1].record(wl.stream)
        st = wl.gmap.stats()
        launches += int(st.kernel_launches)
        raycast_ms += float(st.raycast_ms)
        raycast_launches += int(st.chunks)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    n_entries = res.n_entries
    gather_verified = None
    if wl.world > 1:
        gather_verified = wl.verify_gather(res, gathered)
    if wl.px is not None:
        wl.px.close()
        wl.px = None

    # ---- end-to-end arm: the host-buffer C-ABI call with what the reference's facade consumes, (voxel, information) ----
    wl.gmap.set_result_fields(first_pixel=False, dist_sq=False)
    for _ in range(max(1, warmup - 1)):
        wl.score_host()
    barrier()
    e2e_ms = 0.0
    hres = None
    for s in range(steps):
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hres = wl.score_host()
        e2e_ms += (time.perf_counter() - t0) * 1e3
    barrier()
    h2d = wl.nvp * 48
    d2h = (wl.nvp + 1) * 8 + wl.nvp * 8 + hres.n_entries * 8
    wl.gmap.set_result_fields(first_pixel=True, dist_sq=True)
    return dict(dev_ms=dev_ms, e2e_ms=e2e_ms, raycast_ms=raycast_ms, raycast_launches=raycast_launches, launches=launches,
                h2d=h2d, d2h=d2h, n_entries=int(n_entries), clocks=clocks, gather_verified=gather_verified)


def parity_sample(wl: "Workload", budget_s: float):
    """cpu_baseline + parity of the bench inputs themselves: the oracle on a bounded sample of this rank's viewpoints
    against a full-field engine call on the same viewpoints."""
    from oracle import oracle_py as O

    cb = cpu_baseline(wl.leaves, wl.cam, wl.vps, budget_s)
    n = cb.pop("_n")
    r = cb.pop("_result")
    cb.pop("_dt")
    tree = cb.pop("_tree")
    wl.gmap.set_result_fields(first_pixel=True, dist_sq=True)
    hr = wl.gmap.score_viewpoints(wl.cam.K, wl.w, wl.window, wl.vps[:n], wl.opts, copy=True)
    kept = (hr.offsets[1:n + 1] - hr.offsets[:n]).astype(np.int64)
    same_counts = bool(np.array_equal(kept, r["kept_count"].astype(np.int64)))
    rel = np.abs(hr.total[:n] - r["total"]) / np.maximum(np.abs(r["total"]), 1e-30)
    chk = O.set_checksums(hr.offsets[:n + 1], hr.leaf[:int(hr.offsets[n])], hr.first_pixel[:int(hr.offsets[n])])
    cb["parity_on_sample"] = {"viewpoints": int(n), "hit_list_sizes_equal": same_counts,
                              "voxel_sets_and_first_pixels_equal": bool(np.array_equal(chk, r["set_checksum"])),
                              "max_rel_score_err": float(rel.max()) if n else 0.0,
                              "how": "per viewpoint: kept count, 64-bit checksum over (leaf, first pixel) of the "
                                     "voxel set, total within 1e-5 relative of the oracle"}
    return cb, tree


def shim_bench(wl: "Workload", steps: int):
    """e2e_shim: the same workload through the reference's C++ call surface (include/q3dr/occupied_tree_cuda.hpp),
    timed by tests/cpp/bench_shim (built by __graft_entry__.build) in its own process."""
    import shutil
    import tempfile

    exe = os.path.join(ROOT, "tests", "cpp", "bench_shim")
    if not os.path.exists(exe):
        return {"unavailable": "tests/cpp/bench_shim not built (python -c 'import __graft_entry__ as g; g.build()')"}
    d = tempfile.mkdtemp(prefix="q3dr_shim_")
    try:
        wl.leaves.boxes.astype(np.float32).tofile(os.path.join(d, "boxes.f32"))
        wl.leaves.weight.astype(np.float32).tofile(os.path.join(d, "weight.f32"))
        wl.leaves.normal.astype(np.float32).tofile(os.path.join(d, "normal.f32"))
        np.asarray(wl.depth, dtype=np.uint32).tofile(os.path.join(d, "depth.u32"))
        wl.vps.astype(np.float32).tofile(os.path.join(d, "rt.f32"))
        K = wl.cam.K
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"{wl.leaves.n} {wl.nvp} {wl.w} {wl.h} {float(K[0])!r} {float(K[1])!r} {float(K[2])!r} {float(K[3])!r}\n")
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(host_threads())
        env["CUDA_VISIBLE_DEVICES"] = env.get("CUDA_VISIBLE_DEVICES", str(wl.local_rank))
        r = subprocess.run([exe, d, str(steps), "1"], capture_output=True, text=True, timeout=600, env=env)
        if r.returncode != 0:
            return {"unavailable": f"bench_shim rc={r.returncode}: {r.stderr[-300:]}"}
        out = json.loads(r.stdout.strip().splitlines()[-1])
        out["what"] = ("ViewpointRaycastCuda::getRaycastHitVoxelsWithInformationScoreBatch[View] of the header-only shim, host "
                       "poses in -> per-viewpoint VoxelWithInformationSet[View] + total out, wall clock; c_abi = the "
                       "q3dr_score_viewpoints call underneath (same process, same inputs)")
        out["value"] = out["batch_view_rays_per_sec"]
        out["unit"] = UNIT
        return out
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def roofline_block(wl, t, clocks, sm_count, steps):
    """The raycast kernel is issue-bound (ncu: ~80 % issue-slot utilisation, DRAM < 1 %): its roofline is the issue rate
    of the SMs.  achieved = warp instructions per ray of the committed ncu capture x rays per launch / the launch time
    measured live; peak = SMs x 4 schedulers x the SM clock sampled during the run.  The byte model of SURVEY.md 8(d)
    is reported next to it."""
    b_ray = algorithmic_bytes_per_ray(wl.leaves.n)
    peak_hbm, peak_src = measured_peak_gbs()
    rays_per_launch = wl.rays_rank * steps / max(t["raycast_launches"], 1)
    launch_ms = t["raycast_ms"] / max(t["raycast_launches"], 1)
    ncu = committed_ncu(wl.name)
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
    block = {"kernel": "raycast2_kernel<kModeScore>", "rays_per_launch": rays_per_launch, "launch_ms": launch_ms,
             "launches_timed": t["raycast_launches"], "kernel_share_of_step": t["raycast_ms"] / max(t["dev_ms"], 1e-9)}
    wi = ncu.get("warp_instructions_per_ray")
    if wi and sm_mhz:
        peak_issue = sm_count * 4 * sm_mhz * 1e6
        ach_issue = wi * rays_per_launch / (launch_ms * 1e-3)
        block.update({"bound": "issue", "achieved": ach_issue, "peak": peak_issue, "unit": "warp-inst/s",
                      "frac": ach_issue / peak_issue,
                      "peak_source": f"{sm_count} SMs x 4 warp schedulers x {sm_mhz:.0f} MHz (nvidia-smi, sampled in the timed region)",
                      "how": "warp instructions per ray (smsp__inst_executed.sum / rays of the committed ncu capture, "
                             f"{ncu.get('source', 'profiles/traffic.json')}) x rays per launch / live launch time"})
    else:
        block.update({"bound": "issue", "achieved": None, "peak": None, "unit": "warp-inst/s", "frac": None,
                      "how": "no committed ncu capture for this workload / no clock sample"})
    dram = ncu.get("dram_bytes_per_launch")
    rpl = ncu.get("rays_per_launch")
    block["traffic"] = (dram * rays_per_launch / rpl) if (dram and rpl) else None
    block["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture, scaled by "
                             "rays per launch (capture: %s rays, bench: %.0f rays)" % (rpl, rays_per_launch))
    ach_hbm = b_ray * rays_per_launch / (launch_ms * 1e-3) / 1e9
    block["hbm_model"] = {"algorithmic_bytes_per_ray": b_ray, "achieved": ach_hbm, "peak": peak_hbm, "unit": "GB/s",
                          "frac": ach_hbm / peak_hbm, "peak_source": peak_src,
                          "note": "SURVEY.md 8(d): (2 ceil(log2 N) + 1) x 32 B + 4 B per ray, every ray walking a binary tree out of "
                                  "HBM.  The kernel walks a 4-wide tree once per ray PACKET out of L1/L2 (dram traffic above), so this "
                                  "fraction exceeds 1 and bounds nothing; kept because section 8(d) defines it"}
    block["ncu"] = {k: v for k, v in ncu.items() if k not in ("source",)}
    return block


def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_times(t):
        tt = torch.tensor([t["dev_ms"], t["e2e_ms"], t["raycast_ms"]], dtype=torch.float64, device=dev)
        agg = torch.tensor([float(t["launches"]), float(t["h2d"]), float(t["d2h"]), float(t["n_entries"])], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        return [float(x) for x in tt.tolist()], [float(x) for x in agg.tolist()]

    wl = Workload(args.workload, rank, world, local_rank, args)
    t = time_workload(wl, args.steps, args.warmup, flush, barrier, sample_clocks=True)
    (dev_ms, e2e_ms, _), (launches_all, h2d_all, d2h_all, entries_all) = reduce_times(t)
    sm_count = int(wl.gmap.info().sm_count)
    line = None
    if rank == 0:
        total_rays = wl.rays_rank * world
        ms_per_step = dev_ms / args.steps
        value = total_rays / (ms_per_step * 1e-3)
        e2e_value = total_rays / (e2e_ms / args.steps * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl.leaves, wl.cam, wl.nvp, world),
            "viewpoints_per_sec": wl.nvp * world / (ms_per_step * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                    "ms_per_step": e2e_ms / args.steps,
                    "note": "q3dr_score_viewpoints: pinned host poses -> H2D -> kernels -> D2H of the CSR hit lists "
                            "(leaf, information: what getRaycastHitVoxelsWithInformationScore returns) + scores"},
            "gpu_launches": int(launches_all),
            "clocks": t["clocks"],
            "roofline": roofline_block(wl, t, t["clocks"], sm_count, args.steps),
            "result_entries": int(entries_all),
        }
        line["config"]["gather"] = wl.gather_note
        if world > 1:
            line["gather_verified"] = t["gather_verified"]
        if wl.mesh is not None:
            mi = wl.mesh.info()
            line["config"]["normals"] = (f"image-space (enable_opengl branch): surface mesh of {int(mi.n_triangles)} triangles, "
                                         f"one mesh ray per kept pixel inside the scoring pass")
            line["config"]["workload"] += ", normals from the mesh at the kept pixel"
        else:
            line["config"]["normals"] = "per voxel (enable_opengl = false)"
        if args.cpu_baseline:
            cb, tree = parity_sample(wl, args.cpu_budget)
            line["reference_cuda"] = reference_cuda_baseline(wl.leaves, wl.cam, wl.vps, tree)
            line["cpu_baseline"] = cb
            del tree
        if args.shim and world == 1 and wl.mesh is None:
            line["e2e_shim"] = shim_bench(wl, max(2, min(args.steps, 5)))
    if world > 1 and t["gather_verified"] is False:
        wl.close()
        raise SystemExit("bench.py: the gathered results differ from the ranks' own results")
    wl.close()
    del wl
    torch.cuda.empty_cache()

    # ---- the north-star configuration (BASELINE.json configs[2]) next to the headline workload ----
    if args.north_star and args.workload == "C2" and args.normals == "voxel":
        ns_steps = max(10, min(args.steps, 20))
        ns = Workload("C3", rank, world, local_rank, args)
        tn = time_workload(ns, ns_steps, 3, flush, barrier, sample_clocks=(rank == 0))
        (n_dev_ms, n_e2e_ms, _), (n_launches, n_h2d, n_d2h, n_entries) = reduce_times(tn)
        if rank == 0:
            total_rays = ns.rays_rank * world
            blk = {
                "what": "BASELINE.json configs[2] / north_star: ~10 M-leaf map, 640x480, 1250 viewpoints per GPU "
                        "(10 000 candidates on 8 GPUs), same step definition and timing as the headline workload",
                "config": workload_config("C3", ns.leaves, ns.cam, ns.nvp, world),
                "steps": ns_steps, "warmup": 3,
                "value": total_rays / (n_dev_ms / ns_steps * 1e-3), "unit": UNIT, "ms_per_step": n_dev_ms / ns_steps,
                "viewpoints_per_sec": ns.nvp * world / (n_dev_ms / ns_steps * 1e-3),
                "e2e": {"value": total_rays / (n_e2e_ms / ns_steps * 1e-3), "unit": UNIT, "ms_per_step": n_e2e_ms / ns_steps,
                        "h2d_bytes_per_step": int(n_h2d), "d2h_bytes_per_step": int(n_d2h)},
                "gpu_launches": int(n_launches), "result_entries": int(n_entries),
                "roofline": roofline_block(ns, tn, tn["clocks"], sm_count, ns_steps),
                "target": "north_star: >= 1e10 rays/s scored at 8 B200 on a 10 M-occupied-voxel map",
            }
            blk["config"]["gather"] = ns.gather_note
            if world > 1:
                blk["gather_verified"] = tn["gather_verified"]
            # parity of THIS configuration: the oracle on a few viewpoints of rank 0's slice (the 10 M-leaf reference tree
            # takes a while to build, so the sample is small)
            cb, tree = parity_sample(ns, min(args.cpu_budget, 6.0))
            blk["cpu_baseline"] = cb
            del tree
            line["north_star" if world > 1 else "c3"] = blk
        if world > 1 and tn["gather_verified"] is False:
            raise SystemExit("bench.py: north-star block: the gathered results differ from the ranks' own results")
        ns.close()
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_multi(args):
    """--mode multi: the path as the reference's single C++ process would drive it -- q3dr_multi_*: one process, one host
    thread per GPU, map replicated, viewpoints sharded, host poses in, host CSR results out, and (exchange on) every GPU
    holding all GPUs' results at the end of the call.  Wall clock around the C-ABI call."""
    import torch

    from quad3dr_b200 import engine as E

    n_gpus = max(1, int(args.gpus))
    if torch.cuda.device_count() < n_gpus:
        raise SystemExit(f"bench.py --mode multi: {n_gpus} GPUs requested, {torch.cuda.device_count()} visible")
    leaves, depth, cam, vps = build_inputs(args.workload, 0, 1, nvp=WORKLOADS[args.workload][1] * n_gpus)
    nvp, w, h = vps.shape[0], cam.width, cam.height
    multi = E.MultiGpuMap(leaves.boxes, leaves.weight, leaves.normal, depth, devices=list(range(n_gpus)))
    multi.set_result_fields(first_pixel=False, dist_sq=False)
    opts = E.default_opts()
    window = (0, w, 0, h)
    probe = multi.score_viewpoints(cam.K, w, window, vps, opts, copy=False)
    per_gpu_vp = max(r.n_viewpoints for _, r in probe)
    per_gpu_entries = max(r.n_entries for _, r in probe)
    if n_gpus > 1 and os.environ.get("Q3DR_GATHER", "peer") == "peer":
        multi.enable_gather(per_gpu_vp, int(per_gpu_entries * 1.25) + 4096)
        gather_note = "q3dr_multi_enable_gather: per finished chunk copy-engine pushes into every GPU's HBM (peer access), completion flags"
    else:
        gather_note = "none: results stay per GPU in pinned host memory"
    for _ in range(max(args.warmup, 3)):
        multi.score_viewpoints(cam.K, w, window, vps, opts, copy=False)
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    flushes = [torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", g)) for g in range(n_gpus)]
    t_wall0 = time.time()
    ms = 0.0
    parts = None
    for s_ in range(args.steps):
        for f in flushes:
            f.fill_(s_ & 0xFF)
        for g in range(n_gpus):
            torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        parts = multi.score_viewpoints(cam.K, w, window, vps, opts, copy=False)
        ms += (time.perf_counter() - t0) * 1e3
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    entries = sum(r.n_entries for _, r in parts)
    verified = None
    if n_gpus > 1 and gather_note.startswith("q3dr_multi_enable_gather"):
        # every GPU's copy of every part against the part's own host result
        from quad3dr_b200.dist import CudaArrayView

        verified = True
        for on in range(n_gpus):
            for src, (_, ref) in enumerate(parts):
                v = multi.gathered(on, src)
                ok = v.n_viewpoints == ref.n_viewpoints and v.n_entries == ref.n_entries
                if ok and ref.n_entries:
                    dev = torch.device("cuda", on)
                    leaf = torch.as_tensor(CudaArrayView(v.leaf, v.n_entries, "<i4"), device=dev).cpu().numpy()
                    info = torch.as_tensor(CudaArrayView(v.information, v.n_entries, "<i4"), device=dev).cpu().numpy()
                    ok = bool(np.array_equal(leaf, ref.leaf) and np.array_equal(info, ref.information.view(np.int32)))
                verified = verified and ok
    rays = nvp * w * h
    ms_per_step = ms / args.steps
    cfg = workload_config(args.workload, leaves, cam, nvp // n_gpus, n_gpus)
    cfg["parallelism"] = f"one process, one host thread per GPU (q3dr_multi_*), viewpoint-sharded x{n_gpus}, map replicated"
    cfg["timing"] = "wall clock around q3dr_multi_score_viewpoints (host poses in, host CSR results out), summed over steps"
    cfg["gather"] = gather_note
    cfg["normals"] = "per voxel (enable_opengl = false)"
    line = {
        "metric": METRIC, "value": rays / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "multi", "config": cfg,
        "viewpoints_per_sec": nvp / (ms_per_step * 1e-3),
        "e2e": {"value": rays / (ms_per_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nvp * 48,
                "d2h_bytes_per_step": (nvp + n_gpus) * 8 + nvp * 8 + entries * 8, "ms_per_step": ms_per_step,
                "note": "the timed call IS the host-buffer call: value == e2e in this mode"},
        "gpu_launches": None, "clocks": clocks, "result_entries": int(entries), "gather_verified": verified,
    }
    multi.close()
    emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    Everything except that line is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of wall time for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--normals", default="voxel", choices=["voxel", "mesh"],
                    help="mesh: image-space normals from a synthetic surface mesh (row N2); the cpu_baseline leg is "
                         "skipped (the oracle's mesh render is a brute force over all triangles)")
    ap.add_argument("--mesh-cell", type=float, default=0.25, help="quad size of the synthetic surface mesh in metres")
    ap.add_argument("--no-north-star", dest="north_star", action="store_false",
                    help="skip the C3 block (10 M-leaf map) that accompanies the default C2 line")
    ap.add_argument("--no-shim", dest="shim", action="store_false", help="skip the e2e_shim block (C++ call surface)")
    ap.add_argument("--mode", default="ranks", choices=["ranks", "multi"],
                    help="ranks: one process per GPU (torchrun contract, the default); multi: ONE process drives all --gpus "
                         "through q3dr_multi_* like the reference's C++ process would (launch with plain python)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl == "ours" and args.gpus > 1 and world == 1 and args.mode == "ranks"):
        _quiet_stdout()  # (the torchrun re-launch below passes its children's stdout through untouched)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.mode == "multi":
        if world > 1:
            raise SystemExit("bench.py --mode multi is a single process: launch it with plain python, not torchrun")
        run_multi(args)
        return
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if world > 1 or args.normals == "mesh":
        args.cpu_baseline = False  # rank 0 at N = 1 only; voxel normals only (the north-star block keeps its own sample)
    run_ours(args)


if __name__ == "__main__":
    main()
