"""C5 of BASELINE.json: scaling sweep, 1e6 - 1e10 rays over the 1 M-leaf and 10 M-leaf maps, strong scaling over the
ranks it is launched with (python tools/sweep_c5.py, or torchrun --nproc-per-node N tools/sweep_c5.py).  Device-
resident arm: poses in HBM, results left in HBM, CUDA events, max over ranks.  One JSON line per (map, rays) on rank 0."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quad3dr_b200 import engine as E, synth
from quad3dr_b200.dist import shard_range

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
W, H = 640, 480
PEAK = 6555.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
# instructions per ray of the committed ncu captures (profiles/traffic.json): the kernel is issue-bound, the issue
# fraction is the one that means something; the byte model of SURVEY.md 8(d) is kept because C5 names it
WI = {"1M": 52.5, "10M": 58.8}
try:
    _t = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "traffic.json")))
    WI = {"1M": _t["C2"]["warp_instructions_per_ray"], "10M": _t["C3"]["warp_instructions_per_ray"]}
except Exception:
    pass
ISSUE_PEAK = 148 * 4 * 1.965e9
plans = [("1M-leaf map (octree depth 14 class)", synth.map_1m, [1e6, 1e7, 1e8, 1e9, 1e10]),
         ("10M-leaf map (octree depth 16 class)", synth.map_10m, [1e7, 1e8, 1e9, 1e10])]
if len(sys.argv) > 1 and sys.argv[1] == "--quick":
    plans = [(plans[0][0], plans[0][1], [1e6, 1e8])]
for name, mk, ray_counts in plans:
    leaves = mk()
    order, depth = E.splitmedian_order(leaves.boxes)
    leaves = leaves.permuted(order)
    gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth, device=lr)
    cam = synth.make_camera(W, H)
    b_ray = (2 * int(math.ceil(math.log2(leaves.n))) + 1) * 32 + 4
    for rays in ray_counts:
        nvp_total = max(world, int(round(rays / (W * H))))
        lo, hi = shard_range(nvp_total, rank, world)
        # viewpoints repeat with period 2000: the sweep is about throughput, not about distinct poses
        base = synth.make_viewpoints(leaves, min(nvp_total, 2000), config_id=5)
        idx = np.arange(lo, hi) % base.shape[0]
        vps = np.ascontiguousarray(base[idx])
        d_rt = torch.from_numpy(vps).to(dev)
        stream = torch.cuda.current_stream()
        opts = E.default_opts()
        steps = 3 if rays <= 1e9 else 1
        for _ in range(2 if rays <= 1e9 else 1):
            gmap.score_viewpoints_device(cam.K, W, (0, W, 0, H), d_rt.data_ptr(), hi - lo, opts, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            res = gmap.score_viewpoints_device(cam.K, W, (0, W, 0, H), d_rt.data_ptr(), hi - lo, opts, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            total_rays = nvp_total * W * H
            rps = total_rays / (float(ms) * 1e-3)
            print(json.dumps({"map": name, "n_leaves": leaves.n, "n_gpus": world, "rays": total_rays, "viewpoints": nvp_total,
                              "ms": float(ms), "rays_per_sec": rps, "algorithmic_GBps_per_gpu": rps / world * b_ray / 1e9,
                              "frac_of_measured_hbm_peak": rps / world * b_ray / 1e9 / PEAK, "bytes_per_ray": b_ray,
                              "frac_of_issue_peak": rps / world * WI["1M" if leaves.n < 5_000_000 else "10M"] / ISSUE_PEAK}), flush=True)
    gmap.close()
    del gmap
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
