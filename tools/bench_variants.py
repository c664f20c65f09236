"""Times every packet shape of the camera-window kernel (Q3DR_RAY_VARIANT, q3dr_api.cu: kRayVariants) on one workload
and checks that all of them deliver the same result bit for bit.   python tools/bench_variants.py [C2|C3|C4] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
from quad3dr_b200 import engine as E

workload = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variants = [v for v in os.environ.get("VARIANTS", "0,1,2,3,4,5").split(",")]  # "6" or "6:4" = variant 6 with Q3DR_TILE_RUN=4
leaves, depth, cam, vps = B.build_inputs(workload, 0, 1)
dev = torch.device("cuda", 0)
d_rt = torch.from_numpy(vps).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()
ref = None
out = []
for v in variants:
    vv, _, run = str(v).partition(":")
    os.environ["Q3DR_RAY_VARIANT"] = vv
    if run:
        os.environ["Q3DR_TILE_RUN"] = run
    else:
        os.environ.pop("Q3DR_TILE_RUN", None)
    gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth, device=0)
    opts = E.default_opts()
    win = (0, cam.width, 0, cam.height)
    for _ in range(2):
        gmap.score_viewpoints_device(cam.K, cam.width, win, d_rt.data_ptr(), vps.shape[0], opts, stream=stream.cuda_stream)
    ray_ms = tot_ms = 0.0
    chunks = 0
    for s in range(steps):
        flush.fill_(s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gmap.score_viewpoints_device(cam.K, cam.width, win, d_rt.data_ptr(), vps.shape[0], opts, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        st = gmap.stats()
        ray_ms += st.raycast_ms
        tot_ms += e0.elapsed_time(e1)
        chunks += st.chunks
    hr = gmap.result_to_host()
    sig = (hr.n_entries, int(hr.leaf.astype(np.int64).sum()), int(hr.first_pixel.astype(np.int64).sum()),
           int(hr.information.view(np.uint32).astype(np.int64).sum()), int(hr.dist_sq.view(np.uint32).astype(np.int64).sum()))
    if ref is None:
        ref = sig
    rays = vps.shape[0] * cam.width * cam.height
    rec = {"variant": v, "workload": workload, "raycast_ms_per_step": ray_ms / steps, "step_ms": tot_ms / steps,
           "rays_per_sec": rays / (tot_ms / steps * 1e-3), "chunks_per_step": chunks / steps, "same_result_as_first": sig == ref}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    gmap.close()
if not all(r["same_result_as_first"] for r in out):
    raise SystemExit("variants disagree")
