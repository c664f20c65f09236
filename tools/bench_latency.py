"""Per-call latency of the host-buffer entry point for small batches: what the reference's planner sees when it scores
candidates one at a time (viewpoint_planner_graph.cpp:542-745 calls getRaycastHitVoxelsWithInformationScore per sampled
pose).  python tools/bench_latency.py [workload] -> one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quad3dr_b200 import engine as E, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
mk, _, w, h = synth.CONFIGS[wl]
leaves = mk()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(w, h)
vps = synth.make_viewpoints(leaves, 256, config_id=2)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
win = (0, w, 0, h)
out = {}
for n in (1, 2, 4, 16, 64, 256):
    for _ in range(3):
        gmap.score_viewpoints(cam.K, w, win, vps[:n], copy=False)
    reps = max(4, 256 // n)
    ts = []
    for r in range(reps):
        a = (r * n) % (256 - n + 1)
        t = time.perf_counter()
        gmap.score_viewpoints(cam.K, w, win, vps[a:a + n], copy=False)
        ts.append(time.perf_counter() - t)
    st = gmap.stats()
    med = float(np.median(ts))
    out[str(n)] = {"ms_per_call": med * 1e3, "us_per_viewpoint": med / n * 1e6, "rays_per_s": n * w * h / med,
                   "device_ms_last": st.total_ms, "raycast_ms_last": st.raycast_ms}
print(json.dumps({"workload": wl, "width": w, "height": h, "per_batch": out}))
