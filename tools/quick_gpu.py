"""Quick GPU sanity + timing script (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quad3dr_b200 import synth, engine as E

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
nvp_override = int(sys.argv[2]) if len(sys.argv) > 2 else None
mk, nvp, w, h = synth.CONFIGS[cfg]
if nvp_override: nvp = nvp_override
t = time.time(); leaves = mk(); print("map", leaves.n, round(time.time() - t, 2), "s")
t = time.time(); order, depth = E.splitmedian_order(leaves.boxes); leaves = leaves.permuted(order); print("order", round(time.time() - t, 2), "s")
cam = synth.make_camera(w, h)
vps = synth.make_viewpoints(leaves, nvp, config_id=1)
t = time.time(); m = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth); print("map_create", round(time.time() - t, 2), "s")
i = m.info(); print("nodes", i.n_nodes, "depth", i.tree_depth, "node MB", i.node_bytes / 1e6, "sm", i.sm_count)
for it in range(3):
    t = time.time()
    res = m.score_viewpoints(cam.K, w, (0, w, 0, h), vps, copy=False)
    dt = time.time() - t
    st = m.stats()
    rays = nvp * w * h
    print(f"iter {it}: wall {dt*1e3:.1f} ms  raycast {st.raycast_ms:.1f} ms  score {st.score_ms:.1f} ms  total {st.total_ms:.1f} ms  "
          f"chunks {st.chunks}x{st.viewpoints_per_chunk}  rays/s kernel {rays/st.raycast_ms*1e3:.3e} e2e {rays/dt:.3e}  "
          f"entries {res.n_entries} mean hits/vp {res.n_entries/nvp:.0f}")
print("dev MB", m.info().device_bytes / 1e6)
