"""Throughput of the voxel-set consumers (row N4) on the C2 result resident in HBM, next to the oracle on one host
core.  python tools/bench_setops.py [workload] -> one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle_py as O
from quad3dr_b200 import engine as E, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
mk, nvp, w, h = synth.CONFIGS[wl]
nvp = min(nvp, 1000)
leaves = mk()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(w, h)
vps = synth.make_viewpoints(leaves, nvp, config_id=2)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
res = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
F = np.float32(1 / 3)
obs = E.ObservedVoxels(gmap)
for v in range(0, nvp, max(nvp // 16, 1)):  # a realistic, partly filled map
    obs.add_viewpoint(F, v)
obs.new_information(F, n=nvp)
ts = []
for _ in range(10):
    t = time.perf_counter(); got = obs.new_information(F, n=nvp); ts.append(time.perf_counter() - t)
t_gpu = min(ts)
ref_obs = obs.download()
ns = min(nvp, 64)
t = time.perf_counter()
ref = [O.new_information(res.viewpoint(c)["leaf"], res.viewpoint(c)["info"], ref_obs, leaves.weight, F)[1] for c in range(ns)]
t_cpu = time.perf_counter() - t
entries = int(res.n_entries)
entries_s = int(res.offsets[ns])
rel = max(abs(got[c] - ref[c]) / max(abs(ref[c]), 1e-30) for c in range(ns))
ts2 = []
others = np.arange(nvp, dtype=np.uint32)
for _ in range(5):
    t = time.perf_counter(); E.overlap_information(gmap, 0, others); ts2.append(time.perf_counter() - t)
print(json.dumps({
    "workload": wl, "viewpoints": nvp, "entries": entries,
    "new_information": {"entries_per_s": entries / t_gpu, "ms": t_gpu * 1e3, "algorithmic_GBps": 16 * entries / t_gpu / 1e9,
                        "bytes_per_entry": 16, "note": "whole call through the C ABI incl. D2H of the sums"},
    "overlap_information": {"entries_per_s": entries / min(ts2), "ms": min(ts2) * 1e3},
    "cpu_oracle_1core": {"entries_per_s": entries_s / t_cpu, "sample_viewpoints": ns},
    "max_rel_err_vs_oracle": float(rel)}))
