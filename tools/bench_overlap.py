"""Throughput of isValidObjectPosition batches (row N3) on the C2 map, next to the oracle on one host core.
python tools/bench_overlap.py -> one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle_py as O
from quad3dr_b200 import engine as E, synth

leaves = synth.map_1m()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
tree = O.OracleTree(leaves.boxes)
root = tree.root_box()
rng = np.random.default_rng(1)
n = 2_000_000
pos = rng.uniform(root[:3], root[3:] + np.array([0, 0, 5]), size=(n, 3)).astype(np.float32)
pos[: n // 2, 2] = rng.uniform(0.3, 8, size=n // 2)  # near the ground and the buildings
obj = np.array([-0.6, -0.6, -0.3, 0.6, 0.6, 0.3], np.float32)
bvh = root.copy()
free_h = float(root[5]) + 1.0
gmap.valid_object_positions(pos[:1000], obj, bvh, free_h)
ts = []
for _ in range(5):
    t = time.perf_counter(); got = gmap.valid_object_positions(pos, obj, bvh, free_h); ts.append(time.perf_counter() - t)
ns = 20000
t = time.perf_counter()
ref = np.array([tree.valid_object_position(pos[i], obj, bvh, free_h) for i in range(ns)])
t_cpu = time.perf_counter() - t
print(json.dumps({"positions": n, "valid_fraction": float(got.mean()), "gpu_positions_per_s": n / min(ts), "gpu_ms": min(ts) * 1e3,
                  "note": "whole C-ABI call: H2D of positions, kernel, D2H of flags",
                  "cpu_oracle_1core_positions_per_s": ns / t_cpu, "parity_on_sample": bool(np.array_equal(got[:ns], ref))}))
