"""Timings of the rows next to the hot path on the C2 map, one B200 + the box's host cores:
N1  .ot file -> octree leaves (host) -> BVH boxes (device leaf rule) -> map (flatten + upload)
N3  q3dr_generate_viewpoint_entries: a batch of candidate poses through validity, raycast + score, distance gate
N4  q3dr_map_downweight_real_viewpoints: real images lower the voxel weights
python tools/bench_rows.py [n_candidates] [n_images] -> one JSON line (profiles/r02_rows_c2.json)."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import octree_ot as OT  # test infrastructure: only used to WRITE the synthetic .ot fixture
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = {"workload": "C2"}

mk, _, w, h = synth.CONFIGS["C2"]
leaves = mk()
cam = synth.make_camera(w, h)

# ---- N1: the map as an OctoMap .ot stream (leaves at depth 16 of a 0.25 m tree), read back and turned into a map
res = 0.25
ctr = 0.5 * (leaves.boxes[:, :3] + leaves.boxes[:, 3:])
keys = np.floor(ctr / res).astype(np.int64) + 32768
assert keys.min() >= 0 and keys.max() < 65536
t = time.perf_counter()
wts = leaves.weight.astype(np.float32)
tree = OT.octree_from_boxes(keys, np.zeros(len(keys), dtype=np.int64), lambda i: dict(occ=0.9, count=3, weight=float(wts[i])))
path = os.path.join(tempfile.mkdtemp(), "c2.ot")
OT.write_ot(path, tree, res)
t_write = time.perf_counter() - t
t = time.perf_counter()
ol = E.read_octree_leaves(path)
t_read = time.perf_counter() - t
n_ol = int(ol["info"].n_leaves)
bb = np.array([-1e30] * 3 + [1e30] * 3, dtype=np.float32)
E.octree_leaves_to_boxes(ol["center"][:1000], ol["size"][:1000], ol["occupancy"][:1000], ol["observation_count"][:1000], bb, 1e30)  # warm up
t = time.perf_counter()
boxes, src = E.octree_leaves_to_boxes(ol["center"], ol["size"], ol["occupancy"], ol["observation_count"], bb, 1e30)
t_rule = time.perf_counter() - t
t = time.perf_counter()
order, depth = E.splitmedian_order(boxes)
t_order = time.perf_counter() - t
boxes_o = boxes[order]
weight_o = ol["weight"][src][order]
t = time.perf_counter()
gmap = E.OccupancyMap(boxes_o, weight_o, np.zeros((len(boxes_o), 3), np.float32), depth)
t_map = time.perf_counter() - t
out["N1"] = {"file_MB": os.path.getsize(path) / 1e6, "octree_leaves": n_ol, "bvh_leaves": int(len(boxes)),
             "read_ot_s": t_read, "octree_leaves_per_s": n_ol / t_read, "leaf_rule_device_s": t_rule,
             "splitmedian_order_s": t_order, "map_create_s": t_map,
             "note": "read = parse + walk in begin_tree() order on one host thread; leaf rule incl. H2D / D2H; "
                     "map_create = 4-wide SAH flatten (OpenMP) + upload; fixture written by the Python restatement in %.1f s" % t_write}
gmap.close()
os.remove(path)

# ---- the bench's own C2 map for N3 / N4
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
rng = np.random.default_rng(5)
root = np.concatenate([leaves.boxes[:, :3].min(0), leaves.boxes[:, 3:].max(0)])


def quats_of(vps):
    q = []
    for v in range(vps.shape[0]):
        R = vps[v].reshape(3, 4)[:, :3].astype(np.float64)
        qw = np.sqrt(max(1e-12, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
        q.append([qw, (R[2, 1] - R[1, 2]) / (4 * qw), (R[0, 2] - R[2, 0]) / (4 * qw), (R[1, 0] - R[0, 1]) / (4 * qw)])
    q = np.array(q, dtype=np.float32)
    return q / np.linalg.norm(q, axis=1, keepdims=True)


# ---- N3: candidates = exploration steps around existing viewpoints, as the planner samples them
seeds = synth.make_viewpoints(leaves, 64, config_id=71)
cvp = synth.make_viewpoints(leaves, n_cand, config_id=72)
cand_pos = np.stack([cvp[:, 3], cvp[:, 7], cvp[:, 11]], axis=1).astype(np.float32)
cand_pos += rng.normal(0, 0.15, cand_pos.shape).astype(np.float32)
cand_quat = quats_of(cvp)
ex_pos = np.stack([seeds[:, 3], seeds[:, 7], seeds[:, 11]], axis=1).astype(np.float32)
ex_quat = quats_of(seeds)
bbx = np.concatenate([root[:3] - 30, root[3:] + 30]).astype(np.float32)
copts = E.default_candidate_opts(object_bbox=[-0.4, -0.4, -0.3, 0.4, 0.4, 0.3], bvh_bbox=list(bbx), obstacle_free_height=1e30,
                                 exploration_bbox=list(np.concatenate([root[:3] + 2, root[3:] - 2])))
win = (0, w, 0, h)
gmap.set_result_fields(first_pixel=False, dist_sq=False)  # what the planner's facade consumes: (voxel, information)
gmap.generate_viewpoint_entries(cam.K, w, win, cand_pos, cand_quat, ex_pos, ex_quat, copts, copy=False)  # warm up (pools grown)
ts = []
for _ in range(3):
    t = time.perf_counter()
    status, res_c = gmap.generate_viewpoint_entries(cam.K, w, win, cand_pos, cand_quat, ex_pos, ex_quat, copts, copy=False)
    ts.append(time.perf_counter() - t)
counts = np.bincount(status, minlength=5)
cast = int(n_cand - counts[E.CANDIDATE_INVALID_POSITION])
t_c = min(ts)
out["N3"] = {"candidates": n_cand, "existing_viewpoints": int(ex_pos.shape[0]), "ms_per_call": 1e3 * t_c,
             "us_per_candidate": 1e6 * t_c / n_cand, "cast": cast, "rays_per_s": cast * w * h / t_c,
             "accepted": int(counts[E.CANDIDATE_ACCEPTED]), "too_close": int(counts[E.CANDIDATE_TOO_CLOSE]),
             "invalid_position": int(counts[E.CANDIDATE_INVALID_POSITION]),
             "too_few_voxels": int(counts[E.CANDIDATE_TOO_FEW_VOXELS]), "too_little_information": int(counts[E.CANDIDATE_TOO_LITTLE_INFORMATION]),
             "note": "host poses in -> status per candidate + CSR results of all cast candidates on the host, wall clock, best of 3; "
                     "the reference makes the same decisions with one raycast call per candidate (17 ms each on its CUDA path)"}

# ---- N4: real images (depth maps with invalid regions) lower the weights
rvp = synth.make_viewpoints(leaves, n_img, config_id=73)
dm = rng.uniform(1.0, 40.0, size=(n_img, h, w)).astype(np.float32)
for v in range(n_img):
    for _ in range(8):
        y0, x0 = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
        dm[v, y0:y0 + int(rng.integers(16, h // 2)), x0:x0 + int(rng.integers(16, w // 2))] = 0.0
gmap.set_result_fields(first_pixel=True, dist_sq=True)
opts = E.default_opts(raycast_min_range=5.0, raycast_max_range=3.0e38)
gmap.downweight_real_viewpoints(cam.K, w, h, rvp[:2], dm[:2], 0.5, opts)  # warm up
t = time.perf_counter()
_, n_upd = gmap.downweight_real_viewpoints(cam.K, w, h, rvp, dm, 0.5, opts)
t_d = time.perf_counter() - t
out["N4"] = {"images": n_img, "width": w, "height": h, "ms_per_call": 1e3 * t_d, "ms_per_image": 1e3 * t_d / n_img,
             "rays_per_s": n_img * w * h / t_d, "weight_updates": int(n_upd),
             "note": "host poses + depth maps in (H2D of %d MB inside), weights out, wall clock; the reference: one CPU raycast "
                     "pass per image" % (dm.nbytes >> 20)}
gmap.close()
print(json.dumps(out))
