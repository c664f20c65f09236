"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): the smoke test plus a
multi-chunk scoring call (several chunks share the dedup tables and the result pools), the compat entry points, the
mesh-normal scoring pass, the set consumers and the real-image down-weighting.
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
compute-sanitizer is closed on the GPU pool this was developed on (profiles/r02_memcheck.log); the substitute is
    Q3DR_GUARD=1 python tools/sanitize_case.py
which puts guard words behind every device buffer of the map and checks them afterwards (tests/test_guards.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["Q3DR_CHUNK"] = "3"  # 11 viewpoints -> 4 chunks
import numpy as np

import __graft_entry__ as G
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

G.smoke()
leaves = synth.map_tiny(unknown_interior=True)
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
w, h = 96, 72
cam = synth.make_camera(w, h)
vps = synth.make_viewpoints(leaves, 11, config_id=3)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
win = (0, w, 0, h)
a = gmap.score_viewpoints(cam.K, w, win, vps)
gmap.set_result_fields(first_pixel=False, dist_sq=False)
b = gmap.score_viewpoints(cam.K, w, win, vps)
gmap.set_result_fields(first_pixel=True, dist_sq=True)
assert a.n_entries == b.n_entries and np.array_equal(a.leaf, b.leaf)
# a viewpoint exactly on a voxel plane takes the checking instantiation of the kernel
on_plane = vps[:2].copy()
on_plane[:, 3] = leaves.boxes[0, 0]
gmap.score_viewpoints(cam.K, w, win, on_plane)
gmap.raycast_window(cam.K, vps[0], 3, 77, 5, 61, 0.0, 60.0)
gmap.raycast_unique(cam.K, vps[1], 0, w, 0, h, 0.0, 60.0)
rng = np.random.default_rng(0)
gmap.intersect_rays(rng.uniform(-5, 5, (500, 3)), rng.normal(size=(500, 3)), 0.0, -1.0)
tv, tn = synth.make_surface_mesh(leaves, cell=2.0)
mesh = E.NormalMesh(tv, tn)
gmap.score_viewpoints_mesh_normals(mesh, cam.K, w, h, win, vps)
obs = E.ObservedVoxels(gmap)
gmap.score_viewpoints(cam.K, w, win, vps)
obs.new_information(1 / 3, n=11)
obs.add_viewpoint(1 / 3, 2)
E.overlap_information(gmap, 0, [1, 2, 3])
gmap.overlap_boxes(np.array([[-1, -1, 0, 1, 1, 2]], dtype=np.float32))
depth_maps = np.ones((2, h, w), dtype=np.float32)
depth_maps[:, 10:40, 20:70] = 0
gmap.downweight_real_viewpoints(cam.K, w, h, vps[:2], depth_maps, 0.5, mesh=mesh)
if os.environ.get("Q3DR_GUARD", "") not in ("", "0"):
    bad = gmap.check_guards()
    assert bad == 0, f"{bad} device buffers were written past their end"
    print("guard words behind all device buffers intact")
print("sanitize case OK:", a.n_entries, "entries,", int(gmap.stats().chunks), "chunks in the last call")
