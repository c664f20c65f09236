"""Short driver for ncu: one chunk of C2 viewpoints scored with image-space normals from the mesh (row N2).
python tools/prof_mesh.py [n_viewpoints] [cell]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quad3dr_b200 import engine as E, synth

nvp = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cell = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
leaves = synth.map_1m()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(640, 480)
vps = synth.make_viewpoints(leaves, nvp, config_id=2)
tv, tn = synth.make_surface_mesh(leaves, cell=cell)
mesh = E.NormalMesh(tv, tn)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
for _ in range(2):
    res = gmap.score_viewpoints_mesh_normals(mesh, cam.K, 640, 480, (0, 640, 0, 480), vps, copy=False)
st = gmap.stats()
print("entries", int(res.hit_count.sum()), "raycast_ms", st.raycast_ms, "score_ms", st.score_ms, "chunks", st.chunks)
mesh.render_normals(cam.K, 640, 480, vps[:8])
print("render_ms(8 images)", mesh.info().last_render_ms)
