"""Exercises EVERY kernel of libq3dr_raycast.so once on the C2 map, inside a cudaProfilerStart/Stop range (a warm-up
pass runs first, outside it).  Meant for
    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/all_kernels python tools/prof_all_kernels.py
and summarised by tools/ncu_table.py into profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuda.bindings import runtime as rt
from quad3dr_b200 import engine as E, synth

nvp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
leaves = synth.map_1m()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(640, 480)
w, h = 640, 480
win = (0, w, 0, h)
vps = synth.make_viewpoints(leaves, nvp, config_id=2)
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
tv, tn = synth.make_surface_mesh(leaves, cell=0.25)
mesh = E.NormalMesh(tv, tn)
rng = np.random.default_rng(1)
root = np.concatenate([leaves.boxes[:, :3].min(0), leaves.boxes[:, 3:].max(0)]).astype(np.float32)
pos = rng.uniform(root[:3], root[3:] + np.array([0, 0, 5]), size=(500_000, 3)).astype(np.float32)
pos[:250_000, 2] = rng.uniform(0.3, 8, size=250_000)
obj = np.array([-0.6, -0.6, -0.3, 0.6, 0.6, 0.3], np.float32)
qboxes = np.concatenate([pos[:20000] - 0.7, pos[:20000] + 0.7], axis=1).astype(np.float32)
# octree leaves as generateBVHTree sees them: the map's own leaves + as many free ones
oc_center = np.concatenate([(leaves.boxes[:, :3] + leaves.boxes[:, 3:]) / 2] * 2).astype(np.float32)
oc_size = np.concatenate([leaves.boxes[:, 3] - leaves.boxes[:, 0]] * 2).astype(np.float32)
oc_occ = np.concatenate([leaves.occupancy, np.full(leaves.n, 0.1, np.float32)])
oc_cnt = np.concatenate([leaves.observation_count, np.full(leaves.n, 3, np.uint16)]).astype(np.uint32)
dim = 24
gw = rng.uniform(0, 1, size=(dim, dim, dim)).astype(np.float32)
inc = np.float32(root[3] - root[0]) / np.float32(dim + 1)
origins = np.tile(vps[0].reshape(3, 4)[:, 3], (200_000, 1)).astype(np.float32)
dirs = rng.normal(0, 1, (200_000, 3)).astype(np.float32)
dirs[:, 2] = -np.abs(dirs[:, 2])
F = np.float32(1 / 3)


def everything():
    res = gmap.score_viewpoints(cam.K, w, win, vps, copy=False)                        # raycast2<score>, scoring pipeline
    obs = E.ObservedVoxels(gmap)                                                       # setops_kernels.cuh
    obs.add_viewpoint(F, 0)
    obs.add_viewpoint(F, nvp // 2)
    obs.new_information(F, n=nvp)
    E.overlap_information(gmap, 0, np.arange(nvp, dtype=np.uint32))
    v = res.viewpoint(1)
    obs.voxel_set(F, v["leaf"].copy(), v["info"].copy(), add=False)
    obs.close()
    gmap.raycast_window(cam.K, vps[0], 0, w, 0, h, 0.0, 60.0)                          # raycast2<pixels>
    gmap.raycast_unique(cam.K, vps[1], 0, w, 0, h, 0.0, 60.0)                          # + raycast_rays (pixel list)
    gmap.intersect_rays(origins, dirs, 0.0, 60.0)                                      # raycast_rays (ray list)
    gmap.overlap_boxes(qboxes)                                                         # overlap_kernels.cuh
    gmap.valid_object_positions(pos, obj, root, float(root[5]) + 1.0)
    E.octree_leaves_to_boxes(oc_center, oc_size, oc_occ, oc_cnt, root, float(root[5]) + 1.0)   # octree_kernels.cuh
    gmap.update_weights_from_grid(root[:3], inc, [dim, dim, dim], gw, 0.1, leaves.observation_count.astype(np.uint32))
    gmap.update_weights(leaves.weight)
    gmap.score_viewpoints_mesh_normals(mesh, cam.K, w, h, win, vps, copy=False)        # mesh_kept_normals, score<code>
    img = mesh.render_normals(cam.K, w, h, vps[:8])                                    # mesh_render
    gmap.score_viewpoints_normal_images(img, cam.K, w, win, vps[:8], copy=False)       # score<image>


everything()
rt.cudaDeviceSynchronize()
rt.cudaProfilerStart()
everything()
rt.cudaDeviceSynchronize()
rt.cudaProfilerStop()
print("done")
