"""Row N2 throughput on one B200: scoring with image-space normals from a mesh (fused: one mesh ray per kept pixel)
next to the voxel-normal path on the same map and poses, the full normal-image render, and the oracle's brute force on
a small sample.  python tools/bench_mesh_normals.py [workload] [cell] -> one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import oracle_py as O
from quad3dr_b200 import engine as E, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
cell = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
mk, nvp, w, h = synth.CONFIGS[wl]
nvp = min(nvp, 1000)
leaves = mk()
order, depth = E.splitmedian_order(leaves.boxes)
leaves = leaves.permuted(order)
cam = synth.make_camera(w, h)
vps = synth.make_viewpoints(leaves, nvp, config_id=2)
tv, tn = synth.make_surface_mesh(leaves, cell=cell)
t = time.perf_counter()
mesh = E.NormalMesh(tv, tn)
t_build = time.perf_counter() - t
gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
d_rt = torch.from_numpy(vps).cuda()
win = (0, w, 0, h)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, steps=4, warm=2):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms))


ms_voxel = timed(lambda: gmap.score_viewpoints_device(cam.K, w, win, d_rt.data_ptr(), nvp, stream=stream.cuda_stream))
st_v = gmap.stats()
ms_mesh = timed(lambda: gmap.score_viewpoints_mesh_normals_device(mesh, cam.K, w, h, win, d_rt.data_ptr(), nvp,
                                                                    stream=stream.cuda_stream))
st_m = gmap.stats()
res = gmap.result_to_host()
entries = int(res.hit_count.sum())
rays = nvp * w * h
# full images for a few poses (the GUI / computePoissonMeshNormalVector(position) use)
nr = 16
mesh.render_normals(cam.K, w, h, vps[:nr])
img = mesh.render_normals(cam.K, w, h, vps[:nr])
render_ms = mesh.info().last_render_ms
# parity on a sample against the oracle: brute force over all triangles only at the kept pixels of 2 viewpoints
tree = O.OracleTree(leaves.boxes)
ns = int(os.environ.get("PARITY_VIEWPOINTS", "1"))
t = time.perf_counter()
ok = True
max_rel = 0.0
oracle_rays = 0
for v in range(ns):
    plain = tree.score_viewpoint(leaves.weight, leaves.normal, cam.K, w, vps[v], height=h,
                                 opts=O.default_opts(ignore_voxels_with_zero_information=0))
    pix = plain["pixel"]
    codes = O.mesh_cast_pixels(tv, tn, cam.K, w, h, vps[v], pix % w, pix // w)
    oracle_rays += len(pix)
    sparse = np.full((h, w), 0xFF808080, np.uint32)
    sparse[pix // w, pix % w] = codes
    ok &= bool(np.array_equal(img[v][pix // w, pix % w], codes))
    ref = tree.score_viewpoint_image_normals(leaves.weight, sparse, cam.K, w, vps[v])
    got = res.viewpoint(v)
    ok &= bool(np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"]))
    max_rel = max(max_rel, abs(got["total"] - ref["total"]) / abs(ref["total"]))
t_oracle = time.perf_counter() - t
print(json.dumps({
    "workload": wl, "viewpoints": nvp, "rays": rays, "n_leaves": int(leaves.n), "triangles": int(tv.shape[0]),
    "mesh": {"nodes": int(mesh.info().n_nodes), "depth": int(mesh.info().tree_depth),
             "device_MB": mesh.info().device_bytes / 1e6, "host_build_s": t_build},
    "voxel_normals": {"ms": ms_voxel, "rays_per_s": rays / ms_voxel * 1e3, "score_ms": st_v.score_ms},
    "mesh_normals_fused": {"ms": ms_mesh, "rays_per_s": rays / ms_mesh * 1e3, "score_ms": st_m.score_ms,
                           "mesh_rays": entries, "extra_ms": ms_mesh - ms_voxel,
                           "mesh_rays_per_s": entries / max(ms_mesh - ms_voxel, 1e-6) * 1e3},
    "render_full_images": {"images": nr, "ms": render_ms, "pixels_per_s": nr * w * h / render_ms * 1e3},
    "parity_sample": {"viewpoints": ns, "ok": ok, "max_rel_total_err": max_rel, "oracle_mesh_rays": oracle_rays,
                      "oracle_s": t_oracle},
}))
