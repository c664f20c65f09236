"""Generates tests/golden/*.npz from the CPU oracle (oracle/q3dr_oracle.cpp).

The reference holds no golden vector for this path and cannot be built here (SURVEY.md sections 4, 8c), so these
fixtures pin OUR oracle against regressions and give the GPU tests inputs + expected outputs that do not depend
on the synthetic generators staying unchanged.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from quad3dr_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, leaves, w, h, nvp, config_id, max_range=60.0):
    tree0 = O.OracleTree(leaves.boxes)
    order = tree0.dfs_order()
    depth = tree0.leaf_depths()[order]
    leaves = leaves.permuted(order)
    tree = O.OracleTree(leaves.boxes)
    assert np.array_equal(tree.dfs_order(), np.arange(leaves.n, dtype=np.uint32))
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(leaves, nvp, config_id=config_id)
    out = dict(boxes=leaves.boxes, weight=leaves.weight, normal=leaves.normal, depth=depth.astype(np.uint32),
               input_order=order, K=cam.K, width=np.int32(w), height=np.int32(h), Rt=vps, max_range=np.float32(max_range))
    leaf_px, dsq_px, xyz_px, depth_px = [], [], [], []
    offs, l_leaf, l_info, l_pix, l_dsq, totals, totals32, hits = [0], [], [], [], [], [], [], []
    for v in range(nvp):
        r = tree.raycast_window(cam.K, vps[v], 0, w, 0, h, 0.0, max_range)
        leaf_px.append(r["leaf"]); dsq_px.append(r["dsq"].view(np.uint32)); xyz_px.append(r["xyz"]); depth_px.append(r["depth"])
        s = tree.score_viewpoint(leaves.weight, leaves.normal, cam.K, w, vps[v], height=h,
                                 opts=O.default_opts(raycast_max_range=max_range))
        l_leaf.append(s["leaf"]); l_info.append(s["info"]); l_pix.append(s["pixel"]); l_dsq.append(s["dsq"].view(np.uint32))
        offs.append(offs[-1] + len(s["leaf"])); totals.append(s["total"]); totals32.append(s["total_f32"]); hits.append(s["hit_count"])
    out.update(px_leaf=np.stack(leaf_px), px_dsq_bits=np.stack(dsq_px), px_xyz=np.stack(xyz_px), px_depth=np.stack(depth_px),
               offsets=np.asarray(offs, np.uint64), leaf=np.concatenate(l_leaf), info=np.concatenate(l_info),
               pixel=np.concatenate(l_pix), dsq_bits=np.concatenate(l_dsq), total=np.asarray(totals, np.float32),
               total_f32=np.asarray(totals32, np.float32), hit_count=np.asarray(hits, np.uint32))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, leaves.n, "leaves", nvp, "viewpoints", os.path.getsize(path) // 1024, "KiB")


def make_mesh(name, leaves, w, h, nvp, config_id, cell):
    """Image-space normals (row N2): mesh, poses, the oracle's normal images and the scores computed with them."""
    tree0 = O.OracleTree(leaves.boxes)
    order = tree0.dfs_order()
    depth = tree0.leaf_depths()[order]
    leaves = leaves.permuted(order)
    tree = O.OracleTree(leaves.boxes)
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(leaves, nvp, config_id=config_id)
    tv, tn = synth.make_surface_mesh(leaves, cell=cell)
    images = np.stack([O.mesh_render_normals(tv, tn, cam.K, w, h, vps[v]) for v in range(nvp)])
    offs, l_leaf, l_info, l_pix, totals = [0], [], [], [], []
    for v in range(nvp):
        s = tree.score_viewpoint_image_normals(leaves.weight, images[v], cam.K, w, vps[v])
        l_leaf.append(s["leaf"]); l_info.append(s["info"]); l_pix.append(s["pixel"])
        offs.append(offs[-1] + len(s["leaf"])); totals.append(s["total"])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, boxes=leaves.boxes, weight=leaves.weight, depth=depth.astype(np.uint32), K=cam.K,
                        width=np.int32(w), height=np.int32(h), Rt=vps, tri_vertices=tv, tri_normals=tn, images=images,
                        offsets=np.asarray(offs, np.uint64), leaf=np.concatenate(l_leaf), info=np.concatenate(l_info),
                        pixel=np.concatenate(l_pix), total=np.asarray(totals, np.float32))
    print(name, leaves.n, "leaves", tv.shape[0], "triangles", nvp, "viewpoints", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mesh":
        make_mesh("tiny_mesh_normals", synth.map_tiny(), 64, 48, 4, config_id=106, cell=1.0)
        sys.exit(0)
    make("tiny_c1like", synth.map_tiny(), 64, 48, 8, config_id=101)
    make("tiny_unknown_c4like", synth.map_tiny(unknown_interior=True), 64, 48, 6, config_id=104)
    make("tiny_shortrange", synth.map_tiny(), 48, 36, 4, config_id=105, max_range=6.0)
