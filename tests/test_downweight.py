"""Row N4: ViewpointPlannerData::updateWeightsWithRealViewpoints (viewpoint_planner_data.cpp:498-571) on the device
against the oracle's literal restatement: per-pixel raycast without dedup (a12's remove_duplicates = false mode),
depth-map lookup at the hit pixel, sequential weight = max(0, weight - f * information) in scan order, image after
image."""
import ctypes as C

import numpy as np
import pytest

from _common import bits, scene
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def _depth_maps(n, h, w, seed):
    """valid depths with invalid regions of every kind the reference tests for: 0, negative, +inf, -inf; NaN is 'valid'"""
    rng = np.random.default_rng(seed)
    d = rng.uniform(1.0, 40.0, size=(n, h, w)).astype(np.float32)
    for v in range(n):
        for _ in range(6):
            y0, x0 = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
            hh, ww = int(rng.integers(4, h // 2)), int(rng.integers(4, w // 2))
            d[v, y0:y0 + hh, x0:x0 + ww] = rng.choice(np.array([0.0, -1.0, np.inf, -np.inf, np.nan], dtype=np.float32))
    return d


def test_downweight_argument_checks_without_a_gpu():
    lib = E.load_library()
    assert lib.q3dr_map_downweight_real_viewpoints(None, None, 4, 4, None, None, 0, None, 0.5, None, None, None, None, None) \
        == E.Q3DR_ERR_INVALID_ARG


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,max_range", [("small", 160, 120, 3.0e38), ("tiny_unknown", 96, 72, 25.0)])
def test_downweight_equals_oracle(name, w, h, max_range):
    sc = scene(name)
    cam = synth.make_camera(w, h)
    n = 5
    vps = synth.make_viewpoints(sc.leaves, n, config_id=61)
    depth = _depth_maps(n, h, w, 7)
    opts = E.default_opts(raycast_min_range=5.0, raycast_max_range=max_range)  # the reference's defaults for this pass
    from oracle import oracle_py as O

    oopts = O.default_opts(raycast_min_range=5.0, raycast_max_range=-1.0 if max_range > 1e18 else max_range)
    w_ref = sc.leaves.weight.copy()
    total = 0
    for v in range(n):
        w_ref, k = sc.oracle.downweight_real_viewpoint(w_ref, sc.leaves.normal, cam.K, w, h, vps[v], depth[v], 0.5, oopts)
        total += k
    gmap = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    w_gpu, k_gpu = gmap.downweight_real_viewpoints(cam.K, w, h, vps, depth, 0.5, opts)
    assert k_gpu == total and total > 2000
    touched = w_ref != sc.leaves.weight
    assert touched.sum() > 200
    assert np.array_equal(bits(w_gpu[~touched]), bits(sc.leaves.weight[~touched]))  # nothing else moves
    rel = np.abs(w_gpu[touched] - w_ref[touched]) / np.maximum(np.abs(w_ref[touched]), 1e-30)
    assert rel.max() <= 1e-5  # chains of expf / acosf products; most voxels are exact
    assert (bits(w_gpu[touched]) == bits(w_ref[touched])).mean() > 0.5
    assert (w_gpu >= 0).all()
    # the map scores with the lowered weights from now on, and keeps them across a reset (host copy updated)
    res = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:2])
    ref = sc.oracle.score_viewpoint(w_gpu, sc.leaves.normal, cam.K, w, vps[0], height=h)
    assert np.array_equal(res.viewpoint(0)["leaf"], ref["leaf"])
    assert abs(res.viewpoint(0)["total"] - ref["total"]) <= 1e-5 * abs(ref["total"])
    gmap.reset()
    res2 = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:2])
    assert np.array_equal(bits(res2.information), bits(res.information))
    gmap.close()


@pytest.mark.gpu
def test_downweight_with_image_space_normals():
    """enable_opengl = true: the normal of the PIXEL scores every update (viewpoint_planner_data.cpp:535-538), so the
    factor changes along a voxel's chain and the pixel order matters; mesh mode == image mode on the mesh's own images."""
    sc = scene("small")
    w, h = 128, 96
    cam = synth.make_camera(w, h)
    n = 3
    vps = synth.make_viewpoints(sc.leaves, n, config_id=62)
    depth = _depth_maps(n, h, w, 8)
    tv, tn = synth.make_surface_mesh(sc.leaves, cell=2.0)
    mesh = E.NormalMesh(tv, tn)
    images = mesh.render_normals(cam.K, w, h, vps)
    opts = E.default_opts(raycast_max_range=-1.0)
    from oracle import oracle_py as O

    w_ref = sc.leaves.weight.copy()
    total = 0
    for v in range(n):
        w_ref, k = sc.oracle.downweight_real_viewpoint(w_ref, sc.leaves.normal, cam.K, w, h, vps[v], depth[v], 0.4,
                                                       O.default_opts(raycast_max_range=-1.0), normal_image=images[v])
        total += k
    a = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    b = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    w_img, k_img = a.downweight_real_viewpoints(cam.K, w, h, vps, depth, 0.4, opts, normal_images=images)
    w_mesh, k_mesh = b.downweight_real_viewpoints(cam.K, w, h, vps, depth, 0.4, opts, mesh=mesh)
    assert k_img == total == k_mesh and total > 1000
    assert np.array_equal(bits(w_img), bits(w_mesh))
    touched = w_ref != sc.leaves.weight
    rel = np.abs(w_img[touched] - w_ref[touched]) / np.maximum(np.abs(w_ref[touched]), 1e-30)
    assert rel.max() <= 1e-5 and np.array_equal(bits(w_img[~touched]), bits(sc.leaves.weight[~touched]))
    with pytest.raises(E.Q3drError):
        a.downweight_real_viewpoints(cam.K, w, h, vps, depth, 0.4, opts, mesh=mesh, normal_images=images)
    a.close()
    b.close()
