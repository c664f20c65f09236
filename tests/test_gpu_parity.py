"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle on the same inputs.

Bar: leaf ids, depths, dist_sq bits, intersection bits, first pixels and hit sets bit-exact; per-voxel
information and per-viewpoint totals within 1e-5 relative (only expf/acosf differ between libm and CUDA).
"""
import numpy as np
import pytest

from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

from _common import bits, scene

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: scores within 1e-5 relative


def _assert_pixels_equal(got, ref):
    assert np.array_equal(got["leaf"], ref["leaf"])
    assert np.array_equal(bits(got["dist_sq"]), bits(ref["dsq"]))
    assert np.array_equal(got["depth"], ref["depth"])
    # intersection: numerically equal (sign of zero is not part of the contract)
    assert np.array_equal(got["intersection"], ref["xyz"])


@pytest.mark.parametrize("name,w,h,nvp", [("tiny", 64, 48, 8), ("tiny_unknown", 64, 48, 4), ("small", 160, 120, 6)])
def test_raycast_window_matches_oracle(name, w, h, nvp):
    sc = scene(name)
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, nvp, config_id=7)
    for v in range(nvp):
        for max_range in (60.0, -1.0):
            ref = sc.oracle.raycast_window(cam.K, vps[v], 0, w, 0, h, 0.0, max_range)
            got = sc.gpu_map.raycast_window(cam.K, vps[v], 0, w, 0, h, 0.0, max_range)
            _assert_pixels_equal(got, ref)
            xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
            assert np.array_equal(got["screen_x"], xs.ravel())
            assert np.array_equal(got["screen_y"], ys.ravel())


def test_raycast_subwindow_and_ragged_tiles():
    sc = scene("tiny")
    cam = synth.make_camera(67, 45)  # not a multiple of the 8x4 packet
    vps = synth.make_viewpoints(sc.leaves, 3, config_id=11)
    for v in range(3):
        for (x0, x1, y0, y1) in [(0, 67, 0, 45), (5, 38, 7, 30), (66, 67, 44, 45)]:
            ref = sc.oracle.raycast_window(cam.K, vps[v], x0, x1, y0, y1, 0.0, 60.0)
            got = sc.gpu_map.raycast_window(cam.K, vps[v], x0, x1, y0, y1, 0.0, 60.0)
            _assert_pixels_equal(got, ref)


@pytest.mark.parametrize("name,w,h,nvp", [("tiny", 64, 48, 8), ("small_unknown", 160, 120, 8)])
@pytest.mark.parametrize("ignore_zero", [1, 0])
def test_score_viewpoints_matches_oracle(name, w, h, nvp, ignore_zero):
    sc = scene(name)
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, nvp, config_id=3)
    res = sc.gpu_map.score_viewpoints(cam.K, w, (0, w, 0, h), vps, E.default_opts(ignore_voxels_with_zero_information=ignore_zero))
    assert res.n_viewpoints == nvp
    for v in range(nvp):
        ref = sc.oracle.score_viewpoint(
            sc.leaves.weight, sc.leaves.normal, cam.K, w, vps[v], height=h,
            opts=O.default_opts(ignore_voxels_with_zero_information=ignore_zero),
        )
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], ref["leaf"]), f"hit set differs for viewpoint {v}"
        assert np.array_equal(got["pixel"], ref["pixel"])
        assert np.array_equal(bits(got["dsq"]), bits(ref["dsq"]))
        assert got["hit_count"] == ref["hit_count"]
        np.testing.assert_allclose(got["info"], ref["info"], rtol=REL_TOL, atol=0)
        assert abs(got["total"] - ref["total"]) <= REL_TOL * abs(ref["total"])


# ----------------------------------------------------------------------------------------------------
# golden fixtures (inputs + oracle outputs committed under tests/golden/)
# ----------------------------------------------------------------------------------------------------
import glob
import os

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if "mesh_normals" not in p)  # the mesh fixture has its own tests (test_mesh_normals.py)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_reproduces_golden(path):
    g = np.load(path)
    m = E.OccupancyMap(g["boxes"], g["weight"], g["normal"], g["depth"])
    w, h, mr = int(g["width"]), int(g["height"]), float(g["max_range"])
    for v in range(g["Rt"].shape[0]):
        px = m.raycast_window(g["K"], g["Rt"][v], 0, w, 0, h, 0.0, mr)
        assert np.array_equal(px["leaf"], g["px_leaf"][v])
        assert np.array_equal(bits(px["dist_sq"]), g["px_dsq_bits"][v])
        assert np.array_equal(px["depth"], g["px_depth"][v])
        assert np.array_equal(px["intersection"], g["px_xyz"][v])
    res = m.score_viewpoints(g["K"], w, (0, w, 0, h), g["Rt"], E.default_opts(raycast_max_range=mr))
    assert np.array_equal(res.offsets, g["offsets"])
    assert np.array_equal(res.leaf, g["leaf"])
    assert np.array_equal(res.first_pixel, g["pixel"])
    assert np.array_equal(bits(res.dist_sq), g["dsq_bits"])
    assert np.array_equal(res.hit_count, g["hit_count"])
    np.testing.assert_allclose(res.information, g["info"], rtol=REL_TOL)
    np.testing.assert_allclose(res.total, g["total"], rtol=REL_TOL)
    m.close()


# ----------------------------------------------------------------------------------------------------
# edge cases the domain has: axis-parallel rays (0 * inf = NaN hazard), origins on voxel faces / grid planes,
# origins inside leaves, duplicated boxes (exact ties), ranges, far-from-origin coordinates (coarser rounding
# => many equal dist_sq), degenerate maps
# ----------------------------------------------------------------------------------------------------

def _compare_rays(sc_or_boxes, origins, dirs, max_range=-1.0, gmap=None, oracle=None):
    ref = oracle.intersect_rays(origins, dirs, 0.0, max_range)
    got = gmap.intersect_rays(origins, dirs, 0.0, max_range)
    assert np.array_equal(got["leaf"], ref["leaf"])
    assert np.array_equal(bits(got["dist_sq"]), bits(ref["dsq"]))
    assert np.array_equal(got["depth"], ref["depth"])
    assert np.array_equal(got["intersection"], ref["xyz"])
    return ref


def test_arbitrary_rays_including_axis_parallel_and_on_grid_origins():
    sc = scene("tiny_unknown")
    rng = np.random.default_rng(5)
    n = 4096
    o = rng.uniform(-9, 9, size=(n, 3)).astype(np.float32)
    o[:, 2] = rng.uniform(0.1, 6, size=n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    # axis-parallel and plane-parallel directions: exact zeros in the direction
    d[0:256] = [1, 0, 0]
    d[256:512] = [0, -1, 0]
    d[512:768] = [0, 0, -1]
    d[768:1024, 2] = 0.0
    d[1024:1280, 0] = 0.0
    # origins exactly on grid planes / voxel faces (multiples of the resolution)
    o[128:256] = np.round(o[128:256] * 4) / 4
    o[384:512] = np.round(o[384:512] * 4) / 4
    o[640:768] = np.round(o[640:768] * 4) / 4
    o[1280:2048] = np.round(o[1280:2048] * 4) / 4
    o[2048:2304, 2] = 0.0  # exactly on the ground top face
    ref = _compare_rays(None, o, d, -1.0, sc.gpu_map, sc.oracle)
    assert (ref["leaf"] >= 0).sum() > n // 4
    _compare_rays(None, o, d, 3.0, sc.gpu_map, sc.oracle)


def test_origin_inside_leaves_and_on_faces():
    sc = scene("tiny_unknown")
    b = sc.leaves.boxes
    big = np.where((b[:, 3] - b[:, 0]) >= 1.0)[0][:64]  # unknown-space cubes
    assert len(big) > 0
    rng = np.random.default_rng(6)
    c = (b[big, :3] + b[big, 3:]) / 2
    inside = c + rng.uniform(-0.2, 0.2, size=c.shape).astype(np.float32)
    on_face = c.copy()
    on_face[:, 0] = b[big, 3]  # exactly on the +x face: inclusive bounds => still "inside"
    corner = b[big, 3:].copy()
    o = np.concatenate([inside, on_face, corner]).astype(np.float32)
    d = rng.normal(size=o.shape).astype(np.float32)
    ref = _compare_rays(None, o, d, -1.0, sc.gpu_map, sc.oracle)
    assert np.all(ref["dsq"][: len(big)] == 0)
    # cameras placed on grid points / inside unknown space: window API + scoring
    cam = synth.make_camera(64, 48)
    vps = synth.make_viewpoints(sc.leaves, 6, config_id=31)
    vps[:, [3, 7, 11]] = np.round(vps[:, [3, 7, 11]] * 4) / 4  # origins exactly on voxel face planes
    vps[4, [3, 7, 11]] = inside[0]
    vps[5, [3, 7, 11]] = on_face[1]
    for v in range(6):
        ref = sc.oracle.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, 60.0)
        got = sc.gpu_map.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, 60.0)
        _assert_pixels_equal(got, ref)
    res = sc.gpu_map.score_viewpoints(cam.K, 64, (0, 64, 0, 48), vps)
    for v in range(6):
        ref = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, cam.K, 64, vps[v], height=48)
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"])


def test_axis_aligned_camera_poses():
    """Identity-like rotations put exact zeros into many ray directions (centre row / column)."""
    sc = scene("tiny")
    cam = synth.make_camera(64, 48)
    poses = []
    for R in (np.eye(3), np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]]), np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]]),
              np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]])):
        for t in ([0.5, -6.0, 2.0], [1.0, 2.0, 9.0]):
            poses.append(np.concatenate([np.asarray(R, np.float32), np.asarray(t, np.float32)[:, None]], axis=1).reshape(12))
    for Rt in poses:
        ref = sc.oracle.raycast_window(cam.K, Rt, 0, 64, 0, 48, 0.0, 60.0)
        got = sc.gpu_map.raycast_window(cam.K, Rt, 0, 64, 0, 48, 0.0, 60.0)
        _assert_pixels_equal(got, ref)


def test_duplicate_and_overlapping_boxes_tie_break():
    rng = np.random.default_rng(8)
    c = rng.integers(-6, 6, size=(400, 3)).astype(np.float32)
    s = rng.choice([0.5, 1.0, 2.0], size=(400, 1)).astype(np.float32)
    boxes = np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32)
    boxes = np.concatenate([boxes, boxes[:150], boxes[:50]])  # exact duplicates => exact dist_sq ties
    order, depth = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    assert np.array_equal(oracle.dfs_order(), np.arange(len(boxes)))
    gmap = E.OccupancyMap(boxes, None, None, depth)
    o = rng.uniform(-12, 12, size=(8192, 3)).astype(np.float32)
    d = rng.normal(size=(8192, 3)).astype(np.float32)
    d = (rng.uniform(-6, 6, size=(8192, 3)) - o).astype(np.float32)  # aim into the cloud of boxes
    ref = _compare_rays(None, o, d, -1.0, gmap, oracle)
    hit = ref["leaf"][ref["leaf"] >= 0]
    assert len(hit) > 2000
    # independent of the oracle: among bit-identical boxes the highest index must have been reported
    _, inv = np.unique(boxes, axis=0, return_inverse=True)
    inv = inv.ravel()
    top = np.zeros(inv.max() + 1, dtype=np.int64)
    np.maximum.at(top, inv, np.arange(len(boxes)))
    assert np.array_equal(hit, top[inv[hit]])
    assert (np.bincount(inv)[inv[hit]] > 1).sum() > 500  # plenty of real ties were exercised
    gmap.close()


def test_far_from_origin_coordinates():
    """Map and cameras translated by kilometres: fp32 rounding gets coarse, equal dist_sq become common."""
    leaves = synth.map_tiny()
    shift = np.array([4096.0, -2048.0, 1024.0], np.float32)
    boxes = (leaves.boxes + np.concatenate([shift, shift])).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    gmap = E.OccupancyMap(boxes, leaves.weight[order], leaves.normal[order], depth)
    cam = synth.make_camera(96, 72)
    vps = synth.make_viewpoints(leaves, 6, config_id=41)
    vps[:, [3, 7, 11]] += shift
    for v in range(6):
        ref = oracle.raycast_window(cam.K, vps[v], 0, 96, 0, 72, 0.0, 60.0)
        got = gmap.raycast_window(cam.K, vps[v], 0, 96, 0, 72, 0.0, 60.0)
        _assert_pixels_equal(got, ref)
    gmap.close()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9])
def test_tiny_maps(n):
    rng = np.random.default_rng(n)
    c = rng.uniform(-3, 3, size=(n, 3)).astype(np.float32)
    boxes = np.concatenate([c - 0.5, c + 0.5], axis=1).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    gmap = E.OccupancyMap(boxes, None, None, depth)
    o = rng.uniform(-8, 8, size=(512, 3)).astype(np.float32)
    d = (c[rng.integers(0, n, size=512)] - o + rng.normal(scale=0.3, size=(512, 3))).astype(np.float32)
    _compare_rays(None, o, d, -1.0, gmap, oracle)
    gmap.close()


def test_empty_inputs_and_bad_arguments():
    sc = scene("tiny")
    cam = synth.make_camera(64, 48)
    res = sc.gpu_map.score_viewpoints(cam.K, 64, (0, 64, 0, 48), np.zeros((0, 12), np.float32))
    assert res.n_viewpoints == 0 and res.n_entries == 0 and np.array_equal(res.offsets, [0])
    assert len(sc.gpu_map.intersect_rays(np.zeros((0, 3)), np.zeros((0, 3)))) == 0
    with pytest.raises(E.Q3drError) as ei:
        sc.gpu_map.raycast_window(cam.K, np.zeros(12, np.float32), 5, 5, 0, 48)
    assert ei.value.status == E.Q3DR_ERR_INVALID_ARG
    # a camera looking at the sky: all misses, empty hit list, zero score
    up = np.array([1, 0, 0, 0.3, 0, -1, 0, 0.3, 0, 0, 1, 30.0], np.float32)
    px = sc.gpu_map.raycast_window(cam.K, up, 0, 64, 0, 48, 0.0, 60.0)
    assert np.all(px["leaf"] == -1) and np.all(px["dist_sq"] == 0) and np.all(px["intersection"] == 0)
    res = sc.gpu_map.score_viewpoints(cam.K, 64, (0, 64, 0, 48), up[None])
    assert res.n_entries == 0 and res.total[0] == 0 and res.hit_count[0] == 0


def test_raycast_unique_matches_first_pixel_dedup():
    sc = scene("tiny")
    cam = synth.make_camera(64, 48)
    vps = synth.make_viewpoints(sc.leaves, 3, config_id=13)
    for v in range(3):
        ref = sc.oracle.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, 60.0)
        uniq, first = np.unique(ref["leaf"], return_index=True)
        first, uniq = first[uniq >= 0], uniq[uniq >= 0]
        got = sc.gpu_map.raycast_unique(cam.K, vps[v], 0, 64, 0, 48, 0.0, 60.0)
        assert np.array_equal(got["leaf"], uniq)
        assert np.array_equal(got["screen_x"], (first % 64).astype(np.float32))
        assert np.array_equal(got["screen_y"], (first // 64).astype(np.float32))
        assert np.array_equal(bits(got["dist_sq"]), bits(ref["dsq"][first]))
        assert np.array_equal(got["intersection"], ref["xyz"][first])
        assert np.array_equal(got["depth"], ref["depth"][first])


def test_update_weights_and_options():
    sc = scene("tiny")
    cam = synth.make_camera(64, 48)
    vps = synth.make_viewpoints(sc.leaves, 2, config_id=17)
    w2 = (sc.leaves.weight * 0.5 + 0.125).astype(np.float32)
    w2[::7] = 0.0  # zero-weight voxels drop out of the voxel set when ignore_zero = 1
    gmap = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    gmap.update_weights(w2)
    for kw in (dict(), dict(incidence_ignore_dot_product_sign=1), dict(incidence_angle_threshold_degrees=10.0,
               incidence_angle_inv_falloff_factor_degrees=45.0, voxel_sensor_size_ratio_threshold=0.02,
               voxel_sensor_size_ratio_inv_falloff_factor=0.01), dict(raycast_max_range=9.0)):
        res = gmap.score_viewpoints(cam.K, 64, (0, 64, 0, 48), vps, E.default_opts(**kw))
        for v in range(2):
            ref = sc.oracle.score_viewpoint(w2, sc.leaves.normal, cam.K, 64, vps[v], height=48, opts=O.default_opts(**kw))
            got = res.viewpoint(v)
            assert np.array_equal(got["leaf"], ref["leaf"])
            assert got["hit_count"] == ref["hit_count"]
            np.testing.assert_allclose(got["info"], ref["info"], rtol=REL_TOL)
            assert abs(got["total"] - ref["total"]) <= REL_TOL * max(abs(ref["total"]), 1e-30)
    gmap.reset()
    res2 = gmap.score_viewpoints(cam.K, 64, (0, 64, 0, 48), vps, E.default_opts(raycast_max_range=9.0))
    assert np.array_equal(res2.leaf, res.leaf) and np.array_equal(bits(res2.information), bits(res.information))
    gmap.close()


def test_full_size_map_full_resolution_against_oracle():
    """C2's map (1,028,168 leaves) at 640x480 for a few viewpoints: per-pixel and per-viewpoint parity."""
    sc = scene("1m")
    cam = synth.make_camera(640, 480)
    vps = synth.make_viewpoints(sc.leaves, 4, config_id=2)
    res = sc.gpu_map.score_viewpoints(cam.K, 640, (0, 640, 0, 480), vps)
    for v in range(4):
        ref = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, cam.K, 640, vps[v], height=480)
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], ref["leaf"])
        assert np.array_equal(got["pixel"], ref["pixel"])
        assert np.array_equal(bits(got["dsq"]), bits(ref["dsq"]))
        np.testing.assert_allclose(got["info"], ref["info"], rtol=REL_TOL)
        assert abs(got["total"] - ref["total"]) <= REL_TOL * abs(ref["total"])
    refpx = sc.oracle.raycast_window(cam.K, vps[0], 0, 640, 0, 480, 0.0, 60.0)
    gotpx = sc.gpu_map.raycast_window(cam.K, vps[0], 0, 640, 0, 480, 0.0, 60.0)
    _assert_pixels_equal(gotpx, refpx)


# ----------------------------------------------------------------------------------------------------
# paths of the camera-window kernel (raycast2_kernels.cuh): analytic pruning bound at coarse fp32 resolution,
# mixed-octant packets, packets outside the fast loops' proven domain (exact fallback)
# ----------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("shift", [[1.0e5, -3.0e5, 2.0e5], [2.0e6, 1.0e6, -4.0e6], [3.0e7, 0.0, 1.0e7]])
def test_analytic_bound_at_coarse_resolution(shift):
    """Kilometres to tens of thousands of kilometres from the origin: ulp(coordinate) grows to metres, dist_sq values
    collide massively, and the pruning bound's |origin| term carries the correctness."""
    leaves = synth.map_tiny(unknown_interior=True)
    shift = np.asarray(shift, np.float32)
    boxes = (leaves.boxes + np.concatenate([shift, shift])).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    gmap = E.OccupancyMap(boxes, leaves.weight[order], leaves.normal[order], depth)
    cam = synth.make_camera(96, 72)
    vps = synth.make_viewpoints(leaves, 5, config_id=43)
    vps[:, [3, 7, 11]] += shift
    for v in range(5):
        for max_range in (60.0, -1.0, 7.0):
            ref = oracle.raycast_window(cam.K, vps[v], 0, 96, 0, 72, 0.0, max_range)
            got = gmap.raycast_window(cam.K, vps[v], 0, 96, 0, 72, 0.0, max_range)
            _assert_pixels_equal(got, ref)
    res = gmap.score_viewpoints(cam.K, 96, (0, 96, 0, 72), vps)
    for v in range(5):
        ref = oracle.score_viewpoint(leaves.weight[order], leaves.normal[order], cam.K, 96, vps[v], height=72)
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"])
        assert np.array_equal(bits(got["dsq"]), bits(ref["dsq"]))
    gmap.close()


def test_tiny_scale_scene():
    """Millimetre-sized boxes close to the camera: the bound's absolute term (2^-13) is large against the scene."""
    rng = np.random.default_rng(21)
    c = rng.uniform(-0.02, 0.02, size=(3000, 3)).astype(np.float32)
    s = rng.choice([0.0005, 0.001, 0.002], size=(3000, 1)).astype(np.float32)
    boxes = np.concatenate([c - s, c + s], axis=1).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    gmap = E.OccupancyMap(boxes, None, None, depth)
    cam = synth.make_camera(64, 48)
    for k in range(4):
        eye = rng.uniform(-0.05, 0.05, size=3).astype(np.float32)
        eye[2] = 0.06
        f = -eye / np.linalg.norm(eye)
        r = np.cross(f, [0.0, 1.0, 0.1]); r /= np.linalg.norm(r)
        u = np.cross(f, r)
        Rt = np.concatenate([np.stack([r, u, f], axis=1), eye[:, None]], axis=1).astype(np.float32).reshape(12)
        for max_range in (-1.0, 0.08):
            ref = oracle.raycast_window(cam.K, Rt, 0, 64, 0, 48, 0.0, max_range)
            got = gmap.raycast_window(cam.K, Rt, 0, 64, 0, 48, 0.0, max_range)
            _assert_pixels_equal(got, ref)
        assert (ref["leaf"] >= 0).sum() > 200
    gmap.close()


def test_mixed_octant_packets_everywhere():
    """A wide-angle camera looking straight down an axis: the optical axis crosses two sign planes, so every packet
    around the image centre mixes octants; a tiny image makes almost all packets mixed or ragged."""
    sc = scene("tiny_unknown")
    for (w, h) in [(16, 16), (24, 8), (9, 23)]:
        cam = synth.make_camera(w, h)
        K = cam.K.copy()
        K[0] = K[1] = 0.35 * w  # ~110 degrees
        for R in (np.eye(3), np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]]), np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]])):
            for t in ([0.37, -6.21, 2.13], [1.11, 2.52, 9.3], [-3.3, 0.4, 1.7]):
                Rt = np.concatenate([np.asarray(R, np.float32), np.asarray(t, np.float32)[:, None]], axis=1).reshape(12)
                ref = sc.oracle.raycast_window(K, Rt, 0, w, 0, h, 0.0, 60.0)
                got = sc.gpu_map.raycast_window(K, Rt, 0, w, 0, h, 0.0, 60.0)
                _assert_pixels_equal(got, ref)
                res = sc.gpu_map.score_viewpoints(K, w, (0, w, 0, h), Rt[None])
                sref = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, K, w, Rt, height=h)
                assert np.array_equal(res.viewpoint(0)["leaf"], sref["leaf"])
                assert np.array_equal(res.viewpoint(0)["pixel"], sref["pixel"])


def test_packets_outside_the_fast_domain_fall_back_to_the_exact_path():
    """Huge coordinates in the map (not 'bounded'), a huge max_range and a scaled (non-rotation) pose matrix: the
    fast loops' preconditions fail and every packet takes the literal reference test -- results must not change."""
    leaves = synth.map_tiny()
    boxes = leaves.boxes.copy()
    far = np.array([[1e20, 1e20, 1e20, 1.0001e20, 1.0001e20, 1.0001e20]], np.float32)  # one absurdly far voxel
    boxes = np.concatenate([boxes, far]).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    gmap = E.OccupancyMap(boxes, None, None, depth)
    cam = synth.make_camera(64, 48)
    vps = synth.make_viewpoints(leaves, 3, config_id=47)
    for v in range(3):
        for max_range in (60.0, -1.0, 3e19):
            ref = oracle.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, max_range)
            got = gmap.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, max_range)
            _assert_pixels_equal(got, ref)
    gmap.close()
    # bounded map, but a pose whose rotation block is scaled by 1e-18 (squared norm of the direction ~ 1e-36) and a
    # huge max_range: still the reference's answer
    sc = scene("tiny")
    for v in range(3):
        Rt = vps[v].reshape(3, 4).copy()
        Rt[:, :3] *= np.float32(1e-18)
        ref = sc.oracle.raycast_window(cam.K, Rt.reshape(12), 0, 64, 0, 48, 0.0, 60.0)
        got = sc.gpu_map.raycast_window(cam.K, Rt.reshape(12), 0, 64, 0, 48, 0.0, 60.0)
        _assert_pixels_equal(got, ref)
        ref = sc.oracle.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, 5e18)
        got = sc.gpu_map.raycast_window(cam.K, vps[v], 0, 64, 0, 48, 0.0, 5e18)
        _assert_pixels_equal(got, ref)


def test_scoring_with_many_chunks_and_dense_hits():
    """More viewpoints than one chunk holds (Q3DR_CHUNK is not set: the library picks), dense hit lists from a camera
    close to a wall, both filter settings; the flat block numbering must reproduce per-viewpoint order and counts."""
    sc = scene("small_unknown")
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 700, config_id=59)
    for iz in (1, 0):
        res = sc.gpu_map.score_viewpoints(cam.K, w, (0, w, 0, h), vps, E.default_opts(ignore_voxels_with_zero_information=iz))
        assert res.n_viewpoints == 700 and res.offsets[-1] == res.n_entries
        for v in (0, 1, 255, 256, 257, 511, 699):
            ref = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, cam.K, w, vps[v], height=h,
                                            opts=O.default_opts(ignore_voxels_with_zero_information=iz))
            got = res.viewpoint(v)
            assert np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"])
            assert got["hit_count"] == ref["hit_count"]
            np.testing.assert_allclose(got["info"], ref["info"], rtol=REL_TOL)
            assert abs(got["total"] - ref["total"]) <= REL_TOL * max(abs(ref["total"]), 1e-30)


def test_fuzz_parity_short():
    """A short run of tools/fuzz_parity.py (random degenerate scenes, cameras on faces / inside boxes / axis-aligned,
    random windows and ranges, random ray lists): every pixel and every voxel set equal to the oracle's."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "12", "7"], capture_output=True, text=True,
                       timeout=300, cwd=root)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout + r.stderr
