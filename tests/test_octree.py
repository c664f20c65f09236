"""Row N1: generateBVHTree's leaf rule (octree leaves -> BVH leaf boxes) and updateWeights' grid writes."""
import numpy as np
import pytest

from _common import bits
from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def _octree_leaves(seed, n):
    """Octomap-style leaves: resolution 0.25, depth-16 keys; sizes r * 2^j; a mix of occupied / free / unknown."""
    rng = np.random.default_rng(seed)
    r = 0.25
    j = rng.choice([0, 0, 0, 1, 2, 3, 5], size=n)
    size = (r * 2.0 ** j).astype(np.float32)
    cell = rng.integers(-200, 200, size=(n, 3))
    cell[:, 2] = rng.integers(-4, 120, size=n)
    center = ((np.floor(cell / 2.0 ** j[:, None]) + 0.5) * size[:, None]).astype(np.float32)
    occupancy = rng.choice([0.12, 0.5, 0.51, 0.52, 0.97], size=n).astype(np.float32)
    count = rng.choice([0, 0, 1, 2, 7, 30], size=n).astype(np.uint32)
    return center, size, occupancy, count


def test_oracle_leaf_rule_known_answers():
    bb = np.array([-10, -10, -1, 10, 10, 20], np.float32)
    c = np.array([[0.125, 0.125, 0.125]], np.float32)
    one = lambda occ, cnt, ctr=c, s=0.25, h=15.0: O.octree_leaves_to_boxes(ctr, [s], [occ], [cnt], bb, h)
    assert len(one(0.97, 3)[0]) == 1                    # occupied
    assert len(one(0.2, 3)[0]) == 0                     # free and known: dropped
    assert len(one(0.2, 0)[0]) == 1                     # free but unknown: kept
    assert len(one(0.51, 1)[0]) == 0                    # threshold is strict: 0.51 is not occupied
    np.testing.assert_array_equal(one(0.97, 1)[0][0], [0, 0, 0, 0.25, 0.25, 0.25])
    # clipped by bvh_bbox and cropped at the obstacle-free height
    big = np.array([[8.0, 8.0, 16.0]], np.float32)
    np.testing.assert_array_equal(one(0.97, 1, big, 8.0)[0][0], [4, 4, 12, 10, 10, 15])
    assert len(one(0.97, 1, np.array([[30.0, 0, 0]], np.float32), 1.0)[0]) == 0   # outside bvh_bbox: empty
    assert len(one(0.97, 1, np.array([[0.0, 0, 18.0]], np.float32), 2.0)[0]) == 0  # entirely above the free height


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n", [(1, 1), (2, 1000), (3, 200000)])
def test_leaf_rule_matches_oracle(seed, n):
    center, size, occupancy, count = _octree_leaves(seed, n)
    bb = np.array([-40.3, -35.1, -0.6, 41.7, 38.2, 30.0], np.float32)
    for free_h in (12.4, 100.0):
        ref_boxes, ref_src = O.octree_leaves_to_boxes(center, size, occupancy, count, bb, free_h)
        got_boxes, got_src = E.octree_leaves_to_boxes(center, size, occupancy, count, bb, free_h)
        assert np.array_equal(got_src, ref_src)
        assert np.array_equal(bits(got_boxes), bits(ref_boxes))
    if n >= 1000:
        assert 0.2 * n < len(ref_src) < 0.98 * n


@pytest.mark.gpu
def test_octree_to_map_to_scores():
    """End to end: octree leaves -> leaf rule on the device -> map -> scoring, against the oracle fed by the oracle's
    own leaf rule."""
    center, size, occupancy, count = _octree_leaves(5, 60000)
    bb = np.array([-45, -45, -1, 45, 45, 28], np.float32)
    boxes, src = E.octree_leaves_to_boxes(center, size, occupancy, count, bb, 25.0)
    rb, rs = O.octree_leaves_to_boxes(center, size, occupancy, count, bb, 25.0)
    assert np.array_equal(src, rs) and np.array_equal(bits(boxes), bits(rb))
    weight = np.exp(-0.1 * count[src].astype(np.float32)).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes, weight = np.ascontiguousarray(boxes[order]), np.ascontiguousarray(weight[order])
    normal = np.zeros((len(boxes), 3), np.float32)
    gmap = E.OccupancyMap(boxes, weight, normal, depth)
    oracle = O.OracleTree(boxes)
    cam = synth.make_camera(64, 48)
    rng = np.random.default_rng(9)
    seen = 0
    for _ in range(3):
        eye = rng.uniform(-30, 30, size=3).astype(np.float32); eye[2] = rng.uniform(26, 40)
        f = -eye / np.linalg.norm(eye)
        r = np.cross(f, [0.0, 0.1, 1.0]); r /= np.linalg.norm(r)
        u = np.cross(f, r)
        Rt = np.concatenate([np.stack([r, u, f], axis=1), eye[:, None]], axis=1).astype(np.float32).reshape(12)
        res = gmap.score_viewpoints(cam.K, 64, (0, 64, 0, 48), Rt[None])
        ref = oracle.score_viewpoint(weight, normal, cam.K, 64, Rt, height=48)
        got = res.viewpoint(0)
        assert np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"])
        seen += len(ref["leaf"])
    assert seen > 30
    gmap.close()


@pytest.mark.gpu
def test_grid_weight_update_matches_oracle():
    center, size, occupancy, count = _octree_leaves(7, 4000)
    bb = np.array([-45, -45, -1, 45, 45, 28], np.float32)
    boxes, src = E.octree_leaves_to_boxes(center, size, occupancy, count, bb, 25.0)
    cnt = count[src]
    order, depth = E.splitmedian_order(boxes)
    boxes, cnt = np.ascontiguousarray(boxes[order]), np.ascontiguousarray(cnt[order])
    n = len(boxes)
    w0 = np.full(n, 0.777, np.float32)
    gmap = E.OccupancyMap(boxes, w0, None, depth)
    rng = np.random.default_rng(11)
    # grids as initializeGrid builds them (viewpoint_planner_data.cpp:932-936): origin = bbox min, increment =
    # max extent / (dim + 1); one covering the map, one covering a part of it (leaves outside keep their weight)
    for origin, extent, dim in (([-45.0, -45.0, -1.0], 90.0, 14), ([-10.3, -20.1, 0.2], 31.7, 9)):
        inc = np.float32(extent) / np.float32(dim + 1)
        gw = rng.uniform(0, 1, size=(dim, dim, dim)).astype(np.float32)
        ref = O.update_weights_from_grid(boxes, w0, origin, inc, [dim, dim, dim], gw, 0.1, cnt)
        got = gmap.update_weights_from_grid(origin, inc, [dim, dim, dim], gw, 0.1, cnt)
        touched = ref != w0
        assert touched.sum() > 50 and (~touched).sum() > 0 or dim == 14
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=0)       # expf: CUDA vs libm
        assert np.array_equal(got[~touched], w0[~touched])
        # the weights really are in the map: scoring with them equals scoring with the oracle's
        gmap.update_weights(w0)
    gmap.close()
