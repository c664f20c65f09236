"""CPU tests of the oracle itself: structure independence (brute force over leaves), hand-built known answers,
and the committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from quad3dr_b200 import synth

from _common import bits

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if "mesh_normals" not in p)  # the mesh fixture has its own tests (test_mesh_normals.py)


def _row_of_cubes():
    # three unit cubes along +x; ids 0, 1, 2 (already in splitMedian order)
    b = np.array([[2, -0.5, -0.5, 3, 0.5, 0.5], [4, -0.5, -0.5, 5, 0.5, 0.5], [6, -0.5, -0.5, 7, 0.5, 0.5]], np.float32)
    t = O.OracleTree(b)
    assert np.array_equal(t.dfs_order(), [0, 1, 2])
    return t


def test_known_answer_head_on():
    t = _row_of_cubes()
    r = t.intersect_rays([[0, 0, 0]], [[1, 0, 0]], 0.0, -1.0)
    assert r["leaf"][0] == 0 and r["dsq"][0] == 4.0
    assert np.array_equal(r["xyz"][0], [2, 0, 0])
    assert r["depth"][0] == 1  # splitMedian over 3 leaves: left = {0} at depth 1, right = {1, 2} at depth 2


def test_known_answer_depths():
    t = _row_of_cubes()
    # splitMedian over 3 leaves: mid = 1 -> left {0} (depth 1), right {1,2} -> depth 2 each
    assert np.array_equal(t.leaf_depths(), [1, 2, 2])


def test_known_answer_direction_is_normalised():
    t = _row_of_cubes()
    r = t.intersect_rays([[0, 0, 0]], [[10, 0, 0]], 0.0, -1.0)
    assert r["leaf"][0] == 0 and r["dsq"][0] == 4.0


def test_known_answer_miss_and_away():
    t = _row_of_cubes()
    r = t.intersect_rays([[0, 0, 0], [0, 3, 0]], [[-1, 0, 0], [1, 0, 0]], 0.0, -1.0)
    assert np.array_equal(r["leaf"], [-1, -1])
    assert np.array_equal(r["dsq"], [0, 0])


def test_known_answer_max_range_is_inclusive_and_min_range_ignored():
    t = _row_of_cubes()
    o, d = [[0, 0, 0]], [[1, 0, 0]]
    assert t.intersect_rays(o, d, 0.0, 1.5)["leaf"][0] == -1
    assert t.intersect_rays(o, d, 0.0, 2.0)["leaf"][0] == 0  # dist_sq == max_range^2 is accepted (prune only on >)
    assert t.intersect_rays(o, d, 5.0, -1.0)["leaf"][0] == 0  # min_range is never read (bvh.h:381)


def test_known_answer_origin_inside_leaf():
    t = _row_of_cubes()
    r = t.intersect_rays([[4.5, 0.25, 0]], [[0.3, 0.1, -0.2]], 0.0, -1.0)
    assert r["leaf"][0] == 1 and r["dsq"][0] == 0.0
    assert np.array_equal(r["xyz"][0], np.array([4.5, 0.25, 0], np.float32))
    # inclusive bounds: an origin exactly on a face counts as inside
    r = t.intersect_rays([[4.0, 0.0, 0.0]], [[-1, 0.1, 0.1]], 0.0, -1.0)
    assert r["leaf"][0] == 1 and r["dsq"][0] == 0.0


def test_known_answer_tie_goes_to_later_leaf():
    # two identical boxes + a third: equal dist_sq => the leaf later in depth-first order wins
    b = np.array([[2, -1, -1, 3, 1, 1], [2, -1, -1, 3, 1, 1], [8, -1, -1, 9, 1, 1]], np.float32)
    t = O.OracleTree(b)
    assert np.array_equal(t.dfs_order(), [0, 1, 2])
    r = t.intersect_rays([[0, 0.1, 0.2]], [[1, 0.01, 0.02]], 0.0, -1.0)
    assert r["leaf"][0] == 1
    rb = t.raycast_window_brute  # noqa: F841  (brute force agrees, see below)


def test_score_known_answer():
    # voxel in front of the camera, normal facing it: incidence = 1; resolution factor from the formula
    box = np.array([9.875, -0.125, -0.125, 10.125, 0.125, 0.125], np.float32)
    info = O.voxel_information(box, 0.5, [-1, 0, 0], [0, 0, 0], fx=544.0, width=640)
    rel = np.float32(np.float32(np.float32(544.0) * np.float32(0.25)) / np.float32(10.0)) / np.float32(640)
    assert rel >= np.float32(0.002)
    assert info == np.float32(0.5)
    # back-facing normal -> 0 ; zero normal -> incidence 1
    assert O.voxel_information(box, 0.5, [1, 0, 0], [0, 0, 0], 544.0, 640) == 0.0
    assert O.voxel_information(box, 0.5, [0, 0, 0], [0, 0, 0], 544.0, 640) == np.float32(0.5)
    assert O.voxel_information(box, 0.5, [1, 0, 0], [0, 0, 0], 544.0, 640,
                               O.default_opts(incidence_ignore_dot_product_sign=1)) == np.float32(0.5)
    # far away: rel < thr -> exp falloff
    far = box + np.array([190, 0, 0, 190, 0, 0], np.float32)
    rel = np.float32(np.float32(np.float32(544.0) * np.float32(0.25)) / np.float32(200.0)) / np.float32(640)
    assert rel < np.float32(0.002)
    expect = np.float32(np.exp(np.float32(-np.float32(1.0) / np.float32(0.001)) * (np.float32(0.002) - rel))) * np.float32(0.5)
    got = O.voxel_information(far, 0.5, [-1, 0, 0], [0, 0, 0], 544.0, 640)
    assert abs(got - expect) <= 2e-7 * abs(expect)
    # 60 degrees incidence: exp(-(60-30)/30) = exp(-1)
    n = np.array([-0.5, np.sqrt(3) / 2, 0], np.float32)
    got = O.voxel_information(box, 1.0, n, [0, 0, 0], 544.0, 640)
    assert abs(got - np.exp(-1.0)) < 1e-5


@pytest.mark.parametrize("name", ["tiny", "tiny_unknown"])
def test_tree_result_depends_only_on_leaves(name):
    """recursion == argmin dist_sq over leaves, ties to the later leaf in depth-first order."""
    leaves = synth.map_tiny(unknown_interior=(name == "tiny_unknown"))
    tree = O.OracleTree(leaves.boxes)
    cam = synth.make_camera(48, 36)
    vps = synth.make_viewpoints(leaves, 4, config_id=21)
    for v in range(4):
        for max_range in (60.0, 7.5, -1.0):
            a = tree.raycast_window(cam.K, vps[v], 0, 48, 0, 36, 0.0, max_range)
            b = tree.raycast_window_brute(cam.K, vps[v], 0, 48, 0, 36, 0.0, max_range)
            assert np.array_equal(a["leaf"], b["leaf"])
            assert np.array_equal(bits(a["dsq"]), bits(b["dsq"]))


def test_leaf_order_is_a_fixed_point():
    leaves = synth.map_tiny()
    t0 = O.OracleTree(leaves.boxes)
    order = t0.dfs_order()
    t1 = O.OracleTree(leaves.boxes[order])
    assert np.array_equal(t1.dfs_order(), np.arange(leaves.n))
    assert np.array_equal(t1.leaf_depths(), t0.leaf_depths()[order])
    assert t0.depth == int(np.ceil(np.log2(leaves.n)))
    assert t0.num_nodes == 2 * leaves.n - 1


def test_voxel_index_order():
    """all-node iterator (bvh.h:226-235): stack, push left then right -> root, right subtree, left subtree."""
    t = _row_of_cubes()
    # nodes: root(0) -> right inner(1) -> its right leaf {2}(2), its left leaf {1}(3) -> left leaf {0}(4)
    assert np.array_equal(t.voxel_index(), [4, 3, 2])


def test_dedup_keeps_first_pixel_and_sums():
    leaves = synth.map_tiny()
    tree = O.OracleTree(leaves.boxes)
    cam = synth.make_camera(64, 48)
    vp = synth.make_viewpoints(leaves, 1, config_id=5)[0]
    px = tree.raycast_window(cam.K, vp, 0, 64, 0, 48, 0.0, 60.0)
    s = tree.score_viewpoint(leaves.weight, leaves.normal, cam.K, 64, vp, height=48,
                             opts=O.default_opts(ignore_voxels_with_zero_information=0))
    hit = px["leaf"][px["leaf"] >= 0]
    uniq, first = np.unique(px["leaf"], return_index=True)
    first = first[uniq >= 0]
    uniq = uniq[uniq >= 0]
    assert len(hit) > 0
    assert np.array_equal(s["leaf"], uniq)
    assert np.array_equal(s["pixel"], first)
    assert s["hit_count"] == len(uniq)
    assert abs(s["total"] - np.float32(s["info"].astype(np.float64).sum())) <= 1e-6 * s["total"]
    s1 = tree.score_viewpoint(leaves.weight, leaves.normal, cam.K, 64, vp, height=48)
    keep = s["info"] > 0
    assert np.array_equal(s1["leaf"], s["leaf"][keep])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    tree = O.OracleTree(g["boxes"])
    assert np.array_equal(tree.dfs_order(), np.arange(g["boxes"].shape[0]))
    w, h, mr = int(g["width"]), int(g["height"]), float(g["max_range"])
    for v in range(g["Rt"].shape[0]):
        r = tree.raycast_window(g["K"], g["Rt"][v], 0, w, 0, h, 0.0, mr)
        assert np.array_equal(r["leaf"], g["px_leaf"][v])
        assert np.array_equal(bits(r["dsq"]), g["px_dsq_bits"][v])
        assert np.array_equal(r["depth"], g["px_depth"][v])
        s = tree.score_viewpoint(g["weight"], g["normal"], g["K"], w, g["Rt"][v], height=h,
                                 opts=O.default_opts(raycast_max_range=mr))
        a, b = int(g["offsets"][v]), int(g["offsets"][v + 1])
        assert np.array_equal(s["leaf"], g["leaf"][a:b])
        assert np.array_equal(s["pixel"], g["pixel"][a:b])
        np.testing.assert_allclose(s["info"], g["info"][a:b], rtol=2e-7)
        assert abs(s["total"] - g["total"][v]) <= 2e-7 * abs(g["total"][v])
