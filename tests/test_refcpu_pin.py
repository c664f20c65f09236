"""Pins the oracle to the REFERENCE'S OWN CPU SOURCES, compiled here.

oracle/ref_cpu/Makefile compiles viewpoint_planner/src/bvh/bvh.h + include/bh/math/geometry.h from the reference
checkout (nothing copied) against stand-in Eigen / Boost headers into oracle/_ref/libq3dr_refcpu.so.  These tests run
the reference's Tree::build / Tree::intersects / node iterator / Tree::intersects(bbox) / BoundingBox3D and demand
bit-identical answers from the oracle's restatement (oracle/q3dr_oracle.cpp) -- leaf order (= the tie-break order
every hit decision rests on), union boxes, hit leaf, dist_sq, hit point and depth per ray, voxel indices, overlap
lists.  CPU only; the .so travels to the GPU box prebuilt.  What stays a convention: the association of Eigen's
three-term reductions (oracle/ref_cpu/stubs/Eigen/Dense says which one the stand-in uses) and strict IEEE instead of
the reference's fast-math pragma.
"""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from oracle import refcpu_py as R
from quad3dr_b200 import synth

if not R.available() and os.path.isdir("/root/reference"):
    import subprocess

    subprocess.run(["make", "-C", os.path.join(os.path.dirname(os.path.dirname(R.LIB_PATH)), "ref_cpu"), "-s"], check=False)

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libq3dr_refcpu.so not built (needs the reference checkout)")


def test_harness_library_exports_and_nothing_of_the_reference_is_copied():
    L = R.lib()
    for sname in R.SYMBOLS:
        assert hasattr(L, sname), sname
    here = os.path.join(os.path.dirname(os.path.dirname(R.LIB_PATH)), "ref_cpu")
    assert sorted(os.listdir(here)) == ["Makefile", "ref_cpu_harness.cpp", "stubs"]
    src = open(os.path.join(here, "ref_cpu_harness.cpp")).read()
    assert "#include <bvh.h>" in src  # compiled from where it lies in the reference checkout, via -I
    # the stand-in headers are stand-ins: no reference file has these names, and they say what they are
    for root, _, files in os.walk(os.path.join(here, "stubs")):
        for f in files:
            assert "stand-in" in open(os.path.join(root, f)).read().lower(), f


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


SCENES = {
    "tiny": synth.map_tiny,
    "tiny_unknown": lambda: synth.map_tiny(unknown_interior=True),
    "small": synth.map_small,
    "small_unknown": lambda: synth.map_small(unknown_interior=True),
    "1m": synth.map_1m,
}


@pytest.mark.parametrize("name", list(SCENES))
def test_oracle_tree_order_equals_reference_tree_build(name):
    """a2: leaf order, depths and node counts of the reference's Tree::build == the oracle's splitMedian restatement."""
    leaves = SCENES[name]()
    ref = R.RefCpuTree(leaves.boxes)
    orc = O.OracleTree(leaves.boxes)
    r_order, r_depth, r_boxes = ref.leaf_order()
    assert np.array_equal(r_order, orc.dfs_order())
    assert np.array_equal(r_depth, orc.leaf_depths()[r_order])  # the oracle indexes by input leaf, the walk by rank
    assert np.array_equal(bits(r_boxes), bits(leaves.boxes[r_order]))
    assert ref.num_leaves == leaves.n and ref.num_nodes == orc.num_nodes and ref.depth == orc.depth
    assert np.array_equal(bits(ref.root_box()), bits(orc.root_box()))
    # a20: position of every leaf in the reference's all-node iterator == the oracle's voxel index
    assert np.array_equal(ref.voxel_index(), orc.voxel_index()[r_order])


def test_order_with_equal_centres_is_stable_like_the_reference():
    """Q17: equal centres keep input order (std::stable_sort); duplicated and nested boxes share centres."""
    rng = np.random.default_rng(5)
    c = rng.integers(-6, 6, size=(400, 3)).astype(np.float32) * 0.5
    s = rng.choice([0.25, 0.5, 1.0], size=(400, 1)).astype(np.float32)
    boxes = np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32)
    boxes = np.concatenate([boxes, boxes[:57], boxes[100:130]])  # exact duplicates
    ref = R.RefCpuTree(boxes)
    orc = O.OracleTree(boxes)
    assert np.array_equal(ref.leaf_order()[0], orc.dfs_order())
    assert np.array_equal(ref.leaf_order()[1], orc.leaf_depths()[orc.dfs_order()])


def _ordered(leaves):
    orc0 = O.OracleTree(leaves.boxes)
    return leaves.permuted(orc0.dfs_order())


def _compare_rays(ref, orc, o, d, min_range, max_range):
    a = ref.intersect_rays(o, d, min_range, max_range)
    b = orc.intersect_rays(o, d, min_range, max_range)
    assert np.array_equal(a["leaf"], b["leaf"])
    hit = a["leaf"] >= 0
    assert np.array_equal(bits(a["dsq"][hit]), bits(b["dsq"][hit]))
    assert np.array_equal(bits(a["xyz"][hit]), bits(b["xyz"][hit]))
    assert np.array_equal(a["depth"][hit], b["depth"][hit])
    return a


@pytest.mark.parametrize("name", ["tiny", "tiny_unknown", "small", "small_unknown"])
def test_oracle_ray_query_equals_reference_tree_intersects(name):
    """a6-a8: Tree::intersects(ray) of the reference == the oracle, bit for bit, for camera rays of several poses."""
    leaves = _ordered(SCENES[name]())
    ref = R.RefCpuTree(leaves.boxes)
    orc = O.OracleTree(leaves.boxes)
    assert np.array_equal(ref.leaf_order()[0], np.arange(leaves.n))  # leaf rank == leaf index on both sides
    cam = synth.make_camera(64, 48)
    vps = synth.make_viewpoints(leaves, 6, config_id=41)
    n_hit = 0
    for v in range(vps.shape[0]):
        o, d = O.camera_rays(cam.K, vps[v], 0, cam.width, 0, cam.height)
        for max_range in (60.0, -1.0, 7.5):
            n_hit += int((_compare_rays(ref, orc, o, d, 0.0, max_range)["leaf"] >= 0).sum())
        _compare_rays(ref, orc, o[:500], d[:500], 5.0, 60.0)  # min_range is accepted and ignored by both (Q1)
    assert n_hit > 5000


def test_reference_quirks_the_oracle_restates():
    """Q3 ties (duplicated boxes: the later leaf wins), Q4 origin inside a leaf / on faces / on corners, short ranges,
    axis-parallel rays over boxes of positive extent -- the reference's own code decides, the oracle must agree."""
    leaves = _ordered(synth.map_tiny())
    boxes = np.concatenate([leaves.boxes, leaves.boxes[::7], leaves.boxes[::11]])  # exact duplicates => distance ties
    orc0 = O.OracleTree(boxes)
    boxes = boxes[orc0.dfs_order()]
    ref = R.RefCpuTree(boxes)
    orc = O.OracleTree(boxes)
    assert np.array_equal(ref.leaf_order()[0], np.arange(boxes.shape[0]))
    rng = np.random.default_rng(11)
    pick = rng.integers(0, boxes.shape[0], size=600)
    mn, mx = boxes[pick, :3], boxes[pick, 3:]
    inside = mn + (mx - mn) * rng.random((600, 3)).astype(np.float32)
    on_face = inside.copy()
    on_face[np.arange(600), rng.integers(0, 3, 600)] = mn[np.arange(600), rng.integers(0, 3, 600)]
    corners = np.where(rng.random((600, 3)) < 0.5, mn, mx).astype(np.float32)
    outside = (inside + rng.normal(0, 6, (600, 3))).astype(np.float32)
    dirs = rng.normal(0, 1, (600, 3)).astype(np.float32)
    axis = np.zeros((600, 3), dtype=np.float32)
    axis[np.arange(600), rng.integers(0, 3, 600)] = rng.choice([-1.0, 1.0], 600)
    ties = 0
    for origins in (inside, on_face, corners, outside):
        for d in (dirs, axis):
            for max_range in (-1.0, 60.0, 0.75):
                a = _compare_rays(ref, orc, origins, d, 0.0, max_range)
                ties += int((a["dsq"][a["leaf"] >= 0] == 0).sum())
    assert ties > 500


@pytest.mark.parametrize("name", ["tiny_unknown", "small"])
def test_oracle_box_overlap_equals_reference(name):
    """N3: Tree::intersects(bbox) -- same leaves in the same visiting order."""
    leaves = _ordered(SCENES[name]())
    ref = R.RefCpuTree(leaves.boxes)
    orc = O.OracleTree(leaves.boxes)
    rng = np.random.default_rng(3)
    root = orc.root_box()
    total = 0
    for _ in range(300):
        c = root[:3] + (root[3:] - root[:3]) * rng.random(3)
        e = rng.choice([0.1, 0.5, 2.0, 8.0]) * rng.random(3)
        box = np.concatenate([c - e, c + e]).astype(np.float32)
        a, b = ref.overlap_box(box), orc.overlap_box(box)
        assert np.array_equal(a, b)
        total += len(a)
    # touching counts: a query box sharing exactly a face plane with a leaf
    k = leaves.n // 2
    b = leaves.boxes[k].copy()
    touch = np.array([b[3], b[1], b[2], b[3] + 1.0, b[4], b[5]], dtype=np.float32)
    assert np.array_equal(ref.overlap_box(touch), orc.overlap_box(touch)) and k in ref.overlap_box(touch)
    assert total > 1000


def test_oracle_leaf_box_rule_equals_reference_boundingbox():
    """a4 / N1: BoundingBox3D(center, size), constrainTo, isEmpty as generateBVHTree uses them."""
    rng = np.random.default_rng(9)
    bb = np.array([-20, -20, -1, 20, 20, 15], dtype=np.float32)
    n = 3000
    res = 0.25
    level = rng.integers(0, 5, n)
    size = (res * 2.0 ** level).astype(np.float32)
    center = ((rng.integers(-24, 24, (n, 3)) + 0.5) * size[:, None]).astype(np.float32)
    occ = np.full(n, 0.97, dtype=np.float32)
    cnt = np.ones(n, dtype=np.uint32)
    boxes, src = O.octree_leaves_to_boxes(center, size, occ, cnt, bb, 1e30)
    kept = np.zeros(n, dtype=bool)
    kept[src] = True
    j = 0
    for i in range(n):
        box, empty = R.leaf_box(center[i], size[i], bb)
        assert kept[i] == (not empty)
        if kept[i]:
            assert np.array_equal(bits(box), bits(boxes[j]))
            j += 1
    assert 300 < j < n
