"""Pins the CPU oracle -- and through it the engine -- against outputs of the REFERENCE ITSELF.

The reference's CPU path cannot be built here (Eigen / Boost / octomap absent), but its device code
(viewpoint_planner/src/bvh/bvh.cu + bvh.cuh) compiles from the reference checkout with nvcc alone
(oracle/ref_cuda/Makefile -> oracle/_ref/libq3dr_refcuda_*.so, built in the CPU container, shipped to the GPU
box).  Its recursion (bvh.cu:205-262) is the same algorithm as the CPU recursion (bvh.h:842-909) with ONE
arithmetic difference: squaredNorm() associates sequentially (include/bh/cuda_math.h:135-141) where Eigen's
fixed-size reduction associates a0 + (a1 + a2).  So:

  * the oracle with that one switch flipped must equal the reference's device code BIT FOR BIT (leaf, depth,
    dist_sq bits, intersection bits) on the ray-list entry point -- this pins traversal order, tie-breaking
    (later leaf wins), the inside-a-box rule, the max_range rule and the slab arithmetic to the reference;
  * the engine (Eigen association = the contract) must agree with the reference's device code everywhere except
    on last-ulp near-ties, which are counted and bounded.
"""
from __future__ import annotations

import numpy as np
import pytest

from _common import bits, scene
from oracle import oracle_py as O
from oracle import refcuda_py as R
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available("ieee"), reason="oracle/_ref not built (needs /root/reference)")]


@pytest.fixture()
def sequential_oracle():
    O.set_sqnorm_sequential(True)
    yield
    O.set_sqnorm_sequential(False)


def _rays_of(sc, w, h, nvp, config_id):
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, nvp, config_id=config_id)
    os_, ds_ = [], []
    for v in range(nvp):
        o, d = O.camera_rays(cam.K, vps[v], 0, w, 0, h)
        os_.append(o)
        ds_.append(d)
    return np.concatenate(os_), np.concatenate(ds_)


def _assert_bit_equal(ref, orc):
    assert np.array_equal(ref["leaf"], orc["leaf"])
    assert np.array_equal(bits(ref["dsq"]), bits(orc["dsq"]))
    assert np.array_equal(ref["depth"], orc["depth"])
    assert np.array_equal(bits(ref["xyz"]), bits(orc["xyz"]))


@pytest.mark.parametrize("name,w,h,nvp,max_range", [("tiny", 64, 48, 6, 60.0), ("tiny_unknown", 64, 48, 6, -1.0),
                                                    ("small", 160, 120, 4, 60.0), ("small_unknown", 96, 72, 4, 7.5),
                                                    ("1m", 160, 120, 3, 60.0)])
def test_oracle_equals_reference_device_code_bit_for_bit(sequential_oracle, name, w, h, nvp, max_range):
    sc = scene(name)
    ref_tree = R.RefCudaTree(sc.oracle, "ieee")
    o, d = _rays_of(sc, w, h, nvp, config_id=51)
    ref = ref_tree.intersect(o, d, 0.0, max_range)
    orc = sc.oracle.intersect_rays_raw(o, d, 0.0, max_range)
    _assert_bit_equal(ref, orc)
    hits = int((ref["leaf"] >= 0).sum())
    assert hits > (len(o) // 10 if max_range < 0 or max_range >= 60 else 100)  # short range: mostly misses, some hits


def test_reference_device_code_edge_semantics(sequential_oracle):
    """Ties (duplicated boxes), origins inside leaves / on faces, short ranges: the rules the oracle restates."""
    rng = np.random.default_rng(8)
    c = rng.integers(-6, 6, size=(400, 3)).astype(np.float32)
    s = rng.choice([0.5, 1.0, 2.0], size=(400, 1)).astype(np.float32)
    boxes = np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32)
    boxes = np.concatenate([boxes, boxes[:150], boxes[:50]])  # exact duplicates => exact dist_sq ties
    order, _ = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    oracle = O.OracleTree(boxes)
    ref_tree = R.RefCudaTree(oracle, "ieee")
    n = 8192
    o = rng.uniform(-12, 12, size=(n, 3)).astype(np.float32)
    d = (rng.uniform(-6, 6, size=(n, 3)) - o).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # origins inside boxes, exactly on faces and on corners
    o[:200] = (boxes[:200, :3] + boxes[:200, 3:]) / 2
    o[200:400] = boxes[200:400, 3:]
    o[400:600, 0] = boxes[400:600, 0]
    for max_range in (-1.0, 4.0):
        ref = ref_tree.intersect(o, d, 0.0, max_range)
        orc = oracle.intersect_rays_raw(o, d, 0.0, max_range)
        _assert_bit_equal(ref, orc)
    hit = ref["leaf"][ref["leaf"] >= 0]
    # later leaf wins among bit-identical boxes -- in the reference's own device code
    _, inv = np.unique(boxes, axis=0, return_inverse=True)
    inv = inv.ravel()
    top = np.zeros(inv.max() + 1, dtype=np.int64)
    np.maximum.at(top, inv, np.arange(len(boxes)))
    ref_all = ref_tree.intersect(o, d, 0.0, -1.0)
    hit = ref_all["leaf"][ref_all["leaf"] >= 0]
    assert np.array_equal(hit, top[inv[hit]])
    assert (ref_all["dsq"][:200] == 0).all() and (ref_all["leaf"][:200] >= 0).all()


def test_reference_upload_drops_its_last_batch():
    """Documents the reference defect the harness works around (bvh.cu:40-75): nodes are uploaded in batches of
    num_of_nodes / 20 and the last partial batch is never flushed.  A splitMedian tree has 2N - 1 nodes (odd), so
    there is always a remainder of 1..19 nodes -- the deepest leaves in breadth-first order -- that stay uninitialised
    on the device.  The unpatched tree is NOT run here: walking uninitialised child pointers is undefined."""
    for name in ("tiny", "small", "1m"):
        nn = scene(name).oracle.num_nodes
        assert nn % 2 == 1
        batch = nn // 20
        dropped = nn - (nn // batch) * batch
        assert 1 <= dropped <= 19


@pytest.mark.parametrize("name,w,h,nvp", [("small", 160, 120, 4), ("1m", 320, 240, 2)])
def test_engine_agrees_with_reference_device_code_up_to_last_ulp_ties(name, w, h, nvp):
    sc = scene(name)
    ref_tree = R.RefCudaTree(sc.oracle, "ieee")
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, nvp, config_id=52)
    total = differ = 0
    for v in range(nvp):
        o, d = O.camera_rays(cam.K, vps[v], 0, w, 0, h)
        ref = ref_tree.intersect(o, d, 0.0, 60.0)
        got = sc.gpu_map.raycast_window(cam.K, vps[v], 0, w, 0, h, 0.0, 60.0)
        # hit / miss can flip only for a dist_sq within an ulp of max_range^2: not with these inputs
        assert np.array_equal(got["leaf"] >= 0, ref["leaf"] >= 0)
        ne = got["leaf"] != ref["leaf"]
        total += len(ne)
        differ += int(ne.sum())
        # dist_sq: same leaf => same value up to the one re-association (two roundings each: <= 2 ulp apart); where the leaf differs the two
        # candidates are a near-tie, a few ulp apart at most
        db = np.abs(bits(got["dist_sq"]).astype(np.int64) - bits(ref["dsq"]).astype(np.int64))
        assert db[~ne].max() <= 2
        assert db.max() <= 4
        # where the leaf differs, both candidates are at (last-ulp) equal distance: a legitimate tie
        if ne.any():
            idx = np.where(ne)[0]
            both = sc.oracle.intersect_rays_raw(o[idx], d[idx], 0.0, 60.0)  # contract arithmetic
            assert np.array_equal(both["leaf"], got["leaf"][idx])
    assert differ <= 2e-3 * total, (differ, total)


def test_reference_raycast_entry_point_statistical_agreement():
    """The reference's camera entry point (bvh.cu:647-720) does not normalise the ray direction and is built with
    -use_fast_math by its CMake; it is not bit-comparable, but it must see the same voxels almost everywhere."""
    if not R.available("fast"):
        pytest.skip("fast flavour not built")
    sc = scene("small")
    w, h = 160, 120
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 4, config_id=53)
    ref_tree = R.RefCudaTree(sc.oracle, "fast")
    agree = total = 0
    for v in range(4):
        ref = ref_tree.raycast(cam.K, vps[v], 0, w, 0, h, 0.0, 60.0)
        got = sc.gpu_map.raycast_window(cam.K, vps[v], 0, w, 0, h, 0.0, 60.0)
        assert np.array_equal(ref["screen_xy"][:, 0], got["screen_x"])
        agree += int((ref["leaf"] == got["leaf"]).sum())
        total += len(ref["leaf"])
        hit = (ref["leaf"] >= 0) & (got["leaf"] >= 0)
        assert np.allclose(ref["xyz"][hit], got["intersection"][hit], atol=2e-3)
    assert agree >= 0.995 * total, (agree, total)
