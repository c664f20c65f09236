"""The first-pixel table of the dedup is never reset between chunks: entries carry a tag that decreases from chunk to
chunk (raycast2_kernels.cuh, score_device_impl).  These tests walk the corners of that scheme on the GPU:
  * more than 1022 chunks on one map (the tag space wraps and the table is cleared once),
  * windows beyond 2^22 pixels (untagged mode: the scoring pass resets what it read) and switching between the modes,
  * result fields switched off and on again between calls.
Every result must equal the oracle's / the first call's bit for bit."""
import numpy as np
import pytest

from _common import bits, scene
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def _same(a, b):
    return (a.n_entries == b.n_entries and np.array_equal(a.offsets, b.offsets) and np.array_equal(a.leaf, b.leaf)
            and np.array_equal(a.first_pixel, b.first_pixel) and np.array_equal(bits(a.information), bits(b.information))
            and np.array_equal(bits(a.dist_sq), bits(b.dist_sq)) and np.array_equal(bits(a.total), bits(b.total))
            and np.array_equal(a.hit_count, b.hit_count))


@pytest.mark.gpu
def test_tag_space_wraps_without_a_trace(monkeypatch):
    sc = scene("tiny")
    gmap = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    w, h = 48, 36
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 6, config_id=81)
    monkeypatch.setenv("Q3DR_CHUNK", "2")  # 3 chunks per call, different viewpoints share table slots over time
    first = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    assert first.n_entries > 300
    ref = [sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, cam.K, w, vps[v], height=h) for v in range(6)]
    for v in range(6):
        assert np.array_equal(first.viewpoint(v)["leaf"], ref[v]["leaf"]) and np.array_equal(first.viewpoint(v)["pixel"], ref[v]["pixel"])
    # 400 calls x 3 chunks = 1200 chunks > 1022 tags; rotate the viewpoints so that slots see different images
    for it in range(400):
        k = it % 6
        rolled = np.roll(vps, -k, axis=0)
        res = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), rolled)
        if it % 37 == 0 or 330 <= it <= 345:  # around the wrap (chunk 1022 falls into call 340) and spot checks
            for v in range(6):
                a, b = res.viewpoint(v), first.viewpoint((v + k) % 6)
                assert np.array_equal(a["leaf"], b["leaf"]) and np.array_equal(a["pixel"], b["pixel"]), (it, v)
                assert np.array_equal(bits(a["info"]), bits(b["info"])) and a["hit_count"] == b["hit_count"], (it, v)
    assert _same(gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps), first)
    gmap.close()


@pytest.mark.gpu
def test_untagged_mode_for_huge_windows_and_mode_switches():
    sc = scene("tiny")
    gmap = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    vps = synth.make_viewpoints(sc.leaves, 2, config_id=82)
    small = synth.make_camera(64, 48)
    big_w, big_h = 2304, 1856  # 4.28 M pixels > 2^22: pixel indices no longer fit under a tag
    big = synth.make_camera(big_w, big_h)
    a0 = gmap.score_viewpoints(small.K, 64, (0, 64, 0, 48), vps)
    b0 = gmap.score_viewpoints(big.K, big_w, (0, big_w, 0, big_h), vps[:1])
    a1 = gmap.score_viewpoints(small.K, 64, (0, 64, 0, 48), vps)
    b1 = gmap.score_viewpoints(big.K, big_w, (0, big_w, 0, big_h), vps[:1])
    assert _same(a0, a1) and _same(b0, b1)
    # the huge window against the oracle (first pixels well beyond 2^22 occur)
    ref = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, big.K, big_w, vps[0], height=big_h)
    got = b0.viewpoint(0)
    assert np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"])
    assert big_w * big_h > (1 << 22) and len(got["leaf"]) > 1000
    assert abs(got["total"] - ref["total"]) <= 1e-5 * abs(ref["total"])
    # a sub-window of the same huge image is tagged again (small pixel count): same voxels as the matching crop
    sub = gmap.score_viewpoints(big.K, big_w, (1000, 1400, 800, 1100), vps[:1])
    px = ref["pixel"]
    assert sub.n_entries > 0
    ref_sub = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, big.K, big_w, vps[0], window=(1000, 1400, 800, 1100), height=big_h)
    assert np.array_equal(sub.viewpoint(0)["leaf"], ref_sub["leaf"]) and np.array_equal(sub.viewpoint(0)["pixel"], ref_sub["pixel"])
    gmap.close()


@pytest.mark.gpu
def test_result_fields_off_and_on(monkeypatch):
    sc = scene("small")
    gmap = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 9, config_id=83)
    monkeypatch.setenv("Q3DR_CHUNK", "4")
    full = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    gmap.set_result_fields(first_pixel=False, dist_sq=False)
    lean = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    assert lean.first_pixel.size == 0 and lean.dist_sq.size == 0
    assert np.array_equal(lean.leaf, full.leaf) and np.array_equal(bits(lean.information), bits(full.information))
    assert np.array_equal(lean.offsets, full.offsets) and np.array_equal(bits(lean.total), bits(full.total))
    gmap.set_result_fields(first_pixel=True, dist_sq=False)
    half = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    assert np.array_equal(half.first_pixel, full.first_pixel) and half.dist_sq.size == 0
    gmap.set_result_fields(first_pixel=True, dist_sq=True)
    assert _same(gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps), full)
    with pytest.raises(E.Q3drError):
        E._check(E.load_library().q3dr_map_set_result_fields(gmap._h, 8))
    gmap.close()


@pytest.mark.gpu
def test_viewpoints_on_box_planes_take_the_checking_kernel_on_every_route():
    """Camera centres that share a coordinate with a voxel plane need the reference's inside / outside rule at those
    boxes.  They are sorted out on the host (small host-resident batches), by flag_viewpoints_kernel (large batches,
    device-resident poses) -- all three routes must give the oracle's answer, next to ordinary viewpoints."""
    import torch

    sc = scene("small")
    gmap = E.OccupancyMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth)
    w, h = 64, 48
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 70, config_id=84)
    rng = np.random.default_rng(2)
    special = [3, 11, 40, 69]
    for i, v in enumerate(special):
        leaf = int(rng.integers(0, sc.leaves.n))
        axis = i % 3
        vps[v, 4 * axis + 3] = sc.leaves.boxes[leaf, axis if i % 2 else 3 + axis]  # exactly a min / max plane of some voxel
    ref = {v: sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, cam.K, w, vps[v], height=h) for v in special + [0, 1]}
    big = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps)           # 70 > 64: flag kernel + two launches
    small = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:12])    # host lookup, one launch (checking instantiation)
    plain = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:3])     # host lookup, one launch (fast instantiation)
    d_rt = torch.from_numpy(np.ascontiguousarray(vps)).cuda()
    gmap.score_viewpoints_device(cam.K, w, (0, w, 0, h), d_rt.data_ptr(), 70)
    dev = gmap.result_to_host()
    for v, r in ref.items():
        for res in (big, dev) + ((small,) if v < 12 else ()) + ((plain,) if v < 3 else ()):
            got = res.viewpoint(v)
            assert np.array_equal(got["leaf"], r["leaf"]) and np.array_equal(got["pixel"], r["pixel"]), v
            assert np.array_equal(bits(got["dsq"]), bits(r["dsq"])), v
    assert _same(big, dev)
    gmap.close()


@pytest.mark.gpu
def test_detached_host_arrays_survive_later_calls_and_are_recycled():
    """q3dr_result_detach_host: the shim's voxel-set views keep the engine's own host arrays instead of copying them."""
    leaves = synth.map_small()
    order, depth = E.splitmedian_order(leaves.boxes)
    leaves = leaves.permuted(order)
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(leaves, 12, config_id=41)
    gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
    with pytest.raises(E.Q3drError):
        gmap.detach_host_result()  # nothing scored yet
    r1 = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:7])
    hb1 = gmap.detach_host_result()
    leaf1, info1 = hb1.arrays()
    assert hb1.n_entries == r1.n_entries and hb1.capacity >= hb1.n_entries
    assert np.array_equal(leaf1, r1.leaf) and np.array_equal(info1.view(np.uint32), r1.information.view(np.uint32))
    with pytest.raises(E.Q3drError):
        gmap.detach_host_result()  # already gone
    # the map scores something else: the detached arrays are not its buffers any more
    r2 = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[5:])
    assert r2.n_entries != r1.n_entries or not np.array_equal(r2.leaf, r1.leaf)
    assert np.array_equal(leaf1, r1.leaf) and np.array_equal(info1.view(np.uint32), r1.information.view(np.uint32))
    hb2 = gmap.detach_host_result()
    assert hb2.leaf != hb1.leaf and hb2.information != hb1.information
    leaf2, _ = hb2.arrays()
    assert np.array_equal(leaf2, r2.leaf)
    # a released block is what the map takes next (no fresh pinned allocation per call)
    p1 = (hb1.leaf, hb1.information)
    hb1.release()
    assert hb1.leaf is None and hb1.capacity == 0
    gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:7])
    hb3 = gmap.detach_host_result()
    assert (hb3.leaf, hb3.information) == p1
    leaf3, info3 = hb3.arrays()
    assert np.array_equal(leaf3, r1.leaf) and np.array_equal(info3.view(np.uint32), r1.information.view(np.uint32))
    # detaching works after the device-resident call + q3dr_result_to_host as well
    import torch

    d_rt = torch.from_numpy(np.ascontiguousarray(vps[:7])).cuda()
    gmap.score_viewpoints_device(cam.K, w, (0, w, 0, h), d_rt.data_ptr(), 7)
    r4 = gmap.result_to_host()
    hb4 = gmap.detach_host_result()
    assert np.array_equal(hb4.arrays()[0], r4.leaf) and np.array_equal(r4.leaf, r1.leaf)
    for hb in (hb2, hb3, hb4):
        hb.release()
    gmap.close()
    assert E.load_library().q3dr_host_pool_trim() == E.Q3DR_OK  # the pooled blocks are freed; a later map allocates afresh
    gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
    r5 = gmap.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:7])
    assert np.array_equal(r5.leaf, r1.leaf)
    gmap.close()
