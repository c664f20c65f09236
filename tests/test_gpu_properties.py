"""Size-independent properties checked at BASELINE.json's full sizes (where the oracle would take minutes):
self-consistency between the per-pixel and the de-duplicated paths, chunking / batching independence,
idempotence, window decomposition, sortedness, fp64 totals."""
import os

import numpy as np
import pytest

from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

from _common import bits, scene

pytestmark = pytest.mark.gpu


def _c2():
    sc = scene("1m")
    cam = synth.make_camera(640, 480)
    return sc, cam


def test_c2_batch_properties():
    sc, cam = _c2()
    n = 96
    vps = synth.make_viewpoints(sc.leaves, n, config_id=2)
    win = (0, 640, 0, 480)
    res = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps)
    assert res.n_viewpoints == n and res.offsets[0] == 0 and res.offsets[-1] == res.n_entries
    assert np.all(np.diff(res.offsets.astype(np.int64)) >= 0)
    for v in range(n):
        a, b = int(res.offsets[v]), int(res.offsets[v + 1])
        l = res.leaf[a:b]
        assert np.all(np.diff(l) > 0), "lists are strictly ascending (sorted, no duplicate voxel)"
        assert np.all((l >= 0) & (l < sc.leaves.n))
        assert np.all(res.information[a:b] > 0)  # ignore_zero = 1
        assert res.hit_count[v] >= b - a
        assert np.all(res.first_pixel[a:b] < 640 * 480)
        tot = np.float32(res.information[a:b].astype(np.float64).sum())
        assert abs(res.total[v] - tot) <= 2e-7 * max(tot, 1e-30)
    # idempotence: the tables are reset by the kernels themselves
    res2 = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps)
    for f in ("offsets", "leaf", "first_pixel", "hit_count"):
        assert np.array_equal(getattr(res, f), getattr(res2, f))
    assert np.array_equal(bits(res.information), bits(res2.information))
    assert np.array_equal(bits(res.total), bits(res2.total))
    # batching independence: any split of the viewpoint list gives the same per-viewpoint records
    parts = [sc.gpu_map.score_viewpoints(cam.K, 640, win, vps[a:b]) for a, b in ((0, 1), (1, 40), (40, n))]
    assert np.array_equal(np.concatenate([p.leaf for p in parts]), res.leaf)
    assert np.array_equal(np.concatenate([bits(p.information) for p in parts]), bits(res.information))
    assert np.array_equal(np.concatenate([bits(p.total) for p in parts]), bits(res.total))
    # permutation equivariance
    perm = np.random.default_rng(0).permutation(n)
    rp = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps[perm])
    assert np.array_equal(bits(rp.total), bits(res.total[perm]))
    assert np.array_equal(rp.hit_count, res.hit_count[perm])


def test_c2_chunk_size_independence():
    sc, cam = _c2()
    vps = synth.make_viewpoints(sc.leaves, 40, config_id=3)
    win = (0, 640, 0, 480)
    base = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps)
    old = os.environ.get("Q3DR_CHUNK")
    try:
        for chunk in ("1", "7", "64"):
            os.environ["Q3DR_CHUNK"] = chunk
            r = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps)
            assert np.array_equal(r.offsets, base.offsets) and np.array_equal(r.leaf, base.leaf)
            assert np.array_equal(r.first_pixel, base.first_pixel)
            assert np.array_equal(bits(r.information), bits(base.information))
            assert np.array_equal(bits(r.total), bits(base.total))
    finally:
        if old is None:
            os.environ.pop("Q3DR_CHUNK", None)
        else:
            os.environ["Q3DR_CHUNK"] = old


def test_c2_dedup_consistent_with_per_pixel_path():
    """hit set == unique(leaf per pixel); kept pixel == first occurrence in row-major order; dist_sq carried over."""
    sc, cam = _c2()
    vps = synth.make_viewpoints(sc.leaves, 5, config_id=4)
    win = (0, 640, 0, 480)
    res = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps, E.default_opts(ignore_voxels_with_zero_information=0))
    for v in range(5):
        px = sc.gpu_map.raycast_window(cam.K, vps[v], *win, 0.0, 60.0)
        uniq, first = np.unique(px["leaf"], return_index=True)
        first, uniq = first[uniq >= 0], uniq[uniq >= 0]
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], uniq)
        assert np.array_equal(got["pixel"], first)
        assert np.array_equal(bits(got["dsq"]), bits(px["dist_sq"][first]))
        assert got["hit_count"] == len(uniq)
        hit = px["leaf"] >= 0
        assert np.all(px["dist_sq"][hit] <= np.float32(3600.0))
        # the hit point lies on/in the hit voxel's box (up to rounding) and dist_sq is its squared distance
        b = sc.leaves.boxes[px["leaf"][hit]]
        p = px["intersection"][hit]
        assert np.all(p >= b[:, :3] - 1e-3) and np.all(p <= b[:, 3:] + 1e-3)
        o = vps[v][[3, 7, 11]]
        assert np.allclose(((p - o) ** 2).sum(1), px["dist_sq"][hit], rtol=1e-4, atol=1e-6)


def test_c2_window_decomposition():
    """Casting the image as four sub-windows gives the same pixels; the union of their hit sets is the full set."""
    sc, cam = _c2()
    vp = synth.make_viewpoints(sc.leaves, 1, config_id=6)[0]
    full = sc.gpu_map.raycast_window(cam.K, vp, 0, 640, 0, 480, 0.0, 60.0)["leaf"].reshape(480, 640)
    tiles = [(0, 317, 0, 241), (317, 640, 0, 241), (0, 317, 241, 480), (317, 640, 241, 480)]
    sets = []
    for (x0, x1, y0, y1) in tiles:
        part = sc.gpu_map.raycast_window(cam.K, vp, x0, x1, y0, y1, 0.0, 60.0)["leaf"].reshape(y1 - y0, x1 - x0)
        assert np.array_equal(part, full[y0:y1, x0:x1])
        r = sc.gpu_map.score_viewpoints(cam.K, 640, (x0, x1, y0, y1), vp[None], E.default_opts(ignore_voxels_with_zero_information=0))
        sets.append(r.leaf)
    whole = sc.gpu_map.score_viewpoints(cam.K, 640, (0, 640, 0, 480), vp[None], E.default_opts(ignore_voxels_with_zero_information=0))
    assert np.array_equal(np.unique(np.concatenate(sets)), whole.leaf)


def test_device_resident_entry_point_matches_host_entry_point():
    torch = pytest.importorskip("torch")
    sc, cam = _c2()
    vps = synth.make_viewpoints(sc.leaves, 24, config_id=8)
    win = (0, 640, 0, 480)
    host = sc.gpu_map.score_viewpoints(cam.K, 640, win, vps)
    d_rt = torch.from_numpy(vps).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dres = sc.gpu_map.score_viewpoints_device(cam.K, 640, win, d_rt.data_ptr(), 24, stream=s.cuda_stream)
    s.synchronize()
    from quad3dr_b200 import dist as qdist

    off, tot, leaf, info = qdist.device_result_tensors(dres, torch.device("cuda", 0))
    assert np.array_equal(off.cpu().numpy().astype(np.uint64), host.offsets)
    assert np.array_equal(leaf.cpu().numpy(), host.leaf)
    assert np.array_equal(bits(info.cpu().numpy()), bits(host.information))
    assert np.array_equal(bits(tot.cpu().numpy()), bits(host.total))
    back = sc.gpu_map.result_to_host()
    assert np.array_equal(back.leaf, host.leaf)
    st = sc.gpu_map.stats()
    assert st.rays == 24 * 640 * 480 and st.kernel_launches >= 4 and st.raycast_ms > 0


# ----------------------------------------------------------------------------------------------------
# C3 (10.2 M leaves) and C4 (unknown-space cubes, 1280x960): the oracle's splitMedian build alone takes minutes at
# these sizes, so parity is checked through properties plus a brute-force spot check that needs no tree at all
# ----------------------------------------------------------------------------------------------------

def _brute_force_pixels(boxes, K, Rt, w, pix, max_range):
    """Closest hit of a few pixels by testing EVERY leaf with the reference's arithmetic (numpy fp32, vectorised over
    the leaves): the definition of the result, independent of any tree."""
    f32 = np.float32
    out = []
    R = Rt.reshape(3, 4).astype(f32)
    o = R[:, 3]
    for p in pix:
        x, y = f32(p % w), f32(p // w)
        c0, c1 = (x - K[2]) / K[0], (y - K[3]) / K[1]
        d = np.array([R[i, 0] * c0 + (R[i, 1] * c1 + R[i, 2]) for i in range(3)], dtype=f32)
        n2 = d[0] * d[0] + (d[1] * d[1] + d[2] * d[2])
        d = (d / np.sqrt(n2)).astype(f32)
        inv = (f32(1) / d).astype(f32)
        t0 = (boxes[:, :3] - o) * inv
        t1 = (boxes[:, 3:] - o) * inv
        tmin = np.maximum(np.minimum(t0, t1).max(axis=1), f32(0))
        tmax = np.minimum(np.maximum(t0, t1).min(axis=1), np.finfo(f32).max)
        inside = np.all((o >= boxes[:, :3]) & (o <= boxes[:, 3:]), axis=1)
        hit = (tmax > tmin) | inside
        pt = o[None, :] + d[None, :] * tmin[:, None]
        e = o[None, :] - pt
        dsq = e[:, 0] * e[:, 0] + (e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2])
        dsq = np.where(inside, f32(0), dsq)
        ok = hit & (dsq <= f32(max_range) * f32(max_range))
        if not ok.any():
            out.append((-1, f32(0)))
            continue
        best = dsq[ok].min()
        idx = np.where(ok & (dsq == best))[0].max()  # equal dist_sq: the later leaf wins
        out.append((int(idx), best))
    return out


@pytest.mark.parametrize("name,w,h", [("10m", 640, 480), ("1m_unknown", 1280, 960)])
def test_full_size_maps_properties_and_brute_force_spot_check(name, w, h):
    if name == "10m":
        leaves = synth.map_10m()
    else:
        leaves = synth.map_1m(unknown_interior=True)
    order, depth = E.splitmedian_order(leaves.boxes)
    leaves = leaves.permuted(order)
    gmap = E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth)
    cam = synth.make_camera(w, h)
    n = 12
    vps = synth.make_viewpoints(leaves, n, config_id=3 if name == "10m" else 4)
    win = (0, w, 0, h)
    res = gmap.score_viewpoints(cam.K, w, win, vps)
    assert res.n_viewpoints == n and res.offsets[-1] == res.n_entries and res.n_entries > 1000
    rng = np.random.default_rng(5)
    for v in range(n):
        e = res.viewpoint(v)
        assert np.all(np.diff(e["leaf"]) > 0) and np.all(e["info"] > 0) and np.all(e["pixel"] < w * h)
    # the de-duplicated lists are exactly what the per-pixel path implies: unique leaves, first pixel, its dist_sq
    for v in (0, 7):
        px = gmap.raycast_window(cam.K, vps[v], 0, w, 0, h, 0.0, 60.0)
        uniq, first = np.unique(px["leaf"], return_index=True)
        first, uniq = first[uniq >= 0], uniq[uniq >= 0]
        full = gmap.score_viewpoints(cam.K, w, win, vps[v:v + 1], E.default_opts(ignore_voxels_with_zero_information=0))
        e = full.viewpoint(0)
        assert np.array_equal(e["leaf"], uniq) and np.array_equal(e["pixel"], first)
        assert np.array_equal(bits(e["dsq"]), bits(px["dist_sq"][first]))
        # brute force over ALL leaves for a handful of pixels (hits, misses, image corners)
        pix = np.concatenate([rng.integers(0, w * h, size=10), [0, w - 1, w * (h - 1), w * h - 1]])
        for p, (leaf, dsq) in zip(pix, _brute_force_pixels(leaves.boxes, cam.K, vps[v], w, pix, 60.0)):
            assert px["leaf"][p] == leaf, (v, p)
            if leaf >= 0:
                assert bits(px["dist_sq"][p:p + 1])[0] == bits(np.array([dsq], np.float32))[0]
    # batching independence and idempotence
    parts = [gmap.score_viewpoints(cam.K, w, win, vps[a:b]) for a, b in ((0, 5), (5, n))]
    assert np.array_equal(np.concatenate([p.leaf for p in parts]), res.leaf)
    assert np.array_equal(np.concatenate([bits(p.total) for p in parts]), bits(res.total))
    gmap.close()
