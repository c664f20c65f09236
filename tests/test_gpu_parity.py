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
