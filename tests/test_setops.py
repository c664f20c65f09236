"""Consumers of the voxel sets: computeNewInformation, addObservedVoxelsToViewpointPath,
the stereo-pair overlap sum.  Oracle known-answer tests on the CPU; GPU parity against the oracle on real hit lists,
including a greedy path selection driven end to end on the device-resident result."""
import numpy as np
import pytest

from _common import bits, scene
from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

F = np.float32(1.0 / 3.0)  # viewpoint_information_factor (viewpoint_planner.h:256)
NAN = np.float32(np.nan)


def test_oracle_known_answers():
    weight = np.array([1.0, 0.5, 0.25, 2.0], np.float32)
    observed = np.array([NAN, 0.125, 0.25, NAN], np.float32)
    leaf = np.array([0, 1, 2, 3], np.int32)
    info = np.array([0.75, 0.75, 0.75, 1.5], np.float32)
    # voxel 0: absent -> f*0.75 = 0.25 ; voxel 1: min(0.25, 0.5-0.125) = 0.25 ; voxel 2: min(0.25, 0) = 0 ; voxel 3: 0.5
    v, v64 = O.new_information(leaf, info, observed, weight, F)
    assert v == pytest.approx(1.0, rel=1e-6) and v64 == pytest.approx(1.0, rel=1e-6)
    obs = observed.copy()
    v2, _ = O.observed_add(leaf, info, obs, weight, F)
    assert v2 == pytest.approx(1.0, rel=1e-6)
    np.testing.assert_allclose(obs, [0.25, 0.375, 0.25, 0.5], rtol=1e-6)
    # saturation: adding again clamps at the voxel weight
    for _ in range(8):
        O.observed_add(leaf, info, obs, weight, F)
    np.testing.assert_allclose(obs, np.minimum(weight, obs), rtol=0)
    assert obs[1] == 0.5 and obs[2] == 0.25
    # overlap: equality on voxel AND information (occupied_tree.h:73-79)
    ov, _ = O.overlap_information([0, 2, 3], [0.5, 0.25, 1.0], [0, 1, 2, 3], [0.5, 9.0, 0.125, 1.0])
    assert ov == 1.5


@pytest.fixture(scope="module")
def scored():
    sc = scene("small_unknown")
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 48, config_id=61)
    vps[7] = vps[3]  # identical viewpoints: full overlap with bit-identical informations
    res = sc.gpu_map.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    return sc, res


@pytest.mark.gpu
def test_new_information_and_observed_updates_match_oracle(scored):
    sc, res = scored
    weight = sc.leaves.weight
    obs_gpu = E.ObservedVoxels(sc.gpu_map)
    obs_ref = np.full(sc.leaves.n, NAN, np.float32)
    order = [5, 11, 3, 7, 3, 20, 5]  # repeats exercise the accumulate + clamp branch
    for step, v in enumerate(order):
        # all candidates against the current map
        got = obs_gpu.new_information(F, n=res.n_viewpoints)
        for c in range(res.n_viewpoints):
            e = res.viewpoint(c)
            ref32, ref64 = O.new_information(e["leaf"], e["info"], obs_ref, weight, F)
            assert abs(got[c] - ref64) <= 1e-5 * max(abs(ref64), 1e-30)
            assert abs(got[c] - ref32) <= 1e-4 * max(abs(ref32), 1e-30)  # order freedom of the fp32 running sum
        sub = obs_gpu.new_information(F, viewpoints=[9, 2, 9])
        assert sub[0] == sub[2] == got[9] and sub[1] == got[2]
        # fold viewpoint v in
        e = res.viewpoint(v)
        ref32, ref64 = O.observed_add(e["leaf"], e["info"], obs_ref, weight, F)
        new_info = obs_gpu.add_viewpoint(F, v)
        assert abs(new_info - ref64) <= 1e-5 * max(abs(ref64), 1e-30)
        # the map itself is bit-exact (per-voxel arithmetic, no order dependence)
        assert np.array_equal(bits(obs_gpu.download()), bits(obs_ref))
    obs_gpu.clear()
    assert np.isnan(obs_gpu.download()).all()
    obs_gpu.close()


@pytest.mark.gpu
def test_overlap_information_matches_oracle(scored):
    sc, res = scored
    others = np.arange(res.n_viewpoints, dtype=np.uint32)
    for ref_v in (3, 0, 17):
        got = E.overlap_information(sc.gpu_map, ref_v, others)
        a = res.viewpoint(ref_v)
        for c in range(res.n_viewpoints):
            b = res.viewpoint(c)
            r32, r64 = O.overlap_information(b["leaf"], b["info"], a["leaf"], a["info"])
            assert abs(got[c] - r64) <= 1e-5 * max(abs(r64), 1e-30)
    full = E.overlap_information(sc.gpu_map, 3, [7, 3])
    assert abs(full[0] - res.total[3]) <= 1e-5 * res.total[3] and abs(full[1] - res.total[3]) <= 1e-5 * res.total[3]
    # scratch is restored: a second query gives the same answer
    again = E.overlap_information(sc.gpu_map, 0, others)
    assert np.array_equal(again, E.overlap_information(sc.gpu_map, 0, others))


@pytest.mark.gpu
def test_greedy_selection_on_device_equals_cpu_selection(scored):
    """The path search's inner loop (viewpoint_planner_path.cpp:478-548 without the lazy ordering): repeatedly take the
    candidate with the largest new information and fold it in.  Same picks and same totals as the oracle."""
    sc, res = scored
    weight = sc.leaves.weight
    obs_gpu = E.ObservedVoxels(sc.gpu_map)
    obs_ref = np.full(sc.leaves.n, NAN, np.float32)
    picks_gpu, picks_ref = [], []
    for _ in range(10):
        got = obs_gpu.new_information(F, n=res.n_viewpoints)
        ref = np.array([O.new_information(res.viewpoint(c)["leaf"], res.viewpoint(c)["info"], obs_ref, weight, F)[1]
                        for c in range(res.n_viewpoints)])
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-30)
        vg, vr = int(np.argmax(got)), int(np.argmax(ref))
        picks_gpu.append(vg)
        picks_ref.append(vr)
        obs_gpu.add_viewpoint(F, vg)
        e = res.viewpoint(vg)
        O.observed_add(e["leaf"], e["info"], obs_ref, weight, F)
    assert picks_gpu == picks_ref
    assert len(set(picks_gpu)) > 5
    obs_gpu.close()


@pytest.mark.gpu
def test_explicit_voxel_sets_match_resident_ones(scored):
    """A stored voxel set handed over from the host gives the same sums and the same map as the resident viewpoint."""
    sc, res = scored
    a, b = E.ObservedVoxels(sc.gpu_map), E.ObservedVoxels(sc.gpu_map)
    rng = np.random.default_rng(4)
    for v in (2, 9, 2, 30):
        e = res.viewpoint(v)
        perm = rng.permutation(len(e["leaf"]))  # any order
        assert a.voxel_set(F, e["leaf"][perm], e["info"][perm], add=False) == pytest.approx(b.new_information(F, viewpoints=[v])[0], rel=1e-6)
        assert a.voxel_set(F, e["leaf"][perm], e["info"][perm], add=True) == pytest.approx(b.add_viewpoint(F, v), rel=1e-6)
        assert np.array_equal(bits(a.download()), bits(b.download()))
    assert a.voxel_set(F, [], [], add=True) == 0.0
    with pytest.raises(E.Q3drError):
        a.voxel_set(F, [sc.leaves.n], [1.0], add=False)
    a.close()
    b.close()


@pytest.mark.gpu
def test_setops_argument_checks(scored):
    sc, res = scored
    obs = E.ObservedVoxels(sc.gpu_map)
    with pytest.raises(E.Q3drError):
        obs.new_information(F, viewpoints=[res.n_viewpoints])
    with pytest.raises(E.Q3drError):
        obs.add_viewpoint(F, res.n_viewpoints + 3)
    with pytest.raises(E.Q3drError):
        E.overlap_information(sc.gpu_map, res.n_viewpoints, [0])
    obs.close()
