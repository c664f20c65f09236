"""Row N3: batched candidate generation (q3dr_generate_viewpoint_entries) makes exactly the decisions of the
reference's one-at-a-time loop (ViewpointPlanner::tryToAddViewpointEntry, viewpoint_planner_graph.cpp:542-655),
restated literally in oracle/oracle_py.py::generate_viewpoint_entries_sequential."""
import ctypes as C

import numpy as np
import pytest

from _common import scene
from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def test_candidate_opts_defaults_are_the_planners():
    o = E.default_candidate_opts()
    assert C.sizeof(E.CandidateOpts) == 112  # 28 four-byte fields
    assert o.discard_dist_knn == 10 and o.exploration_step == 3.0 and o.exploration_dilation_squared == 1
    assert o.exploration_dilation_speed == 2.0 and o.exploration_angular_dist_threshold_degrees == 30.0
    assert o.min_voxel_count == 100 and o.min_information == 10.0
    lib = E.load_library()
    assert lib.q3dr_generate_viewpoint_entries(None, None, 1, 0, 1, 0, 1, None, None, 0, None, None, 0, None, None, None, None) \
        == E.Q3DR_ERR_INVALID_ARG


def _rand_quats(rng, n):
    q = rng.normal(size=(n, 4)).astype(np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("knn,ang_deg", [(10, 30.0), (2, 180.0)])
def test_batch_equals_sequential_loop(knn, ang_deg):
    sc = scene("small")
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    rng = np.random.default_rng(17)
    root = sc.oracle.root_box()
    # candidates as the planner makes them: exploration steps around existing viewpoints (clusters => the distance gate
    # fires), several orientations per position (viewpoint_exploration_num_orientations), some inside obstacles
    vps = synth.make_viewpoints(sc.leaves, 12, config_id=51)
    seeds = np.stack([vps[:, 3], vps[:, 7], vps[:, 11]], axis=1)
    cand_pos, cand_quat = [], []
    for s in seeds:
        for axis in range(3):
            for sign in (-1.0, 1.0):
                p = s.copy()
                p[axis] += sign * 3.0
                for _ in range(3):
                    cand_pos.append(p + rng.normal(0, 0.2, 3).astype(np.float32) * (rng.random() < 0.5))
                    cand_quat.append(None)
    cand_pos = np.array(cand_pos, dtype=np.float32)
    n = cand_pos.shape[0]
    # orientations that look at the scene: reuse the generator's rotations (as quaternions) with random re-use
    from_rt = []
    for v in range(vps.shape[0]):
        R = vps[v].reshape(3, 4)[:, :3].astype(np.float64)
        qw = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
        from_rt.append(np.array([qw, (R[2, 1] - R[1, 2]) / (4 * qw), (R[0, 2] - R[2, 0]) / (4 * qw), (R[1, 0] - R[0, 1]) / (4 * qw)], dtype=np.float32))
    cand_quat = np.stack([from_rt[int(rng.integers(0, len(from_rt)))] for _ in range(n)]).astype(np.float32)
    flip = rng.random(n) < 0.25
    cand_quat[flip] = _rand_quats(rng, int(flip.sum()))
    existing_pos = seeds[:5]
    existing_quat = np.stack(from_rt[:5])
    bb = np.concatenate([root[:3] - 30, root[3:] + 30]).astype(np.float32)
    fields = dict(object_bbox=[-0.4, -0.4, -0.3, 0.4, 0.4, 0.3], bvh_bbox=list(bb), obstacle_free_height=1e30,
                  discard_dist_knn=knn, exploration_step=3.0, exploration_step_max=7.0, exploration_dilation_squared=1,
                  exploration_dilation_start_bbox_distance=0.5, exploration_dilation_speed=2.0,
                  exploration_bbox=list(np.concatenate([root[:3] + 2, root[3:] - 2])),
                  exploration_angular_dist_threshold_degrees=ang_deg, min_voxel_count=60, min_information=8.0)
    copts = E.default_candidate_opts(**fields)
    status, res = sc.gpu_map.generate_viewpoint_entries(cam.K, w, (0, w, 0, h), cand_pos, cand_quat, existing_pos, existing_quat, copts)
    ref_status, ref_results = O.generate_viewpoint_entries_sequential(
        sc.oracle, sc.leaves.weight, sc.leaves.normal, cam.K, w, h, cand_pos, cand_quat, existing_pos, existing_quat, fields)
    assert np.array_equal(status, ref_status)
    counts = np.bincount(status, minlength=5)
    # every kind of decision occurs in the batch (how often depends on the scene)
    assert counts[E.CANDIDATE_ACCEPTED] > 10 and counts[E.CANDIDATE_TOO_CLOSE] > 3 and counts[E.CANDIDATE_INVALID_POSITION] > 3, counts
    assert counts[E.CANDIDATE_TOO_FEW_VOXELS] + counts[E.CANDIDATE_TOO_LITTLE_INFORMATION] > 0, counts
    assert res.n_viewpoints == n
    for j in range(n):
        got = res.viewpoint(j)
        if status[j] == E.CANDIDATE_INVALID_POSITION:
            assert len(got["leaf"]) == 0
        if ref_results[j] is not None:  # the sequential loop cast it: same voxel set, same score
            assert np.array_equal(got["leaf"], ref_results[j]["leaf"])
            assert abs(got["total"] - ref_results[j]["total"]) <= 1e-5 * abs(ref_results[j]["total"]) + 1e-30
