"""q3dr_multi: one process, several GPUs (map replicated, viewpoints sharded, a host thread per GPU)."""
import numpy as np
import pytest

from _common import scene
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


@pytest.mark.gpu
@pytest.mark.parametrize("n_vp", [37, 1, 0])
def test_multi_equals_single_gpu(n_vp):
    n_dev = E.device_count()
    sc = scene("small_unknown")
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, max(n_vp, 1), config_id=71)[:n_vp]
    single = sc.gpu_map.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    # all visible devices; on a one-GPU box the same device twice still exercises the sharding and the threads
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    multi = E.MultiGpuMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth, devices=devices)
    assert multi.n_devices == len(devices)
    parts = multi.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    assert len(parts) == len(devices)
    seen = 0
    for g, (first, res) in enumerate(parts):
        assert first == seen
        for v in range(res.n_viewpoints):
            a, b = res.viewpoint(v), single.viewpoint(first + v)
            assert np.array_equal(a["leaf"], b["leaf"]) and np.array_equal(a["pixel"], b["pixel"])
            assert np.array_equal(a["info"], b["info"]) and a["total"] == b["total"] and a["hit_count"] == b["hit_count"]
        seen += res.n_viewpoints
    assert seen == n_vp
    sizes = [r.n_viewpoints for _, r in parts]
    assert max(sizes) - min(sizes) <= 1
    multi.close()


@pytest.mark.gpu
def test_multi_with_mesh_normals_equals_single_gpu():
    """row N2 through the single-process multi-GPU entry point: one mesh per device"""
    n_dev = E.device_count()
    sc = scene("small")
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, 11, config_id=72)
    tv, tn = synth.make_surface_mesh(sc.leaves, cell=2.0)
    mesh = E.NormalMesh(tv, tn)
    single = sc.gpu_map.score_viewpoints_mesh_normals(mesh, cam.K, w, h, (0, w, 0, h), vps)
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    multi = E.MultiGpuMap(sc.leaves.boxes, sc.leaves.weight, sc.leaves.normal, sc.depth, devices=devices)
    multi.set_normal_mesh(tv, tn, h)
    parts = multi.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    for first, res in parts:
        for v in range(res.n_viewpoints):
            a, b = res.viewpoint(v), single.viewpoint(first + v)
            assert np.array_equal(a["leaf"], b["leaf"]) and np.array_equal(a["info"], b["info"]) and a["total"] == b["total"]
    multi.set_normal_mesh(None, None, 0)
    plain = multi.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    ref = sc.gpu_map.score_viewpoints(cam.K, w, (0, w, 0, h), vps)
    for first, res in plain:
        for v in range(res.n_viewpoints):
            assert np.array_equal(res.viewpoint(v)["info"], ref.viewpoint(first + v)["info"])
    multi.close()
