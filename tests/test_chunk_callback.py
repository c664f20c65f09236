"""q3dr_map_set_chunk_callback: the chunk reports of a scoring call tile its viewpoints and its entries, in order, and
agree with the final CSR offsets (external consumers of finished chunks hang on them; the library's own multi-GPU
exchange, q3dr_exchange, is fed from the same place in the scoring call)."""
import os

import numpy as np
import pytest

from quad3dr_b200 import synth

from _common import scene


@pytest.mark.gpu
def test_chunk_reports_tile_the_result(monkeypatch):
    sc = scene("small")
    cam = synth.make_camera(96, 72)
    vps = synth.make_viewpoints(sc.leaves, 23, config_id=21)
    monkeypatch.setenv("Q3DR_CHUNK", "5")
    seen = []
    sc.gpu_map.set_chunk_callback(lambda a, b, c, d, part: seen.append((a, b, c, d, part.n_viewpoints, part.n_entries,
                                                                        int(part.leaf or 0))))
    try:
        res = sc.gpu_map.score_viewpoints(cam.K, cam.width, (0, cam.width, 0, cam.height), vps)
    finally:
        sc.gpu_map.set_chunk_callback(None)
    assert [s[1] for s in seen] == [5, 5, 5, 5, 3]
    v = e = 0
    for first_vp, n_vp, first_entry, n_entries, nv_so_far, ne_so_far, leaf_ptr in seen:
        assert first_vp == v and first_entry == e
        v += n_vp
        e += n_entries
        assert nv_so_far == v and ne_so_far == e and leaf_ptr != 0
        assert int(res.offsets[v]) == e  # the chunk's entries are exactly the CSR range of its viewpoints
    assert v == 23 and e == res.n_entries
    # without a callback nothing is reported any more
    n = len(seen)
    sc.gpu_map.score_viewpoints(cam.K, cam.width, (0, cam.width, 0, cam.height), vps[:7])
    assert len(seen) == n


def test_chunk_callback_can_be_set_and_cleared_without_a_gpu_call():
    """host-only: installing / removing the callback on a NULL map fails loudly, not silently"""
    import ctypes as C

    from quad3dr_b200 import engine as E

    lib = E.load_library()
    assert lib.q3dr_map_set_chunk_callback(None, C.cast(None, E.CHUNK_CALLBACK), None) == 1  # Q3DR_ERR_INVALID_ARG
