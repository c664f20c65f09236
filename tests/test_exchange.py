"""The result exchange between GPUs behind the C ABI (q3dr_exchange_*).

CPU: buffer layout arithmetic and argument checks.  GPU: two ranks as two host threads -- on a one-GPU box both on
device 0 (the pushes are then device-local copies, the protocol is the same), on a multi-GPU box on devices 0 and 1 --
must end every step holding each other's results bit for bit; an overflowing step fails on BOTH ranks instead of
hanging.  (tools/check_peer_gather.py is the same check across processes under torchrun, tests/test_peer_gather_gpu.py.)
"""
import ctypes as C
import threading

import numpy as np
import pytest

from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def test_layout_is_aligned_ordered_and_disjoint():
    for n_ranks, nv, ne in [(1, 0, 0), (2, 1, 1), (8, 1250, 40_000_000), (3, 1000, 12345)]:
        lay = E.exchange_layout(n_ranks, nv, ne)
        offs = [lay.header_off, lay.offsets_off, lay.total_off, lay.leaf_off, lay.info_off, lay.slot_bytes]
        assert all(o % 256 == 0 for o in offs)
        assert offs == sorted(offs)
        assert lay.offsets_off - lay.header_off >= 32
        assert lay.total_off - lay.offsets_off >= 8 * (nv + 1)
        assert lay.leaf_off - lay.total_off >= 4 * nv
        assert lay.info_off - lay.leaf_off >= 4 * ne
        assert lay.slot_bytes - lay.info_off >= 4 * ne
        assert lay.flags_off == n_ranks * lay.slot_bytes
        assert lay.buffer_bytes >= lay.flags_off + 8 * n_ranks


def test_exchange_argument_checks_without_a_gpu():
    lib = E.load_library()
    lay = E.ExchangeLayout()
    assert lib.q3dr_exchange_layout_of(0, 1, 1, C.byref(lay)) == E.Q3DR_ERR_INVALID_ARG
    h = C.c_void_p()
    assert lib.q3dr_exchange_create(None, 2, 0, 1, 1, C.byref(h)) == E.Q3DR_ERR_INVALID_ARG and not h.value
    assert lib.q3dr_exchange_begin_step(None) == E.Q3DR_ERR_INVALID_ARG
    assert lib.q3dr_exchange_end_step(None, 0, None) == E.Q3DR_ERR_INVALID_ARG
    assert lib.q3dr_exchange_destroy(None) == E.Q3DR_OK
    assert lib.q3dr_multi_enable_gather(None, 1, 1) == E.Q3DR_ERR_INVALID_ARG


def _dev_array(ptr, n, dt, device):
    import torch

    from quad3dr_b200.dist import CudaArrayView

    if n == 0:
        return np.empty(0, dtype=dt)
    typestr = np.dtype(dt).str
    return torch.as_tensor(CudaArrayView(ptr, n, typestr), device=torch.device("cuda", device)).cpu().numpy()


def _warm_up(maps, slices, cam, w, h, devs):
    """One plain scoring call per rank at the largest size of the test: the result pools are then grown, and no rank
    calls cudaMalloc inside an exchange step.  Only needed because these tests put several ranks on ONE device: a
    device-synchronising call of one rank would wait for the flag-wait kernel of another rank on the same device, which
    waits for the first rank's flag (one rank per device, the supported deployment, has no such coupling)."""
    import torch

    for r, m in enumerate(maps):
        torch.cuda.set_device(devs[r])
        d_rt = torch.from_numpy(np.ascontiguousarray(slices[r])).to(torch.device("cuda", devs[r]))
        m.score_viewpoints_device(cam.K, w, (0, w, 0, h), d_rt.data_ptr(), len(slices[r]))
        m.result_to_host()
        torch.cuda.synchronize(devs[r])


@pytest.mark.gpu
def test_two_ranks_exchange_their_results(monkeypatch):
    import torch

    monkeypatch.setenv("Q3DR_CHUNK", "4")  # several chunks per call: several pushes per step
    monkeypatch.setenv("Q3DR_EXCHANGE_TIMEOUT_MS", "8000")
    devs = [0, 1] if E.device_count() >= 2 else [0, 0]
    leaves = synth.map_small()
    order, depth = E.splitmedian_order(leaves.boxes)
    leaves = leaves.permuted(order)
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(leaves, 23, config_id=31)
    slices = [vps[:13], vps[13:]]
    maps = [E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth, device=d) for d in devs]
    maps[1].set_result_fields(first_pixel=False, dist_sq=False)  # the exchange ships (leaf, information) either way
    _warm_up(maps, slices, cam, w, h, devs)
    exs = [E.Exchange(maps[r], 2, r, 13, 13 * w * h) for r in range(2)]
    E.Exchange.connect(exs)
    own = [None, None]
    errors = []

    def rank_main(r, n_use, barrier):
        try:
            torch.cuda.set_device(devs[r])
            stream = torch.cuda.Stream(device=devs[r])
            d_rt = torch.from_numpy(np.ascontiguousarray(slices[r])).to(torch.device("cuda", devs[r]))
            torch.cuda.synchronize(devs[r])
            barrier.wait()
            exs[r].begin_step()
            maps[r].score_viewpoints_device(cam.K, w, (0, w, 0, h), d_rt.data_ptr(), n_use, stream=stream.cuda_stream)
            exs[r].end_step(0, stream.cuda_stream)
            own[r] = maps[r].result_to_host()
        except Exception as e:  # noqa: BLE001 -- carried to the main thread
            errors.append((r, e))

    for step, n_use in enumerate([(13, 10), (9, 10), (13, 0)]):
        barrier = threading.Barrier(2)
        ts = [threading.Thread(target=rank_main, args=(r, n_use[r], barrier)) for r in range(2)]
        [t.start() for t in ts]
        [t.join(timeout=120) for t in ts]
        assert not errors, errors
        for on in range(2):
            for src in range(2):
                v = exs[on].view(src)
                ref = own[src]
                assert v.n_viewpoints == n_use[src] == ref.n_viewpoints and v.n_entries == ref.n_entries, (step, on, src)
                if n_use[src] == 0:
                    continue
                assert np.array_equal(_dev_array(v.offsets, n_use[src] + 1, np.uint64, devs[on]), ref.offsets)
                assert np.array_equal(_dev_array(v.total, n_use[src], np.uint32, devs[on]), ref.total.view(np.uint32))
                assert np.array_equal(_dev_array(v.leaf, v.n_entries, np.int32, devs[on]), ref.leaf)
                assert np.array_equal(_dev_array(v.information, v.n_entries, np.uint32, devs[on]), ref.information.view(np.uint32))
        assert own[0].n_entries > 500
    for ex in exs:
        ex.close()

    # a result that does not fit the receive slot of ONE rank: both ranks fail, nobody hangs, and a fresh exchange works
    exs = [E.Exchange(maps[r], 2, r, 13, 64 if r == 0 else 13 * w * h) for r in range(2)]
    E.Exchange.connect(exs)
    barrier = threading.Barrier(2)
    ts = [threading.Thread(target=rank_main, args=(r, 10, barrier)) for r in range(2)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert sorted(r for r, _ in errors) == [0, 1], errors
    assert all(isinstance(e, E.Q3drError) and "exchange" in str(e) for _, e in errors)
    errors.clear()
    for ex in exs:
        ex.close()
    for m in maps:
        m.close()


@pytest.mark.gpu
def test_four_ranks_exchange_over_per_target_streams(monkeypatch):
    """Four ranks (host threads; devices round-robin over what the box has): every rank pushes to three peers and
    itself on four different streams, several chunks per step, ragged slice sizes including an empty one."""
    import torch

    monkeypatch.setenv("Q3DR_CHUNK", "3")
    monkeypatch.setenv("Q3DR_EXCHANGE_TIMEOUT_MS", "8000")
    G = 4
    n_dev = max(1, E.device_count())
    devs = [r % n_dev for r in range(G)]
    leaves = synth.map_small()
    order, depth = E.splitmedian_order(leaves.boxes)
    leaves = leaves.permuted(order)
    w, h = 80, 60
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(leaves, 32, config_id=37)
    slices = [vps[8 * r: 8 * r + 8] for r in range(G)]
    maps = [E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal, depth, device=d) for d in devs]
    _warm_up(maps, slices, cam, w, h, devs)
    exs = [E.Exchange(maps[r], G, r, 8, 8 * w * h) for r in range(G)]
    E.Exchange.connect(exs)
    own = [None] * G
    errors = []

    def rank_main(r, n_use, barrier):
        try:
            torch.cuda.set_device(devs[r])
            stream = torch.cuda.Stream(device=devs[r])
            d_rt = torch.from_numpy(np.ascontiguousarray(slices[r])).to(torch.device("cuda", devs[r]))
            torch.cuda.synchronize(devs[r])
            barrier.wait()
            exs[r].begin_step()
            maps[r].score_viewpoints_device(cam.K, w, (0, w, 0, h), d_rt.data_ptr(), n_use, stream=stream.cuda_stream)
            exs[r].end_step(0, stream.cuda_stream)
            own[r] = maps[r].result_to_host()
        except Exception as e:  # noqa: BLE001 -- carried to the main thread
            errors.append((r, e))

    for step, n_use in enumerate([(8, 7, 8, 5), (0, 8, 3, 8), (8, 8, 8, 8)]):
        barrier = threading.Barrier(G)
        ts = [threading.Thread(target=rank_main, args=(r, n_use[r], barrier)) for r in range(G)]
        [t.start() for t in ts]
        [t.join(timeout=120) for t in ts]
        assert not errors, errors
        for on in range(G):
            for src in range(G):
                v = exs[on].view(src)
                ref = own[src]
                assert v.n_viewpoints == n_use[src] == ref.n_viewpoints and v.n_entries == ref.n_entries, (step, on, src)
                if n_use[src] == 0:
                    continue
                assert np.array_equal(_dev_array(v.offsets, n_use[src] + 1, np.uint64, devs[on]), ref.offsets)
                assert np.array_equal(_dev_array(v.total, n_use[src], np.uint32, devs[on]), ref.total.view(np.uint32))
                assert np.array_equal(_dev_array(v.leaf, v.n_entries, np.int32, devs[on]), ref.leaf)
                assert np.array_equal(_dev_array(v.information, v.n_entries, np.uint32, devs[on]), ref.information.view(np.uint32))
    for ex in exs:
        ex.close()
    for m in maps:
        m.close()


@pytest.mark.gpu
def test_multi_gather_leaves_all_parts_on_every_gpu():
    n_dev = E.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    leaves = synth.map_small()
    order, depth = E.splitmedian_order(leaves.boxes)
    leaves = leaves.permuted(order)
    w, h = 96, 72
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(leaves, 29, config_id=33)
    multi = E.MultiGpuMap(leaves.boxes, leaves.weight, leaves.normal, depth, devices=devices)
    G = multi.n_devices
    multi.enable_gather(29, 29 * w * h)
    for n_use in (29, 7, 29):
        parts = multi.score_viewpoints(cam.K, w, (0, w, 0, h), vps[:n_use])
        for on in range(G):
            for src, (first, ref) in enumerate(parts):
                v = multi.gathered(on, src)
                assert v.n_viewpoints == ref.n_viewpoints and v.n_entries == ref.n_entries
                if ref.n_viewpoints == 0:
                    continue
                assert np.array_equal(_dev_array(v.offsets, ref.n_viewpoints + 1, np.uint64, devices[on]), ref.offsets)
                assert np.array_equal(_dev_array(v.leaf, v.n_entries, np.int32, devices[on]), ref.leaf)
                assert np.array_equal(_dev_array(v.information, v.n_entries, np.uint32, devices[on]), ref.information.view(np.uint32))
                assert np.array_equal(_dev_array(v.total, ref.n_viewpoints, np.uint32, devices[on]), ref.total.view(np.uint32))
    multi.enable_gather(0, 0)
    with pytest.raises(E.Q3drError):
        multi.gathered(0, 0)
    multi.close()
