"""Row N1, host half: OctoMap ".ot" stream -> octree leaves in begin_tree() order (quad3dr_b200/csrc/octree_io.cpp),
against an independent Python restatement of the format and a synthetic writer (oracle/octree_ot.py), and end to end:
file -> leaves -> generateBVHTree's leaf rule -> map -> scores equal to the oracle's on the same leaves."""
import os

import numpy as np
import pytest

from oracle import octree_ot as OT
from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _random_tree(seed, n_cubes=400, res=0.25, span=64):
    """aligned cubes of mixed sizes around the origin, no overlaps; occupied / free / unknown payloads"""
    rng = np.random.default_rng(seed)
    taken = set()
    keys, levels = [], []
    for _ in range(n_cubes * 4):
        if len(keys) >= n_cubes:
            break
        lv = int(rng.choice([0, 0, 0, 1, 1, 2, 3]))
        e = 1 << lv
        k = (rng.integers(-span // e, span // e, 3) * e + OT.TREE_MAX_VAL).astype(np.int64)
        cells = {(int(k[0]) + i, int(k[1]) + j, int(k[2]) + l) for i in range(0, e, max(1, e // 2)) for j in range(0, e, max(1, e // 2))
                 for l in range(0, e, max(1, e // 2))}
        # reject cubes overlapping an existing one at any level
        blocks = {((int(k[0]) >> s) << s, (int(k[1]) >> s) << s, (int(k[2]) >> s) << s, s) for s in range(0, 5)}
        if blocks & taken or any((int(k[0]) + i, int(k[1]) + j, int(k[2]) + l, 0) in taken for i in range(e) for j in range(e) for l in range(e)
                                 ) if e <= 2 else blocks & taken:
            continue
        ok = True
        for s in range(lv):  # finer cubes already inside this one?
            step = 1 << s
            for i in range(0, e, step):
                for j in range(0, e, step):
                    for l in range(0, e, step):
                        if (int(k[0]) + i, int(k[1]) + j, int(k[2]) + l, s) in taken:
                            ok = False
        if not ok:
            continue
        taken.add((int(k[0]), int(k[1]), int(k[2]), lv))
        keys.append(k)
        levels.append(lv)
    keys = np.array(keys)
    levels = np.array(levels)
    kind = rng.integers(0, 3, len(keys))  # 0 occupied, 1 free + known, 2 unknown
    occ = np.where(kind == 0, 0.97, np.where(kind == 1, 0.12, 0.5)).astype(np.float32)
    cnt = np.where(kind == 2, 0, rng.integers(1, 20, len(keys))).astype(np.uint32)
    wgt = rng.random(len(keys)).astype(np.float32)
    tree = OT.octree_from_boxes(keys, levels, lambda i: dict(occ=float(occ[i]), count=int(cnt[i]), weight=float(wgt[i])))
    return tree, res


@pytest.mark.parametrize("kind,code", [("augmented", E.OCTREE_AUGMENTED), ("occupancy", E.OCTREE_OCCUPANCY), ("logodds", E.OCTREE_LOGODDS)])
def test_reader_equals_python_restatement(tmp_path, kind, code):
    tree, res = _random_tree(3)
    path = os.path.join(tmp_path, f"map_{kind}.ot")
    OT.write_ot(path, tree, res, kind)
    header, ref = OT.read_ot(path, kind)
    for payload in (code, E.OCTREE_AUTO):
        got = E.read_octree_leaves(path, payload)
        assert got["info"].id.decode() == header["id"] and got["info"].payload == code
        assert got["info"].n_nodes == int(header["size"]) == OT.count_nodes(tree) and got["info"].resolution == res
        assert got["info"].n_leaves == ref["size"].shape[0] > 300 and got["info"].tree_depth == 16
        assert np.array_equal(bits(got["center"]), bits(ref["center"]))
        assert np.array_equal(bits(got["size"]), bits(ref["size"]))
        assert np.array_equal(got["depth"], ref["depth"])
        assert np.array_equal(got["observation_count"], ref["count"])
        assert np.array_equal(bits(got["weight"]), bits(ref["weight"]))
        if kind == "logodds":
            assert np.allclose(got["occupancy"], ref["occupancy"], rtol=0, atol=1e-6)
        else:
            assert np.array_equal(bits(got["occupancy"]), bits(ref["occupancy"]))
    # leaf geometry: centres sit on the (k + 0.5) * size lattice and cubes do not overlap
    k = got["center"] / got["size"][:, None] - 0.5
    assert np.array_equal(k, np.round(k))
    lo = (got["center"] - got["size"][:, None] / 2) / res
    vol = float(((got["size"] / res) ** 3).sum())
    cells = set()
    for i in range(lo.shape[0]):
        e = int(round(got["size"][i] / res))
        if e == 1:
            cells.add(tuple(np.round(lo[i]).astype(int)))
    assert len(cells) == int((got["size"] == np.float32(res)).sum()) and vol > 0


def test_reader_rejects_bad_streams(tmp_path):
    tree, res = _random_tree(4, n_cubes=50)
    path = os.path.join(tmp_path, "ok.ot")
    OT.write_ot(path, tree, res, "augmented")
    data = open(path, "rb").read()
    bad = os.path.join(tmp_path, "bad.ot")
    for blob in (data[:-7], data + b"xx", data.replace(b"# Octomap OcTree file", b"# Something else file"), data.replace(b"size ", b"size 1")):
        open(bad, "wb").write(blob)
        with pytest.raises(E.Q3drError):
            E.read_octree_leaves(bad)
    with pytest.raises(E.Q3drError):
        E.read_octree_leaves(os.path.join(tmp_path, "missing.ot"))
    with pytest.raises(E.Q3drError):
        E.read_octree_leaves(path, E.OCTREE_OCCUPANCY)  # the other node type of the same id does not fit the stream
    # an empty tree is a valid stream
    empty = os.path.join(tmp_path, "empty.ot")
    open(empty, "wb").write(b"# Octomap OcTree file\nid Ait_OccupancyMap\nsize 0\nres 0.1\ndata\n")
    assert E.read_octree_leaves(empty)["info"].n_leaves == 0


def test_synthetic_map_round_trips_through_an_ot_file(tmp_path):
    """the bench's box-building leaves written as an octree file come back as the same cubes (any order: the file
    order is the octree's, not the generator's)"""
    leaves = synth.map_tiny(unknown_interior=True)
    res = leaves.resolution
    size = leaves.boxes[:, 3] - leaves.boxes[:, 0]
    cube = np.all(np.isclose(leaves.boxes[:, 3:] - leaves.boxes[:, :3], size[:, None]), axis=1)
    boxes, size = leaves.boxes[cube], size[cube]
    lv = np.round(np.log2(size / res)).astype(int)
    keys = np.round(boxes[:, :3] / res).astype(np.int64) + OT.TREE_MAX_VAL
    rng = np.random.default_rng(1)
    cnt = rng.integers(0, 9, boxes.shape[0])
    tree = OT.octree_from_boxes(keys, lv, lambda i: dict(occ=0.97 if cnt[i] else 0.5, count=int(cnt[i]), weight=0.5))
    path = os.path.join(tmp_path, "tiny.ot")
    OT.write_ot(path, tree, float(res), "augmented")
    got = E.read_octree_leaves(path)
    assert got["info"].n_leaves == boxes.shape[0]
    back = np.concatenate([got["center"] - got["size"][:, None] / 2, got["center"] + got["size"][:, None] / 2], axis=1)
    a = np.array(sorted(map(tuple, np.round(back / res * 2).astype(np.int64))))
    b = np.array(sorted(map(tuple, np.round(boxes / res * 2).astype(np.int64))))
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_ot_file_to_scores_end_to_end(tmp_path):
    """.ot file -> leaves -> leaf rule on the device -> reference leaf order -> map -> scores == oracle on those leaves"""
    tree, res = _random_tree(7, n_cubes=900, span=48)
    path = os.path.join(tmp_path, "scene.ot")
    OT.write_ot(path, tree, res, "augmented")
    lv = E.read_octree_leaves(path)
    bb = np.array([-20, -20, -20, 20, 20, 9], dtype=np.float32)
    boxes, src = E.octree_leaves_to_boxes(lv["center"], lv["size"], lv["occupancy"], lv["observation_count"], bb, 8.0)
    rb, rs = O.octree_leaves_to_boxes(lv["center"], lv["size"], lv["occupancy"], lv["observation_count"], bb, 8.0)
    assert np.array_equal(src, rs) and np.array_equal(bits(boxes), bits(rb)) and 100 < boxes.shape[0] < lv["size"].shape[0]
    weight = np.exp(-0.1 * lv["observation_count"][src].astype(np.float32)).astype(np.float32)
    order, depth = E.splitmedian_order(boxes)
    boxes, weight = boxes[order], weight[order]
    gmap = E.OccupancyMap(boxes, weight, None, depth)
    oracle = O.OracleTree(boxes)
    cam = synth.make_camera(96, 72)
    rt = np.array([[1, 0, 0, 0.3], [0, 0, 1, -30.1], [0, -1, 0, 0.7]], dtype=np.float32).reshape(1, 12)
    res_gpu = gmap.score_viewpoints(cam.K, 96, (0, 96, 0, 72), rt)
    ref = oracle.score_viewpoint(weight, np.zeros((boxes.shape[0], 3), np.float32), cam.K, 96, rt[0], height=72)
    assert np.array_equal(res_gpu.viewpoint(0)["leaf"], ref["leaf"]) and len(ref["leaf"]) > 20
    assert abs(res_gpu.viewpoint(0)["total"] - ref["total"]) <= 1e-5 * abs(ref["total"])
    gmap.close()
