"""OctoMap ".ot" streams: an independent pure-Python restatement of the format (reader) and a synthetic writer.

TEST INFRASTRUCTURE ONLY (tests/test_octree_io.py): the product's reader is quad3dr_b200/csrc/octree_io.cpp.  Both
restate OctoMap's published file format and iterator arithmetic -- OctoMap is an un-vendored dependency of the
reference (viewpoint_planner/CMakeLists.txt:9), absent here, so parity against the library itself is UNPINNED:
  * stream: AbstractOcTree::write header ("# Octomap OcTree file", id / size / res / data), then
    OcTreeBaseImpl::writeNodesRecurs: node payload, a byte with one bit per existing child, the children 0 .. 7
  * payloads of the reference's nodes: viewpoint_planner/src/octree/occupancy_node.cpp:139-211
  * leaf walk: begin_tree() = pre-order, children 0 .. 7 (the iterator pushes them 7 .. 0 on a stack)
  * geometry: computeChildKey + keyToCoord(key, depth) in double, cast to float; size = res * 2^(16 - depth)
    (what generateBVHTree reads through it.getCoordinate() / it.getSize(), viewpoint_planner_data.cpp:834-838)
A tree is a nested dict: {"occ": float, "count": int, "weight": float, "children": {index: subtree}}.
"""
from __future__ import annotations

import math
import struct
from typing import Dict, List, Tuple

import numpy as np

TREE_DEPTH = 16
TREE_MAX_VAL = 32768
PAYLOAD = {"occupancy": 8, "augmented": 20, "logodds": 4}


def _payload_bytes(node: dict, kind: str) -> bytes:
    if kind == "logodds":
        p = min(max(float(node["occ"]), 1e-6), 1 - 1e-6)
        return struct.pack("<f", math.log(p / (1 - p)))
    b = struct.pack("<fI", float(node["occ"]), int(node["count"]))
    if kind == "augmented":
        b += struct.pack("<fQ", float(node.get("weight", 0.0)), int(node.get("count_sum", 0)))
    return b


def count_nodes(tree: dict) -> int:
    return 1 + sum(count_nodes(c) for c in tree.get("children", {}).values())


def write_ot(path: str, tree: dict, resolution: float, kind: str = "augmented", tree_id: str = None, comments: int = 2):
    """AbstractOcTree::write + OcTreeBaseImpl::writeNodesRecurs."""
    tree_id = tree_id or ("OcTree" if kind == "logodds" else "Ait_OccupancyMap")
    out = bytearray()

    def rec(node):
        out.extend(_payload_bytes(node, kind))
        ch = node.get("children", {})
        mask = 0
        for i in ch:
            mask |= 1 << i
        out.append(mask)
        for i in range(8):
            if i in ch:
                rec(ch[i])

    rec(tree)
    with open(path, "wb") as f:
        f.write(b"# Octomap OcTree file\n")
        for _ in range(comments):
            f.write(b"# (feel free to add / change comments, but leave the first line as it is!)\n#\n")
        f.write(f"id {tree_id}\nsize {count_nodes(tree)}\nres {resolution!r}\ndata\n".encode())
        f.write(bytes(out))


def key_to_coord(key: int, depth: int, res: float) -> float:
    """OcTreeBaseImpl::keyToCoord(key, depth): double arithmetic, the caller casts to float."""
    if depth == 0:
        return 0.0
    if depth == TREE_DEPTH:
        return (float(int(key) - TREE_MAX_VAL) + 0.5) * res
    return (math.floor((float(key) - float(TREE_MAX_VAL)) / float(1 << (TREE_DEPTH - depth))) + 0.5) * (res * float(1 << (TREE_DEPTH - depth)))


def child_key(pos: int, offset: int, parent: Tuple[int, int, int]) -> Tuple[int, int, int]:
    """octomap::computeChildKey"""
    k = []
    for a in range(3):
        if pos & (1 << a):
            k.append(parent[a] + offset)
        else:
            k.append(parent[a] - offset - (0 if offset else 1))
    return tuple(k)


def read_ot(path: str, kind: str):
    """(header dict, leaves) with leaves = arrays center (n,3) f32, size, occupancy, count, weight, depth in begin_tree() order."""
    data = open(path, "rb").read()
    pos = 0
    header: Dict[str, str] = {}
    first = True
    while True:
        nl = data.index(b"\n", pos)
        line = data[pos:nl].decode()
        pos = nl + 1
        if first:
            assert line.startswith("# Octomap OcTree file")
            first = False
            continue
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        if tok[0] == "data":
            break
        header[tok[0]] = tok[1]
    res = float(header["res"])
    pb = PAYLOAD[kind]
    leaves: List[tuple] = []
    n_nodes = 0
    stack = [((TREE_MAX_VAL,) * 3, 0)]
    while stack and int(header["size"]) > 0:
        key, depth = stack.pop()
        if kind == "logodds":
            (l,) = struct.unpack_from("<f", data, pos)
            occ, cnt, w = 1.0 - 1.0 / (1.0 + math.exp(l)), 1, 0.0
        else:
            occ, cnt = struct.unpack_from("<fI", data, pos)
            w = struct.unpack_from("<f", data, pos + 8)[0] if kind == "augmented" else 0.0
        pos += pb
        mask = data[pos]
        pos += 1
        n_nodes += 1
        if mask == 0:
            leaves.append(([np.float32(key_to_coord(key[a], depth, res)) for a in range(3)],
                           np.float32(res * float(1 << (TREE_DEPTH - depth))), np.float32(occ), cnt, np.float32(w), depth))
            continue
        off = TREE_MAX_VAL >> (depth + 1)
        for i in range(7, -1, -1):
            if mask & (1 << i):
                stack.append((child_key(i, off, key), depth + 1))
    assert pos == len(data) and n_nodes == int(header["size"]), "stream does not match its header"
    n = len(leaves)
    out = dict(center=np.array([l[0] for l in leaves], dtype=np.float32).reshape(n, 3),
               size=np.array([l[1] for l in leaves], dtype=np.float32), occupancy=np.array([l[2] for l in leaves], dtype=np.float32),
               count=np.array([l[3] for l in leaves], dtype=np.uint32), weight=np.array([l[4] for l in leaves], dtype=np.float32),
               depth=np.array([l[5] for l in leaves], dtype=np.uint32))
    return header, out


def octree_from_boxes(center_keys: np.ndarray, levels: np.ndarray, payload) -> dict:
    """Builds the nested-dict octree whose leaves are the aligned cubes (key of the cube's minimum corner at the finest
    level, level j = cube of 2^j finest voxels per edge); payload(i) -> dict(occ, count, weight) of leaf i.  Inner nodes
    get the max occupancy / min count of their children (updateInnerOccupancy's rule, not read by the engine)."""
    root: dict = {"children": {}}
    for i in range(center_keys.shape[0]):
        depth = TREE_DEPTH - int(levels[i])
        node = root
        key = [int(k) for k in center_keys[i]]  # any finest-level key inside the cube
        for d in range(depth):
            bit = TREE_DEPTH - 1 - d
            idx = sum((((key[a] >> bit) & 1) << a) for a in range(3))
            node = node.setdefault("children", {}).setdefault(idx, {})
        assert not node.get("children"), "cubes overlap"
        node.update(payload(i))

    def fill(node):
        ch = node.get("children", {})
        if ch:
            for c in ch.values():
                fill(c)
            node["occ"] = max(c["occ"] for c in ch.values())
            node["count"] = min(c["count"] for c in ch.values())
            node["weight"] = 0.0

    fill(root)
    return root
