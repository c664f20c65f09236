"""Box-vs-map overlap queries: Tree::intersects(bbox) and isValidObjectPosition."""
import numpy as np
import pytest

from _common import scene
from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


def _cubes():
    c = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0], [5, 5, 5]], np.float32)
    return np.concatenate([c - 0.5, c + 0.5], axis=1).astype(np.float32)


def test_oracle_known_answers():
    t = O.OracleTree(_cubes())
    assert sorted(t.overlap_box([-0.4, -0.4, -0.4, 0.4, 0.4, 0.4])) == [0]
    # touching counts: max == other.min is not "<"
    assert sorted(t.overlap_box([0.5, -0.4, -0.4, 0.5, 0.4, 0.4])) == [0, 1]
    assert sorted(t.overlap_box([0.6, 0.6, 0.6, 4.4, 4.4, 4.4])) == []
    assert sorted(t.overlap_box([-9, -9, -9, 9, 9, 9])) == [0, 1, 2, 3, 4]
    obj = np.array([-0.25, -0.25, -0.25, 0.25, 0.25, 0.25], np.float32)
    bvh = np.array([-3, -3, -3, 8, 8, 8], np.float32)
    assert not t.valid_object_position([0.6, 0.0, 0.0], obj, bvh, 50.0)   # overlaps cubes 0 and 1
    assert t.valid_object_position([3.5, 3.0, 3.0], obj, bvh, 50.0)       # free space inside the root box
    assert not t.valid_object_position([7.0, 7.0, 7.0], obj, bvh, 50.0)   # outside the root box, below the free height
    assert t.valid_object_position([7.0, 7.0, 7.0], obj, bvh, 6.5)        # above the free height and inside bvh_bbox
    assert not t.valid_object_position([9.0, 7.0, 7.0], obj, bvh, 6.5)    # above the free height, outside bvh_bbox


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_unknown", "small"])
def test_overlap_boxes_match_oracle(name):
    sc = scene(name)
    rng = np.random.default_rng(31)
    root = sc.oracle.root_box()
    n = 600
    c = rng.uniform(root[:3] - 1, root[3:] + 1, size=(n, 3))
    c[: n // 2, 2] = rng.uniform(0, 4, size=n // 2)  # plenty near the ground / walls
    s = rng.choice([0.05, 0.3, 0.75, 2.0, 6.0], size=(n, 1))
    boxes = np.concatenate([c - s, c + s], axis=1).astype(np.float32)
    boxes[0] = np.concatenate([root[:3] - 1, root[3:] + 1])       # everything
    boxes[1] = sc.leaves.boxes[17]                                  # exactly one voxel's box: touches its neighbours
    boxes[2, :3] = boxes[2, 3:] = sc.leaves.boxes[5, 3:]           # a point on a voxel corner
    boxes[3] = [1e6, 1e6, 1e6, 2e6, 2e6, 2e6]                       # far away
    off, leaf = sc.gpu_map.overlap_boxes(boxes)
    assert off[0] == 0 and off[-1] == len(leaf)
    for i in range(n):
        ref = sc.oracle.overlap_box(boxes[i])
        got = leaf[off[i]:off[i + 1]]
        # leaves are handed to the engine in the reference's left-to-right order, so its visiting order is ascending
        assert np.array_equal(ref, np.sort(ref))
        assert np.array_equal(got, ref), i
    assert off[1] - off[0] == sc.leaves.n and off[4] - off[3] == 0 and off[2] - off[1] > 1
    o2, l2 = sc.gpu_map.overlap_boxes(np.zeros((0, 6), np.float32))
    assert len(o2) == 1 and len(l2) == 0


@pytest.mark.gpu
def test_valid_object_positions_match_oracle():
    sc = scene("small_unknown")
    rng = np.random.default_rng(37)
    root = sc.oracle.root_box()
    bvh = (root + np.array([2, 2, 0, -2, -2, 6], np.float32)).astype(np.float32)
    obj = np.array([-0.6, -0.6, -0.3, 0.6, 0.6, 0.3], np.float32)  # drone_bbox_
    n = 5000
    pos = rng.uniform(root[:3] - 2, root[3:] + np.array([2, 2, 12]), size=(n, 3)).astype(np.float32)
    pos[:200] = ((sc.leaves.boxes[:200, :3] + sc.leaves.boxes[:200, 3:]) / 2)      # inside voxels
    pos[200:400, 2] = sc.leaves.boxes[200:400, 5] + np.float32(0.3)                # box bottom exactly on a voxel top
    for free_h in (float(root[5]) + 1.0, 6.0, 2.5):
        got = sc.gpu_map.valid_object_positions(pos, obj, bvh, free_h)
        ref = np.array([sc.oracle.valid_object_position(pos[i], obj, bvh, free_h) for i in range(n)])
        assert np.array_equal(got, ref)
        assert 0.02 * n < got.sum() < 0.98 * n
