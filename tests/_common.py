"""Shared helpers for the parity tests: a map in tie-break (splitMedian) order + its oracle tree."""
from __future__ import annotations

import functools

import numpy as np

from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth


class Scene:
    def __init__(self, leaves: synth.OccupancyLeaves):
        order, depth = E.splitmedian_order(leaves.boxes)
        self.order = order
        self.leaves = leaves.permuted(order)
        self.depth = depth
        self.oracle = O.OracleTree(self.leaves.boxes)
        # leaves handed over in the reference's leaf order are a fixed point of Tree::build
        assert np.array_equal(self.oracle.dfs_order(), np.arange(self.leaves.n, dtype=np.uint32))
        self._map = None

    @property
    def gpu_map(self) -> E.OccupancyMap:
        if self._map is None:
            self._map = E.OccupancyMap(self.leaves.boxes, self.leaves.weight, self.leaves.normal, self.depth)
        return self._map


@functools.lru_cache(maxsize=None)
def scene(name: str) -> Scene:
    factory = {
        "tiny": synth.map_tiny,
        "tiny_unknown": lambda: synth.map_tiny(unknown_interior=True),
        "small": synth.map_small,
        "small_unknown": lambda: synth.map_small(unknown_interior=True),
        "1m": synth.map_1m,
        "1m_unknown": lambda: synth.map_1m(unknown_interior=True),
    }[name]
    return Scene(factory())


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
