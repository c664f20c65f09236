"""ctypes binding of oracle/_ref/libq3dr_refcpu.so: the REFERENCE's own CPU code -- bvh::Tree
(viewpoint_planner/src/bvh/bvh.h) over bh::BoundingBox3D / bh::Ray (include/bh/math/geometry.h) -- compiled from the
reference checkout by oracle/ref_cpu/Makefile against stand-in Eigen / Boost headers.

TEST INFRASTRUCTURE ONLY (tests/test_refcpu_pin.py).  The product never imports this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libq3dr_refcpu.so")
SYMBOLS = ["refcpu_tree_build", "refcpu_tree_free", "refcpu_tree_depth", "refcpu_tree_num_nodes", "refcpu_tree_num_leaves",
           "refcpu_leaf_order", "refcpu_root_box", "refcpu_voxel_index", "refcpu_intersect_rays", "refcpu_overlap_box",
           "refcpu_leaf_box"]

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(LIB_PATH)
    fp = C.POINTER(C.c_float)
    i32p = C.POINTER(C.c_int32)
    u32p = C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.refcpu_tree_build.argtypes = [fp, C.c_size_t]
    L.refcpu_tree_build.restype = vp
    L.refcpu_tree_free.argtypes = [vp]
    L.refcpu_tree_free.restype = None
    for name in ("refcpu_tree_depth", "refcpu_tree_num_nodes", "refcpu_tree_num_leaves"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_size_t
    L.refcpu_leaf_order.argtypes = [vp, u32p, u32p, fp]
    L.refcpu_leaf_order.restype = None
    L.refcpu_root_box.argtypes = [vp, fp]
    L.refcpu_root_box.restype = None
    L.refcpu_voxel_index.argtypes = [vp, u32p]
    L.refcpu_voxel_index.restype = None
    L.refcpu_intersect_rays.argtypes = [vp, fp, fp, C.c_size_t, C.c_float, C.c_float, i32p, fp, fp, u32p]
    L.refcpu_intersect_rays.restype = None
    L.refcpu_overlap_box.argtypes = [vp, fp, C.c_size_t, i32p]
    L.refcpu_overlap_box.restype = C.c_size_t
    L.refcpu_leaf_box.argtypes = [fp, C.c_float, fp, fp]
    L.refcpu_leaf_box.restype = C.c_int
    _lib = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefCpuTree:
    """bvh::Tree<Obj, float> built by the reference's own Tree::build from boxes in the given order."""

    def __init__(self, boxes):
        b = _f32(boxes).reshape(-1, 6)
        self.n = b.shape[0]
        self._h = lib().refcpu_tree_build(_f(b), self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().refcpu_tree_free(self._h)
            self._h = None

    @property
    def depth(self) -> int:
        return int(lib().refcpu_tree_depth(self._h))

    @property
    def num_nodes(self) -> int:
        return int(lib().refcpu_tree_num_nodes(self._h))

    @property
    def num_leaves(self) -> int:
        return int(lib().refcpu_tree_num_leaves(self._h))

    def leaf_order(self):
        """(order, depth, boxes): input index, depth and box of the k-th leaf, left to right."""
        order = np.empty(self.n, dtype=np.uint32)
        depth = np.empty(self.n, dtype=np.uint32)
        boxes = np.empty((self.n, 6), dtype=np.float32)
        lib().refcpu_leaf_order(self._h, order.ctypes.data_as(C.POINTER(C.c_uint32)), depth.ctypes.data_as(C.POINTER(C.c_uint32)),
                                _f(boxes))
        return order, depth, boxes

    def root_box(self) -> np.ndarray:
        out = np.empty(6, dtype=np.float32)
        lib().refcpu_root_box(self._h, _f(out))
        return out

    def voxel_index(self) -> np.ndarray:
        out = np.full(self.n, 0xFFFFFFFF, dtype=np.uint32)
        lib().refcpu_voxel_index(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def intersect_rays(self, origins, directions, min_range=0.0, max_range=-1.0):
        """Tree::intersects(bh::Ray(o, d), min_range, max_range): leaf = left-to-right leaf rank (-1 = miss)."""
        o = _f32(origins).reshape(-1, 3)
        d = _f32(directions).reshape(-1, 3)
        n = o.shape[0]
        leaf = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        xyz = np.empty((n, 3), dtype=np.float32)
        depth = np.empty(n, dtype=np.uint32)
        lib().refcpu_intersect_rays(self._h, _f(o), _f(d), n, min_range, max_range, leaf.ctypes.data_as(C.POINTER(C.c_int32)),
                                    _f(dsq), _f(xyz), depth.ctypes.data_as(C.POINTER(C.c_uint32)))
        return dict(leaf=leaf, dsq=dsq, xyz=xyz, depth=depth)

    def overlap_box(self, box) -> np.ndarray:
        b = _f32(box).reshape(6)
        n = lib().refcpu_overlap_box(self._h, _f(b), 0, None)
        out = np.empty(n, dtype=np.int32)
        if n:
            lib().refcpu_overlap_box(self._h, _f(b), n, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out


def leaf_box(center, size, constrain=None):
    """BoundingBox3D(center, size) [.constrainTo(constrain)]: (box, isEmpty())."""
    out = np.empty(6, dtype=np.float32)
    c = _f32(center).reshape(3)
    k = None if constrain is None else _f32(constrain).reshape(6)
    empty = lib().refcpu_leaf_box(_f(c), float(size), None if k is None else _f(k), _f(out))
    return out, bool(empty)
