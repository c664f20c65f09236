"""ctypes binding of oracle/_ref/libq3dr_refcuda_{ieee,fast}.so: the REFERENCE's own device code
(viewpoint_planner/src/bvh/bvh.cu, compiled from the reference checkout by oracle/ref_cuda/Makefile).

TEST / BASELINE INFRASTRUCTURE ONLY (tests/ and bench.py).  The product never imports this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SYMBOLS = ["refcuda_last_error", "refcuda_create", "refcuda_destroy", "refcuda_intersect", "refcuda_raycast",
           "refcuda_time_raycasts"]


def lib_path(flavour: str) -> str:
    assert flavour in ("ieee", "fast")
    return os.path.join(REF_DIR, f"libq3dr_refcuda_{flavour}.so")


def available(flavour: str = "ieee") -> bool:
    return os.path.exists(lib_path(flavour))


_libs = {}


def lib(flavour: str):
    if flavour in _libs:
        return _libs[flavour]
    L = C.CDLL(lib_path(flavour))
    fp = C.POINTER(C.c_float)
    i32p = C.POINTER(C.c_int32)
    u32p = C.POINTER(C.c_uint32)
    L.refcuda_last_error.restype = C.c_char_p
    L.refcuda_create.argtypes = [fp, i32p, i32p, C.c_int32, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    L.refcuda_destroy.argtypes = [C.c_void_p]
    L.refcuda_destroy.restype = None
    L.refcuda_intersect.argtypes = [C.c_void_p, fp, fp, C.c_size_t, C.c_float, C.c_float, i32p, fp, fp, u32p]
    L.refcuda_raycast.argtypes = [C.c_void_p, fp, fp] + [C.c_uint32] * 4 + [C.c_float, C.c_float, i32p, fp, fp, u32p, fp]
    L.refcuda_time_raycasts.argtypes = [C.c_void_p, fp, fp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_float,
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    _libs[flavour] = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class RefCudaTree:
    """bvh::CudaTree<float> of the reference over the oracle's splitMedian tree (OracleTree.export_nodes())."""

    def __init__(self, oracle_tree, flavour: str = "ieee", patch_tail: bool = True, stack_bytes: int = 32 * 1024):
        self.L = lib(flavour)
        boxes, left, right, leaf, root = oracle_tree.export_nodes()
        self.node_leaf = leaf  # node index -> input leaf index (-1 inner)
        self.nn = boxes.shape[0]
        h = C.c_void_p()
        st = self.L.refcuda_create(_f(boxes), _i(left), _i(right), root, self.nn, oracle_tree.depth, int(patch_tail),
                                   stack_bytes, C.byref(h))
        if st != 0:
            raise RuntimeError("refcuda_create: " + self.L.refcuda_last_error().decode())
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.refcuda_destroy(self._h)
            self._h = None

    def _leaf_of(self, node_index):
        out = np.full(node_index.shape, -1, dtype=np.int32)
        hit = node_index >= 0
        out[hit] = self.node_leaf[node_index[hit]]
        return out

    def intersect(self, origins, directions, min_range=0.0, max_range=-1.0):
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(directions, dtype=np.float32)
        n = o.shape[0]
        node = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        xyz = np.empty((n, 3), dtype=np.float32)
        depth = np.empty(n, dtype=np.uint32)
        st = self.L.refcuda_intersect(self._h, _f(o), _f(d), n, min_range, max_range, _i(node), _f(dsq), _f(xyz),
                                      depth.ctypes.data_as(C.POINTER(C.c_uint32)))
        if st != 0:
            raise RuntimeError("refcuda_intersect: " + self.L.refcuda_last_error().decode())
        return dict(node=node, leaf=self._leaf_of(node), dsq=dsq, xyz=xyz, depth=depth)

    def raycast(self, K, Rt, x0, x1, y0, y1, min_range=0.0, max_range=-1.0):
        K = np.ascontiguousarray(K, dtype=np.float32)
        Rt = np.ascontiguousarray(Rt, dtype=np.float32).reshape(12)
        n = (x1 - x0) * (y1 - y0)
        node = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        xyz = np.empty((n, 3), dtype=np.float32)
        depth = np.empty(n, dtype=np.uint32)
        sxy = np.empty((n, 2), dtype=np.float32)
        st = self.L.refcuda_raycast(self._h, _f(K), _f(Rt), x0, x1, y0, y1, min_range, max_range, _i(node), _f(dsq),
                                    _f(xyz), depth.ctypes.data_as(C.POINTER(C.c_uint32)), _f(sxy))
        if st != 0:
            raise RuntimeError("refcuda_raycast: " + self.L.refcuda_last_error().decode())
        return dict(node=node, leaf=self._leaf_of(node), dsq=dsq, xyz=xyz, depth=depth, screen_xy=sxy)

    def time_raycasts(self, K, Rt, width, height, max_range=60.0, kernel_only=True):
        K = np.ascontiguousarray(K, dtype=np.float32)
        Rt = np.ascontiguousarray(Rt, dtype=np.float32).reshape(-1, 12)
        api = C.c_double(0)
        ker = C.c_double(0)
        hits = C.c_uint64(0)
        st = self.L.refcuda_time_raycasts(self._h, _f(K), _f(Rt), Rt.shape[0], width, height, max_range, C.byref(api),
                                          C.byref(ker) if kernel_only else None, C.byref(hits))
        if st != 0:
            raise RuntimeError("refcuda_time_raycasts: " + self.L.refcuda_last_error().decode())
        return dict(api_ms=api.value, kernel_ms=ker.value, hits=int(hits.value), rays=Rt.shape[0] * width * height)
