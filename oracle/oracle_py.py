"""ctypes binding of the CPU oracle (oracle/libq3dr_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (quad3dr_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libq3dr_oracle.so")


class ScoreOpts(C.Structure):
    _fields_ = [
        ("voxel_sensor_size_ratio_threshold", C.c_float),
        ("voxel_sensor_size_ratio_inv_falloff_factor", C.c_float),
        ("incidence_angle_threshold_degrees", C.c_float),
        ("incidence_angle_inv_falloff_factor_degrees", C.c_float),
        ("incidence_ignore_dot_product_sign", C.c_int32),
        ("ignore_voxels_with_zero_information", C.c_int32),
        ("raycast_min_range", C.c_float),
        ("raycast_max_range", C.c_float),
    ]


def default_opts(**kw) -> ScoreOpts:
    o = ScoreOpts(0.002, 0.001, 30.0, 30.0, 0, 1, 0.0, 60.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build_library(force: bool = False) -> str:
    src = os.path.join(_HERE, "q3dr_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "libq3dr_oracle.so"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build_library()
    L = C.CDLL(_LIB_PATH)
    fp = C.POINTER(C.c_float)
    u32p = C.POINTER(C.c_uint32)
    i32p = C.POINTER(C.c_int32)
    L.orc_tree_build.restype = C.c_void_p
    L.orc_tree_build.argtypes = [fp, C.c_size_t]
    L.orc_tree_free.argtypes = [C.c_void_p]
    for f in ("orc_tree_num_leaves", "orc_tree_num_nodes", "orc_tree_depth"):
        getattr(L, f).restype = C.c_size_t
        getattr(L, f).argtypes = [C.c_void_p]
    L.orc_tree_dfs_order.argtypes = [C.c_void_p, u32p]
    L.orc_tree_leaf_depths.argtypes = [C.c_void_p, u32p]
    L.orc_tree_voxel_index.argtypes = [C.c_void_p, u32p]
    L.orc_tree_root_box.argtypes = [C.c_void_p, fp]
    L.orc_pose_to_rt.argtypes = [fp, fp, fp]
    L.orc_raycast_window.restype = C.c_uint64
    L.orc_raycast_window.argtypes = [C.c_void_p, fp, fp] + [C.c_uint32] * 4 + [C.c_float] * 2 + [i32p, fp, fp, u32p]
    L.orc_raycast_window_brute.argtypes = [C.c_void_p, fp, fp] + [C.c_uint32] * 4 + [C.c_float] * 2 + [i32p, fp]
    L.orc_intersect_rays.argtypes = [C.c_void_p, fp, fp, C.c_size_t, C.c_float, C.c_float, i32p, fp, fp, u32p]
    L.orc_intersect_rays_raw.argtypes = [C.c_void_p, fp, fp, C.c_size_t, C.c_float, C.c_float, i32p, fp, fp, u32p]
    L.orc_camera_rays.argtypes = [fp, fp] + [C.c_uint32] * 4 + [fp, fp]
    L.orc_set_sqnorm_sequential.argtypes = [C.c_int]
    L.orc_tree_export_nodes.argtypes = [C.c_void_p, fp, i32p, i32p, i32p, i32p]
    L.orc_voxel_information.restype = C.c_float
    L.orc_voxel_information.argtypes = [fp, C.c_float, fp, fp, C.c_float, C.c_uint32, C.POINTER(ScoreOpts)]
    L.orc_score_viewpoint.restype = C.c_size_t
    L.orc_score_viewpoint.argtypes = (
        [C.c_void_p, fp, fp, fp, C.c_uint32, fp]
        + [C.c_uint32] * 4
        + [C.POINTER(ScoreOpts), C.c_size_t, i32p, fp, u32p, fp, u32p, fp, fp]
    )
    L.orc_score_batch.restype = C.c_uint64
    L.orc_score_batch.argtypes = [
        C.c_void_p, fp, fp, fp, C.c_uint32, C.c_uint32, fp, C.c_size_t, C.POINTER(ScoreOpts), C.c_int, u32p, u32p, fp,
    ]
    L.orc_octree_leaf_to_box.restype = C.c_int
    L.orc_octree_leaf_to_box.argtypes = [fp, C.c_float, C.c_float, C.c_uint32, fp, C.c_float, C.c_float, C.c_uint32, fp]
    L.orc_update_weights_from_grid.argtypes = [fp, C.c_size_t, fp, C.c_float, i32p, fp, C.c_float, u32p, fp]
    L.orc_overlap_box.restype = C.c_size_t
    L.orc_overlap_box.argtypes = [C.c_void_p, fp, C.c_size_t, i32p]
    L.orc_valid_object_position.restype = C.c_int
    L.orc_valid_object_position.argtypes = [C.c_void_p, fp, fp, fp, C.c_float]
    dp = C.POINTER(C.c_double)
    L.orc_new_information.restype = C.c_float
    L.orc_new_information.argtypes = [i32p, fp, C.c_size_t, fp, fp, C.c_float, dp]
    L.orc_observed_add.restype = C.c_float
    L.orc_observed_add.argtypes = [i32p, fp, C.c_size_t, fp, fp, C.c_float, dp]
    L.orc_overlap_information.restype = C.c_float
    L.orc_overlap_information.argtypes = [i32p, fp, C.c_size_t, i32p, fp, C.c_size_t, dp]
    L.orc_score_batch_checksum.restype = C.c_uint64
    L.orc_score_batch_checksum.argtypes = [
        C.c_void_p, fp, fp, fp, C.c_uint32, C.c_uint32, fp, C.c_size_t, C.POINTER(ScoreOpts), C.c_int, u32p, u32p, fp,
        C.POINTER(C.c_uint64),
    ]
    L.orc_downweight_real_viewpoint.restype = C.c_uint64
    L.orc_downweight_real_viewpoint.argtypes = [C.c_void_p, fp, fp, u32p, fp, C.c_uint32, C.c_uint32, fp, fp,
                                                C.POINTER(ScoreOpts), C.c_float]
    L.orc_decode_normal.argtypes = [C.c_uint32, fp]
    L.orc_mesh_face_normals.argtypes = [fp, C.c_size_t, fp]
    L.orc_mesh_render_normals.argtypes = [fp, fp, C.c_size_t, fp, C.c_uint32, C.c_uint32, fp, C.c_float, C.c_float,
                                          C.c_uint32, u32p]
    L.orc_mesh_cast_pixels.argtypes = [fp, fp, C.c_size_t, fp, C.c_uint32, C.c_uint32, fp, C.c_float, C.c_float,
                                       C.c_uint32, u32p, u32p, C.c_size_t, u32p]
    L.orc_score_viewpoint_image_normals.restype = C.c_size_t
    L.orc_score_viewpoint_image_normals.argtypes = (
        [C.c_void_p, fp, u32p, C.c_uint32, fp, C.c_uint32, fp]
        + [C.c_uint32] * 4
        + [C.POINTER(ScoreOpts), C.c_size_t, i32p, fp, u32p, fp, u32p, fp]
    )
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


class OracleTree:
    """The reference's bvh::Tree (splitMedian build + recursive closest-hit), on the CPU."""

    def __init__(self, boxes: np.ndarray):
        self.boxes = _f32(boxes)
        assert self.boxes.ndim == 2 and self.boxes.shape[1] == 6
        self.n = self.boxes.shape[0]
        self._h = lib().orc_tree_build(_f(self.boxes), self.n)
        if not self._h:
            raise ValueError("orc_tree_build failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tree_free(self._h)
            self._h = None

    @property
    def num_nodes(self) -> int:
        return lib().orc_tree_num_nodes(self._h)

    @property
    def depth(self) -> int:
        return lib().orc_tree_depth(self._h)

    def dfs_order(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint32)
        lib().orc_tree_dfs_order(self._h, _u(out))
        return out

    def leaf_depths(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint32)
        lib().orc_tree_leaf_depths(self._h, _u(out))
        return out

    def voxel_index(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint32)
        lib().orc_tree_voxel_index(self._h, _u(out))
        return out

    def root_box(self) -> np.ndarray:
        out = np.empty(6, dtype=np.float32)
        lib().orc_tree_root_box(self._h, _f(out))
        return out

    def raycast_window(self, K, Rt, x0, x1, y0, y1, min_range=0.0, max_range=-1.0):
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(12)
        n = (x1 - x0) * (y1 - y0)
        leaf = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        xyz = np.empty((n, 3), dtype=np.float32)
        depth = np.empty(n, dtype=np.uint32)
        visits = lib().orc_raycast_window(
            self._h, _f(K), _f(Rt), x0, x1, y0, y1, min_range, max_range, _i(leaf), _f(dsq), _f(xyz), _u(depth)
        )
        return dict(leaf=leaf, dsq=dsq, xyz=xyz, depth=depth, visits=int(visits))

    def raycast_window_brute(self, K, Rt, x0, x1, y0, y1, min_range=0.0, max_range=-1.0):
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(12)
        n = (x1 - x0) * (y1 - y0)
        leaf = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        lib().orc_raycast_window_brute(self._h, _f(K), _f(Rt), x0, x1, y0, y1, min_range, max_range, _i(leaf), _f(dsq))
        return dict(leaf=leaf, dsq=dsq)

    def intersect_rays(self, origins, directions, min_range=0.0, max_range=-1.0):
        o = _f32(origins)
        d = _f32(directions)
        n = o.shape[0]
        leaf = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        xyz = np.empty((n, 3), dtype=np.float32)
        depth = np.empty(n, dtype=np.uint32)
        lib().orc_intersect_rays(self._h, _f(o), _f(d), n, min_range, max_range, _i(leaf), _f(dsq), _f(xyz), _u(depth))
        return dict(leaf=leaf, dsq=dsq, xyz=xyz, depth=depth)

    def intersect_rays_raw(self, origins, directions, min_range=0.0, max_range=-1.0):
        """Rays used as given (no normalisation), the form of the reference's CudaTree::intersectsRecursive."""
        o = _f32(origins)
        d = _f32(directions)
        n = o.shape[0]
        leaf = np.empty(n, dtype=np.int32)
        dsq = np.empty(n, dtype=np.float32)
        xyz = np.empty((n, 3), dtype=np.float32)
        depth = np.empty(n, dtype=np.uint32)
        lib().orc_intersect_rays_raw(self._h, _f(o), _f(d), n, min_range, max_range, _i(leaf), _f(dsq), _f(xyz), _u(depth))
        return dict(leaf=leaf, dsq=dsq, xyz=xyz, depth=depth)

    def export_nodes(self):
        """The binary splitMedian tree: (boxes nn x 6, left, right, leaf, root)."""
        nn = self.num_nodes
        boxes = np.empty((nn, 6), dtype=np.float32)
        left = np.empty(nn, dtype=np.int32)
        right = np.empty(nn, dtype=np.int32)
        leaf = np.empty(nn, dtype=np.int32)
        root = C.c_int32(0)
        lib().orc_tree_export_nodes(self._h, _f(boxes), _i(left), _i(right), _i(leaf), C.byref(root))
        return boxes, left, right, leaf, int(root.value)

    def overlap_box(self, box) -> np.ndarray:
        """Tree::intersects(bbox): overlapping leaves in the reference's visiting order."""
        b = _f32(box, (6,))
        n = lib().orc_overlap_box(self._h, _f(b), 0, None)
        out = np.empty(n, dtype=np.int32)
        if n:
            lib().orc_overlap_box(self._h, _f(b), n, _i(out))
        return out

    def valid_object_position(self, position, object_bbox, bvh_bbox, obstacle_free_height) -> bool:
        return bool(lib().orc_valid_object_position(self._h, _f(_f32(position, (3,))), _f(_f32(object_bbox, (6,))),
                                                    _f(_f32(bvh_bbox, (6,))), float(obstacle_free_height)))

    def score_viewpoint(self, weights, normals, K, width, Rt, window=None, height=None, opts: Optional[ScoreOpts] = None):
        """getRaycastHitVoxelsWithInformationScore for one viewpoint (sorted by leaf index)."""
        opts = opts or default_opts()
        w = _f32(weights, (self.n,))
        nr = _f32(normals, (self.n, 3))
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(12)
        if window is None:
            assert height is not None
            window = (0, width, 0, height)
        x0, x1, y0, y1 = window
        cap = (x1 - x0) * (y1 - y0)
        leaf = np.empty(cap, dtype=np.int32)
        info = np.empty(cap, dtype=np.float32)
        pix = np.empty(cap, dtype=np.uint32)
        dsq = np.empty(cap, dtype=np.float32)
        hc = C.c_uint32(0)
        tot = C.c_float(0)
        tot32 = C.c_float(0)
        k = lib().orc_score_viewpoint(
            self._h, _f(w), _f(nr), _f(K), width, _f(Rt), x0, x1, y0, y1, C.byref(opts), cap,
            _i(leaf), _f(info), _u(pix), _f(dsq), C.byref(hc), C.byref(tot), C.byref(tot32),
        )
        assert k != C.c_size_t(-1).value
        return dict(
            leaf=leaf[:k].copy(), info=info[:k].copy(), pixel=pix[:k].copy(), dsq=dsq[:k].copy(),
            hit_count=int(hc.value), total=float(tot.value), total_f32=float(tot32.value),
        )

    def score_viewpoint_image_normals(self, weights, normal_image, K, width, Rt, window=None,
                                      opts: Optional[ScoreOpts] = None):
        """The enable_opengl = true branch: normals decoded from `normal_image` (H x W ARGB32) at the kept pixels."""
        opts = opts or default_opts()
        w = _f32(weights, (self.n,))
        img = np.ascontiguousarray(normal_image, dtype=np.uint32)
        assert img.ndim == 2 and img.shape[1] == width
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(12)
        if window is None:
            window = (0, width, 0, img.shape[0])
        x0, x1, y0, y1 = window
        cap = (x1 - x0) * (y1 - y0)
        leaf = np.empty(cap, dtype=np.int32)
        info = np.empty(cap, dtype=np.float32)
        pix = np.empty(cap, dtype=np.uint32)
        dsq = np.empty(cap, dtype=np.float32)
        hc = C.c_uint32(0)
        tot = C.c_float(0)
        k = lib().orc_score_viewpoint_image_normals(
            self._h, _f(w), _u(img), img.shape[0], _f(K), width, _f(Rt), x0, x1, y0, y1, C.byref(opts), cap,
            _i(leaf), _f(info), _u(pix), _f(dsq), C.byref(hc), C.byref(tot),
        )
        assert k != C.c_size_t(-1).value
        return dict(leaf=leaf[:k].copy(), info=info[:k].copy(), pixel=pix[:k].copy(), dsq=dsq[:k].copy(),
                    hit_count=int(hc.value), total=float(tot.value))

    def downweight_real_viewpoint(self, weights, normals, K, width, height, Rt, depth_map, factor: float,
                                  opts: Optional[ScoreOpts] = None, normal_image=None):
        """updateWeightsWithRealViewpoints for one image: returns (new weights, number of updates)."""
        opts = opts or default_opts()
        w = _f32(weights, (self.n,)).copy()
        nr = _f32(normals, (self.n, 3))
        d = _f32(depth_map, (height, width))
        img = None if normal_image is None else np.ascontiguousarray(normal_image, dtype=np.uint32).reshape(height, width)
        k = lib().orc_downweight_real_viewpoint(self._h, _f(w), _f(nr), None if img is None else _u(img), _f(_f32(K, (4,))),
                                                width, height, _f(_f32(Rt).reshape(12)), _f(d), C.byref(opts), float(factor))
        return w, int(k)

    def score_batch(self, weights, normals, K, width, height, Rt, opts: Optional[ScoreOpts] = None, threads: int = 1):
        opts = opts or default_opts()
        w = _f32(weights, (self.n,))
        nr = _f32(normals, (self.n, 3))
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(-1, 12)
        n = Rt.shape[0]
        kept = np.empty(n, dtype=np.uint32)
        hit = np.empty(n, dtype=np.uint32)
        tot = np.empty(n, dtype=np.float32)
        chk = np.empty(n, dtype=np.uint64)
        rays = lib().orc_score_batch_checksum(
            self._h, _f(w), _f(nr), _f(K), width, height, _f(Rt), n, C.byref(opts), threads, _u(kept), _u(hit), _f(tot),
            chk.ctypes.data_as(C.POINTER(C.c_uint64)),
        )
        return dict(kept_count=kept, hit_count=hit, total=tot, rays=int(rays), set_checksum=chk)


def voxel_information(box, weight, normal, cam, fx, width, opts: Optional[ScoreOpts] = None) -> float:
    opts = opts or default_opts()
    return float(
        lib().orc_voxel_information(
            _f(_f32(box, (6,))), float(weight), _f(_f32(normal, (3,))), _f(_f32(cam, (3,))), float(fx), int(width),
            C.byref(opts),
        )
    )


def pose_to_rt(quat_wxyz, trans) -> np.ndarray:
    out = np.empty(12, dtype=np.float32)
    lib().orc_pose_to_rt(_f(_f32(quat_wxyz, (4,))), _f(_f32(trans, (3,))), _f(out))
    return out


def camera_rays(K, Rt, x0, x1, y0, y1):
    """(origins, normalised directions) of the window's pixel rays, row-major, as the query sees them."""
    K = _f32(K, (4,))
    Rt = _f32(Rt).reshape(12)
    n = (x1 - x0) * (y1 - y0)
    o = np.empty((n, 3), dtype=np.float32)
    d = np.empty((n, 3), dtype=np.float32)
    lib().orc_camera_rays(_f(K), _f(Rt), x0, x1, y0, y1, _f(o), _f(d))
    return o, d


def set_sqnorm_sequential(on: bool) -> None:
    """Pin-test switch (see q3dr_oracle.h); always reset to False afterwards."""
    lib().orc_set_sqnorm_sequential(1 if on else 0)


def octree_leaves_to_boxes(center, size, occupancy, count, bvh_bbox, obstacle_free_height, occ_thr=0.51, obs_thr=1):
    """generateBVHTree's leaf rule, leaf by leaf: (boxes (k, 6), source index (k,))."""
    center = _f32(center).reshape(-1, 3)
    size = _f32(size)
    occupancy = _f32(occupancy)
    count = np.ascontiguousarray(count, dtype=np.uint32)
    bb = _f32(bvh_bbox, (6,))
    boxes, src = [], []
    out = np.empty(6, dtype=np.float32)
    for i in range(center.shape[0]):
        c = np.ascontiguousarray(center[i])
        if lib().orc_octree_leaf_to_box(_f(c), float(size[i]), float(occupancy[i]), int(count[i]), _f(bb),
                                        float(obstacle_free_height), float(occ_thr), int(obs_thr), _f(out)):
            boxes.append(out.copy())
            src.append(i)
    return (np.array(boxes, dtype=np.float32).reshape(-1, 6), np.array(src, dtype=np.uint32))


def update_weights_from_grid(boxes, weight, grid_origin, grid_increment, grid_dim, grid_weight, lam, count):
    """updateWeights' BVH writes (brute force over cells x leaves); returns the new weights."""
    boxes = _f32(boxes).reshape(-1, 6)
    w = _f32(weight).copy()
    dim = np.ascontiguousarray(grid_dim, dtype=np.int32)
    lib().orc_update_weights_from_grid(_f(boxes), boxes.shape[0], _f(_f32(grid_origin, (3,))), float(grid_increment), _i(dim),
                                       _f(_f32(grid_weight)), float(lam), _u(np.ascontiguousarray(count, dtype=np.uint32)), _f(w))
    return w


def new_information(leaf, info, observed, weight, factor):
    """(fp32 running sum, fp64 sum) of computeNewInformation for one voxel set."""
    leaf = np.ascontiguousarray(leaf, dtype=np.int32)
    info = _f32(info)
    s = C.c_double(0)
    v = lib().orc_new_information(_i(leaf), _f(info), len(leaf), _f(_f32(observed)), _f(_f32(weight)), float(factor), C.byref(s))
    return float(v), float(s.value)


def observed_add(leaf, info, observed, weight, factor):
    """addObservedVoxelsToViewpointPath; `observed` (float32, contiguous) is updated in place."""
    leaf = np.ascontiguousarray(leaf, dtype=np.int32)
    info = _f32(info)
    assert observed.dtype == np.float32 and observed.flags.c_contiguous
    s = C.c_double(0)
    v = lib().orc_observed_add(_i(leaf), _f(info), len(leaf), _f(observed), _f(_f32(weight)), float(factor), C.byref(s))
    return float(v), float(s.value)


def overlap_information(leaf1, info1, leaf2, info2):
    leaf1 = np.ascontiguousarray(leaf1, dtype=np.int32)
    leaf2 = np.ascontiguousarray(leaf2, dtype=np.int32)
    info1 = _f32(info1)
    info2 = _f32(info2)
    s = C.c_double(0)
    v = lib().orc_overlap_information(_i(leaf1), _f(info1), len(leaf1), _i(leaf2), _f(info2), len(leaf2), C.byref(s))
    return float(v), float(s.value)


def set_checksums(offsets, leaf, pixel) -> np.ndarray:
    """The checksum of orc_score_batch_checksum for CSR voxel sets (numpy, vectorised)."""
    leaf = np.asarray(leaf).astype(np.uint64)
    pixel = np.asarray(pixel).astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (leaf << np.uint64(32)) | pixel
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        c = np.concatenate([[np.uint64(0)], np.cumsum(z, dtype=np.uint64)])
        off = np.asarray(offsets).astype(np.int64)
        return c[off[1:]] - c[off[:-1]]


def decode_normal(argb: int) -> np.ndarray:
    """computePoissonMeshNormalVector's pixel decode (viewpoint_offscreen_renderer.cpp:519-526)."""
    out = np.empty(3, dtype=np.float32)
    lib().orc_decode_normal(int(argb) & 0xFFFFFFFF, _f(out))
    return out


def mesh_face_normals(tri_vertices) -> np.ndarray:
    v = _f32(tri_vertices).reshape(-1, 9)
    out = np.empty_like(v)
    lib().orc_mesh_face_normals(_f(v), v.shape[0], _f(out))
    return out


def mesh_render_normals(tri_vertices, tri_normals, K, width, height, Rt, near=0.5, far=1e5,
                        clear_argb=0xFFFFFFFF) -> np.ndarray:
    """Brute-force restatement of drawPoissonMeshNormals for one pose: (height, width) uint32 ARGB."""
    v = _f32(tri_vertices).reshape(-1, 9)
    n = _f32(tri_normals).reshape(-1, 9)
    assert n.shape == v.shape
    out = np.empty((height, width), dtype=np.uint32)
    lib().orc_mesh_render_normals(_f(v), _f(n), v.shape[0], _f(_f32(K, (4,))), width, height, _f(_f32(Rt).reshape(12)),
                                  near, far, clear_argb, _u(out))
    return out


def mesh_cast_pixels(tri_vertices, tri_normals, K, width, height, Rt, xs, ys, near=0.5, far=1e5,
                     clear_argb=0xFFFFFFFF) -> np.ndarray:
    v = _f32(tri_vertices).reshape(-1, 9)
    n = _f32(tri_normals).reshape(-1, 9)
    xs = np.ascontiguousarray(xs, dtype=np.uint32)
    ys = np.ascontiguousarray(ys, dtype=np.uint32)
    out = np.empty(xs.shape[0], dtype=np.uint32)
    lib().orc_mesh_cast_pixels(_f(v), _f(n), v.shape[0], _f(_f32(K, (4,))), width, height, _f(_f32(Rt).reshape(12)),
                               near, far, clear_argb, _u(xs), _u(ys), xs.shape[0], _u(out))
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---- batched candidate generation (row N3): the reference's one-at-a-time loop, literally ----

def _exploration_step(p, o) -> np.float32:
    """ViewpointPlanner::computeExplorationStep (viewpoint_planner_graph.cpp:672-690), binary32 throughout"""
    f = np.float32
    step = f(o["exploration_step"])
    bb = np.asarray(o["exploration_bbox"], dtype=np.float32)
    p = np.asarray(p, dtype=np.float32)
    outside = bool(np.any(p < bb[:3]) or np.any(p > bb[3:]))
    if outside:
        dist_sq = f(0)
        for i in range(3):
            if p[i] < bb[i] or p[i] > bb[3 + i]:
                d = min(abs(f(p[i] - bb[i])), abs(f(p[i] - bb[3 + i])))
                dist_sq = f(dist_sq + f(d * d))
        dil = f(np.sqrt(dist_sq) - f(o["exploration_dilation_start_bbox_distance"]))
        if dil > 0:
            if o["exploration_dilation_squared"]:
                step = f(step + f(f(f(o["exploration_dilation_speed"]) * dil) * dil))
            else:
                step = f(step + f(f(o["exploration_dilation_speed"]) * dil))
    return min(step, f(o["exploration_step_max"]))


def _angular_distance(a, b) -> np.float32:
    """Eigen::QuaternionBase::angularDistance (3.3): 2 atan2(|vec(a conj(b))|, |w|); (w, x, y, z)"""
    f = np.float32
    a = [f(x) for x in a]
    bw, bx, by, bz = f(b[0]), f(-b[1]), f(-b[2]), f(-b[3])
    w = f(f(f(f(a[0] * bw) - f(a[1] * bx)) - f(a[2] * by)) - f(a[3] * bz))
    x = f(f(f(f(a[0] * bx) + f(a[1] * bw)) + f(a[2] * bz)) - f(a[3] * by))
    y = f(f(f(f(a[0] * by) + f(a[2] * bw)) + f(a[3] * bx)) - f(a[1] * bz))
    z = f(f(f(f(a[0] * bz) + f(a[3] * bw)) + f(a[1] * by)) - f(a[2] * bx))
    return f(f(2) * f(np.arctan2(f(np.sqrt(f(f(x * x) + f(f(y * y) + f(z * z))))), abs(w))))


def generate_viewpoint_entries_sequential(tree: "OracleTree", weights, normals, K, width, height, cand_pos, cand_quat,
                                          existing_pos, existing_quat, o: dict, opts: Optional[ScoreOpts] = None):
    """tryToAddViewpointEntry (viewpoint_planner_graph.cpp:542-655) candidate after candidate: validity, distance gate
    against the k nearest viewpoints accepted so far (exact neighbours, ascending squared distance), raycast + score
    of THIS candidate, thresholds, add.  o: dict with the fields of q3dr_candidate_opts.  Returns (status, results)
    where results[j] is the score_viewpoint dict of candidate j or None if it was never cast."""
    f = np.float32
    pos = [np.asarray(p, dtype=np.float32) for p in np.asarray(existing_pos, dtype=np.float32).reshape(-1, 3)]
    quat = [np.asarray(q, dtype=np.float32) for q in np.asarray(existing_quat, dtype=np.float32).reshape(-1, 4)]
    cand_pos = np.asarray(cand_pos, dtype=np.float32).reshape(-1, 3)
    cand_quat = np.asarray(cand_quat, dtype=np.float32).reshape(-1, 4)
    ang_thr = f(f(f(o["exploration_angular_dist_threshold_degrees"]) * f(np.pi)) / f(180.0))
    status, results = [], []
    for j in range(cand_pos.shape[0]):
        p, q = cand_pos[j], cand_quat[j]
        results.append(None)
        if not tree.valid_object_position(p, o["object_bbox"], o["bvh_bbox"], o["obstacle_free_height"]):
            status.append(1)
            continue
        step = _exploration_step(p, o)
        discard = False
        if pos:
            P = np.stack(pos)
            d = (P - p).astype(np.float32)
            dsq = ((d[:, 0] * d[:, 0]).astype(np.float32) + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
            dsq = (dsq + (d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
            knn = np.argsort(dsq, kind="stable")[: int(o["discard_dist_knn"])]
            for i in knn:
                step = min(step, _exploration_step(pos[i], o))
                thr = f(f(f(0.6) * step) * step)
                if dsq[i] < thr and _angular_distance(q, quat[i]) < ang_thr:
                    discard = True
                    break
        if discard:
            status.append(2)
            continue
        rt = pose_to_rt(q, p)
        r = tree.score_viewpoint(weights, normals, K, width, rt, height=height, opts=opts)
        results[-1] = r
        if len(r["leaf"]) < int(o["min_voxel_count"]):
            status.append(3)
            continue
        if f(r["total"]) < f(o["min_information"]):
            status.append(4)
            continue
        status.append(0)
        pos.append(p)
        quat.append(q)
    return np.array(status, dtype=np.uint8), results
