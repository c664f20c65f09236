"""ctypes plumbing over the C ABI (include/q3dr_raycast.h -> quad3dr_b200/libq3dr_raycast.so).

This module holds no arithmetic of the hot path: it marshals numpy buffers into the C entry points and
raises when the CUDA library is missing or reports an error.  There is deliberately no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libq3dr_raycast.so")

Q3DR_OK = 0
Q3DR_ERR_INVALID_ARG = 1
Q3DR_ERR_CUDA = 2
Q3DR_ERR_NO_DEVICE = 3
Q3DR_ERR_OUT_OF_MEMORY = 4
Q3DR_ERR_INTERNAL = 5


class Q3drError(RuntimeError):
    """Non-zero status of a C-ABI call (the shim's equivalent of OccupiedTreeType::Error)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"q3dr status {status}: {message}")
        self.status = status
        self.message = message


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


class _Result(C.Structure):
    _fields_ = [
        ("n_viewpoints", C.c_uint64),
        ("n_entries", C.c_uint64),
        ("offsets", C.c_void_p),
        ("leaf", C.c_void_p),
        ("information", C.c_void_p),
        ("first_pixel", C.c_void_p),
        ("dist_sq", C.c_void_p),
        ("total", C.c_void_p),
        ("hit_count", C.c_void_p),
    ]


class HostBlock(C.Structure):
    """q3dr_host_block: the (leaf, information) host arrays of a result after q3dr_result_detach_host"""

    _fields_ = [("leaf", C.c_void_p), ("information", C.c_void_p), ("n_entries", C.c_uint64), ("capacity", C.c_uint64)]

    def arrays(self):
        """numpy views (no copy) of the valid entries; they die with release()"""
        n = int(self.n_entries)
        if n == 0:
            return np.empty(0, np.int32), np.empty(0, np.float32)
        leaf = np.ctypeslib.as_array(C.cast(self.leaf, C.POINTER(C.c_int32)), shape=(n,))
        info = np.ctypeslib.as_array(C.cast(self.information, C.POINTER(C.c_float)), shape=(n,))
        return leaf, info

    def release(self):
        _check(load_library().q3dr_host_block_release(C.byref(self)))


class _OverlapResult(C.Structure):
    _fields_ = [("n_boxes", C.c_uint64), ("n_entries", C.c_uint64), ("offsets", C.c_void_p), ("leaf", C.c_void_p)]


class CandidateOpts(C.Structure):
    """q3dr_candidate_opts: the planner options tryToAddViewpointEntry reads"""

    _fields_ = [("object_bbox", C.c_float * 6), ("bvh_bbox", C.c_float * 6), ("obstacle_free_height", C.c_float),
                ("discard_dist_knn", C.c_uint32), ("exploration_step", C.c_float), ("exploration_step_max", C.c_float),
                ("exploration_dilation_squared", C.c_int32), ("exploration_dilation_start_bbox_distance", C.c_float),
                ("exploration_dilation_speed", C.c_float), ("exploration_bbox", C.c_float * 6),
                ("exploration_angular_dist_threshold_degrees", C.c_float), ("min_voxel_count", C.c_uint32),
                ("min_information", C.c_float)]


CANDIDATE_ACCEPTED, CANDIDATE_INVALID_POSITION, CANDIDATE_TOO_CLOSE, CANDIDATE_TOO_FEW_VOXELS, CANDIDATE_TOO_LITTLE_INFORMATION = range(5)


def default_candidate_opts(**kw) -> "CandidateOpts":
    o = CandidateOpts()
    load_library().q3dr_candidate_opts_default(C.byref(o))
    for k, v in kw.items():
        if isinstance(v, (list, tuple, np.ndarray)):
            v = (C.c_float * 6)(*[float(x) for x in v])
        setattr(o, k, v)
    return o


class OctreeInfo(C.Structure):
    """q3dr_octree_info"""

    _fields_ = [("id", C.c_char * 64), ("resolution", C.c_double), ("n_nodes", C.c_uint64), ("n_leaves", C.c_uint64),
                ("tree_depth", C.c_uint32), ("payload", C.c_uint32)]


OCTREE_AUTO, OCTREE_OCCUPANCY, OCTREE_AUGMENTED, OCTREE_LOGODDS = 0, 1, 2, 3


class ExchangeLayout(C.Structure):
    """q3dr_exchange_layout: byte offsets inside one receive buffer of the result exchange."""

    _fields_ = [(n, C.c_uint64) for n in ("header_off", "offsets_off", "total_off", "leaf_off", "info_off", "slot_bytes",
                                          "flags_off", "buffer_bytes")]


IPC_HANDLE_BYTES = 64
RESULT_FIRST_PIXEL = 1
RESULT_DIST_SQ = 2


class _MultiResult(C.Structure):
    _fields_ = [("n_parts", C.c_uint32), ("n_viewpoints", C.c_uint64), ("first_viewpoint", C.c_void_p), ("parts", C.c_void_p)]


# q3dr_chunk_callback(user, first_viewpoint, n_viewpoints, first_entry, n_entries, partial_device_result)
CHUNK_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_Result))


class RenderOpts(C.Structure):
    """q3dr_render_opts: clip planes and clear colour of the reference's off-screen renderer."""

    _fields_ = [("near_plane", C.c_float), ("far_plane", C.c_float), ("clear_argb", C.c_uint32)]


class MeshInfo(C.Structure):
    _fields_ = [
        ("n_triangles", C.c_uint64),
        ("n_nodes", C.c_uint64),
        ("tree_depth", C.c_uint32),
        ("device", C.c_int32),
        ("device_bytes", C.c_uint64),
        ("last_render_ms", C.c_float),
    ]


class MapInfo(C.Structure):
    _fields_ = [
        ("n_leaves", C.c_uint64),
        ("n_nodes", C.c_uint64),
        ("tree_depth", C.c_uint32),
        ("ref_depth", C.c_uint32),
        ("node_bytes", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("device", C.c_int32),
        ("sm_count", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("raycast_ms", C.c_float),
        ("score_ms", C.c_float),
        ("total_ms", C.c_float),
        ("chunks", C.c_uint32),
        ("viewpoints_per_chunk", C.c_uint32),
    ]


PIXEL_HIT_DTYPE = np.dtype(
    [
        ("intersection", np.float32, (3,)),
        ("leaf", np.int32),
        ("depth", np.uint32),
        ("dist_sq", np.float32),
        ("screen_x", np.float32),
        ("screen_y", np.float32),
    ]
)
assert PIXEL_HIT_DTYPE.itemsize == 32

_lib = None


def load_library() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the sm_100a extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C quad3dr_b200/csrc`). "
            "quad3dr_b200 has no CPU or PyTorch fallback."
        )
    L = C.CDLL(LIB_PATH)
    fp = C.POINTER(C.c_float)
    u32p = C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.q3dr_abi_version.restype = C.c_int
    L.q3dr_status_string.restype = C.c_char_p
    L.q3dr_status_string.argtypes = [C.c_int]
    L.q3dr_last_error.restype = C.c_char_p
    L.q3dr_device_count.argtypes = [C.POINTER(C.c_int)]
    L.q3dr_score_opts_default.argtypes = [C.POINTER(ScoreOpts)]
    L.q3dr_score_opts_default.restype = None
    L.q3dr_splitmedian_order.argtypes = [fp, C.c_size_t, u32p, u32p]
    L.q3dr_pose_to_rt.argtypes = [fp, fp, fp]
    L.q3dr_voxel_indices.argtypes = [C.c_size_t, u32p]
    L.q3dr_debug_flat_tree.argtypes = [fp, C.c_size_t, C.c_int, C.POINTER(C.c_uint64), u32p, u32p, u32p]
    L.q3dr_map_create.argtypes = [fp, fp, fp, u32p, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.q3dr_map_update_weights.argtypes = [vp, fp]
    L.q3dr_map_update_normals.argtypes = [vp, fp]
    L.q3dr_map_get_info.argtypes = [vp, C.POINTER(MapInfo)]
    L.q3dr_map_last_stats.argtypes = [vp, C.POINTER(Stats)]
    L.q3dr_map_reset.argtypes = [vp]
    L.q3dr_map_destroy.argtypes = [vp]
    L.q3dr_raycast_window.argtypes = [vp, fp, fp] + [C.c_uint32] * 4 + [C.c_float, C.c_float, vp]
    L.q3dr_intersect_rays.argtypes = [vp, fp, fp, C.c_size_t, C.c_float, C.c_float, vp]
    L.q3dr_raycast_unique.argtypes = (
        [vp, fp, fp] + [C.c_uint32] * 4 + [C.c_float, C.c_float, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    )
    L.q3dr_score_viewpoints.argtypes = (
        [vp, fp, C.c_uint32] + [C.c_uint32] * 4 + [fp, C.c_size_t, C.POINTER(ScoreOpts), C.POINTER(_Result)]
    )
    L.q3dr_score_viewpoints_device.argtypes = (
        [vp, fp, C.c_uint32] + [C.c_uint32] * 4 + [vp, C.c_size_t, C.POINTER(ScoreOpts), vp, C.POINTER(_Result)]
    )
    L.q3dr_result_to_host.argtypes = [vp, C.POINTER(_Result)]
    L.q3dr_map_set_chunk_callback.argtypes = [vp, CHUNK_CALLBACK, vp]
    L.q3dr_map_set_result_fields.argtypes = [vp, C.c_uint32]
    L.q3dr_result_detach_host.argtypes = [vp, C.POINTER(HostBlock)]
    L.q3dr_host_block_release.argtypes = [C.POINTER(HostBlock)]
    L.q3dr_map_check_guards.argtypes = [vp, C.POINTER(C.c_int)]
    L.q3dr_candidate_opts_default.argtypes = [C.POINTER(CandidateOpts)]
    L.q3dr_candidate_opts_default.restype = None
    L.q3dr_generate_viewpoint_entries.argtypes = ([vp, fp, C.c_uint32] + [C.c_uint32] * 4 + [fp, fp, C.c_size_t, fp, fp, C.c_size_t,
                                                  C.POINTER(ScoreOpts), C.POINTER(CandidateOpts), C.POINTER(C.c_uint8),
                                                  C.POINTER(_Result)])
    L.q3dr_octree_read.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(vp)]
    L.q3dr_octree_get_info.argtypes = [vp, C.POINTER(OctreeInfo)]
    L.q3dr_octree_leaves.argtypes = [vp, fp, fp, fp, u32p, fp, u32p]
    L.q3dr_octree_destroy.argtypes = [vp]
    L.q3dr_map_downweight_real_viewpoints.argtypes = [vp, fp, C.c_uint32, C.c_uint32, fp, fp, C.c_size_t, C.POINTER(ScoreOpts),
                                                      C.c_float, vp, C.POINTER(RenderOpts), u32p, fp, C.POINTER(C.c_uint64)]
    L.q3dr_exchange_layout_of.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(ExchangeLayout)]
    L.q3dr_exchange_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.q3dr_exchange_destroy.argtypes = [vp]
    L.q3dr_exchange_disconnect.argtypes = [vp]
    L.q3dr_exchange_export.argtypes = [vp, vp]
    L.q3dr_exchange_import.argtypes = [vp, vp]
    L.q3dr_exchange_connect.argtypes = [C.POINTER(vp), C.c_uint32]
    L.q3dr_exchange_begin_step.argtypes = [vp]
    L.q3dr_exchange_end_step.argtypes = [vp, C.c_int, vp]
    L.q3dr_exchange_view.argtypes = [vp, C.c_uint32, C.POINTER(_Result)]
    L.q3dr_multi_enable_gather.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.q3dr_multi_gathered.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(_Result)]
    L.q3dr_overlap_boxes.argtypes = [vp, fp, C.c_size_t, C.POINTER(_OverlapResult)]
    L.q3dr_valid_object_positions.argtypes = [vp, fp, C.c_size_t, fp, fp, C.c_float, C.POINTER(C.c_uint8)]
    L.q3dr_octree_leaves_to_boxes.argtypes = [fp, fp, fp, u32p, C.c_size_t, fp, C.c_float, C.c_float, C.c_uint32, C.c_int,
                                              fp, u32p, C.POINTER(C.c_size_t)]
    L.q3dr_map_update_weights_from_grid.argtypes = [vp, fp, C.c_float, C.POINTER(C.c_int32), fp, C.c_float, u32p, fp]
    L.q3dr_multi_create.argtypes = [fp, fp, fp, u32p, C.c_size_t, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.q3dr_multi_destroy.argtypes = [vp]
    L.q3dr_multi_device_count.argtypes = [vp]
    L.q3dr_multi_map.argtypes = [vp, C.c_int]
    L.q3dr_multi_map.restype = vp
    L.q3dr_multi_update_weights.argtypes = [vp, fp]
    L.q3dr_multi_set_normal_mesh.argtypes = [vp, fp, fp, C.c_size_t, C.c_uint32, C.POINTER(RenderOpts)]
    L.q3dr_multi_score_viewpoints.argtypes = (
        [vp, fp, C.c_uint32] + [C.c_uint32] * 4 + [fp, C.c_size_t, C.POINTER(ScoreOpts), C.POINTER(_MultiResult)]
    )
    L.q3dr_observed_create.argtypes = [vp, C.POINTER(vp)]
    L.q3dr_observed_clear.argtypes = [vp]
    L.q3dr_observed_destroy.argtypes = [vp]
    L.q3dr_observed_download.argtypes = [vp, fp]
    L.q3dr_new_information.argtypes = [vp, vp, C.c_float, u32p, C.c_size_t, fp]
    L.q3dr_observed_add_viewpoint.argtypes = [vp, vp, C.c_float, C.c_uint32, fp]
    L.q3dr_overlap_information.argtypes = [vp, C.c_uint32, u32p, C.c_size_t, fp]
    L.q3dr_observed_voxel_set.argtypes = [vp, vp, C.c_float, C.POINTER(C.c_int32), fp, C.c_size_t, C.c_int, fp]
    L.q3dr_render_opts_default.argtypes = [C.POINTER(RenderOpts)]
    L.q3dr_render_opts_default.restype = None
    L.q3dr_mesh_create.argtypes = [fp, fp, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.q3dr_mesh_destroy.argtypes = [vp]
    L.q3dr_mesh_get_info.argtypes = [vp, C.POINTER(MeshInfo)]
    L.q3dr_mesh_render_normals.argtypes = [vp, fp, C.c_uint32, C.c_uint32, fp, C.c_size_t, C.POINTER(RenderOpts), u32p]
    L.q3dr_decode_normal.argtypes = [C.c_uint32, fp]
    L.q3dr_score_viewpoints_mesh_normals.argtypes = (
        [vp, vp, C.POINTER(RenderOpts), fp] + [C.c_uint32] * 6 + [fp, C.c_size_t, C.POINTER(ScoreOpts), C.POINTER(_Result)]
    )
    L.q3dr_score_viewpoints_mesh_normals_device.argtypes = (
        [vp, vp, C.POINTER(RenderOpts), fp] + [C.c_uint32] * 6 + [vp, C.c_size_t, C.POINTER(ScoreOpts), vp, C.POINTER(_Result)]
    )
    L.q3dr_score_viewpoints_normal_images.argtypes = (
        [vp, u32p, fp] + [C.c_uint32] * 6 + [fp, C.c_size_t, C.POINTER(ScoreOpts), C.POINTER(_Result)]
    )
    _lib = L
    return L


def _check(status: int):
    if status != Q3DR_OK:
        L = load_library()
        msg = L.q3dr_last_error().decode() or L.q3dr_status_string(status).decode()
        raise Q3drError(status, msg)


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_opts(**kw) -> ScoreOpts:
    o = ScoreOpts()
    load_library().q3dr_score_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def device_count() -> int:
    n = C.c_int(0)
    st = load_library().q3dr_device_count(C.byref(n))
    return int(n.value) if st == Q3DR_OK else 0


def splitmedian_order(boxes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(order, depth): order[k] = input index of the k-th leaf of the reference's Tree::build."""
    b = _f32(boxes)
    n = b.shape[0]
    order = np.empty(n, dtype=np.uint32)
    depth = np.empty(n, dtype=np.uint32)
    _check(
        load_library().q3dr_splitmedian_order(
            _fp(b), n, order.ctypes.data_as(C.POINTER(C.c_uint32)), depth.ctypes.data_as(C.POINTER(C.c_uint32))
        )
    )
    return order, depth


def octree_leaves_to_boxes(center, size, occupancy, observation_count, bvh_bbox, obstacle_free_height: float,
                           occupancy_threshold: float = 0.51, observation_threshold: int = 1, device: int = 0):
    """generateBVHTree's leaf rule on the device: (boxes (k, 6), index of the source octree leaf (k,))."""
    center = _f32(center).reshape(-1, 3)
    n = center.shape[0]
    size = _f32(size, (n,))
    occupancy = _f32(occupancy, (n,))
    cnt = np.ascontiguousarray(observation_count, dtype=np.uint32)
    boxes = np.empty((n, 6), dtype=np.float32)
    src = np.empty(n, dtype=np.uint32)
    k = C.c_size_t(0)
    _check(load_library().q3dr_octree_leaves_to_boxes(
        _fp(center), _fp(size), _fp(occupancy), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), n, _fp(_f32(bvh_bbox, (6,))),
        float(obstacle_free_height), float(occupancy_threshold), int(observation_threshold), int(device),
        _fp(boxes), src.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(k)))
    return boxes[:k.value].copy(), src[:k.value].copy()


def read_octree_leaves(path: str, payload: int = OCTREE_AUTO) -> dict:
    """OctoMap ".ot" stream -> its leaves in begin_tree() order (host only): dict with info (OctreeInfo), center (n, 3),
    size, occupancy, observation_count, weight, depth -- the inputs of octree_leaves_to_boxes."""
    L = load_library()
    h = C.c_void_p()
    _check(L.q3dr_octree_read(os.fsencode(path), payload, C.byref(h)))
    try:
        info = OctreeInfo()
        _check(L.q3dr_octree_get_info(h, C.byref(info)))
        n = int(info.n_leaves)
        out = dict(info=info, center=np.empty((n, 3), np.float32), size=np.empty(n, np.float32),
                   occupancy=np.empty(n, np.float32), observation_count=np.empty(n, np.uint32),
                   weight=np.empty(n, np.float32), depth=np.empty(n, np.uint32))
        u = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        _check(L.q3dr_octree_leaves(h, _fp(out["center"]), _fp(out["size"]), _fp(out["occupancy"]),
                                    u(out["observation_count"]), _fp(out["weight"]), u(out["depth"])))
        return out
    finally:
        L.q3dr_octree_destroy(h)


def voxel_indices(n: int) -> np.ndarray:
    """Voxel index (position in the reference's all-node iterator) of every leaf, left to right."""
    out = np.empty(n, dtype=np.uint32)
    _check(load_library().q3dr_voxel_indices(n, out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


def pose_to_rt(quat_wxyz, trans) -> np.ndarray:
    out = np.empty(12, dtype=np.float32)
    _check(load_library().q3dr_pose_to_rt(_fp(_f32(quat_wxyz, (4,))), _fp(_f32(trans, (3,))), _fp(out)))
    return out


def debug_flat_tree(boxes: np.ndarray, top_levels: int = 4) -> dict:
    b = _f32(boxes)
    nn = C.c_uint64(0)
    d = C.c_uint32(0)
    ms = C.c_uint32(0)
    tn = C.c_uint32(0)
    _check(load_library().q3dr_debug_flat_tree(_fp(b), b.shape[0], top_levels, C.byref(nn), C.byref(d), C.byref(ms), C.byref(tn)))
    return dict(n_nodes=int(nn.value), depth=int(d.value), max_stack=int(ms.value), top_nodes=int(tn.value))


class ScoreResult:
    """CSR result of a batched scoring call, copied out of the map-owned host buffers."""

    def __init__(self, r: _Result, copy: bool = True):
        nv, ne = int(r.n_viewpoints), int(r.n_entries)

        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.empty(0, dtype=dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt)
            return a.copy() if copy else a

        self.n_viewpoints = nv
        self.n_entries = ne
        self.offsets = arr(r.offsets, nv + 1, np.uint64)
        self.leaf = arr(r.leaf, ne, np.int32)
        self.information = arr(r.information, ne, np.float32)
        self.first_pixel = arr(r.first_pixel, ne, np.uint32)
        self.dist_sq = arr(r.dist_sq, ne, np.float32)
        self.total = arr(r.total, nv, np.float32)
        self.hit_count = arr(r.hit_count, nv, np.uint32)

    def viewpoint(self, v: int) -> dict:
        a, b = int(self.offsets[v]), int(self.offsets[v + 1])
        return dict(
            leaf=self.leaf[a:b], info=self.information[a:b], pixel=self.first_pixel[a:b], dsq=self.dist_sq[a:b],
            total=float(self.total[v]), hit_count=int(self.hit_count[v]),
        )


class DeviceResult:
    """Device pointers (ints) of a device-resident scoring result, owned by the map."""

    def __init__(self, r: _Result):
        self.n_viewpoints = int(r.n_viewpoints)
        self.n_entries = int(r.n_entries)
        self.offsets = r.offsets
        self.leaf = r.leaf
        self.information = r.information
        self.first_pixel = r.first_pixel
        self.dist_sq = r.dist_sq
        self.total = r.total
        self.hit_count = r.hit_count


class OccupancyMap:
    """Device-resident flattened occupancy map (q3dr_map).  Leaves must be in tie-break order
    (see splitmedian_order); leaf k is reported as index k."""

    def __init__(self, boxes, weights=None, normals=None, depth=None, device: int = 0):
        L = load_library()
        b = _f32(boxes)
        if b.ndim != 2 or b.shape[1] != 6:
            raise ValueError("boxes must be (n, 6)")
        n = b.shape[0]
        w = _f32(weights, (n,)) if weights is not None else None
        nr = _f32(normals, (n, 3)) if normals is not None else None
        dp = np.ascontiguousarray(depth, dtype=np.uint32) if depth is not None else None
        h = C.c_void_p(None)
        _check(
            L.q3dr_map_create(
                _fp(b), _fp(w) if w is not None else None, _fp(nr) if nr is not None else None,
                dp.ctypes.data_as(C.POINTER(C.c_uint32)) if dp is not None else None, n, device, C.byref(h),
            )
        )
        self._h = h
        self.n = n

    def close(self):
        if getattr(self, "_h", None):
            load_library().q3dr_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> MapInfo:
        i = MapInfo()
        _check(load_library().q3dr_map_get_info(self._h, C.byref(i)))
        return i

    def stats(self) -> Stats:
        s = Stats()
        _check(load_library().q3dr_map_last_stats(self._h, C.byref(s)))
        return s

    def update_weights(self, weights):
        _check(load_library().q3dr_map_update_weights(self._h, _fp(_f32(weights, (self.n,)))))

    def update_normals(self, normals):
        _check(load_library().q3dr_map_update_normals(self._h, _fp(_f32(normals, (self.n, 3)))))

    def reset(self):
        _check(load_library().q3dr_map_reset(self._h))

    def raycast_window(self, K, Rt, x0, x1, y0, y1, min_range=0.0, max_range=-1.0) -> np.ndarray:
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(12)
        out = np.empty((y1 - y0) * (x1 - x0), dtype=PIXEL_HIT_DTYPE)
        _check(
            load_library().q3dr_raycast_window(
                self._h, _fp(K), _fp(Rt), x0, x1, y0, y1, min_range, max_range, out.ctypes.data_as(C.c_void_p)
            )
        )
        return out

    def intersect_rays(self, origins, directions, min_range=0.0, max_range=-1.0) -> np.ndarray:
        o = _f32(origins)
        d = _f32(directions)
        if o.shape != d.shape or o.ndim != 2 or o.shape[1] != 3:
            raise ValueError("origins/directions must both be (n, 3)")
        out = np.empty(o.shape[0], dtype=PIXEL_HIT_DTYPE)
        _check(
            load_library().q3dr_intersect_rays(
                self._h, _fp(o), _fp(d), o.shape[0], min_range, max_range, out.ctypes.data_as(C.c_void_p)
            )
        )
        return out

    def raycast_unique(self, K, Rt, x0, x1, y0, y1, min_range=0.0, max_range=-1.0) -> np.ndarray:
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(12)
        cap = min((y1 - y0) * (x1 - x0), self.n)
        out = np.empty(cap, dtype=PIXEL_HIT_DTYPE)
        cnt = C.c_size_t(0)
        _check(
            load_library().q3dr_raycast_unique(
                self._h, _fp(K), _fp(Rt), x0, x1, y0, y1, min_range, max_range, out.ctypes.data_as(C.c_void_p), cap,
                C.byref(cnt),
            )
        )
        return out[: cnt.value].copy()

    def score_viewpoints(self, K, image_width: int, window: Sequence[int], Rt, opts: Optional[ScoreOpts] = None,
                         copy: bool = True) -> ScoreResult:
        """Host poses in, host CSR result out (H2D + kernels + D2H inside the call)."""
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(-1, 12)
        opts = opts or default_opts()
        x0, x1, y0, y1 = window
        r = _Result()
        _check(
            load_library().q3dr_score_viewpoints(
                self._h, _fp(K), image_width, x0, x1, y0, y1, _fp(Rt), Rt.shape[0], C.byref(opts), C.byref(r)
            )
        )
        return ScoreResult(r, copy=copy)

    def score_viewpoints_device(self, K, image_width: int, window: Sequence[int], d_rt_ptr: int, n: int,
                                opts: Optional[ScoreOpts] = None, stream: int = 0) -> DeviceResult:
        """Poses already in HBM (raw device pointer), results stay in HBM."""
        K = _f32(K, (4,))
        opts = opts or default_opts()
        x0, x1, y0, y1 = window
        r = _Result()
        _check(
            load_library().q3dr_score_viewpoints_device(
                self._h, _fp(K), image_width, x0, x1, y0, y1, C.c_void_p(d_rt_ptr), n, C.byref(opts),
                C.c_void_p(stream), C.byref(r),
            )
        )
        return DeviceResult(r)

    def score_viewpoints_mesh_normals(self, mesh: "NormalMesh", K, image_width: int, image_height: int,
                                      window: Sequence[int], Rt, opts: Optional[ScoreOpts] = None,
                                      ropts: Optional["RenderOpts"] = None, copy: bool = True) -> ScoreResult:
        """enable_opengl = true: normals from the mesh at every voxel's kept pixel (host buffers in and out)."""
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(-1, 12)
        opts = opts or default_opts()
        ropts = ropts or default_render_opts()
        x0, x1, y0, y1 = window
        r = _Result()
        _check(
            load_library().q3dr_score_viewpoints_mesh_normals(
                self._h, mesh._h, C.byref(ropts), _fp(K), image_width, image_height, x0, x1, y0, y1, _fp(Rt),
                Rt.shape[0], C.byref(opts), C.byref(r),
            )
        )
        return ScoreResult(r, copy=copy)

    def score_viewpoints_mesh_normals_device(self, mesh: "NormalMesh", K, image_width: int, image_height: int,
                                             window: Sequence[int], d_rt_ptr: int, n: int,
                                             opts: Optional[ScoreOpts] = None, ropts: Optional["RenderOpts"] = None,
                                             stream: int = 0) -> DeviceResult:
        K = _f32(K, (4,))
        opts = opts or default_opts()
        ropts = ropts or default_render_opts()
        x0, x1, y0, y1 = window
        r = _Result()
        _check(
            load_library().q3dr_score_viewpoints_mesh_normals_device(
                self._h, mesh._h, C.byref(ropts), _fp(K), image_width, image_height, x0, x1, y0, y1,
                C.c_void_p(d_rt_ptr), n, C.byref(opts), C.c_void_p(stream), C.byref(r),
            )
        )
        return DeviceResult(r)

    def score_viewpoints_normal_images(self, images, K, image_width: int, window: Sequence[int], Rt,
                                       opts: Optional[ScoreOpts] = None, copy: bool = True) -> ScoreResult:
        """enable_opengl = true with the caller's normal images: (n, H, W) uint32 ARGB32."""
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(-1, 12)
        img = np.ascontiguousarray(images, dtype=np.uint32)
        if img.ndim != 3 or img.shape[0] != Rt.shape[0] or img.shape[2] != image_width:
            raise ValueError("images must be (n, image_height, image_width)")
        opts = opts or default_opts()
        x0, x1, y0, y1 = window
        r = _Result()
        _check(
            load_library().q3dr_score_viewpoints_normal_images(
                self._h, img.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(K), image_width, img.shape[1], x0, x1, y0, y1,
                _fp(Rt), Rt.shape[0], C.byref(opts), C.byref(r),
            )
        )
        return ScoreResult(r, copy=copy)

    def check_guards(self) -> int:
        """Q3DR_GUARD=1 only: number of device buffers with an overwritten guard (a write past the end)."""
        n = C.c_int(0)
        _check(load_library().q3dr_map_check_guards(self._h, C.byref(n)))
        return int(n.value)

    def set_result_fields(self, first_pixel: bool = True, dist_sq: bool = True):
        """Which per-entry arrays scoring calls deliver besides (leaf, information); see q3dr_map_set_result_fields."""
        f = (RESULT_FIRST_PIXEL if first_pixel else 0) | (RESULT_DIST_SQ if dist_sq else 0)
        _check(load_library().q3dr_map_set_result_fields(self._h, f))

    def set_chunk_callback(self, fn=None):
        """fn(first_viewpoint, n_viewpoints, first_entry, n_entries, DeviceResult) is called on the host as soon as the
        kernels finishing a chunk have been issued on the call's stream (q3dr_map_set_chunk_callback); None removes it."""
        if fn is None:
            self._chunk_cb = None
            _check(load_library().q3dr_map_set_chunk_callback(self._h, C.cast(None, CHUNK_CALLBACK), None))
            return

        def tramp(_user, first_vp, n_vp, first_entry, n_entries, partial):
            fn(int(first_vp), int(n_vp), int(first_entry), int(n_entries), DeviceResult(partial.contents))

        self._chunk_cb = CHUNK_CALLBACK(tramp)  # keep the thunk alive as long as it is installed
        _check(load_library().q3dr_map_set_chunk_callback(self._h, self._chunk_cb, None))

    def update_weights_from_grid(self, grid_origin, grid_increment: float, grid_dim, grid_weight, lam: float,
                                 observation_count) -> np.ndarray:
        """ViewpointPlannerData::updateWeights' writes into the BVH leaves; returns the resulting weights."""
        n = int(self.info().n_leaves)
        dim = np.ascontiguousarray(grid_dim, dtype=np.int32)
        gw = _f32(grid_weight).reshape(-1)
        assert gw.size == int(dim[0]) * int(dim[1]) * int(dim[2])
        cnt = np.ascontiguousarray(observation_count, dtype=np.uint32)
        assert cnt.shape == (n,)
        out = np.empty(n, dtype=np.float32)
        _check(load_library().q3dr_map_update_weights_from_grid(
            self._h, _fp(_f32(grid_origin, (3,))), float(grid_increment), dim.ctypes.data_as(C.POINTER(C.c_int32)), _fp(gw),
            float(lam), cnt.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(out)))
        return out

    def downweight_real_viewpoints(self, K, width: int, height: int, Rt, depth_maps, factor: float,
                                   opts: Optional[ScoreOpts] = None, mesh: Optional["NormalMesh"] = None,
                                   ropts: Optional["RenderOpts"] = None, normal_images=None):
        """ViewpointPlannerData::updateWeightsWithRealViewpoints for images sharing one depth camera: the map's weights
        are lowered in place; returns (weights, number of updates)."""
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(-1, 12)
        n = Rt.shape[0]
        d = _f32(depth_maps, (n, height, width))
        opts = opts or default_opts()
        img = None if normal_images is None else np.ascontiguousarray(normal_images, dtype=np.uint32).reshape(n, height, width)
        if mesh is not None and ropts is None:
            ropts = default_render_opts()
        out = np.empty(self.n, dtype=np.float32)
        cnt = C.c_uint64(0)
        _check(load_library().q3dr_map_downweight_real_viewpoints(
            self._h, _fp(K), width, height, _fp(Rt), _fp(d), n, C.byref(opts), float(factor),
            mesh._h if mesh is not None else None, C.byref(ropts) if ropts is not None else None,
            img.ctypes.data_as(C.POINTER(C.c_uint32)) if img is not None else None, _fp(out), C.byref(cnt)))
        return out, int(cnt.value)

    def generate_viewpoint_entries(self, K, image_width: int, window: Sequence[int], cand_pos, cand_quat, existing_pos,
                                   existing_quat, copts: "CandidateOpts", opts: Optional[ScoreOpts] = None, copy: bool = True):
        """tryToAddViewpointEntry for a batch of candidate poses in order: (status uint8 (n,), ScoreResult over all n);
        copy = False: the ScoreResult's arrays are views of the map's host buffers (valid until the next call)."""
        K = _f32(K, (4,))
        cp = _f32(cand_pos).reshape(-1, 3)
        cq = _f32(cand_quat).reshape(-1, 4)
        ep = _f32(existing_pos).reshape(-1, 3)
        eq = _f32(existing_quat).reshape(-1, 4)
        opts = opts or default_opts()
        x0, x1, y0, y1 = window
        status = np.empty(cp.shape[0], dtype=np.uint8)
        r = _Result()
        _check(load_library().q3dr_generate_viewpoint_entries(
            self._h, _fp(K), image_width, x0, x1, y0, y1, _fp(cp), _fp(cq), cp.shape[0], _fp(ep), _fp(eq), ep.shape[0],
            C.byref(opts), C.byref(copts), status.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(r)))
        return status, ScoreResult(r, copy=copy)

    def overlap_boxes(self, boxes) -> Tuple[np.ndarray, np.ndarray]:
        """bvh::Tree::intersects(bbox) for many boxes: CSR (offsets uint64 (n+1), leaf int32) of the overlapping
        leaves, ascending leaf index per box."""
        b = _f32(boxes).reshape(-1, 6)
        r = _OverlapResult()
        _check(load_library().q3dr_overlap_boxes(self._h, _fp(b), b.shape[0], C.byref(r)))
        n, ne = int(r.n_boxes), int(r.n_entries)
        off = np.ctypeslib.as_array(C.cast(r.offsets, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        leaf = (np.ctypeslib.as_array(C.cast(r.leaf, C.POINTER(C.c_int32)), shape=(ne,)).copy()
                if ne else np.empty(0, np.int32))
        return off, leaf

    def valid_object_positions(self, positions, object_bbox, bvh_bbox, obstacle_free_height: float) -> np.ndarray:
        """ViewpointPlannerData::isValidObjectPosition for many positions (no-fly zones excluded)."""
        p = _f32(positions).reshape(-1, 3)
        out = np.empty(p.shape[0], dtype=np.uint8)
        _check(load_library().q3dr_valid_object_positions(
            self._h, _fp(p), p.shape[0], _fp(_f32(object_bbox, (6,))), _fp(_f32(bvh_bbox, (6,))),
            float(obstacle_free_height), out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out.astype(bool)

    def result_to_host(self, copy: bool = True) -> ScoreResult:
        r = _Result()
        _check(load_library().q3dr_result_to_host(self._h, C.byref(r)))
        return ScoreResult(r, copy=copy)

    def detach_host_result(self) -> HostBlock:
        """Takes the (leaf, information) host arrays of the last result away from the map (q3dr_result_detach_host):
        they stay valid until HostBlock.release(), whatever the map does next."""
        hb = HostBlock()
        _check(load_library().q3dr_result_detach_host(self._h, C.byref(hb)))
        return hb


def exchange_layout(n_ranks: int, max_viewpoints: int, max_entries: int) -> ExchangeLayout:
    lay = ExchangeLayout()
    _check(load_library().q3dr_exchange_layout_of(n_ranks, max_viewpoints, max_entries, C.byref(lay)))
    return lay


class Exchange:
    """q3dr_exchange: this rank's end of the result exchange between GPUs (copy-engine pushes into peer HBM per finished
    chunk, completion flags instead of a barrier).  One per map.  Separate processes: export_handles() on every rank,
    ship all handles to all ranks (any transport), import_handles(); threads of one process: Exchange.connect(all)."""

    def __init__(self, gmap: "OccupancyMap", n_ranks: int, rank: int, max_viewpoints: int, max_entries: int):
        h = C.c_void_p()
        _check(load_library().q3dr_exchange_create(gmap._h, n_ranks, rank, max_viewpoints, max_entries, C.byref(h)))
        self._h = h
        self.map = gmap
        self.n_ranks = n_ranks
        self.rank = rank

    def close(self):
        if getattr(self, "_h", None):
            load_library().q3dr_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def disconnect(self):
        if getattr(self, "_h", None):
            _check(load_library().q3dr_exchange_disconnect(self._h))

    def export_handles(self) -> bytes:
        buf = C.create_string_buffer(2 * IPC_HANDLE_BYTES)
        _check(load_library().q3dr_exchange_export(self._h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def import_handles(self, all_handles: Sequence[bytes]):
        blob = b"".join(all_handles)
        if len(blob) != self.n_ranks * 2 * IPC_HANDLE_BYTES:
            raise ValueError("need 2 handles of 64 bytes from every rank, in rank order")
        buf = C.create_string_buffer(blob, len(blob))
        _check(load_library().q3dr_exchange_import(self._h, C.cast(buf, C.c_void_p)))

    @staticmethod
    def connect(exchanges: Sequence["Exchange"]):
        arr = (C.c_void_p * len(exchanges))(*[e._h for e in exchanges])
        _check(load_library().q3dr_exchange_connect(arr, len(exchanges)))

    def begin_step(self):
        _check(load_library().q3dr_exchange_begin_step(self._h))

    def end_step(self, local_status: int = 0, stream: int = 0):
        _check(load_library().q3dr_exchange_end_step(self._h, int(local_status), C.c_void_p(stream)))

    def view(self, src: int) -> DeviceResult:
        r = _Result()
        _check(load_library().q3dr_exchange_view(self._h, src, C.byref(r)))
        return DeviceResult(r)


def default_render_opts(**kw) -> RenderOpts:
    o = RenderOpts()
    load_library().q3dr_render_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def decode_normal(argb: int) -> np.ndarray:
    out = np.empty(3, dtype=np.float32)
    _check(load_library().q3dr_decode_normal(int(argb) & 0xFFFFFFFF, _fp(out)))
    return out


class NormalMesh:
    """Device-resident triangle mesh of the image-space normals (q3dr_mesh): the Poisson mesh the reference
    uploads to its off-screen renderer.  tri_vertices / tri_normals: (nf, 3, 3) corners."""

    def __init__(self, tri_vertices, tri_normals=None, device: int = 0):
        v = _f32(tri_vertices).reshape(-1, 9)
        nr = _f32(tri_normals).reshape(-1, 9) if tri_normals is not None else None
        if nr is not None and nr.shape != v.shape:
            raise ValueError("tri_normals must match tri_vertices")
        h = C.c_void_p(None)
        _check(load_library().q3dr_mesh_create(_fp(v), _fp(nr) if nr is not None else None, v.shape[0], device, C.byref(h)))
        self._h = h
        self.n = v.shape[0]

    def close(self):
        if getattr(self, "_h", None):
            load_library().q3dr_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> MeshInfo:
        i = MeshInfo()
        _check(load_library().q3dr_mesh_get_info(self._h, C.byref(i)))
        return i

    def render_normals(self, K, width: int, height: int, Rt, ropts: Optional[RenderOpts] = None) -> np.ndarray:
        """drawPoissonMeshNormals for n poses: (n, height, width) uint32 ARGB32."""
        K = _f32(K, (4,))
        Rt = _f32(Rt).reshape(-1, 12)
        ropts = ropts or default_render_opts()
        out = np.empty((Rt.shape[0], height, width), dtype=np.uint32)
        _check(
            load_library().q3dr_mesh_render_normals(
                self._h, _fp(K), width, height, _fp(Rt), Rt.shape[0], C.byref(ropts),
                out.ctypes.data_as(C.POINTER(C.c_uint32)),
            )
        )
        return out


class ObservedVoxels:
    """One ViewpointPath::observed_voxel_map on the device (q3dr_observed): accumulated information per voxel.

    All methods refer to the viewpoints of the LAST scoring call on `gmap` (its result stays resident in HBM)."""

    def __init__(self, gmap: "OccupancyMap"):
        self.map = gmap
        h = C.c_void_p()
        _check(load_library().q3dr_observed_create(gmap._h, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            load_library().q3dr_observed_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        _check(load_library().q3dr_observed_clear(self._h))

    def download(self) -> np.ndarray:
        out = np.empty(int(self.map.info().n_leaves), dtype=np.float32)
        _check(load_library().q3dr_observed_download(self._h, _fp(out)))
        return out

    def new_information(self, factor: float, viewpoints=None, n: Optional[int] = None) -> np.ndarray:
        """computeNewInformation for the given viewpoints (None: the first n of the result)."""
        if viewpoints is None:
            assert n is not None
            idx = None
        else:
            idx = np.ascontiguousarray(viewpoints, dtype=np.uint32)
            n = len(idx)
        out = np.empty(n, dtype=np.float32)
        _check(load_library().q3dr_new_information(
            self.map._h, self._h, float(factor), None if idx is None else idx.ctypes.data_as(C.POINTER(C.c_uint32)), n, _fp(out)))
        return out

    def voxel_set(self, factor: float, leaf, information, add: bool) -> float:
        """computeNewInformation (add=False) / addObservedVoxelsToViewpointPath (add=True) for an explicit voxel set."""
        leaf = np.ascontiguousarray(leaf, dtype=np.int32)
        info = _f32(information, (len(leaf),))
        v = C.c_float(0)
        _check(load_library().q3dr_observed_voxel_set(self.map._h, self._h, float(factor), leaf.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      _fp(info), len(leaf), int(bool(add)), C.byref(v)))
        return float(v.value)

    def add_viewpoint(self, factor: float, viewpoint: int) -> float:
        """addObservedVoxelsToViewpointPath; returns the novel information."""
        v = C.c_float(0)
        _check(load_library().q3dr_observed_add_viewpoint(self.map._h, self._h, float(factor), int(viewpoint), C.byref(v)))
        return float(v.value)


def overlap_information(gmap: "OccupancyMap", ref: int, others) -> np.ndarray:
    """compute_overlap_information_lambda of each viewpoint in `others` (set 1) against `ref` (set 2)."""
    idx = np.ascontiguousarray(others, dtype=np.uint32)
    out = np.empty(len(idx), dtype=np.float32)
    _check(load_library().q3dr_overlap_information(gmap._h, int(ref), idx.ctypes.data_as(C.POINTER(C.c_uint32)), len(idx), _fp(out)))
    return out


class MultiGpuMap:
    """q3dr_multi: the map replicated on several GPUs of this process, viewpoints sharded, one host thread per GPU."""

    def __init__(self, boxes, weights=None, normals=None, depth=None, devices=None):
        boxes = _f32(boxes)
        n = boxes.shape[0]
        w = None if weights is None else _f32(weights, (n,))
        nr = None if normals is None else _f32(normals, (n, 3))
        d = None if depth is None else np.ascontiguousarray(depth, dtype=np.uint32)
        devs = None if devices is None else (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(load_library().q3dr_multi_create(
            _fp(boxes), None if w is None else _fp(w), None if nr is None else _fp(nr),
            None if d is None else d.ctypes.data_as(C.POINTER(C.c_uint32)), n, devs, 0 if devices is None else len(devices),
            C.byref(h)))
        self._h = h
        self.n_devices = int(load_library().q3dr_multi_device_count(h))

    def close(self):
        if getattr(self, "_h", None):
            load_library().q3dr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_normal_mesh(self, tri_vertices, tri_normals, image_height: int, ropts: Optional[RenderOpts] = None):
        """image-space normals for the following score_viewpoints calls (tri_vertices None: back to voxel normals)"""
        if tri_vertices is None:
            _check(load_library().q3dr_multi_set_normal_mesh(self._h, None, None, 0, 0, None))
            return
        v = _f32(tri_vertices).reshape(-1, 9)
        nr = _f32(tri_normals).reshape(-1, 9) if tri_normals is not None else None
        ropts = ropts or default_render_opts()
        _check(load_library().q3dr_multi_set_normal_mesh(self._h, _fp(v), _fp(nr) if nr is not None else None, v.shape[0],
                                                         image_height, C.byref(ropts)))

    def update_weights(self, weights):
        _check(load_library().q3dr_multi_update_weights(self._h, _fp(_f32(weights))))

    def set_result_fields(self, first_pixel: bool = True, dist_sq: bool = True):
        f = (RESULT_FIRST_PIXEL if first_pixel else 0) | (RESULT_DIST_SQ if dist_sq else 0)
        for i in range(self.n_devices):
            _check(load_library().q3dr_map_set_result_fields(load_library().q3dr_multi_map(self._h, i), f))

    def enable_gather(self, max_viewpoints_per_gpu: int, max_entries_per_gpu: int):
        """after this every score_viewpoints call leaves the results of ALL GPUs on every GPU (0, 0 switches it off)"""
        _check(load_library().q3dr_multi_enable_gather(self._h, max_viewpoints_per_gpu, max_entries_per_gpu))

    def gathered(self, on: int, src: int) -> DeviceResult:
        """the part scored by GPU `src` as GPU `on` holds it after the last call (device pointers on GPU `on`)"""
        r = _Result()
        _check(load_library().q3dr_multi_gathered(self._h, on, src, C.byref(r)))
        return DeviceResult(r)

    def score_viewpoints(self, K, image_width: int, window: Sequence[int], Rt, opts: Optional[ScoreOpts] = None,
                         copy: bool = True):
        """Returns [(first_viewpoint, ScoreResult)] per GPU; viewpoint v of part g is viewpoint first + v of the call.
        copy=False: the arrays alias the maps' pinned host buffers (valid until the next call)."""
        K = _f32(K, (4,))
        Rt = np.ascontiguousarray(_f32(Rt).reshape(-1, 12))
        opts = opts or default_opts()
        x0, x1, y0, y1 = window
        r = _MultiResult()
        _check(load_library().q3dr_multi_score_viewpoints(
            self._h, _fp(K), image_width, x0, x1, y0, y1, _fp(Rt), Rt.shape[0], C.byref(opts), C.byref(r)))
        first = np.ctypeslib.as_array(C.cast(r.first_viewpoint, C.POINTER(C.c_uint64)), shape=(r.n_parts + 1,)).copy()
        parts = C.cast(r.parts, C.POINTER(_Result))
        return [(int(first[g]), ScoreResult(parts[g], copy=copy)) for g in range(r.n_parts)]
