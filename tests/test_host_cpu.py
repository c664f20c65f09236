"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, the host
helpers agree with the oracle, the flattened tree keeps its invariants, and -- with no GPU -- compute entry points
fail loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "q3dr_raycast.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(q3dr_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/q3dr_raycast.h but not exported"
    assert lib.q3dr_abi_version() == 5


def test_struct_layouts_match_header():
    assert E.PIXEL_HIT_DTYPE.itemsize == 32
    assert ctypes.sizeof(E.ScoreOpts) == 32
    assert ctypes.sizeof(E._Result) == 9 * 8
    assert ctypes.sizeof(E.MapInfo) == 48
    assert ctypes.sizeof(E.Stats) == 40
    assert ctypes.sizeof(E.RenderOpts) == 12   # q3dr_render_opts
    assert ctypes.sizeof(E.MeshInfo) == 40     # q3dr_mesh_info
    assert ctypes.sizeof(E.ExchangeLayout) == 64  # q3dr_exchange_layout
    assert ctypes.sizeof(E.HostBlock) == 32       # q3dr_host_block


def test_score_opts_defaults_match_reference_defaults():
    o = E.default_opts()
    assert abs(o.voxel_sensor_size_ratio_threshold - 0.002) < 1e-9
    assert abs(o.voxel_sensor_size_ratio_inv_falloff_factor - 0.001) < 1e-9
    assert o.incidence_angle_threshold_degrees == 30.0 and o.incidence_angle_inv_falloff_factor_degrees == 30.0
    assert o.incidence_ignore_dot_product_sign == 0 and o.ignore_voxels_with_zero_information == 1
    assert o.raycast_min_range == 0.0 and o.raycast_max_range == 60.0


@pytest.mark.parametrize("factory", [synth.map_tiny, lambda: synth.map_tiny(True), synth.map_small])
def test_splitmedian_order_equals_oracle_tree(factory):
    leaves = factory()
    order, depth = E.splitmedian_order(leaves.boxes)
    tree = O.OracleTree(leaves.boxes)
    assert np.array_equal(order, tree.dfs_order())
    assert np.array_equal(depth, tree.leaf_depths()[order])


def test_splitmedian_order_with_duplicates_and_small_n():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 17, 257, 5000):
        c = rng.integers(-4, 4, size=(n, 3)).astype(np.float32) * 0.25  # many equal centres
        s = rng.choice([0.25, 0.5, 1.0], size=(n, 1)).astype(np.float32)
        boxes = np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32)
        order, depth = E.splitmedian_order(boxes)
        tree = O.OracleTree(boxes)
        assert np.array_equal(order, tree.dfs_order()), n
        assert np.array_equal(depth, tree.leaf_depths()[order]), n
        assert sorted(order.tolist()) == list(range(n))


def test_pose_to_rt_equals_oracle():
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = rng.normal(size=4).astype(np.float32)
        q /= np.linalg.norm(q)
        t = rng.normal(size=3).astype(np.float32) * 50
        a, b = E.pose_to_rt(q, t), O.pose_to_rt(q, t)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        R = a.reshape(3, 4)[:, :3].astype(np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)


@pytest.mark.parametrize("factory,top", [(synth.map_tiny, 4), (lambda: synth.map_tiny(True), 2), (synth.map_small, 4)])
def test_flat_tree_invariants(factory, top):
    leaves = factory()
    order, _ = E.splitmedian_order(leaves.boxes)
    info = E.debug_flat_tree(leaves.boxes[order], top_levels=top)  # raises if an invariant is broken
    n = leaves.n
    assert (n - 1) / 3 <= info["n_nodes"] <= n
    assert info["max_stack"] <= 64
    assert info["top_nodes"] <= sum(4 ** k for k in range(top))


def test_flat_tree_degenerate_inputs():
    one = np.array([[0, 0, 0, 1, 1, 1]], np.float32)
    assert E.debug_flat_tree(one)["n_nodes"] == 1
    same = np.tile(one, (37, 1))  # identical boxes: no spatial split possible
    assert E.debug_flat_tree(same)["n_nodes"] >= 12
    rng = np.random.default_rng(0)
    c = rng.uniform(-50, 50, size=(3000, 3)).astype(np.float32)
    s = rng.uniform(0.01, 5.0, size=(3000, 3)).astype(np.float32)
    E.debug_flat_tree(np.concatenate([c - s, c + s], axis=1).astype(np.float32))


def test_invalid_arguments_are_reported():
    lib = E.load_library()
    assert lib.q3dr_splitmedian_order(None, 3, None, None) == E.Q3DR_ERR_INVALID_ARG
    assert b"NULL" in lib.q3dr_last_error()
    h = ctypes.c_void_p(None)
    assert lib.q3dr_map_create(None, None, None, None, 3, 0, ctypes.byref(h)) == E.Q3DR_ERR_INVALID_ARG
    assert lib.q3dr_map_destroy(None) == E.Q3DR_OK
    assert lib.q3dr_status_string(E.Q3DR_ERR_NO_DEVICE).startswith(b"no usable CUDA device")


@pytest.mark.skipif(E.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    leaves = synth.map_tiny()
    with pytest.raises(E.Q3drError) as ei:
        E.OccupancyMap(leaves.boxes, leaves.weight, leaves.normal)
    assert ei.value.status == E.Q3DR_ERR_NO_DEVICE
    # the mesh of the image-space normals has no CPU path either
    with pytest.raises(E.Q3drError) as ei:
        E.NormalMesh(np.array([[[0, 0, 5], [1, 0, 5], [0, 1, 5]]], np.float32))
    assert ei.value.status == E.Q3DR_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under quad3dr_b200/ or include/ may reference it."""
    bad = []
    for base in ("quad3dr_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_py|q3dr_oracle|libq3dr_oracle|from oracle|import oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_synthetic_generators_are_deterministic():
    a, b = synth.map_tiny(), synth.map_tiny()
    assert np.array_equal(a.boxes, b.boxes) and np.array_equal(a.weight, b.weight)
    v1 = synth.make_viewpoints(a, 16, config_id=9)
    v2 = synth.make_viewpoints(b, 16, config_id=9)
    assert np.array_equal(v1, v2)
    R = v1[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].reshape(-1, 3, 3).astype(np.float64)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-5)
    assert np.all(np.abs(np.linalg.det(R) - 1) < 1e-5)
    # no origin on a voxel face, octomap-aligned boxes
    assert not np.any(np.mod(v1[:, [3, 7, 11]].astype(np.float64), 0.25) == 0)
    assert np.all(np.mod(a.boxes.astype(np.float64), 0.25) == 0)
    m = synth.map_tiny(unknown_interior=True)
    sides = m.boxes[:, 3] - m.boxes[:, 0]
    assert sides.max() > 0.25 and np.all((m.normal[sides > 0.25] == 0))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 1000, 13207])
def test_voxel_indices_equal_the_reference_iterator_order(n):
    rng = np.random.default_rng(n)
    c = rng.uniform(-5, 5, size=(n, 3)).astype(np.float32)
    boxes = np.concatenate([c - 0.1, c + 0.1], axis=1).astype(np.float32)
    order, _ = E.splitmedian_order(boxes)
    boxes = np.ascontiguousarray(boxes[order])
    t = O.OracleTree(boxes)
    assert np.array_equal(t.dfs_order(), np.arange(n))
    assert np.array_equal(E.voxel_indices(n), t.voxel_index())


def test_no_fused_packed_multiply_add_in_the_library():
    """ptxas 12.9 contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 even under --fmad false (one rounding instead
    of the reference's two).  The kernels avoid that shape (raycast2_kernels.cuh: packed products, scalar sums); this
    test makes a relapse visible without a GPU."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", E.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "FMUL2" in sass and "FADD2" in sass  # the packed path is really there ...
    assert "FFMA2" not in sass  # ... and nothing of it was contracted


def test_host_block_calls_without_a_gpu():
    lib = E.load_library()
    hb = E.HostBlock()
    assert lib.q3dr_result_detach_host(None, ctypes.byref(hb)) == E.Q3DR_ERR_INVALID_ARG
    assert lib.q3dr_host_block_release(None) == E.Q3DR_OK
    assert lib.q3dr_host_block_release(ctypes.byref(hb)) == E.Q3DR_OK  # all zeros: nothing to do
    assert lib.q3dr_host_pool_trim() == E.Q3DR_OK  # empty pool
