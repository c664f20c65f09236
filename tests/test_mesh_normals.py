"""Image-space normals: the reference's enable_opengl = true branch.

CPU: the oracle's restatement (known answers, decode, golden fixture), the product's traversal compiled for the host
against the oracle's brute force over all triangles, consistency of the image-normal scoring with the pinned
voxel-normal scoring.  GPU: q3dr_mesh_render_normals, q3dr_score_viewpoints_mesh_normals and
q3dr_score_viewpoints_normal_images through the C ABI against the oracle -- codes and voxel sets bit-exact,
informations exact up to expf / acosf (<= 2e-6 relative), totals within 1e-5 relative.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from quad3dr_b200 import synth

from _common import scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "tiny_mesh_normals.npz")
CLEAR = 0xFFFFFFFF


def _code(n):
    """the shader + 8-bit target for a unit normal n (numpy restatement used by the known-answer tests)"""
    c = np.clip(np.float32(0.5) * np.asarray(n, np.float32) + np.float32(0.5), 0, 1)
    q = np.rint(c * np.float32(255)).astype(np.uint32)
    return int(0xFF000000 | (q[0] << 16) | (q[1] << 8) | q[2])


IDENTITY_RT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)  # camera at the origin looking down +z


# ------------------------------------------------------------------------------------------------ CPU: oracle

def test_decode_known_answers_and_host_helper_equals_oracle():
    # clear colour (1,1,1,1) decodes to (1,1,1)/sqrt(3): the reference scores background pixels with that normal
    assert np.allclose(O.decode_normal(CLEAR), np.full(3, 1 / np.sqrt(3)), atol=1e-7)
    # mid grey: 2 (128/255 - 0.5) = 0.0039 per channel, squared norm < 0.5 -> zero vector (incidence factor 1)
    assert np.array_equal(O.decode_normal(0xFF808080), np.zeros(3, np.float32))
    assert np.allclose(O.decode_normal(_code([0, 0, 1])), [0.0039, 0.0039, 1.0], atol=1e-3)
    assert np.allclose(O.decode_normal(_code([-1, 0, 0])), [-1.0, 0.0039, 0.0039], atol=1e-3)
    rng = np.random.default_rng(5)
    codes = np.concatenate([rng.integers(0, 1 << 32, 20000, dtype=np.uint64).astype(np.uint32),
                            np.array([0, CLEAR, 0xFF000000, 0x00FFFFFF, 0xFF7F7F7F, 0xFF80807F], np.uint32)])
    for c in codes:
        a, b = O.decode_normal(int(c)), E.decode_normal(int(c))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), hex(int(c))


def test_render_known_answers():
    K = np.array([40, 40, 16, 12], np.float32)
    W, H = 32, 24
    n_front = np.tile(np.array([0, 0, -1], np.float32), (3, 1))
    n_tilt = np.tile(np.array([0.6, 0, -0.8], np.float32), (3, 1))
    big = np.array([[-50, -50, 10], [50, -50, 10], [0, 80, 10]], np.float32)       # covers the whole view at z = 10
    small = np.array([[-0.5, -0.5, 5], [0.5, -0.5, 5], [0, 0.7, 5]], np.float32)   # in front of it around the axis
    img = O.mesh_render_normals(np.stack([big, small]), np.stack([n_front, n_tilt]), K, W, H, IDENTITY_RT)
    assert img[0, 0] == _code([0, 0, -1])                       # corner pixel: only the far triangle
    assert img[H // 2, W // 2] == _code([0.6, 0, -0.8])         # centre: the nearer triangle wins the depth test
    # order of submission does not matter when depths differ
    img2 = O.mesh_render_normals(np.stack([small, big]), np.stack([n_tilt, n_front]), K, W, H, IDENTITY_RT)
    assert np.array_equal(img, img2)
    # clip planes: far = 7 removes the big triangle (background = clear colour), near = 6 removes the small one
    img3 = O.mesh_render_normals(np.stack([big, small]), np.stack([n_front, n_tilt]), K, W, H, IDENTITY_RT, far=7.0)
    assert img3[0, 0] == CLEAR and img3[H // 2, W // 2] == _code([0.6, 0, -0.8])
    img4 = O.mesh_render_normals(np.stack([big, small]), np.stack([n_front, n_tilt]), K, W, H, IDENTITY_RT, near=6.0)
    assert np.all(img4 == _code([0, 0, -1]))
    # equal depth: the triangle drawn first (lower index) stays, like GL_LESS
    img5 = O.mesh_render_normals(np.stack([big, big]), np.stack([n_front, n_tilt]), K, W, H, IDENTITY_RT)
    assert np.all(img5 == _code([0, 0, -1]))
    # a triangle behind the camera is never seen
    behind = big * np.array([1, 1, -1], np.float32)
    assert np.all(O.mesh_render_normals(behind[None], n_front[None], K, W, H, IDENTITY_RT) == CLEAR)
    # barycentric blend: normals (1,0,0), (0,1,0), (0,0,1) at the corners; the pixel nearest the centroid decodes to
    # roughly (1,1,1)/3 before the 8-bit rounding
    corners = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    tri = np.array([[-3, -3, 6], [3, -3, 6], [0, 3, 6]], np.float32)
    img6 = O.mesh_render_normals(tri[None], corners[None], K, W, H, IDENTITY_RT)
    px = img6[int(12 - 1 * 40 / 6), 16]  # centroid (0, -1, 6) projects to y = cy - 40/6
    rgb = np.array([(px >> 16) & 255, (px >> 8) & 255, px & 255]) / 255.0 * 2 - 1
    assert np.allclose(rgb, [1 / 3, 1 / 3, 1 / 3], atol=0.08) and abs(rgb.sum() - 1.0) < 0.02


def test_face_normals_follow_the_reference_formula():
    tv = np.array([[[0, 0, 0], [2, 0, 0], [0, 3, 0]]], np.float32)
    n = O.mesh_face_normals(tv).reshape(3, 3)
    assert np.array_equal(n, np.tile(np.array([0, 0, 1], np.float32), (3, 1)))  # cross(v2 - v1, v3 - v2) normalised


def test_golden_fixture_reproduced_by_the_oracle():
    g = np.load(GOLDEN)
    W, H = int(g["width"]), int(g["height"])
    tree = O.OracleTree(g["boxes"])
    for v in range(g["Rt"].shape[0]):
        img = O.mesh_render_normals(g["tri_vertices"], g["tri_normals"], g["K"], W, H, g["Rt"][v])
        assert np.array_equal(img, g["images"][v])
        s = tree.score_viewpoint_image_normals(g["weight"], img, g["K"], W, g["Rt"][v])
        a, b = int(g["offsets"][v]), int(g["offsets"][v + 1])
        assert np.array_equal(s["leaf"], g["leaf"][a:b]) and np.array_equal(s["pixel"], g["pixel"][a:b])
        assert np.array_equal(s["info"].view(np.uint32), g["info"][a:b].view(np.uint32))
    assert len(np.unique(g["images"])) > 500  # the fixture exercises real variation, not one flat colour


def test_image_normal_scoring_with_a_constant_image_equals_voxel_normal_scoring():
    """A constant image makes every voxel's normal decode(code): the (pinned) voxel-normal oracle must agree."""
    sc = scene("tiny")
    cam = synth.make_camera(64, 48)
    vps = synth.make_viewpoints(sc.leaves, 3, config_id=7)
    for code in (_code([0, 0, 1]), _code([-0.6, 0.64, 0.48]), 0xFF808080, CLEAR):
        img = np.full((cam.height, cam.width), code, np.uint32)
        nrm = np.tile(O.decode_normal(code), (sc.leaves.n, 1))
        for v in range(vps.shape[0]):
            a = sc.oracle.score_viewpoint_image_normals(sc.leaves.weight, img, cam.K, cam.width, vps[v])
            b = sc.oracle.score_viewpoint(sc.leaves.weight, nrm, cam.K, cam.width, vps[v], height=cam.height)
            assert np.array_equal(a["leaf"], b["leaf"]) and np.array_equal(a["pixel"], b["pixel"])
            assert np.array_equal(a["info"].view(np.uint32), b["info"].view(np.uint32))
            assert a["total"] == b["total"] and a["hit_count"] == b["hit_count"]


# ------------------------------------------------------- CPU: the product's traversal, compiled for the host

HARNESS_SRC = os.path.join(ROOT, "tests", "cpp", "mesh_host_harness.cu")
HARNESS_SO = os.path.join(ROOT, "tests", "cpp", "libmesh_host_harness.so")


def _harness():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    deps = [HARNESS_SRC] + [os.path.join(ROOT, "quad3dr_b200", "csrc", f) for f in ("mesh_build.h", "mesh_kernels.cuh")]
    if not os.path.exists(HARNESS_SO) or any(os.path.getmtime(d) > os.path.getmtime(HARNESS_SO) for d in deps):
        if not os.path.exists(nvcc):
            pytest.skip("nvcc not available to build the host harness")
        r = subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
                            "/usr/bin/g++", "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "-shared", "-o",
                            HARNESS_SO, HARNESS_SRC], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(HARNESS_SO)
    fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    lib.mesh_host_cast_pixels.argtypes = [fp, fp, C.c_size_t, fp, C.c_uint32, C.c_uint32, fp, C.c_float, C.c_float,
                                          C.c_uint32, u32p, u32p, C.c_size_t, u32p, u32p, C.POINTER(C.c_uint64)]
    return lib


def _host_cast(lib, tv, tn, K, W, H, Rt, xs, ys, near=0.5, far=1e5):
    fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    tv = np.ascontiguousarray(tv, np.float32).reshape(-1, 9)
    tn = None if tn is None else np.ascontiguousarray(tn, np.float32).reshape(-1, 9)
    K = np.ascontiguousarray(K, np.float32)
    Rt = np.ascontiguousarray(Rt, np.float32).reshape(12)
    xs, ys = np.ascontiguousarray(xs, np.uint32), np.ascontiguousarray(ys, np.uint32)
    out = np.empty(xs.shape[0], np.uint32)
    depth = C.c_uint32(0)
    rc = lib.mesh_host_cast_pixels(tv.ctypes.data_as(fp), None if tn is None else tn.ctypes.data_as(fp), tv.shape[0],
                                   K.ctypes.data_as(fp), W, H, Rt.ctypes.data_as(fp), near, far, CLEAR,
                                   xs.ctypes.data_as(u32p), ys.ctypes.data_as(u32p), xs.shape[0],
                                   out.ctypes.data_as(u32p), C.byref(depth), None)
    assert rc == 0
    return out


def _soup(rng, kind, nf):
    ctr = rng.uniform(-20, 20, (nf, 1, 3))
    if kind == 0:
        tv = ctr + rng.normal(0, 1.5, (nf, 3, 3))
    elif kind == 1:  # slivers
        tv = ctr + rng.normal(0, 3, (nf, 1, 3)) * rng.uniform(0, 1, (nf, 3, 1)) + rng.normal(0, 1e-3, (nf, 3, 3))
    elif kind == 2:  # grid-aligned, flat along one axis (zero-thickness boxes, exact depth ties, zero-area triangles)
        tv = np.round(ctr) + np.round(rng.uniform(-2, 2, (nf, 3, 3)))
        tv[:, :, int(rng.integers(0, 3))] = np.round(ctr[:, :, 0])
    elif kind == 3:  # huge overlapping triangles
        tv = ctr + rng.normal(0, 30, (nf, 3, 3))
    elif kind == 4:  # far apart and tiny
        tv = ctr * 50 + rng.normal(0, 0.05, (nf, 3, 3))
    else:  # every triangle twice: equal depth everywhere
        base = ctr[: nf // 2 + 1] + rng.normal(0, 2, (nf // 2 + 1, 3, 3))
        tv = np.concatenate([base, base])[:nf]
    return tv.astype(np.float32)


def test_host_compiled_traversal_equals_brute_force_on_the_synthetic_mesh():
    lib = _harness()
    L = synth.map_small()
    tv, tn = synth.make_surface_mesh(L, cell=1.0)
    cam = synth.make_camera(96, 72)
    vps = synth.make_viewpoints(L, 4, config_id=11)
    ys, xs = np.mgrid[0:cam.height, 0:cam.width]
    xs, ys = xs.ravel(), ys.ravel()
    for v in range(vps.shape[0]):
        ref = O.mesh_cast_pixels(tv, tn, cam.K, cam.width, cam.height, vps[v], xs, ys)
        got = _host_cast(lib, tv, tn, cam.K, cam.width, cam.height, vps[v], xs, ys)
        assert np.array_equal(ref, got)
        assert (ref != CLEAR).mean() > 0.2


def test_sah_tree_that_would_outgrow_the_stack_falls_back_to_the_balanced_tree(monkeypatch):
    """Triangles spaced in a geometric progression make the surface-area heuristic build a lopsided tree (29 levels for
    3000 triangles).  A tree deeper than the traversal stack allows (38 levels; lowered to 16 here through the
    builder's test knob) must be replaced by the object-median tree, with the same answers."""
    monkeypatch.setenv("Q3DR_MESH_MAX_DEPTH", "16")
    lib = _harness()
    nf = 3000
    rng = np.random.default_rng(3)
    x = (1.025 ** np.arange(nf)).astype(np.float64)
    tv = np.zeros((nf, 3, 3))
    tv[:, :, 0] = x[:, None] + rng.uniform(-0.1, 0.1, (nf, 3)) * x[:, None]
    tv[:, :, 1] = rng.uniform(-1, 1, (nf, 3)) * x[:, None]
    tv[:, :, 2] = 5.0 + rng.uniform(-1, 1, (nf, 3)) * x[:, None] * 0.2
    tv = tv.astype(np.float32)
    W, H = 64, 48
    cam = synth.make_camera(W, H)
    Rt = np.array([1, 0, 0, 0.3, 0, 1, 0, -0.2, 0, 0, 1, -1.0], np.float32)  # looks along +z
    ys, xs = np.mgrid[0:H, 0:W]
    xs, ys = xs.ravel().astype(np.uint32), ys.ravel().astype(np.uint32)
    fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    out = np.empty(xs.shape[0], np.uint32)
    depth = C.c_uint32(0)
    tvc = np.ascontiguousarray(tv.reshape(-1, 9))
    K = np.ascontiguousarray(cam.K, np.float32)
    rc = lib.mesh_host_cast_pixels(tvc.ctypes.data_as(fp), None, nf, K.ctypes.data_as(fp), W, H, Rt.ctypes.data_as(fp),
                                   0.5, 1e30, CLEAR, xs.ctypes.data_as(u32p), ys.ctypes.data_as(u32p), xs.shape[0],
                                   out.ctypes.data_as(u32p), C.byref(depth), None)
    assert rc == 0
    assert depth.value <= 12  # ceil(log2(3000 / 4)) + slack: the balanced tree, not a chain of hundreds of levels
    ref = O.mesh_cast_pixels(tv, O.mesh_face_normals(tv), cam.K, W, H, Rt, xs, ys, near=0.5, far=1e30)
    assert np.array_equal(ref, out)
    assert (ref != CLEAR).sum() > 50


def test_host_compiled_traversal_equals_brute_force_on_adversarial_soups():
    lib = _harness()
    rng = np.random.default_rng(20261020)
    rays = 0
    for it in range(36):
        nf = int(rng.integers(1, 3000))
        tv = _soup(rng, it % 6, nf)
        tn = rng.normal(0, 1, (nf, 3, 3))
        tn = (tn / np.linalg.norm(tn, axis=-1, keepdims=True)).astype(np.float32)
        W, H = int(rng.integers(8, 70)), int(rng.integers(8, 50))  # odd sizes put a pixel centre on the optical axis
        fx = float(rng.uniform(0.3, 2) * W)
        K = np.array([fx, fx * rng.uniform(0.8, 1.2), W / 2, H / 2], np.float32)
        if it % 4 == 0:  # axis-aligned camera: direction components are exact zeros for the centre row / column
            R = np.eye(3)[rng.permutation(3)] * rng.choice([-1, 1], (3, 1))
        else:
            R, _ = np.linalg.qr(rng.normal(0, 1, (3, 3)))
        t = rng.uniform(-25, 25, 3) if it % 5 else np.round(rng.uniform(-5, 5, 3))
        Rt = np.concatenate([R, t[:, None]], axis=1).astype(np.float32)
        ys, xs = np.mgrid[0:H, 0:W]
        xs, ys = xs.ravel(), ys.ravel()
        near, far = float(rng.choice([0.5, 0.01, 5.0])), float(rng.choice([1e5, 30.0]))
        use_n = None if it % 7 == 3 else tn
        ref = O.mesh_cast_pixels(tv, O.mesh_face_normals(tv) if use_n is None else tn, K, W, H, Rt, xs, ys, near, far)
        got = _host_cast(lib, tv, use_n, K, W, H, Rt, xs, ys, near, far)
        assert np.array_equal(ref, got), (it, int((ref != got).sum()))
        rays += xs.shape[0]
    assert rays > 30000


# ------------------------------------------------------------------------------------------------ GPU

def _mesh_scene(name, cell):
    sc = scene(name)
    tv, tn = synth.make_surface_mesh(sc.leaves, cell=cell)
    return sc, tv, tn


@pytest.mark.gpu
def test_gpu_render_equals_oracle_and_golden():
    g = np.load(GOLDEN)
    W, H = int(g["width"]), int(g["height"])
    mesh = E.NormalMesh(g["tri_vertices"], g["tri_normals"])
    img = mesh.render_normals(g["K"], W, H, g["Rt"])
    assert np.array_equal(img, g["images"])
    info = mesh.info()
    assert info.n_triangles == g["tri_vertices"].shape[0] and info.last_render_ms > 0


@pytest.mark.gpu
def test_gpu_render_equals_brute_force_small_scene_and_soups():
    sc, tv, tn = _mesh_scene("small", 1.0)
    cam = synth.make_camera(160, 120)
    vps = synth.make_viewpoints(sc.leaves, 6, config_id=12)
    mesh = E.NormalMesh(tv, tn)
    img = mesh.render_normals(cam.K, cam.width, cam.height, vps)
    for v in range(vps.shape[0]):
        assert np.array_equal(img[v], O.mesh_render_normals(tv, tn, cam.K, cam.width, cam.height, vps[v]))
    rng = np.random.default_rng(77)
    for it in range(18):
        nf = int(rng.integers(1, 2500))
        soup = _soup(rng, it % 6, nf)
        use_n = None
        R, _ = np.linalg.qr(rng.normal(0, 1, (3, 3)))
        if it % 3 == 0:
            R = np.eye(3)[rng.permutation(3)] * rng.choice([-1, 1], (3, 1))
        Rt = np.concatenate([R, rng.uniform(-25, 25, (3, 1))], axis=1).astype(np.float32)
        W, H = int(rng.integers(9, 90)), int(rng.integers(9, 70))
        K = np.array([0.9 * W, 0.8 * W, W / 2, H / 2], np.float32)
        ro = E.default_render_opts(near_plane=float(rng.choice([0.5, 0.01])), far_plane=float(rng.choice([1e5, 40.0])),
                                   clear_argb=0xFF102030)
        got = E.NormalMesh(soup, use_n).render_normals(K, W, H, Rt[None], ro)[0]
        ref = O.mesh_render_normals(soup, O.mesh_face_normals(soup), K, W, H, Rt, ro.near_plane, ro.far_plane, 0xFF102030)
        assert np.array_equal(got, ref), it


def _check_against_oracle(sc, res, images, cam, vps, window=None):
    for v in range(vps.shape[0]):
        ref = sc.oracle.score_viewpoint_image_normals(sc.leaves.weight, images[v], cam.K, cam.width, vps[v], window)
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], ref["leaf"])
        assert np.array_equal(got["pixel"], ref["pixel"])
        assert got["hit_count"] == ref["hit_count"]
        assert np.allclose(got["info"], ref["info"], rtol=2e-6, atol=0)
        assert abs(got["total"] - ref["total"]) <= 1e-5 * abs(ref["total"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,cell,w,h,nvp", [("tiny", 1.0, 64, 48, 6), ("small", 1.0, 160, 120, 8),
                                               ("small_unknown", 2.0, 96, 72, 5)])
def test_gpu_scoring_with_mesh_normals_equals_oracle(name, cell, w, h, nvp):
    sc, tv, tn = _mesh_scene(name, cell)
    cam = synth.make_camera(w, h)
    vps = synth.make_viewpoints(sc.leaves, nvp, config_id=13)
    images = np.stack([O.mesh_render_normals(tv, tn, cam.K, w, h, vps[v]) for v in range(nvp)])
    mesh = E.NormalMesh(tv, tn)
    full = (0, w, 0, h)
    fused = sc.gpu_map.score_viewpoints_mesh_normals(mesh, cam.K, w, h, full, vps)
    _check_against_oracle(sc, fused, images, cam, vps)
    # the caller's-images entry point with the oracle's images, and with the engine's own render: same result
    via_images = sc.gpu_map.score_viewpoints_normal_images(images, cam.K, w, full, vps)
    own = sc.gpu_map.score_viewpoints_normal_images(mesh.render_normals(cam.K, w, h, vps), cam.K, w, full, vps)
    for r in (via_images, own):
        assert np.array_equal(r.leaf, fused.leaf) and np.array_equal(r.first_pixel, fused.first_pixel)
        assert np.array_equal(r.information.view(np.uint32), fused.information.view(np.uint32))
        assert np.array_equal(r.total.view(np.uint32), fused.total.view(np.uint32))
    # and the normals matter: the voxel-normal result differs
    plain = sc.gpu_map.score_viewpoints(cam.K, w, full, vps)
    assert not np.array_equal(plain.total, fused.total)


@pytest.mark.gpu
def test_gpu_scoring_with_mesh_normals_sub_window_and_options():
    sc, tv, tn = _mesh_scene("small", 1.0)
    cam = synth.make_camera(128, 96)
    vps = synth.make_viewpoints(sc.leaves, 4, config_id=14)
    images = np.stack([O.mesh_render_normals(tv, tn, cam.K, cam.width, cam.height, vps[v]) for v in range(4)])
    mesh = E.NormalMesh(tv, tn)
    window = (17, 113, 9, 80)
    opts_g = E.default_opts(incidence_ignore_dot_product_sign=1, ignore_voxels_with_zero_information=0)
    opts_o = O.default_opts(incidence_ignore_dot_product_sign=1, ignore_voxels_with_zero_information=0)
    res = sc.gpu_map.score_viewpoints_mesh_normals(mesh, cam.K, cam.width, cam.height, window, vps, opts_g)
    for v in range(4):
        ref = sc.oracle.score_viewpoint_image_normals(sc.leaves.weight, images[v], cam.K, cam.width, vps[v], window,
                                                      opts=opts_o)
        got = res.viewpoint(v)
        assert np.array_equal(got["leaf"], ref["leaf"]) and np.array_equal(got["pixel"], ref["pixel"])
        assert np.allclose(got["info"], ref["info"], rtol=2e-6, atol=0)
    # the voxel-normal path still works on the same map afterwards (scratch reuse)
    plain = sc.gpu_map.score_viewpoints(cam.K, cam.width, (0, cam.width, 0, cam.height), vps)
    for v in range(4):
        ref = sc.oracle.score_viewpoint(sc.leaves.weight, sc.leaves.normal, cam.K, cam.width, vps[v], height=cam.height)
        assert np.array_equal(plain.viewpoint(v)["leaf"], ref["leaf"])


@pytest.mark.gpu
def test_gpu_mesh_argument_errors():
    sc = scene("tiny")
    cam = synth.make_camera(32, 24)
    vps = synth.make_viewpoints(sc.leaves, 1)
    with pytest.raises(E.Q3drError):
        E.NormalMesh(np.zeros((0, 3, 3), np.float32))
    with pytest.raises(E.Q3drError):
        E.NormalMesh(np.full((2, 3, 3), np.nan, np.float32))
    mesh = E.NormalMesh(np.array([[[0, 0, 5], [1, 0, 5], [0, 1, 5]]], np.float32))  # one triangle: root is a leaf
    img = mesh.render_normals(cam.K, 32, 24, IDENTITY_RT[None])
    assert set(np.unique(img).tolist()) == {CLEAR, _code([0, 0, 1])}
    with pytest.raises(E.Q3drError):  # window larger than the normal image
        sc.gpu_map.score_viewpoints_mesh_normals(mesh, cam.K, 32, 24, (0, 32, 0, 30), vps)
    with pytest.raises(ValueError):
        sc.gpu_map.score_viewpoints_normal_images(np.zeros((2, 24, 32), np.uint32), cam.K, 32, (0, 32, 0, 24), vps)


@pytest.mark.gpu
def test_gpu_full_size_fused_equals_render_then_lookup():
    """C2-sized property (1.03 M leaves, 1.44 M triangles, 640x480): casting only the kept pixels inside the scoring
    pass gives bit for bit what rendering whole images and looking the kept pixels up gives; a sample of kept pixels is
    checked against the oracle's brute force over all triangles."""
    sc, tv, tn = _mesh_scene("1m", 0.25)
    cam = synth.make_camera(640, 480)
    vps = synth.make_viewpoints(sc.leaves, 6, config_id=2)
    mesh = E.NormalMesh(tv, tn)
    assert mesh.info().n_triangles == tv.shape[0] > 1_000_000
    win = (0, 640, 0, 480)
    fused = sc.gpu_map.score_viewpoints_mesh_normals(mesh, cam.K, 640, 480, win, vps)
    images = mesh.render_normals(cam.K, 640, 480, vps)
    via = sc.gpu_map.score_viewpoints_normal_images(images, cam.K, 640, win, vps)
    assert fused.n_entries == via.n_entries > 50_000
    assert np.array_equal(fused.leaf, via.leaf) and np.array_equal(fused.first_pixel, via.first_pixel)
    assert np.array_equal(fused.information.view(np.uint32), via.information.view(np.uint32))
    assert np.array_equal(fused.total.view(np.uint32), via.total.view(np.uint32))
    rng = np.random.default_rng(3)
    for v in (0, 5):
        pix = fused.viewpoint(v)["pixel"]
        pick = rng.choice(pix, size=min(400, pix.shape[0]), replace=False)
        ref = O.mesh_cast_pixels(tv, tn, cam.K, 640, 480, vps[v], pick % 640, pick // 640)
        assert np.array_equal(images[v][pick // 640, pick % 640], ref)


# --------------------------------------------------- CPU: the conventions, derived a second way from the reference

def _gl_style_render(tv, tn, K, W, H, Rt, near=0.5, far=1e5):
    """The reference's OpenGL pipeline in float64 numpy, matrix by matrix as viewpoint_offscreen_renderer.cpp builds
    them: view = world-to-image of the pose rotated by pi about x (:303-321), proj (:283-289), clip -> NDC -> window,
    samples at pixel centres, depth test GL_LESS, perspective-correct interpolation, shader 0.5 n + 0.5, 8-bit target,
    QOpenGLFramebufferObject::toImage() flipping the rows.  Triangles are rasterised in homogeneous coordinates
    (no explicit clipping).  Independent of the ray formula of mesh_kernels.cuh / the oracle."""
    fx, fy, cx, cy = [float(v) for v in K]
    M = np.asarray(Rt, np.float64).reshape(3, 4)
    R, t = M[:, :3], M[:, 3]
    Rv = R @ np.diag([1.0, -1.0, -1.0])                       # quaternion * AngleAxis(pi, UnitX)
    V = np.eye(4)
    V[:3, :3] = Rv.T                                            # getTransformationWorldToImage
    V[:3, 3] = -Rv.T @ t
    P = np.zeros((4, 4))
    P[0, 0], P[1, 1] = fx / cx, fy / cy
    P[2, 2] = -(far + near) / (far - near)
    P[2, 3] = -2 * far * near / (far - near)
    P[3, 2] = -1
    PV = P @ V
    jj, ii = np.mgrid[0:H, 0:W]                                 # window pixel (ii, jj), jj counted from the BOTTOM
    px = (ii + 0.5) / W * 2 - 1
    py = (jj + 0.5) / H * 2 - 1
    depth = np.full((H, W), np.inf)
    color = np.full((H, W, 3), 255, np.int64)                   # clear colour (1,1,1,1)
    rhs = np.stack([px.ravel(), py.ravel(), np.ones(W * H)])
    for f in range(tv.shape[0]):
        c = (PV @ np.concatenate([tv[f].astype(np.float64), np.ones((3, 1))], axis=1).T)   # columns = corners
        A = np.stack([c[0], c[1], c[3]])                        # x_clip, y_clip, w_clip of the three corners
        if abs(np.linalg.det(A)) < 1e-30:
            continue
        lam = np.linalg.solve(A, rhs)                           # b_i / w at every pixel centre
        s = lam.sum(0)
        inside = (lam >= 0).all(0) & (s > 0)
        if not inside.any():
            continue
        wclip = 1.0 / np.where(inside, s, 1.0)                  # = camera depth (w_clip = -z_eye)
        ok = inside & (wclip >= near) & (wclip <= far) & (wclip < depth.ravel())
        if not ok.any():
            continue
        b = lam * wclip                                         # perspective-correct barycentrics
        n = b.T @ tn[f].astype(np.float64)
        col = np.rint(np.clip(0.5 * n + 0.5, 0, 1) * 255).astype(np.int64)
        d = depth.ravel()
        d[ok] = wclip[ok]
        depth = d.reshape(H, W)
        cflat = color.reshape(-1, 3)
        cflat[ok] = col[ok]
        color = cflat.reshape(H, W, 3)
    color = color[::-1]                                         # toImage(): row 0 on top
    return (0xFF000000 | (color[..., 0] << 16) | (color[..., 1] << 8) | color[..., 2]).astype(np.uint32)


def test_ray_cast_definition_agrees_with_a_gl_style_pipeline():
    """The image definition (pixel-centre rays, camera axes, row order, depth, interpolation, encoding) against the
    reference's own matrices pushed through a rasteriser-style pipeline: identical except where float32 vs float64
    rounding flips an edge / tie / 8-bit boundary."""
    g = np.load(GOLDEN)
    W, H = int(g["width"]), int(g["height"])
    tv, tn = g["tri_vertices"], g["tri_normals"]
    total = exact = close = same_cover = 0
    for v in range(g["Rt"].shape[0]):
        ref = _gl_style_render(tv, tn, g["K"], W, H, g["Rt"][v])
        got = g["images"][v]
        ch = lambda a: np.stack([(a >> 16) & 255, (a >> 8) & 255, a & 255], -1).astype(np.int64)
        diff = np.abs(ch(ref) - ch(got)).max(-1)
        total += W * H
        exact += int((diff == 0).sum())
        close += int((diff <= 1).sum())
        same_cover += int(((ref == CLEAR) == (got == CLEAR)).sum())
    assert same_cover >= 0.995 * total, (same_cover, total)      # hit / background: silhouette pixels only
    assert close >= 0.985 * total and exact >= 0.95 * total, (exact, close, total)
    # an upside-down or mirrored convention would agree on a few per cent of the pixels only
    flipped = g["images"][0][::-1]
    ref0 = _gl_style_render(tv, tn, g["K"], W, H, g["Rt"][0])
    assert (flipped == ref0).mean() < 0.6
