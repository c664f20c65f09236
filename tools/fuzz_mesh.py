"""Randomised parity of the mesh-normal render (row N2) on the GPU against the oracle's brute force over all
triangles: adversarial triangle soups (slivers, zero-area and grid-aligned triangles, exact depth ties, huge and tiny
triangles), axis-aligned and random cameras, odd image sizes, clip planes.  python tools/fuzz_mesh.py [seconds] [seed]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import oracle_py as O
from quad3dr_b200 import engine as E
from test_mesh_normals import _soup

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else int(time.time())
rng = np.random.default_rng(seed)
t0 = time.time()
scenes = rays = bad = 0
while time.time() - t0 < budget:
    it = scenes
    nf = int(rng.integers(1, 6000))
    tv = _soup(rng, it % 6, nf)
    tn = None
    if it % 3:
        tn = rng.normal(0, 1, (nf, 3, 3))
        tn = (tn / np.linalg.norm(tn, axis=-1, keepdims=True)).astype(np.float32)
    W, H = int(rng.integers(8, 120)), int(rng.integers(8, 90))
    fx = float(rng.uniform(0.3, 2) * W)
    K = np.array([fx, fx * rng.uniform(0.8, 1.2), W / 2, H / 2], np.float32)
    nv = 3
    Rts = []
    for k in range(nv):
        if (it + k) % 4 == 0:
            R = np.eye(3)[rng.permutation(3)] * rng.choice([-1, 1], (3, 1))
        else:
            R, _ = np.linalg.qr(rng.normal(0, 1, (3, 3)))
        t = rng.uniform(-25, 25, 3) if (it + k) % 5 else np.round(rng.uniform(-5, 5, 3))
        Rts.append(np.concatenate([R, t[:, None]], axis=1).astype(np.float32).reshape(12))
    Rts = np.stack(Rts)
    ro = E.default_render_opts(near_plane=float(rng.choice([0.5, 0.01, 5.0])), far_plane=float(rng.choice([1e5, 30.0])),
                               clear_argb=int(rng.integers(0, 1 << 32)))
    got = E.NormalMesh(tv, tn).render_normals(K, W, H, Rts, ro)
    ref_n = O.mesh_face_normals(tv) if tn is None else tn
    for k in range(nv):
        ref = O.mesh_render_normals(tv, ref_n, K, W, H, Rts[k], ro.near_plane, ro.far_plane, ro.clear_argb)
        nb = int((ref != got[k]).sum())
        if nb:
            bad += nb
            print("MISMATCH scene", it, "view", k, "pixels", nb, "seed", seed, file=sys.stderr)
        rays += W * H
    scenes += 1
print(json.dumps({"seed": seed, "scenes": scenes, "rays": rays, "mismatches": bad, "seconds": round(time.time() - t0, 1)}))
sys.exit(1 if bad else 0)
