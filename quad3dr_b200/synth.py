"""Synthetic "box-building" occupancy maps and candidate viewpoints.

Pure numpy, seeded with a documented counter-based PRNG (splitmix64), so the same arrays come out
on every machine.  Nothing here touches the GPU or the oracle: these are *inputs*.

Map geometry follows Octomap (viewpoint_planner/src/octree/occupancy_map.h:389-403 of the reference):
resolution r, finest leaf centres at (k + 0.5) r, coarser leaves are aligned cubes of side r 2^j.
Leaves kept in the BVH are everything that is not (free and known)
(viewpoint_planner/src/planner/viewpoint_planner_data.cpp:828-832): occupied shell voxels and,
optionally, unknown-space cubes filling building interiors.

Camera and pose conventions follow viewpoint_planner.cpp:343-352 (fx = fy, cx = W/2, cy = H/2) and
pose.h:122-127 (row-major 3x4 [R|t], image-to-world; camera looks down +z, x right, y down).
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import numpy as np

MAP_SEED = 0x51ADD3D
VIEWPOINT_SEED = 0xC0FFEE

_U64 = np.uint64
_MASK = (1 << 64) - 1


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n uint64 values of the splitmix64 sequence started at `seed` (+ a stream offset)."""
    with np.errstate(over="ignore"):
        base = _U64((seed + 0x632BE59BD9B4E019 * stream) & _MASK)
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = base + idx * _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        z = z ^ (z >> _U64(31))
    return z


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n float64 values uniform in [0, 1) with 53 random bits."""
    return (splitmix64(seed, n, stream) >> _U64(11)).astype(np.float64) * (1.0 / (1 << 53))


@dataclasses.dataclass
class Building:
    x0: int  # voxel coordinates of the outer shell (inclusive lower corner)
    y0: int
    nx: int
    ny: int
    nz: int
    floor_every: int = 12


@dataclasses.dataclass
class OccupancyLeaves:
    """BVH leaf list as generateBVHTree would emit it (viewpoint_planner_data.cpp:820-897)."""

    boxes: np.ndarray  # (n, 6) float32: min xyz, max xyz
    occupancy: np.ndarray  # (n,) float32
    observation_count: np.ndarray  # (n,) uint16
    weight: np.ndarray  # (n,) float32
    normal: np.ndarray  # (n, 3) float32
    resolution: float
    buildings: List[Building]
    ground_half: int

    @property
    def n(self) -> int:
        return int(self.boxes.shape[0])

    def permuted(self, order: np.ndarray) -> "OccupancyLeaves":
        return dataclasses.replace(
            self,
            boxes=np.ascontiguousarray(self.boxes[order]),
            occupancy=np.ascontiguousarray(self.occupancy[order]),
            observation_count=np.ascontiguousarray(self.observation_count[order]),
            weight=np.ascontiguousarray(self.weight[order]),
            normal=np.ascontiguousarray(self.normal[order]),
        )


def _shell_voxels(b: Building) -> Tuple[np.ndarray, np.ndarray]:
    """Integer voxel coordinates (n,3) and outward normals (n,3) of a hollow box building."""
    xs = np.arange(b.x0, b.x0 + b.nx)
    ys = np.arange(b.y0, b.y0 + b.ny)
    zs = np.arange(0, b.nz)
    coords = []
    normals = []

    def add(x, y, z, n):
        g = np.stack(np.meshgrid(x, y, z, indexing="ij"), axis=-1).reshape(-1, 3)
        coords.append(g)
        normals.append(np.tile(np.asarray(n, dtype=np.float32), (g.shape[0], 1)))

    zw = zs[:-1]  # walls below the roof layer
    add(xs[:1], ys, zw, (-1, 0, 0))
    add(xs[-1:], ys, zw, (1, 0, 0))
    add(xs[1:-1], ys[:1], zw, (0, -1, 0))
    add(xs[1:-1], ys[-1:], zw, (0, 1, 0))
    add(xs, ys, zs[-1:], (0, 0, 1))  # roof
    if b.floor_every > 0:
        fz = np.arange(b.floor_every, b.nz - 1, b.floor_every)
        if fz.size:
            add(xs[1:-1], ys[1:-1], fz, (0, 0, 1))
    return np.concatenate(coords), np.concatenate(normals)


def _aligned_cubes(lo: Sequence[int], hi: Sequence[int], max_log2: int = 6) -> np.ndarray:
    """Decompose the integer box [lo, hi) into maximal octree-aligned cubes -> (m, 4) [x, y, z, side]."""
    lo = np.asarray(lo, dtype=np.int64)
    hi = np.asarray(hi, dtype=np.int64)
    out = []
    if np.any(hi <= lo):
        return np.zeros((0, 4), dtype=np.int64)
    side0 = 1 << max_log2
    start = (lo // side0) * side0
    stop = -((-hi) // side0) * side0
    stack = [
        (x, y, z, side0)
        for x in range(start[0], stop[0], side0)
        for y in range(start[1], stop[1], side0)
        for z in range(start[2], stop[2], side0)
    ]
    while stack:
        x, y, z, s = stack.pop()
        c_lo = np.array([x, y, z])
        c_hi = c_lo + s
        if np.any(c_hi <= lo) or np.any(c_lo >= hi):
            continue
        if np.all(c_lo >= lo) and np.all(c_hi <= hi):
            out.append((x, y, z, s))
            continue
        h = s // 2
        for dx in (0, h):
            for dy in (0, h):
                for dz in (0, h):
                    stack.append((x + dx, y + dy, z + dz, h))
    out.sort()
    return np.asarray(out, dtype=np.int64).reshape(-1, 4)


def make_box_building_map(
    ground_half: int,
    buildings: Sequence[Building],
    resolution: float = 0.25,
    unknown_interior: bool = False,
    seed: int = MAP_SEED,
) -> OccupancyLeaves:
    """Ground slab of (2*ground_half)^2 voxels at z index -1 plus hollow box buildings standing on it."""
    r = np.float64(resolution)
    gx = np.arange(-ground_half, ground_half)
    g = np.stack(np.meshgrid(gx, gx, np.array([-1]), indexing="ij"), axis=-1).reshape(-1, 3)
    coords = [g]
    sides = [np.ones(g.shape[0], dtype=np.int64)]
    normals = [np.tile(np.array([0, 0, 1], dtype=np.float32), (g.shape[0], 1))]
    known = [np.ones(g.shape[0], dtype=bool)]
    for b in buildings:
        c, n = _shell_voxels(b)
        coords.append(c)
        sides.append(np.ones(c.shape[0], dtype=np.int64))
        normals.append(n)
        known.append(np.ones(c.shape[0], dtype=bool))
        if unknown_interior:
            zs = [0] + list(range(b.floor_every, b.nz - 1, b.floor_every)) + [b.nz - 1]
            for k in range(len(zs) - 1):
                z_lo = zs[k] + (1 if k > 0 else 0)
                z_hi = zs[k + 1]
                cubes = _aligned_cubes((b.x0 + 1, b.y0 + 1, z_lo), (b.x0 + b.nx - 1, b.y0 + b.ny - 1, z_hi))
                if cubes.shape[0]:
                    coords.append(cubes[:, :3])
                    sides.append(cubes[:, 3])
                    normals.append(np.zeros((cubes.shape[0], 3), dtype=np.float32))
                    known.append(np.zeros(cubes.shape[0], dtype=bool))
    coords = np.concatenate(coords).astype(np.int64)
    sides = np.concatenate(sides)
    normals = np.concatenate(normals)
    known = np.concatenate(known)
    # deterministic input order: x-major grid order (it feeds std::stable_sort in Tree::build)
    order = np.lexsort((sides, coords[:, 2], coords[:, 1], coords[:, 0]))
    coords, sides, normals, known = coords[order], sides[order], normals[order], known[order]
    n = coords.shape[0]
    mn = (coords.astype(np.float64) * r).astype(np.float32)
    mx = ((coords + sides[:, None]).astype(np.float64) * r).astype(np.float32)
    boxes = np.ascontiguousarray(np.concatenate([mn, mx], axis=1), dtype=np.float32)
    cnt = (splitmix64(seed, n) % _U64(20)).astype(np.uint16) + np.uint16(1)
    cnt = np.where(known, cnt, np.uint16(0)).astype(np.uint16)
    occupancy = np.where(known, np.float32(0.97), np.float32(0.5)).astype(np.float32)
    # viewpoint_planner_data.cpp:573-577: weight = grid_weight * exp(-0.1 * observation_count)
    weight = np.exp(np.float32(-0.1) * cnt.astype(np.float32)).astype(np.float32)
    return OccupancyLeaves(
        boxes=boxes,
        occupancy=occupancy,
        observation_count=cnt,
        weight=weight,
        normal=np.ascontiguousarray(normals, dtype=np.float32),
        resolution=float(resolution),
        buildings=list(buildings),
        ground_half=int(ground_half),
    )


# ---------------------------------------------------------------------------------------------
# Named configurations (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------

def map_tiny(unknown_interior: bool = False) -> OccupancyLeaves:
    """~6.6 k leaves: used by unit tests and golden fixtures."""
    return make_box_building_map(
        32, [Building(-12, -8, 24, 16, 14, 6), Building(6, 14, 10, 12, 9, 4)], unknown_interior=unknown_interior
    )


def map_small(unknown_interior: bool = False) -> OccupancyLeaves:
    """~75 k leaves: parity tests that the oracle finishes in seconds."""
    return make_box_building_map(
        120,
        [Building(-50, -30, 60, 40, 40, 8), Building(30, 20, 40, 50, 30, 10), Building(-80, 50, 30, 30, 60, 12)],
        unknown_interior=unknown_interior,
    )


_B1M = [
    Building(-300, -260, 120, 80, 100, 12),
    Building(-60, -40, 140, 90, 110, 12),
    Building(180, 120, 120, 100, 90, 12),
    Building(-280, 160, 100, 120, 60, 12),
]


def map_1m(unknown_interior: bool = False) -> OccupancyLeaves:
    """C1/C2 (and C4 with unknown_interior=True): 720x720 ground + 4 buildings, 1,028,168 leaves
    (< 2^20, so the reference tree has depth 20 as in SURVEY.md section 8d)."""
    return make_box_building_map(360, _B1M, unknown_interior=unknown_interior)


def map_10m() -> OccupancyLeaves:
    """C3: 2530x2530 ground + 8 large buildings, ~10 M leaves (octree depth 16 at r = 0.25 m)."""
    bl = []
    k = 0
    for ix in (-900, -300, 300, 900):
        for iy in (-600, 400):
            bl.append(Building(ix - 100, iy - 75, 200, 150, 140 + 10 * (k % 3), 12))
            k += 1
    return make_box_building_map(1265, bl)


# ---------------------------------------------------------------------------------------------
# Viewpoints
# ---------------------------------------------------------------------------------------------

@dataclasses.dataclass
class Camera:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float

    @property
    def K(self) -> np.ndarray:
        return np.array([self.fx, self.fy, self.cx, self.cy], dtype=np.float32)


def make_camera(width: int, height: int) -> Camera:
    """setVirtualCamera(width, height, focal_length): fx = fy = 0.85 W, cx = W/2, cy = H/2."""
    f = np.float32(0.85) * np.float32(width)
    return Camera(width, height, float(f), float(f), float(np.float32(width) / 2), float(np.float32(height) / 2))


def make_viewpoints(leaves: OccupancyLeaves, n: int, config_id: int = 0, seed: int = VIEWPOINT_SEED) -> np.ndarray:
    """n poses as (n, 12) float32 row-major [R|t].

    Position: uniform in the 5-40 m shell around a randomly chosen building box, z in [2, 60], rejected if
    within 1 m of any building.  Orientation: camera z towards a uniform random point of that building box,
    pitch clamped <= 0, x = z cross up, y = z cross x.
    """
    r = leaves.resolution
    bl = leaves.buildings
    nb = len(bl)
    lo = np.array([[b.x0 * r, b.y0 * r, 0.0] for b in bl])
    hi = np.array([[(b.x0 + b.nx) * r, (b.y0 + b.ny) * r, b.nz * r] for b in bl])
    limit = leaves.ground_half * r - 1.0
    out = np.zeros((n, 12), dtype=np.float32)
    filled = 0
    attempt = 0
    s = seed + config_id
    while filled < n:
        m = max(2 * (n - filled), 64)
        u = uniform01(s, m * 8, stream=attempt).reshape(m, 8)
        attempt += 1
        bi = np.minimum((u[:, 0] * nb).astype(np.int64), nb - 1)
        blo, bhi = lo[bi], hi[bi]
        margin = 40.0
        p = np.empty((m, 3))
        p[:, 0] = blo[:, 0] - margin + u[:, 1] * (bhi[:, 0] - blo[:, 0] + 2 * margin)
        p[:, 1] = blo[:, 1] - margin + u[:, 2] * (bhi[:, 1] - blo[:, 1] + 2 * margin)
        p[:, 2] = 2.0 + u[:, 3] * 58.0
        # tiny irrational offset: no coordinate sits on a voxel face (multiples of r)
        p += np.array([np.sqrt(2.0), np.sqrt(3.0), np.sqrt(5.0)]) * 1e-3
        # distance to the chosen building box
        d = np.maximum(np.maximum(blo - p, p - bhi), 0.0)
        dist = np.sqrt((d * d).sum(axis=1))
        ok = (dist >= 5.0) & (dist <= 40.0) & (np.abs(p[:, 0]) < limit) & (np.abs(p[:, 1]) < limit)
        for j in range(nb):
            dj = np.maximum(np.maximum(lo[j] - p, p - hi[j]), 0.0)
            ok &= np.sqrt((dj * dj).sum(axis=1)) >= 1.0
        tgt = blo + u[:, 4:7] * (bhi - blo)
        z = tgt - p
        z[:, 2] = np.minimum(z[:, 2], 0.0)  # pitch <= 0
        zn = np.linalg.norm(z, axis=1)
        ok &= zn > 1e-3
        zn = np.where(zn > 0, zn, 1.0)
        z = z / zn[:, None]
        up = np.array([0.0, 0.0, 1.0])
        x = np.cross(z, up)
        xn = np.linalg.norm(x, axis=1)
        ok &= xn > 1e-3
        xn = np.where(xn > 0, xn, 1.0)
        x = x / xn[:, None]
        y = np.cross(z, x)
        R = np.stack([x, y, z], axis=2)  # columns = camera axes in world coordinates
        Rt = np.concatenate([R, p[:, :, None]], axis=2).reshape(m, 12)
        Rt = Rt[ok]
        take = min(n - filled, Rt.shape[0])
        out[filled : filled + take] = Rt[:take].astype(np.float32)
        filled += take
    return out


CONFIGS = {
    # name: (map factory, n_viewpoints, width, height)
    "tiny": (map_tiny, 8, 64, 48),
    "small": (map_small, 16, 160, 120),
    "C1": (map_1m, 100, 320, 240),
    "C2": (map_1m, 1000, 640, 480),
    "C3": (map_10m, 10000, 640, 480),
    "C4": (lambda: map_1m(unknown_interior=True), 4000, 1280, 960),
}


# ---------------------------------------------------------------------------------------------
# Surface mesh (the "Poisson mesh" of the image-space normals)
# ---------------------------------------------------------------------------------------------

def _face_grid(origin, eu, ev, normal, nu: int, nv: int, seed: int, stream: int, bump: float, tilt: float):
    """(nu x nv) quads spanning origin + a*eu + b*ev: vertices bumped along the normal, smooth unit vertex normals
    tilted by up to `tilt`; two triangles per quad.  Returns corners (2*nu*nv, 3, 3) and normals of the same shape."""
    origin, eu, ev, normal = (np.asarray(a, dtype=np.float64) for a in (origin, eu, ev, normal))
    a = np.arange(nu + 1, dtype=np.float64) / nu
    b = np.arange(nv + 1, dtype=np.float64) / nv
    A, B = np.meshgrid(a, b, indexing="ij")
    u = uniform01(seed, 3 * (nu + 1) * (nv + 1), stream=stream).reshape(nu + 1, nv + 1, 3)
    P = origin + A[..., None] * eu + B[..., None] * ev + ((u[..., 0] - 0.5) * 2 * bump)[..., None] * normal
    tu, tv = eu / np.linalg.norm(eu), ev / np.linalg.norm(ev)
    N = normal + ((u[..., 1] - 0.5) * 2 * tilt)[..., None] * tu + ((u[..., 2] - 0.5) * 2 * tilt)[..., None] * tv
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    P = P.astype(np.float32)
    N = N.astype(np.float32)
    N /= np.sqrt((N.astype(np.float32) ** 2).sum(-1, keepdims=True, dtype=np.float32))
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    n00, n10, n01, n11 = N[:-1, :-1], N[1:, :-1], N[:-1, 1:], N[1:, 1:]
    tv_ = np.stack([np.stack([p00, p10, p11], axis=-2), np.stack([p00, p11, p01], axis=-2)], axis=-3)
    tn_ = np.stack([np.stack([n00, n10, n11], axis=-2), np.stack([n00, n11, n01], axis=-2)], axis=-3)
    return tv_.reshape(-1, 3, 3), tn_.reshape(-1, 3, 3)


def make_surface_mesh(leaves: OccupancyLeaves, cell: float = 1.0, seed: int = MAP_SEED + 7, bump: float = 0.04,
                      tilt: float = 0.25) -> Tuple[np.ndarray, np.ndarray]:
    """Triangle mesh over the outer surfaces of the map (ground top, building walls and roofs) with `cell`-sized
    quads: what a Poisson reconstruction of the scene would look like to the reference's normal renderer.
    Returns (tri_vertices, tri_normals), each (nf, 3, 3) float32, corner normals unit length."""
    r = leaves.resolution
    g = leaves.ground_half * r
    faces = []

    def add(origin, eu, ev, normal):
        nu = max(1, int(round(np.linalg.norm(eu) / cell)))
        nv = max(1, int(round(np.linalg.norm(ev) / cell)))
        faces.append(_face_grid(origin, eu, ev, normal, nu, nv, seed, len(faces), bump, tilt))

    add((-g, -g, 0.0), (2 * g, 0, 0), (0, 2 * g, 0), (0, 0, 1))
    for b in leaves.buildings:
        x0, y0, x1, y1, z1 = b.x0 * r, b.y0 * r, (b.x0 + b.nx) * r, (b.y0 + b.ny) * r, b.nz * r
        add((x0, y0, 0), (0, y1 - y0, 0), (0, 0, z1), (-1, 0, 0))
        add((x1, y0, 0), (0, y1 - y0, 0), (0, 0, z1), (1, 0, 0))
        add((x0, y0, 0), (x1 - x0, 0, 0), (0, 0, z1), (0, -1, 0))
        add((x0, y1, 0), (x1 - x0, 0, 0), (0, 0, z1), (0, 1, 0))
        add((x0, y0, z1), (x1 - x0, 0, 0), (0, y1 - y0, 0), (0, 0, 1))
    tv = np.ascontiguousarray(np.concatenate([f[0] for f in faces]), dtype=np.float32)
    tn = np.ascontiguousarray(np.concatenate([f[1] for f in faces]), dtype=np.float32)
    return tv, tn
