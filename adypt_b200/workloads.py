"""Deterministic synthetic inputs for the hot path: procedural OBJ/MTL scenes and
ray buffers. Everything is keyed by integer hashes (splitmix64) of (seed, index) -- no global RNG
state -- so the same arrays come out on every machine.

Scenes are written as Wavefront OBJ + MTL with ``%.9g`` floats (exact float32 round trip) because
OBJ is the only ingest format the reference has (``src/Util/Scene.cpp:9-136``). Every face carries an
explicit material: a missing material id is -1 in the reference and indexes out of bounds
(``Scene.cpp:52``), and tinyobj's default ``illum 0`` is a pass-through surface
(``pathtracer.glsl:144-201`` has no case for it).

This module is input generation for tests and ``bench.py``; it is not on the traced path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------------------------
# hashing


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def hash_u01(seed: int, *idx) -> np.ndarray:
    """float64 in [0,1) from (seed, idx...) with 53 random bits; idx arrays broadcast."""
    h = np.uint64(seed)
    h = splitmix64(np.asarray(h))
    for a in idx:
        with np.errstate(over="ignore"):
            h = splitmix64(h ^ (np.asarray(a).astype(np.uint64) * np.uint64(0xD6E8FEB86659FD93)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


# ----------------------------------------------------------------------------------------------
# scenes


@dataclass
class Material:
    name: str
    kd: tuple = (0.7, 0.7, 0.7)
    ke: tuple = (0.0, 0.0, 0.0)
    ks: tuple = (0.0, 0.0, 0.0)
    illum: int = 1
    ns: float = 10.0
    ni: float = 1.0
    d: float = 1.0


@dataclass
class SceneMesh:
    """Indexed triangle soup: verts (V,3) f32, faces (F,3) i32 (0-based), face_mat (F,) i32."""

    name: str
    verts: np.ndarray
    faces: np.ndarray
    face_mat: np.ndarray
    materials: list = field(default_factory=list)

    @property
    def n_tris(self) -> int:
        return int(self.faces.shape[0])

    def positions(self) -> np.ndarray:
        """(F,3,3) f32 vertex positions in face order == reference Triangle order."""
        return self.verts[self.faces]

    def write_obj(self, directory: str) -> str:
        """Write <name>.obj/.mtl into `directory`; returns the OBJ path. Skips if already there."""
        os.makedirs(directory, exist_ok=True)
        obj = os.path.join(directory, self.name + ".obj")
        mtl = os.path.join(directory, self.name + ".mtl")
        if os.path.exists(obj) and os.path.exists(mtl) and os.path.exists(obj + ".ok"):
            return obj
        with open(mtl, "w") as f:
            for m in self.materials:
                f.write(
                    f"newmtl {m.name}\nKd {m.kd[0]:.9g} {m.kd[1]:.9g} {m.kd[2]:.9g}\n"
                    f"Ke {m.ke[0]:.9g} {m.ke[1]:.9g} {m.ke[2]:.9g}\n"
                    f"Ks {m.ks[0]:.9g} {m.ks[1]:.9g} {m.ks[2]:.9g}\n"
                    f"Ns {m.ns:.9g}\nNi {m.ni:.9g}\nd {m.d:.9g}\nillum {m.illum}\n\n"
                )
        v = self.verts.astype(np.float32)
        with open(obj, "w") as f:
            f.write(f"mtllib {self.name}.mtl\n")
            vs = np.char.mod("%.9g", v.astype(np.float64))
            f.write("\n".join("v " + " ".join(r) for r in vs))
            f.write("\n")
            # one usemtl per run of equal material ids
            fm = self.face_mat
            starts = np.flatnonzero(np.r_[True, fm[1:] != fm[:-1]])
            ends = np.r_[starts[1:], fm.shape[0]]
            f1 = self.faces + 1
            for s, e in zip(starts, ends):
                f.write(f"usemtl {self.materials[int(fm[s])].name}\n")
                blk = f1[s:e]
                f.write("\n".join("f %d %d %d" % (a, b, c) for a, b, c in blk.tolist()))
                f.write("\n")
        open(obj + ".ok", "w").close()
        return obj


def _octa_sphere(level: int):
    """Octahedron subdivided `level` times, projected on the unit sphere. 8*4^level faces."""
    verts = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    faces = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
    verts = [np.array(v, dtype=np.float64) for v in verts]
    for _ in range(level):
        cache = {}
        nf = []

        def mid(a, b):
            k = (a, b) if a < b else (b, a)
            if k not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[k] = len(verts) - 1
            return cache[k]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca)]
        faces = nf
    return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int32)


def sphere_lattice(level: int = 5, name: str | None = None) -> SceneMesh:
    """C1: 8 subdivided-octahedron spheres on a 2x2x2 lattice (level 5 => 65 536 triangles)."""
    sv, sf = _octa_sphere(level)
    verts, faces, fmat = [], [], []
    k = 0
    for iz in (-1, 1):
        for iy in (-1, 1):
            for ix in (-1, 1):
                c = np.array([ix, iy, iz], dtype=np.float64) * 1.25
                faces.append(sf + k * sv.shape[0])
                verts.append(sv + c)
                fmat.append(np.full(sf.shape[0], k % 4, dtype=np.int32))
                k += 1
    mats = [
        Material("red", kd=(0.8, 0.2, 0.2)),
        Material("green", kd=(0.2, 0.8, 0.2)),
        Material("blue", kd=(0.2, 0.2, 0.8)),
        Material("grey", kd=(0.7, 0.7, 0.7)),
    ]
    return SceneMesh(
        name or f"sphere_lattice_l{level}",
        np.concatenate(verts).astype(np.float32),
        np.concatenate(faces).astype(np.int32),
        np.concatenate(fmat),
        mats,
    )


# box: 8 corners indexed by bits (x=1,y=2,z=4); 12 outward-facing triangles
_BOX_FACES = np.array(
    [
        (0, 2, 3), (0, 3, 1),  # z-
        (4, 5, 7), (4, 7, 6),  # z+
        (0, 1, 5), (0, 5, 4),  # y-
        (2, 6, 7), (2, 7, 3),  # y+
        (0, 4, 6), (0, 6, 2),  # x-
        (1, 3, 7), (1, 7, 5),  # x+
    ],
    dtype=np.int32,
)

CITY_MATERIALS = [
    Material("ground", kd=(0.55, 0.55, 0.5)),
    Material("wall_a", kd=(0.75, 0.7, 0.6)),
    Material("wall_b", kd=(0.45, 0.5, 0.6)),
    Material("wall_c", kd=(0.7, 0.35, 0.3)),
    Material("glossy", kd=(0.3, 0.3, 0.35), ks=(0.6, 0.6, 0.6), illum=2, ns=80.0),
    Material("mirror", kd=(0.0, 0.0, 0.0), ks=(0.9, 0.9, 0.9), illum=3),
    Material("glass", kd=(0.0, 0.0, 0.0), ks=(1.0, 1.0, 1.0), illum=7, ni=1.5),
    Material("lamp", kd=(0.8, 0.8, 0.8), ke=(6.0, 5.0, 4.0)),
]


def city(cells: int = 183, seed: int = 1, mixed_materials: bool = False, name: str | None = None) -> SceneMesh:
    """C2/C3/C4: axis-aligned boxes stacked 1-4 high on a jittered cells x cells grid + ground quad.

    cells=183 -> ~83.7k boxes (~1.0M triangles); cells=577 -> ~832k boxes (~10M triangles).
    mixed_materials=False: all boxes diffuse (C2/C4). True: adds glossy / mirror / glass / emissive
    boxes so every branch of pathtracer.glsl:144-201 is exercised (C3/C5).
    """
    ii, jj = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    ii = ii.ravel()
    jj = jj.ravel()
    cell_id = ii * cells + jj
    height = 1 + np.floor(hash_u01(seed, cell_id, 0) * 4.0).astype(np.int64)  # 1..4
    height = np.minimum(height, 4)
    # expand to one row per box
    rep_cell = np.repeat(cell_id, height)
    rep_i = np.repeat(ii, height)
    rep_j = np.repeat(jj, height)
    first = np.cumsum(height) - height
    level = np.arange(rep_cell.shape[0]) - np.repeat(first, height)
    cx = rep_i + 0.5 + (hash_u01(seed, rep_cell, 1 + 8 * level) - 0.5) * 0.3
    cz = rep_j + 0.5 + (hash_u01(seed, rep_cell, 2 + 8 * level) - 0.5) * 0.3
    shrink = 0.85 ** level
    hx = (0.22 + 0.16 * hash_u01(seed, rep_cell, 3 + 8 * level)) * shrink
    hz = (0.22 + 0.16 * hash_u01(seed, rep_cell, 4 + 8 * level)) * shrink
    storey = 0.45 + 0.3 * hash_u01(seed, rep_cell, 5)  # per-cell storey height
    y0 = level * storey
    y1 = (level + 1) * storey
    nb = rep_cell.shape[0]
    lo = np.stack([cx - hx, y0, cz - hz], axis=1)
    hi = np.stack([cx + hx, y1, cz + hz], axis=1)
    corners = np.empty((nb, 8, 3), dtype=np.float64)
    for c in range(8):
        for a in range(3):
            corners[:, c, a] = np.where((c >> a) & 1, hi[:, a], lo[:, a])
    verts = corners.reshape(-1, 3)
    faces = (_BOX_FACES[None, :, :] + (np.arange(nb, dtype=np.int64) * 8)[:, None, None]).reshape(-1, 3)
    if mixed_materials:
        r = hash_u01(seed, np.arange(nb), 77)
        mat = np.select(
            [r < 0.02, r < 0.06, r < 0.12, r < 0.20, r < 0.50, r < 0.78],
            [7, 6, 5, 4, 1, 2],
            default=3,
        ).astype(np.int32)
    else:
        mat = (1 + np.floor(hash_u01(seed, np.arange(nb), 77) * 3.0)).astype(np.int32)
        mat = np.minimum(mat, 3)
    fmat = np.repeat(mat, 12)
    # ground quad under everything (slightly larger than the grid)
    g0 = verts.shape[0]
    e = float(cells)
    ground = np.array([(-2.0, 0.0, -2.0), (e + 2.0, 0.0, -2.0), (e + 2.0, 0.0, e + 2.0), (-2.0, 0.0, e + 2.0)])
    verts = np.concatenate([verts, ground])
    faces = np.concatenate([faces, np.array([(g0, g0 + 2, g0 + 1), (g0, g0 + 3, g0 + 2)], dtype=np.int64)])
    fmat = np.concatenate([fmat, np.zeros(2, dtype=np.int32)])
    nm = name or f"city_c{cells}_s{seed}{'_mix' if mixed_materials else ''}"
    return SceneMesh(nm, verts.astype(np.float32), faces.astype(np.int32), fmat.astype(np.int32), list(CITY_MATERIALS))


def tiny_scene(kind: str) -> SceneMesh:
    """Hand-checkable scenes for known-answer tests."""
    m = [Material("m0", kd=(0.8, 0.8, 0.8))]
    if kind == "two_triangles":
        # the smallest scene the reference can build: with ONE triangle the SBVH root is a leaf and
        # WideBVHBuilder::fetch_children dereferences its left child -1 (WideBVHBuilder.cpp:93-108)
        v = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (3, 0, 1), (4, 0, 1), (3, 1, 1)], dtype=np.float32)
        f = np.array([(0, 1, 2), (3, 4, 5)], dtype=np.int32)
    elif kind == "shared_edge":
        # two coplanar triangles sharing the diagonal of the unit square in z=0
        v = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)], dtype=np.float32)
        f = np.array([(0, 1, 2), (0, 2, 3)], dtype=np.int32)
    elif kind == "strip":
        # 12 separated unit quads along x: forces inner nodes
        vs, fs = [], []
        for k in range(12):
            o = 2.0 * k
            b = len(vs)
            vs += [(o, 0, 0), (o + 1, 0, 0), (o + 1, 1, 0), (o, 1, 0)]
            fs += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
        v = np.array(vs, dtype=np.float32)
        f = np.array(fs, dtype=np.int32)
    elif kind == "deep":
        # 8x8x8 lattice of small tetrahedra faces: >= 3 wide-BVH levels
        vs, fs = [], []
        for x in range(8):
            for y in range(8):
                for z in range(8):
                    b = len(vs)
                    o = np.array([x, y, z], dtype=np.float64) * 1.5
                    vs += [tuple(o), tuple(o + (0.7, 0.1, 0.0)), tuple(o + (0.1, 0.8, 0.2)), tuple(o + (0.2, 0.2, 0.9))]
                    fs += [(b, b + 1, b + 2), (b, b + 2, b + 3), (b, b + 3, b + 1), (b + 1, b + 3, b + 2)]
        v = np.array(vs, dtype=np.float32)
        f = np.array(fs, dtype=np.int32)
    else:
        raise ValueError(kind)
    return SceneMesh("tiny_" + kind, v, f, np.zeros(f.shape[0], dtype=np.int32), m)


# ----------------------------------------------------------------------------------------------
# cameras per config (position, yaw, pitch, fov) -- chosen off-axis so no direction component is 0


def lattice_camera():
    return dict(position=(6.1, 4.3, 7.7), yaw=38.0, pitch=-24.0, fov=45.0)


def city_camera(cells: int = 183):
    e = float(cells)
    return dict(position=(e * 0.5 + 0.37, 14.0 + e * 0.06, e * 1.02 + 3.1), yaw=7.0, pitch=-31.0, fov=45.0)


# ----------------------------------------------------------------------------------------------
# incoherent bounce rays (C2/C4)


def _normalize(v):
    return v / np.sqrt((v * v).sum(axis=-1, keepdims=True))


def bounce_rays(positions: np.ndarray, primary_rays: np.ndarray, hit_tri: np.ndarray, hit_uv: np.ndarray,
                per_hit: int = 8, seed: int = 42, tmin: float = 1e-4) -> np.ndarray:
    """Spawn `per_hit` cosine-weighted rays about the geometric normal at every primary hit.

    positions: (F,3,3) f32 triangle corners in scene order; primary_rays: (N,8) f32
    (ox,oy,oz,tmin,dx,dy,dz,pad); hit_tri (N,) i32 scene ids (-1 = miss); hit_uv (N,2) f32.
    Returns (M*per_hit, 8) f32 in spawn order (hit-major), which is incoherent in direction.
    Origin = barycentric point u*p1 + v*p2 + (1-u-v)*p3, as pathtracer.glsl:82-85 does.
    """
    idx = np.flatnonzero(hit_tri >= 0)
    tri = hit_tri[idx]
    p = positions[tri].astype(np.float64)
    u = hit_uv[idx, 0].astype(np.float64)[:, None]
    v = hit_uv[idx, 1].astype(np.float64)[:, None]
    point = p[:, 0] * u + p[:, 1] * v + p[:, 2] * (1.0 - u - v)
    n = _normalize(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]))
    d_in = primary_rays[idx, 4:7].astype(np.float64)
    flip = (n * d_in).sum(axis=1) > 0
    n[flip] = -n[flip]
    m = idx.shape[0]
    ray_id = np.repeat(idx, per_hit)
    k = np.tile(np.arange(per_hit), m)
    r1 = hash_u01(seed, ray_id, 2 * k)
    r2 = hash_u01(seed, ray_id, 2 * k + 1)
    phi = 2.0 * np.pi * r1
    st = np.sqrt(r2)
    ct = np.sqrt(1.0 - r2)
    nn = np.repeat(n, per_hit, axis=0)
    helper = np.where((np.abs(nn[:, 0]) > 0.5)[:, None], np.array([0.0, 1.0, 0.0]), np.array([1.0, 0.0, 0.0]))
    t = _normalize(np.cross(helper, nn))
    b = np.cross(nn, t)
    d = t * (st * np.cos(phi))[:, None] + b * (st * np.sin(phi))[:, None] + nn * ct[:, None]
    d = _normalize(d)
    out = np.zeros((m * per_hit, 8), dtype=np.float32)
    out[:, 0:3] = np.repeat(point, per_hit, axis=0).astype(np.float32)
    out[:, 3] = np.float32(tmin)
    out[:, 4:7] = d.astype(np.float32)
    # never emit an exactly-zero direction component
    z = out[:, 4:7] == 0
    out[:, 4:7][z] = np.float32(1e-12)
    return out


def shadow_rays(positions: np.ndarray, hit_tri: np.ndarray, hit_uv: np.ndarray, tmin: float = 1e-4,
                sun=(0.6, 1.0, 0.2)) -> np.ndarray:
    """C4 any-hit rays: from every hit point towards normalize(sun) (pathtracer.glsl:132)."""
    idx = np.flatnonzero(hit_tri >= 0)
    p = positions[hit_tri[idx]].astype(np.float64)
    u = hit_uv[idx, 0].astype(np.float64)[:, None]
    v = hit_uv[idx, 1].astype(np.float64)[:, None]
    point = p[:, 0] * u + p[:, 1] * v + p[:, 2] * (1.0 - u - v)
    s = np.array(sun, dtype=np.float64)
    s = s / np.linalg.norm(s)
    out = np.zeros((idx.shape[0], 8), dtype=np.float32)
    out[:, 0:3] = point.astype(np.float32)
    out[:, 3] = np.float32(tmin)
    out[:, 4:7] = s.astype(np.float32)
    return out


def random_rays(n: int, lo, hi, seed: int = 5, tmin: float = 1e-4) -> np.ndarray:
    """Uniform origins in the box [lo,hi] with uniform directions on the sphere (stress rays)."""
    i = np.arange(n)
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    o = np.stack([hash_u01(seed, i, a) for a in range(3)], axis=1) * (hi - lo) + lo
    z = 2.0 * hash_u01(seed, i, 3) - 1.0
    phi = 2.0 * np.pi * hash_u01(seed, i, 4)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    out = np.zeros((n, 8), dtype=np.float32)
    out[:, 0:3] = o.astype(np.float32)
    out[:, 3] = np.float32(tmin)
    out[:, 4] = (r * np.cos(phi)).astype(np.float32)
    out[:, 5] = (r * np.sin(phi)).astype(np.float32)
    out[:, 6] = z.astype(np.float32)
    zz = out[:, 4:7] == 0
    out[:, 4:7][zz] = np.float32(1e-12)
    return out
