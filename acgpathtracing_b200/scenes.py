"""Synthetic scenes of the shapes BASELINE.json names (the reference ships no scene assets).

* ``cornell32()``        -- the 32-triangle Cornell box of SURVEY.md Appendix E (configs C1, C2, C5)
* ``material_scene()``   -- Cornell box + dielectric icosphere + conductor sphere (config C3)
* ``tessellated_scene()``-- Cornell box + ~2 M-triangle displaced height field (config C4)

Every scene is produced both as the four arrays the reference's ``TinyObjWrapper`` hands to the renderer
(``getVerticesFloat`` xyzw, ``getIndexBuffer``, ``getMaterialIndices``, ``getMaterials``;
TinyObjWrapper.cpp:138-244) and -- for the small ones -- as OBJ/MTL text with a fixed vertex order, so that
the loader path can be exercised and primitive ids stay stable.  The camera and area light are the
reference's hard-coded ones (PathTracerMain.cpp:154-158, 228-233).
"""
from dataclasses import dataclass, field

import numpy as np

from .types import BSDF_DIFFUSE, BSDF_METALLIC, BSDF_REFRACTION, Material, Float3

# PathTracerMain.cpp:228-233
CAMERA_EYE = (278.0, 273.0, -900.0)
CAMERA_LOOKAT = (278.0, 273.0, 330.0)
CAMERA_UP = (0.0, 1.0, 0.0)
CAMERA_FOVY = 35.0
# PathTracerMain.cpp:154-158
LIGHT_EMISSION = (10.0, 10.0, 10.0)
LIGHT_CORNER = (343.0, 547.0, 227.0)
LIGHT_V1 = (0.0, 0.0, 105.0)
LIGHT_V2 = (-130.0, 0.0, 0.0)


@dataclass
class SceneArrays:
    verts: np.ndarray            # (nV, 4) float32, w = 1   (TinyObjWrapper.cpp:200-221)
    indices: np.ndarray          # (nT, 3) uint32           (:229-244)
    mat_ids: np.ndarray          # (nT,)  uint32            (:177-192)
    materials: list              # list of dict(name, Kd, Ke, Ni, Pr, Pm)
    obj_text: str = ""
    mtl_text: str = ""
    name: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def n_tri(self):
        return int(self.indices.shape[0])

    def material_structs(self):
        """ctypes Material[nM] with the reference's name -> BSDF rule (TinyObjWrapper.cpp:153-164)."""
        arr = (Material * len(self.materials))()
        for i, m in enumerate(self.materials):
            arr[i].diffuse = Float3(*m["Kd"])
            arr[i].emission = Float3(*m["Ke"])
            arr[i].roughness = m.get("Pr", 0.0)
            arr[i].metallic = m.get("Pm", 0.0)
            arr[i].ior = m.get("Ni", 1.0)
            arr[i].bsdfType = bsdf_from_name(m["name"])
        return arr


def bsdf_from_name(name: str) -> int:
    if "Refractive" in name:
        return BSDF_REFRACTION
    if "Metallic" in name:
        return BSDF_METALLIC
    return BSDF_DIFFUSE


class _Builder:
    def __init__(self):
        self.v = []
        self.f = []          # (a, b, c, material index) 0-based
        self.mats = []
        self.groups = []     # (name, first face)

    def material(self, name, Kd, Ke=(0, 0, 0), Ni=1.0, Pr=0.0, Pm=0.0):
        self.mats.append(dict(name=name, Kd=tuple(Kd), Ke=tuple(Ke), Ni=Ni, Pr=Pr, Pm=Pm))
        return len(self.mats) - 1

    def group(self, name):
        self.groups.append((name, len(self.f)))

    def quad(self, p0, p1, p2, p3, mat):
        b = len(self.v)
        self.v += [p0, p1, p2, p3]
        self.f += [(b, b + 1, b + 2, mat), (b, b + 2, b + 3, mat)]

    def mesh(self, verts, tris, mat):
        b = len(self.v)
        self.v += [tuple(float(x) for x in p) for p in verts]
        self.f += [(b + int(t[0]), b + int(t[1]), b + int(t[2]), mat) for t in tris]

    def block(self, top_xz, height, mat):
        (x0, z0), (x1, z1), (x2, z2), (x3, z3) = top_xz
        h = height
        self.quad((x0, h, z0), (x1, h, z1), (x2, h, z2), (x3, h, z3), mat)   # top
        c = [(x0, z0), (x1, z1), (x2, z2), (x3, z3)]
        for i in range(4):
            (ax, az), (bx, bz) = c[i], c[(i + 1) % 4]
            self.quad((ax, 0.0, az), (ax, h, az), (bx, h, bz), (bx, 0.0, bz), mat)

    def finish(self, name, with_text=True):
        verts = np.ones((len(self.v), 4), dtype=np.float32)
        verts[:, :3] = np.asarray(self.v, dtype=np.float32).reshape(-1, 3)
        idx = np.asarray([t[:3] for t in self.f], dtype=np.uint32).reshape(-1, 3)
        mat_ids = np.asarray([t[3] for t in self.f], dtype=np.uint32)
        obj = mtl = ""
        if with_text:
            mtl_lines = []
            for m in self.mats:
                mtl_lines += [f"newmtl {m['name']}", "Kd %.9g %.9g %.9g" % m["Kd"], "Ke %.9g %.9g %.9g" % m["Ke"],
                              "Ni %.9g" % m["Ni"], "Pr %.9g" % m["Pr"], "Pm %.9g" % m["Pm"], ""]
            mtl = "\n".join(mtl_lines)
            lines = [f"mtllib {name}.mtl"]
            lines += ["v %.9g %.9g %.9g" % (float(p[0]), float(p[1]), float(p[2])) for p in verts[:, :3]]
            starts = {first: gname for gname, first in self.groups}
            cur = None
            for k, (a, b, c, m) in enumerate(self.f):
                if k in starts:
                    lines.append(f"o {starts[k]}")
                    cur = None
                if m != cur:
                    lines.append(f"usemtl {self.mats[m]['name']}")
                    cur = m
                lines.append(f"f {a + 1} {b + 1} {c + 1}")
            obj = "\n".join(lines) + "\n"
        return SceneArrays(verts, idx, mat_ids, list(self.mats), obj, mtl, name)


def _cornell_walls(b: _Builder):
    white = b.material("white", (0.8, 0.8, 0.8))
    green = b.material("green", (0.05, 0.8, 0.05))
    red = b.material("red", (0.8, 0.05, 0.05))
    light = b.material("light", (0.78, 0.78, 0.78), Ke=(10.0, 10.0, 10.0))
    X, Y, Z = 556.0, 548.8, 559.2
    b.group("floor")
    b.quad((0, 0, 0), (X, 0, 0), (X, 0, Z), (0, 0, Z), white)
    b.group("ceiling")
    b.quad((0, Y, 0), (0, Y, Z), (X, Y, Z), (X, Y, 0), white)
    b.group("back")
    b.quad((0, 0, Z), (X, 0, Z), (X, Y, Z), (0, Y, Z), white)
    b.group("green_wall")          # x = 0
    b.quad((0, 0, 0), (0, 0, Z), (0, Y, Z), (0, Y, 0), green)
    b.group("red_wall")            # x = 556
    b.quad((X, 0, 0), (X, Y, 0), (X, Y, Z), (X, 0, Z), red)
    return white, green, red, light


def _cornell_light(b: _Builder, light):
    b.group("light")
    b.quad((343.0, 548.6, 227.0), (343.0, 548.6, 332.0), (213.0, 548.6, 332.0), (213.0, 548.6, 227.0), light)


SHORT_BLOCK = [(130.0, 65.0), (82.0, 225.0), (242.0, 274.0), (290.0, 114.0)]
TALL_BLOCK = [(423.0, 247.0), (265.0, 296.0), (314.0, 455.0), (472.0, 406.0)]


def cornell32() -> SceneArrays:
    """32 triangles: 5 wall quads, two 10-triangle blocks, one emissive quad."""
    b = _Builder()
    white, green, red, light = _cornell_walls(b)
    b.group("short_block")
    b.block(SHORT_BLOCK, 165.0, white)
    b.group("tall_block")
    b.block(TALL_BLOCK, 330.0, white)
    _cornell_light(b, light)
    s = b.finish("cornell32")
    assert s.n_tri == 32
    return s


def icosphere(center, radius, subdivisions):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache = {}
        nf = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b_, c in f:
            ab, bc, ca = mid(a, b_), mid(b_, c), mid(c, a)
            nf += [(a, ab, ca), (b_, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    verts = (np.asarray(v) * radius + np.asarray(center, dtype=np.float64)).astype(np.float32)
    return verts, np.asarray(f, dtype=np.uint32)


def material_scene(subdivisions: int = 3) -> SceneArrays:
    """Cornell box with a dielectric sphere (material name contains "Refractive", Ni 1.5) on the short block
    and a conductor sphere ("Metallic") on the floor -- exercises all three BSDF branches
    (pathTracerPrograms.cu:905-983)."""
    b = _Builder()
    white, green, red, light = _cornell_walls(b)
    b.group("short_block")
    b.block(SHORT_BLOCK, 165.0, white)
    b.group("tall_block")
    b.block(TALL_BLOCK, 330.0, white)
    glass = b.material("GlassRefractive", (1.0, 1.0, 1.0), Ni=1.5)
    metal = b.material("MirrorMetallic", (0.95, 0.93, 0.88), Pr=0.2, Pm=1.0)
    gv, gf = icosphere((186.0, 165.0 + 82.0, 169.5), 80.0, subdivisions)
    b.group("glass_sphere")
    b.mesh(gv, gf, glass)
    mv, mf = icosphere((420.0, 90.5, 130.0), 90.0, subdivisions)
    b.group("metal_sphere")
    b.mesh(mv, mf, metal)
    _cornell_light(b, light)
    return b.finish("material_scene")


def _tea4(v0, v1):
    """cuda/random.h:32-46 on numpy uint32 arrays (used to derive displacements without libm RNG)."""
    v0 = v0.astype(np.uint32).copy()
    v1 = v1.astype(np.uint32).copy()
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(4):
            s0 = np.uint32((int(s0) + 0x9E3779B9) & 0xFFFFFFFF)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s0) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s0) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0


def tessellated_scene(grid: int = 1000) -> SceneArrays:
    """Cornell walls + light + a (grid x grid)-cell displaced height field (2*grid^2 triangles; grid=1000 ->
    2.0 M) standing on the floor.  Heights come from a few octaves of value noise whose lattice values are
    tea<4>(i, j) (no libm RNG), so the mesh is reproducible bit for bit."""
    b = _Builder()
    white, green, red, light = _cornell_walls(b)
    _cornell_light(b, light)
    base = b.finish("tess_base", with_text=False)
    n = grid + 1
    ii, jj = np.meshgrid(np.arange(n, dtype=np.uint32), np.arange(n, dtype=np.uint32), indexing="ij")
    h = np.zeros((n, n), dtype=np.float64)
    amp = 120.0
    for octave, cells in enumerate((4, 16, 64, 256)):
        step = max(grid // cells, 1)
        li, lj = ii // step, jj // step
        fi = (ii % step) / step
        fj = (jj % step) / step

        def lat(a, b_):
            return _tea4(a.astype(np.uint32) + np.uint32(1000 * octave), b_.astype(np.uint32)).astype(np.float64) / 2.0 ** 32

        si = fi * fi * (3 - 2 * fi)
        sj = fj * fj * (3 - 2 * fj)
        v00, v10, v01, v11 = lat(li, lj), lat(li + 1, lj), lat(li, lj + 1), lat(li + 1, lj + 1)
        h += amp * ((v00 * (1 - si) + v10 * si) * (1 - sj) + (v01 * (1 - si) + v11 * si) * sj)
        amp *= 0.45
    x = 40.0 + ii.astype(np.float64) * (476.0 / grid)
    z = 60.0 + jj.astype(np.float64) * (460.0 / grid)
    y = 20.0 + h
    hv = np.ones((n * n, 4), dtype=np.float32)
    hv[:, 0] = x.reshape(-1)
    hv[:, 1] = y.reshape(-1)
    hv[:, 2] = z.reshape(-1)
    ci, cj = np.meshgrid(np.arange(grid, dtype=np.uint32), np.arange(grid, dtype=np.uint32), indexing="ij")
    a = (ci * n + cj).reshape(-1)
    bq = a + 1
    c = a + n + 1
    d = a + n
    tris = np.empty((2 * grid * grid, 3), dtype=np.uint32)
    tris[0::2] = np.stack([a, bq, c], axis=1)
    tris[1::2] = np.stack([a, c, d], axis=1)
    off = np.uint32(base.verts.shape[0])
    verts = np.concatenate([base.verts, hv], axis=0)
    indices = np.concatenate([base.indices, tris + off], axis=0)
    mat_ids = np.concatenate([base.mat_ids, np.zeros(tris.shape[0], dtype=np.uint32)])
    return SceneArrays(verts, indices, mat_ids, base.materials, "", "", f"tessellated_{grid}")
