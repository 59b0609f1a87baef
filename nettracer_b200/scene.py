"""Host-side scene and camera containers (the part a NetTracer host would keep: scene set-up,
camera maths, image output).  They flatten to the row-major double arrays of `nt_scene_desc`
(include/nettracer_b200.h).  Rules: SPEC-PROVISIONAL.md §1-§2 — this repository's own spec;
the reference's scene classes are unknown (/root/reference/README:1-3 is all there is).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import abi


@dataclass
class Material:
    color: tuple = (1.0, 1.0, 1.0)
    ka: float = 0.1
    kd: float = 0.8
    ks: float = 0.0
    shininess: float = 1.0
    kr: float = 0.0
    kt: float = 0.0
    ior: float = 1.0

    def row(self):
        return [*self.color, self.ka, self.kd, self.ks, self.shininess, self.kr, self.kt, self.ior]


def _vec(v):
    a = np.asarray(v, dtype=np.float64)
    assert a.shape == (3,)
    return a


class Camera:
    """Look-at pinhole camera resolved on the host into eye/p00/dx/dy (SPEC-PROVISIONAL §2).
    All arithmetic here is numpy float64 in the order the spec writes it; `tan` is used only here."""

    def __init__(self, eye, at, up=(0.0, 1.0, 0.0), vfov_deg=45.0):
        self.eye, self.at, self.up, self.vfov_deg = _vec(eye), _vec(at), _vec(up), float(vfov_deg)

    def resolve(self, width: int, height: int) -> abi.nt_camera:
        def norm(v):
            return v * (1.0 / math.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))

        def cross(a, b):
            return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2],
                             a[0] * b[1] - a[1] * b[0]])

        w = norm(self.eye - self.at)
        u = norm(cross(self.up, w))
        v = cross(w, u)
        hh = math.tan(math.radians(self.vfov_deg) / 2.0)
        hw = hh * width / height
        p00 = -w - hw * u + hh * v
        dx = (2.0 * hw / width) * u
        dy = -(2.0 * hh / height) * v
        cam = abi.nt_camera()
        for name, val in (("eye", self.eye), ("p00", p00), ("dx", dx), ("dy", dy)):
            setattr(cam, name, (C.c_double * 3)(*[float(t) for t in val]))
        return cam


@dataclass
class Scene:
    """Flat scene: spheres, planes, triangles, materials, point lights."""
    spheres: list = field(default_factory=list)        # (cx, cy, cz, r)
    sphere_mat: list = field(default_factory=list)
    planes: list = field(default_factory=list)         # (nx, ny, nz, d), unit normal
    plane_mat: list = field(default_factory=list)
    triangles: object = field(default_factory=list)    # (9,) rows v0 v1 v2; list or ndarray [n,9]
    triangle_mat: object = field(default_factory=list)
    materials: list = field(default_factory=list)      # Material
    lights: list = field(default_factory=list)         # (px, py, pz, r, g, b)
    ambient: tuple = (1.0, 1.0, 1.0)
    background: tuple = (0.0, 0.0, 0.0)

    def add_material(self, m: Material) -> int:
        self.materials.append(m)
        return len(self.materials) - 1

    def add_sphere(self, c, r, mat):
        self.spheres.append((*map(float, c), float(r)))
        self.sphere_mat.append(int(mat))

    def add_plane(self, n, d, mat):
        n = _vec(n)
        ln = math.sqrt((float(n[0]) * float(n[0]) + float(n[1]) * float(n[1])) + float(n[2]) * float(n[2]))
        self.planes.append((*(n / ln), float(d) / ln))
        self.plane_mat.append(int(mat))

    def add_triangle(self, v0, v1, v2, mat):
        self.triangles.append((*map(float, v0), *map(float, v1), *map(float, v2)))
        self.triangle_mat.append(int(mat))

    def add_light(self, p, color=(1.0, 1.0, 1.0)):
        self.lights.append((*map(float, p), *map(float, color)))

    # ---- flattening ----
    def arrays(self) -> dict:
        def f64(x, cols):
            a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, cols))
            return a

        def i32(x):
            return np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))

        return {
            "spheres": f64(self.spheres, 4), "sphere_mat": i32(self.sphere_mat),
            "planes": f64(self.planes, 4), "plane_mat": i32(self.plane_mat),
            "triangles": f64(self.triangles, 9), "triangle_mat": i32(self.triangle_mat),
            "materials": f64([m.row() if isinstance(m, Material) else m for m in self.materials], 10),
            "lights": f64(self.lights, 6),
        }

    def to_desc(self):
        """Returns (nt_scene_desc, keepalive) — keepalive owns the numpy buffers."""
        a = self.arrays()
        assert len(a["sphere_mat"]) == len(a["spheres"])
        assert len(a["plane_mat"]) == len(a["planes"])
        assert len(a["triangle_mat"]) == len(a["triangles"])
        d = abi.nt_scene_desc()
        d.struct_size = C.sizeof(abi.nt_scene_desc)
        d.n_spheres, d.n_planes, d.n_triangles = len(a["spheres"]), len(a["planes"]), len(a["triangles"])
        d.n_materials, d.n_lights = len(a["materials"]), len(a["lights"])
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        for k in ("spheres", "planes", "triangles", "materials", "lights"):
            setattr(d, k, a[k].ctypes.data_as(pd))
        for k in ("sphere_mat", "plane_mat", "triangle_mat"):
            setattr(d, k, a[k].ctypes.data_as(pi))
        d.ambient = (C.c_double * 3)(*map(float, self.ambient))
        d.background = (C.c_double * 3)(*map(float, self.background))
        return d, a


def make_params(width, height, spp, max_depth, camera: abi.nt_camera, precision=abi.NT_F64_STRICT,
                ray_epsilon=0.0, shard_index=0, shard_count=1, band_rows=16,
                layout=abi.NT_LAYOUT_FULL, flags=0) -> abi.nt_render_params:
    p = abi.nt_render_params()
    p.struct_size = C.sizeof(abi.nt_render_params)
    p.width, p.height, p.spp, p.max_depth = int(width), int(height), int(spp), int(max_depth)
    p.precision = int(precision)
    p.ray_epsilon = float(ray_epsilon)
    p.camera = camera
    p.shard_index, p.shard_count, p.band_rows = int(shard_index), int(shard_count), int(band_rows)
    p.layout = int(layout)
    p.flags = int(flags)
    return p


def shard_rows(height, band_rows, shard_index, shard_count) -> int:
    """Pure-host twin of nt_shard_rows (tests check they agree)."""
    nb = (height + band_rows - 1) // band_rows
    return sum(min(b * band_rows + band_rows, height) - b * band_rows
               for b in range(shard_index, nb, shard_count))


def owned_rows(height, band_rows, shard_index, shard_count) -> np.ndarray:
    """Image rows owned by a shard, in compact (virtual-row) order."""
    ys = np.arange(height)
    return ys[(ys // band_rows) % shard_count == shard_index]


def write_ppm(path, rgba: np.ndarray):
    h, w = rgba.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgba[:, :, :3]).tobytes())
