"""Synthetic scenes for BASELINE.json configs[1..4]: seeded with 0x4E54, fully
deterministic (own splitmix64, no dependence on numpy's generators).  configs[0] (the reference's
bundled scene) is BLOCKED: /root/reference holds no scene."""
from __future__ import annotations

import math

import numpy as np

from .scene import Camera, Material, Scene

SEED = 0x4E54
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M64

    def u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self, lo=0.0, hi=1.0):
        return lo + (hi - lo) * ((self.u64() >> 11) * (1.0 / (1 << 53)))


def _splitmix_array(seed, n):
    """Vectorised splitmix64 stream -> float64 in [0,1)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def cornell_box(seed=SEED):
    """configs[1]/[2]: box of 6 axis-aligned planes, 8 spheres, 2 point lights.
    Spheres: 3 mirrors, 3 glass, 2 glossy-diffuse; the floor is slightly reflective.
    Returns (Scene, Camera)."""
    rng = SplitMix64(seed)
    s = Scene(ambient=(1.0, 1.0, 1.0), background=(0.02, 0.02, 0.03))
    white = s.add_material(Material((0.75, 0.75, 0.75), ka=0.08, kd=0.85))
    red = s.add_material(Material((0.75, 0.15, 0.15), ka=0.08, kd=0.85))
    green = s.add_material(Material((0.15, 0.75, 0.15), ka=0.08, kd=0.85))
    floor = s.add_material(Material((0.6, 0.6, 0.65), ka=0.08, kd=0.7, ks=0.2, shininess=40.0, kr=0.2))
    mirror = s.add_material(Material((0.9, 0.9, 0.95), ka=0.02, kd=0.15, ks=0.6, shininess=120.0, kr=0.75))
    glass = s.add_material(Material((0.95, 0.98, 1.0), ka=0.0, kd=0.05, ks=0.5, shininess=200.0,
                                    kr=0.1, kt=0.85, ior=1.5))
    glossy = [s.add_material(Material((0.2, 0.35, 0.85), ka=0.1, kd=0.7, ks=0.4, shininess=30.0)),
              s.add_material(Material((0.9, 0.7, 0.2), ka=0.1, kd=0.7, ks=0.4, shininess=60.0))]
    # room: x in [-6,6], y in [0,10], z in [-8,16]; dot(n,p)=d with inward-facing normals
    s.add_plane((0, 1, 0), 0.0, floor)
    s.add_plane((0, -1, 0), -10.0, white)
    s.add_plane((1, 0, 0), -6.0, red)
    s.add_plane((-1, 0, 0), -6.0, green)
    s.add_plane((0, 0, 1), -8.0, white)
    s.add_plane((0, 0, -1), -16.0, white)
    mats = [mirror, glass, glossy[0], mirror, glass, glossy[1], mirror, glass]
    # 8 non-overlapping spheres on a jittered 4x2 layout, resting on or floating above the floor
    k = 0
    for row in range(2):
        for col in range(4):
            r = rng.uniform(0.9, 1.35)
            cx = -4.2 + col * 2.8 + rng.uniform(-0.25, 0.25)
            cz = -3.5 + row * 4.5 + rng.uniform(-0.4, 0.4)
            cy = r + (rng.uniform(0.0, 2.5) if (k % 3) == 1 else 0.0)
            s.add_sphere((cx, cy, cz), r, mats[k])
            k += 1
    s.add_light((-3.0, 9.2, 4.0), (0.65, 0.62, 0.6))
    s.add_light((3.5, 8.8, -2.0), (0.45, 0.47, 0.5))
    cam = Camera(eye=(0.0, 5.0, 15.0), at=(0.0, 3.2, 0.0), up=(0, 1, 0), vfov_deg=42.0)
    return s, cam


def terrain_mesh(n, extent=100.0, seed=SEED):
    """Displaced-grid mesh of 2*n*n triangles over [-extent/2, extent/2]^2 (y up)."""
    g = np.linspace(-extent / 2, extent / 2, n + 1)
    X, Z = np.meshgrid(g, g, indexing="xy")
    noise = _splitmix_array(seed ^ 0xA5A5, (n + 1) * (n + 1)).reshape(n + 1, n + 1)
    Y = (3.0 * np.sin(X * 0.11) * np.cos(Z * 0.13) + 1.5 * np.sin(X * 0.37 + 1.0) * np.sin(Z * 0.29)
         + 0.6 * np.cos((X + Z) * 0.71) + (extent / n) * 0.35 * (noise - 0.5))
    P = np.stack([X, Y, Z], axis=-1)
    a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    t1 = np.concatenate([a, c, b], axis=-1).reshape(-1, 9)
    t2 = np.concatenate([b, c, d], axis=-1).reshape(-1, 9)
    tris = np.empty((2 * n * n, 9))
    tris[0::2], tris[1::2] = t1, t2
    return tris


def spheres_and_mesh(n_spheres=10_000, mesh_n=708, seed=SEED, extent=100.0):
    """configs[3]/[4]: n_spheres random spheres + a 2*mesh_n^2-triangle terrain (1,002,528 at the
    default), two lights, one ground plane far below as a backstop.  Returns (Scene, Camera)."""
    s = Scene(ambient=(1.0, 1.0, 1.0), background=(0.35, 0.5, 0.75))
    ground = s.add_material(Material((0.45, 0.55, 0.35), ka=0.1, kd=0.8, ks=0.1, shininess=10.0))
    rock = s.add_material(Material((0.55, 0.5, 0.45), ka=0.1, kd=0.8, ks=0.15, shininess=20.0, kr=0.05))
    pal = [s.add_material(Material((0.85, 0.25, 0.2), ka=0.1, kd=0.7, ks=0.4, shininess=50.0)),
           s.add_material(Material((0.9, 0.9, 0.95), ka=0.02, kd=0.2, ks=0.6, shininess=120.0, kr=0.7)),
           s.add_material(Material((0.95, 0.98, 1.0), ka=0.0, kd=0.05, ks=0.5, shininess=200.0,
                                   kr=0.1, kt=0.85, ior=1.45)),
           s.add_material(Material((0.2, 0.4, 0.85), ka=0.1, kd=0.7, ks=0.3, shininess=30.0))]
    s.add_plane((0, 1, 0), -12.0, ground)
    tris = terrain_mesh(mesh_n, extent, seed)
    s.triangles = tris
    s.triangle_mat = np.full(len(tris), rock, dtype=np.int32)
    u = _splitmix_array(seed ^ 0x5A5A, n_spheres * 5).reshape(n_spheres, 5)
    half = extent / 2 * 0.92
    cx = -half + 2 * half * u[:, 0]
    cz = -half + 2 * half * u[:, 1]
    cy = 6.0 + 22.0 * u[:, 2]
    r = 0.12 + 0.45 * u[:, 3]
    s.spheres = np.stack([cx, cy, cz, r], axis=1)
    s.sphere_mat = np.asarray(pal, dtype=np.int32)[(u[:, 4] * len(pal)).astype(np.int64) % len(pal)]
    s.add_light((-40.0, 80.0, 30.0), (0.75, 0.72, 0.68))
    s.add_light((55.0, 60.0, -20.0), (0.35, 0.37, 0.42))
    cam = Camera(eye=(0.0, 34.0, 78.0), at=(0.0, 6.0, 0.0), up=(0, 1, 0), vfov_deg=40.0)
    return s, cam


def random_mixed(n_spheres, n_planes, n_triangles, n_lights=2, seed=1, glassy=True):
    """Small random scenes for parity tests (all three primitive kinds, all material features)."""
    rng = SplitMix64(seed)
    s = Scene(ambient=(0.9, 0.95, 1.0), background=(0.1, 0.12, 0.2))
    mats = []
    for i in range(6):
        kr = rng.uniform(0.0, 0.6) if glassy and i % 3 == 1 else 0.0
        kt = rng.uniform(0.3, 0.8) if glassy and i % 3 == 2 else 0.0
        mats.append(s.add_material(Material(
            (rng.uniform(0.1, 1), rng.uniform(0.1, 1), rng.uniform(0.1, 1)), ka=rng.uniform(0, 0.2),
            kd=rng.uniform(0.3, 0.9), ks=rng.uniform(0, 0.6) if i % 2 else 0.0,
            shininess=rng.uniform(2, 80), kr=kr, kt=kt, ior=rng.uniform(1.1, 1.8))))
    for i in range(n_spheres):
        s.add_sphere((rng.uniform(-6, 6), rng.uniform(-3, 5), rng.uniform(-8, 2)), rng.uniform(0.3, 1.6),
                     mats[rng.u64() % len(mats)])
    for i in range(n_planes):
        n = (rng.uniform(-0.3, 0.3), 1.0, rng.uniform(-0.3, 0.3)) if i == 0 else \
            (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.2, 1))
        s.add_plane(n, -4.0 - 3.0 * i, mats[rng.u64() % len(mats)])
    for i in range(n_triangles):
        c = np.array([rng.uniform(-6, 6), rng.uniform(-3, 5), rng.uniform(-8, 2)])
        vs = [c + np.array([rng.uniform(-1.5, 1.5) for _ in range(3)]) for _ in range(3)]
        s.add_triangle(*vs, mats[rng.u64() % len(mats)])
    for i in range(n_lights):
        s.add_light((rng.uniform(-8, 8), rng.uniform(6, 12), rng.uniform(0, 10)),
                    (rng.uniform(0.3, 0.7),) * 3)
    cam = Camera(eye=(0.5, 1.5, 12.0), at=(0.0, 0.5, -2.0), up=(0, 1, 0), vfov_deg=50.0)
    return s, cam


def mirror_field(n_spheres=3000, seed=7):
    """Parity-test scene: thousands of small mirror and glass spheres that see each other at distances of many radii,
    so that ray trees run along mirror chains.  SPEC-PROVISIONAL section 4 does not re-normalise secondary directions;
    along such chains |d| drifts from 1 (every bounce off a small, distant sphere amplifies the drift) and the sphere
    rule of section 3 then accepts points off the sphere - any acceleration structure has to stay conservative for
    that (nt_bvh_trace.cuh query_start, oracle ray_grow).  > 64 bounded primitives: the BVH path."""
    rng = SplitMix64(seed)
    s = Scene(ambient=(1.0, 1.0, 1.0), background=(0.3, 0.4, 0.6))
    mirror = s.add_material(Material((0.9, 0.9, 0.95), ka=0.02, kd=0.1, ks=0.6, shininess=120.0, kr=0.9))
    glass = s.add_material(Material((0.95, 0.98, 1.0), ka=0.0, kd=0.05, ks=0.5, shininess=200.0, kr=0.1, kt=0.85, ior=1.5))
    matte = s.add_material(Material((0.7, 0.5, 0.3), ka=0.1, kd=0.8, ks=0.2, shininess=20.0, kr=0.3))
    s.add_plane((0, 1, 0), -22.0, matte)
    for i in range(n_spheres):
        c = (rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-45, -5))
        s.add_sphere(c, rng.uniform(0.4, 0.9), [mirror, mirror, mirror, glass][i % 4])
    for i in range(8):
        c = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-45, -5)])
        vs = [c + np.array([rng.uniform(-5, 5) for _ in range(3)]) for _ in range(3)]
        s.add_triangle(*vs, mirror)
    s.add_light((-20.0, 60.0, 30.0), (0.7, 0.7, 0.65))
    s.add_light((40.0, 30.0, 20.0), (0.4, 0.42, 0.5))
    cam = Camera(eye=(0.0, 2.0, 20.0), at=(0.0, 0.0, -25.0), up=(0, 1, 0), vfov_deg=50.0)
    return s, cam


CONFIGS = {
    # name: (scene factory, width, height, spp, max_depth) — BASELINE.json configs[1..4]
    "cfg2_cornell_1080p_1spp_d1": (cornell_box, 1920, 1080, 1, 1),
    "cfg3_cornell_1080p_4spp_d5": (cornell_box, 1920, 1080, 4, 5),
    "cfg4_mesh1m_4k_4spp_d3": (spheres_and_mesh, 3840, 2160, 4, 3),
    "cfg5_mesh1m_8k_16spp_d5": (spheres_and_mesh, 7680, 4320, 16, 5),
}
