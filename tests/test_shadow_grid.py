"""Shadow grids of BVH scenes (nettracer_b200/csrc/nt_shadowgrid.h, diagnostic nt_shadow_grid): the lists must be CONSERVATIVE -
every sphere that touches the segment from a point to the light is listed in the cell the DEVICE computes for that point
(binary32 arithmetic, emulated here in numpy) - because shadow queries test only the listed spheres instead of walking the
sphere set of the tree.  CPU only: the diagnostic builds the same lists nt_scene_create uploads."""
import numpy as np
import pytest

from nettracer_b200 import scenes
from nettracer_b200.renderer import shadow_grid
from nettracer_b200.scene import Material, Scene


def device_cell(params, K, P):
    """nt_bvh_trace.cuh shadow_query_start, operation for operation in binary32.  Returns the cell index or -1."""
    f = np.float32
    L, ax, U, V = params[0:3], params[3:6], params[6:9], params[9:12]
    u0, v0, su, sv = params[12:16]
    d = P.astype(f) - L
    w = f(f(d[0] * ax[0]) + f(d[1] * ax[1])) + f(d[2] * ax[2])
    if not w > 0:
        return -1
    iw = f(1.0) / f(w)
    fu = f(f(f(f(f(d[0] * U[0]) + f(d[1] * U[1])) + f(d[2] * U[2])) * iw - u0) * su)
    fv = f(f(f(f(f(d[0] * V[0]) + f(d[1] * V[1])) + f(d[2] * V[2])) * iw - v0) * sv)
    if not (fu >= 0 and fv >= 0 and fu < K and fv < K):
        return -1
    return int(fv) * K + int(fu)


def spheres_touching_segment(sph, P, L):
    """Exact geometry in binary64: indices of the balls whose distance to the segment P..L is <= their radius."""
    c, r = sph[:, :3], sph[:, 3]
    d = L - P
    t = np.clip(((c - P) @ d) / (d @ d), 0.0, 1.0)
    q = P + t[:, None] * d
    return np.nonzero(np.linalg.norm(c - q, axis=1) <= r)[0]


def check_scene(s, n_points, rng, expect_grid=None):
    sph = np.asarray(s.spheres, dtype=np.float64).reshape(-1, 4)
    lo, hi = (sph[:, :3] - sph[:, 3:4]).min(axis=0), (sph[:, :3] + sph[:, 3:4]).max(axis=0)
    n_checked = n_listed = 0
    for l, light in enumerate(np.asarray(s.lights, dtype=np.float64).reshape(-1, 6)):
        g = shadow_grid(s, l)
        if expect_grid is not None:
            assert (g is not None) == expect_grid[l], (l, g is None)
        if g is None:
            continue
        params, K, off, items = g
        assert off[0] == 0 and off[-1] == items.size and np.all(np.diff(off.astype(np.int64)) >= 0)
        Lp = light[:3]
        # points: uniformly around the scene, and - the hard cases - points just behind spheres as seen from the light
        pts = [rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo)) for _ in range(n_points // 2)]
        for _ in range(n_points - len(pts)):
            j = int(rng.integers(len(sph)))
            u = rng.normal(size=3)
            u /= np.linalg.norm(u)
            edge = sph[j, :3] + sph[j, 3] * rng.uniform(0.9, 1.0) * u  # a point of ball j, mostly near its silhouette
            pts.append(Lp + (edge - Lp) * rng.uniform(1.0, 1.6))      # ... pushed away from the light along the same line
        for P in pts:
            hit = spheres_touching_segment(sph, P, Lp)
            cell = device_cell(params, K, P)
            listed = set() if cell < 0 else set(items[off[cell]:off[cell + 1]].tolist())
            assert set(hit.tolist()) <= listed, (l, P, sorted(set(hit.tolist()) - listed), cell)
            n_checked += 1
            n_listed += len(listed)
    return n_checked, n_listed


def test_shadow_grid_conservative_random_clouds():
    rng = np.random.default_rng(5)
    total = 0
    for trial in range(6):
        s = Scene()
        m = s.add_material(Material((0.5, 0.5, 0.5)))
        ext = float(10 ** rng.uniform(0, 2))
        for _ in range(int(rng.integers(80, 600))):
            s.add_sphere(tuple(rng.uniform(-ext, ext, 3) * np.array([1.0, rng.uniform(0.1, 1.0), 1.0])), float(rng.uniform(0.002, 0.08) * ext), m)
        for _ in range(3):
            v = rng.normal(size=3)
            s.add_light(tuple(v / np.linalg.norm(v) * ext * rng.uniform(2.5, 6.0)), (0.5, 0.5, 0.5))
        n, _ = check_scene(s, 400, rng, expect_grid=[True, True, True])
        total += n
    assert total == 6 * 3 * 400


def test_shadow_grid_configs3_slab_is_tight_and_conservative():
    """configs[3]'s own sphere slab and lights: conservative, and a cell lists a handful of spheres, not hundreds."""
    s, _ = scenes.spheres_and_mesh(n_spheres=10_000, mesh_n=8)
    rng = np.random.default_rng(9)
    n, listed = check_scene(s, 600, rng, expect_grid=[True, True])
    assert listed / n < 12, listed / n
    _, K, off, items = shadow_grid(s, 0)
    assert K == 512 and items.size < 60 * 10_000


def test_shadow_grid_refused_for_lights_inside_or_beside_the_cloud():
    """A light inside the sphere cloud, or so close beside it that a ball reaches behind its plane or projects at a grazing
    angle, gets no grid (K = 0): its shadow queries walk the whole tree."""
    rng = np.random.default_rng(2)
    s = Scene()
    m = s.add_material(Material((0.5, 0.5, 0.5)))
    for _ in range(300):
        s.add_sphere(tuple(rng.uniform(-10, 10, 3)), 0.3, m)
    s.add_light((0.5, 0.2, -0.3), (1, 1, 1))      # inside
    s.add_light((10.5, 0.0, 0.0), (1, 1, 1))      # on the cloud's face: balls at more than 83 degrees from the axis
    s.add_light((60.0, 5.0, 0.0), (1, 1, 1))      # well outside
    assert shadow_grid(s, 0) is None and shadow_grid(s, 1) is None and shadow_grid(s, 2) is not None
    check_scene(s, 300, rng)
    with pytest.raises(Exception):
        shadow_grid(s, 3)
