"""CPU tests of the host-side logic: camera resolution, scene flattening, sharding bookkeeping."""
import math

import numpy as np

from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import deinterleave_host
from nettracer_b200.scene import Camera, Material, Scene, make_params, owned_rows, shard_rows


def test_camera_resolve_matches_spec():
    cam = Camera((0, 0, 5), (0, 0, 0), (0, 1, 0), vfov_deg=90.0)
    c = cam.resolve(200, 100)
    # w = (0,0,1), u = (1,0,0), v = (0,1,0); hh = tan(45deg), hw = 2*hh
    hh = math.tan(math.radians(90.0) / 2)
    assert np.allclose(list(c.p00), [-2 * hh, hh, -1.0])
    assert np.allclose(list(c.dx), [2 * (2 * hh) / 200, 0, 0])
    assert np.allclose(list(c.dy), [0, -2 * hh / 100, 0])
    assert list(c.eye) == [0, 0, 5]
    # centre ray points at the target
    D = np.array(c.p00) + np.array(c.dx) * 100 + np.array(c.dy) * 50
    assert np.allclose(D / np.linalg.norm(D), [0, 0, -1])


def test_scene_flattening_and_generators_are_deterministic():
    s1, _ = scenes.cornell_box()
    s2, _ = scenes.cornell_box()
    a1, a2 = s1.arrays(), s2.arrays()
    for k in a1:
        assert np.array_equal(a1[k], a2[k])
    assert a1["spheres"].shape == (8, 4) and a1["planes"].shape == (6, 4) and a1["lights"].shape == (2, 6)
    assert np.allclose(np.linalg.norm(a1["planes"][:, :3], axis=1), 1.0)
    d, keep = s1.to_desc()
    assert d.n_spheres == 8 and d.n_planes == 6 and d.n_triangles == 0 and d.n_lights == 2
    assert d.spheres[3] == a1["spheres"][0, 3]
    # spheres do not overlap and stay inside the room
    c, r = a1["spheres"][:, :3], a1["spheres"][:, 3]
    for i in range(8):
        for j in range(i + 1, 8):
            assert np.linalg.norm(c[i] - c[j]) > r[i] + r[j]
    m, cam = scenes.spheres_and_mesh(n_spheres=100, mesh_n=16)
    am = m.arrays()
    assert am["triangles"].shape == (2 * 16 * 16, 9) and am["spheres"].shape == (100, 4)
    assert len(scenes.terrain_mesh(708)) == 1_002_528  # configs[3]: the "1M-triangle mesh"


def test_params_and_shard_bookkeeping():
    cam = Camera((0, 1, 5), (0, 0, 0))
    p = make_params(64, 48, 4, 3, cam.resolve(64, 48), shard_index=1, shard_count=3, band_rows=5)
    assert p.struct_size == 152 and p.flags == 0 and p.width == 64 and p.shard_count == 3
    h, band, n = 48, 5, 3
    rows = [owned_rows(h, band, i, n) for i in range(n)]
    assert sorted(np.concatenate(rows).tolist()) == list(range(h))
    assert [len(r) for r in rows] == [shard_rows(h, band, i, n) for i in range(n)]
    # deinterleave puts every compact row back where it belongs
    frame = np.arange(h * 7 * 4, dtype=np.uint32).astype(np.uint8).reshape(h, 7, 4)
    parts = [frame[r] for r in rows]
    assert np.array_equal(deinterleave_host(parts, h, 7, band), frame)


def test_material_row_order():
    m = Material((0.1, 0.2, 0.3), ka=0.4, kd=0.5, ks=0.6, shininess=7, kr=0.8, kt=0.9, ior=1.5)
    assert m.row() == [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 7, 0.8, 0.9, 1.5]
    s = Scene()
    assert s.add_material(m) == 0 and s.arrays()["materials"].shape == (1, 10)
    assert abi.FLOPS["sphere_tests"] == 17


def test_plane_quotient_shortcut_never_rejects_a_hit():
    """nt_trace.cuh plane_below_eps, restated in numpy binary64 (same IEEE operations): with eps_lo = fl(eps (1 - 2^-50))
    made as nt_api.cu makes it, `|num| <= fl(|dn| eps_lo)` with a normal product must imply that the correctly rounded
    quotient fails `t > eps` - probed where it matters, numerators within +-40 ulps of |dn| * eps, over the whole
    exponent range, for ordinary, huge and absurdly small epsilons (the shortcut is switched off below 1e-290)."""
    import numpy as np
    rng = np.random.default_rng(5)
    k = 1.0 - 8.8817841970012523e-16
    rejected = 0
    for eps in (1e-6, 1e-4, 1e-7, 1e-9, 3.3e-12, 1e-290, 5e-291, 0.5, 7.0, 1e-300):
        eps_lo = eps * k if eps >= 1e-290 else 0.0
        for _ in range(4):
            n = 200000
            dn = rng.normal(size=n) * 10.0 ** rng.uniform(-300, 300, size=n)
            f = 1.0 + rng.integers(-40, 40, size=n) * 2.0 ** -53
            num = np.abs(dn) * eps * f * np.where(rng.random(n) < 0.5, 1, -1)
            num = np.where(rng.random(n) < 0.1, np.nextafter(num, np.inf), num)
            with np.errstate(all="ignore"):
                lo = np.abs(dn) * eps_lo
                reject = (np.abs(num) <= lo) & (lo >= 2.2250738585072014e-308)
                hit = num / dn > eps
            assert not (reject & hit).any()
            rejected += int(reject.sum())
    assert rejected > 1_000_000  # the probe does sit on the boundary
