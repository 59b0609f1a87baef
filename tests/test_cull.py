"""CPU tests of the flat-scene culling tables (nettracer_b200/csrc/nt_cull.cpp, exported for inspection as
nt_cull_tables): every table must be CONSERVATIVE — a primitive a query can hit is always in the mask the
query would look up.  Checked here by brute force in numpy float64 against the exact SPEC-PROVISIONAL §3
tests, with the light-buffer cell picked in emulated binary32 exactly as the device does (lbuf_mask in
nt_trace.cuh).  The bit-exact GPU parity tests (test_parity_gpu.py, culling on and off) are the end-to-end proof."""
import numpy as np
import pytest

from nettracer_b200 import scenes
from nettracer_b200.renderer import cull_tables, light_rooms, plane_free_lights, primary_rects
from nettracer_b200.scene import Camera, Material, Scene, make_params

EPS = 1e-6


def lbuf_cell(Lv, k):
    """Mirror of lbuf_mask(): direction light -> P = -Lv, binary32."""
    v = (-Lv).astype(np.float32)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    ax, ay, az = np.abs(x), np.abs(y), np.abs(z)
    fx = (ax >= ay) & (ax >= az)
    fy = ~fx & (ay >= az)
    m = np.where(fx, ax, np.where(fy, ay, az))
    comp = np.where(fx, x, np.where(fy, y, z))
    a = np.where(fx, y, x)
    b = np.where(fx | fy, z, y)
    face = np.where(fx, 0, np.where(fy, 2, 4)) + (comp < 0)
    half = np.float32(0.5 * k)
    sc = half / m
    iu = np.clip((a * sc + half).astype(np.int64), 0, k - 1)
    iv = np.clip((b * sc + half).astype(np.int64), 0, k - 1)
    return face, iv, iu


def sphere_hits(o, d, sph, tmax):
    """SPEC §3 sphere rule, vectorised: hit with eps < t < tmax."""
    oc = o - sph[:3]
    b = (oc * d).sum(-1)
    cc = (oc * oc).sum(-1) - sph[3] * sph[3]
    disc = b * b - cc
    sq = np.sqrt(np.maximum(disc, 0))
    t = -b - sq
    t = np.where(t > EPS, t, -b + sq)
    return (disc >= 0) & (t > EPS) & (t < tmax), t


def tri_hits(o, d, tri, tmax):
    v0, e1, e2 = tri[0:3], tri[3:6] - tri[0:3], tri[6:9] - tri[0:3]
    p = np.cross(d, e2)
    det = (e1 * p).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / det
        tv = o - v0
        u = (tv * p).sum(-1) * inv
        q = np.cross(tv, e1)
        v = (d * q).sum(-1) * inv
        t = (q * e2).sum(-1) * inv
    return (det != 0) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > EPS) & (t < tmax), t


def surface_points(scene, rng, n):
    """Points where shadow queries start: on spheres, on triangles, on the planes' region, and free space."""
    a = scene.arrays()
    pts = [rng.uniform(-12, 12, (n, 3))]
    for s in a["spheres"]:
        v = rng.normal(size=(n // 4, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts.append(s[:3] + s[3] * v)
    for t in a["triangles"]:
        w = rng.dirichlet((1, 1, 1), n // 8)
        pts.append(w @ t.reshape(3, 3))
    return np.concatenate(pts)


@pytest.mark.parametrize("which", ["cornell", "mixed1", "mixed2"])
def test_light_buffer_is_conservative(which):
    scene = scenes.cornell_box()[0] if which == "cornell" else scenes.random_mixed(12, 3, 20, seed=1 if which == "mixed1" else 2)[0]
    a = scene.arrays()
    tab = cull_tables(scene)
    k, ns = tab["k"], len(a["spheres"])
    rng = np.random.default_rng(7)
    P = surface_points(scene, rng, 4000)
    culled = total = 0
    for l, light in enumerate(a["lights"]):
        Lv = light[:3] - P
        dist = np.sqrt((Lv * Lv).sum(-1))
        L = Lv / dist[:, None]
        face, iv, iu = lbuf_cell(Lv, k)
        masks = tab["lbuf"][l, face, iv, iu]
        for j in range(ns + len(a["triangles"])):
            hit, _ = (sphere_hits(P, L, a["spheres"][j], dist) if j < ns else tri_hits(P, L, a["triangles"][j - ns], dist))
            inmask = ((masks >> np.uint64(j)) & np.uint64(1)).astype(bool)
            assert not (hit & ~inmask).any(), f"light {l} primitive {j}: occluder missing from the light buffer"
            culled += int((~inmask).sum())
            total += len(P)
    assert culled > 0.5 * total, "the light buffer should cull most candidates on these scenes"


@pytest.mark.parametrize("which", ["cornell", "mixed1"])
def test_neighbour_masks_are_conservative(which):
    scene = scenes.cornell_box()[0] if which == "cornell" else scenes.random_mixed(12, 3, 20, seed=1)[0]
    a = scene.arrays()
    tab = cull_tables(scene)
    ns = len(a["spheres"])
    rng = np.random.default_rng(11)
    n = 3000
    for i, s in enumerate(a["spheres"]):
        assert not (int(tab["nbr"][i]) >> i) & 1, "own bit must be excluded"
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        o = s[:3] + s[3] * v                       # on sphere i
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        own, t_own = sphere_hits(o, d, s, np.inf)  # rays that hit sphere i again: chords
        for j in range(ns + len(a["triangles"])):
            if j == i:
                continue
            hit, t = (sphere_hits(o, d, a["spheres"][j], np.inf) if j < ns else tri_hits(o, d, a["triangles"][j - ns], np.inf))
            nearer = own & hit & (t <= t_own)
            if nearer.any():
                assert (int(tab["nbr"][i]) >> j) & 1, f"sphere {i}: primitive {j} stops a chord but is not a neighbour"


def test_bounding_spheres_contain_primitives():
    scene = scenes.random_mixed(12, 3, 20, seed=3)[0]
    a = scene.arrays()
    tab = cull_tables(scene)
    ns = len(a["spheres"])
    assert np.array_equal(tab["bsph"][:ns], a["spheres"])
    for j, t in enumerate(a["triangles"]):
        c, r = tab["bsph"][ns + j, :3], tab["bsph"][ns + j, 3]
        assert (np.linalg.norm(t.reshape(3, 3) - c, axis=1) <= r).all()


def test_not_eligible_is_an_error():
    from nettracer_b200.lib import NetTracerError
    scene = scenes.random_mixed(150, 2, 300, seed=4)[0]  # > 64 bounded primitives -> BVH scene
    with pytest.raises(NetTracerError):
        cull_tables(scene)


def test_light_inside_a_sphere_sees_it_everywhere():
    """A light inside a primitive's (dilated) ball: every direction cell must carry that primitive."""
    scene = scenes.cornell_box()[0]
    c = scene.spheres[2]
    scene.add_light((c[0], c[1] + 0.1 * c[3], c[2]), (1.0, 1.0, 1.0))
    tab = cull_tables(scene)
    l = len(scene.lights) - 1
    assert ((tab["lbuf"][l] >> np.uint64(2)) & np.uint64(1)).all()
    # and the other lights' tables are unchanged by it
    ref = cull_tables(scenes.cornell_box()[0])
    assert np.array_equal(tab["lbuf"][:l], ref["lbuf"])


def test_too_many_lights_disables_the_tables():
    from nettracer_b200.lib import NetTracerError
    scene = scenes.cornell_box()[0]
    for i in range(20):
        scene.add_light((0.1 * i, 9.0, 1.0), (0.1, 0.1, 0.1))
    with pytest.raises(NetTracerError):
        cull_tables(scene)  # > NT_CULL_MAX_LIGHTS: nt_scene_create then simply renders without culling


# ---- primary rays: per-primitive pixel rectangles (nt_cull_primary_rects) ----
def _primary_dirs(cam, w, h, n):
    """Directions of all n x n samples of every pixel: D = p00 + (px + (i + .5) / n) dx + (py + (j + .5) / n) dy, plus
    the pixel corners (a superset of the sample positions)."""
    c = cam.resolve(w, h)
    eye, p00, dx, dy = (np.array(list(getattr(c, k))) for k in ("eye", "p00", "dx", "dy"))
    offs = np.concatenate([(np.arange(n) + 0.5) / n, [0.0, 1.0]])
    fx = (np.arange(w)[:, None] + offs[None, :]).reshape(-1)          # [w * (n + 2)]
    fy = (np.arange(h)[:, None] + offs[None, :]).reshape(-1)
    D = p00[None, None, :] + fx[None, :, None] * dx[None, None, :] + fy[:, None, None] * dy[None, None, :]
    px = np.repeat(np.arange(w), len(offs))
    py = np.repeat(np.arange(h), len(offs))
    return eye, D, px, py


def _line_touches_ball(eye, D, ball):
    """Does the forward half-line eye + t D, t > 0, touch the ball?  Exact rule in float64 with a tiny slack."""
    V = ball[:3] - eye
    dd = (D * D).sum(-1)
    vd = (D * V).sum(-1)
    dist2 = (V * V).sum() - vd * vd / dd
    return (dist2 <= ball[3] * ball[3] * (1 + 1e-9)) & ((vd > 0) | ((V * V).sum() <= ball[3] * ball[3]))


CAMERAS = {
    "cornell": None,  # the scene's own camera
    "close": Camera(eye=(0.3, 0.2, 3.0), at=(0.0, 0.0, 0.0), vfov_deg=70.0),          # big outlines, some partly off-screen
    "side": Camera(eye=(9.0, 4.0, 2.0), at=(0.0, 0.0, -3.0), up=(0.1, 1, 0), vfov_deg=35.0),
    "inside": Camera(eye=(0.0, 0.0, 0.0), at=(0.0, 0.0, -1.0), vfov_deg=100.0),        # may sit inside a ball; balls behind the eye
}


@pytest.mark.parametrize("which,camname", [("cornell", "cornell"), ("cornell", "close"), ("mixed1", "cornell"),
                                           ("mixed1", "side"), ("mixed2", "inside"), ("mixed2", "close")])
def test_primary_rects_are_conservative(which, camname):
    scene, cam = scenes.cornell_box() if which == "cornell" else scenes.random_mixed(12, 3, 20, seed=1 if which == "mixed1" else 2)
    cam = CAMERAS[camname] or cam
    w, h, n = 224, 128, 2
    rects = primary_rects(scene, make_params(w, h, n * n, 1, cam.resolve(w, h))).astype(np.int64)
    balls = cull_tables(scene)["bsph"]
    eye, D, px, py = _primary_dirs(cam, w, h, n)
    kept = 0
    for j, ball in enumerate(balls):
        touch = _line_touches_ball(eye, D, ball)                      # [rows, cols] of sample positions
        x0, x1, y0, y1 = rects[j]
        inside = ((py >= y0) & (py <= y1))[:, None] & ((px >= x0) & (px <= x1))[None, :]
        assert not (touch & ~inside).any(), f"primitive {j}: a primary ray touches its ball outside its rectangle {rects[j]}"
        kept += int(inside.sum())
    assert kept < 0.6 * len(balls) * D.shape[0] * D.shape[1] or camname == "inside", "the rectangles should cull most of the image"


def test_primary_rects_are_tight_enough():
    """A ball in the middle of the view: its rectangle is its outline plus a few pixels, not the whole image."""
    scene, cam = scenes.cornell_box()
    w, h = 320, 180
    rects = primary_rects(scene, make_params(w, h, 1, 1, cam.resolve(w, h))).astype(np.int64)
    balls = cull_tables(scene)["bsph"]
    eye, D, px, py = _primary_dirs(cam, w, h, 1)
    for j, ball in enumerate(balls):
        touch = _line_touches_ball(eye, D, ball)
        if not touch.any():
            continue
        ys, xs = np.nonzero(touch)
        x0, x1, y0, y1 = rects[j]
        assert x0 >= max(px[xs].min() - 6, 0) - 1 and x1 <= min(px[xs].max() + 6, w - 1) + 1
        assert y0 >= max(py[ys].min() - 6, 0) - 1 and y1 <= min(py[ys].max() + 6, h - 1) + 1


# ---- shadow queries from bounded primitives: lights no plane can hide (nt_cull_plane_free_lights) ----
def _plane_occludes(P, light, planes):
    """SPEC §3 plane rule for the segment P -> light: eps < t < dist for any plane."""
    Lv = light[:3] - P
    dist = np.sqrt((Lv * Lv).sum(-1))
    L = Lv / dist[:, None]
    occ = np.zeros(len(P), dtype=bool)
    for pl in planes:
        dn = (L * pl[:3]).sum(-1)
        num = pl[3] - (P * pl[:3]).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = num / dn
        occ |= (dn != 0) & (t > EPS) & (t < dist)
    return occ


def _points_on_bounded(scene, rng, n):
    a = scene.arrays()
    pts = []
    for s in a["spheres"]:
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts.append(s[:3] + s[3] * v)
    for t in a["triangles"]:
        pts.append(rng.dirichlet((1, 1, 1), n) @ t.reshape(3, 3))
    return np.concatenate(pts)


def test_plane_free_lights_cornell():
    scene = scenes.cornell_box()[0]
    a = scene.arrays()
    mask = plane_free_lights(scene)
    assert mask == (1 << len(a["lights"])) - 1, "both lights of the box sit inside the room with all spheres"
    P = _points_on_bounded(scene, np.random.default_rng(3), 2000)
    for light in a["lights"]:
        assert not _plane_occludes(P, light, a["planes"]).any()


def test_plane_free_lights_is_conservative():
    """Random scenes and a box with one light moved behind a wall: a set bit must mean that the exact rule never finds
    an occluding plane from any point of a bounded primitive; an unset bit is always allowed."""
    rng = np.random.default_rng(11)
    cases = [scenes.random_mixed(12, 3, 20, seed=sd)[0] for sd in (1, 2, 3)]
    box = scenes.cornell_box()[0]
    l0 = box.lights[0]
    box.lights[0] = (l0[0], l0[1] + 100.0, l0[2], l0[3], l0[4], l0[5])  # far above the ceiling
    cases.append(box)
    seen_clear = False
    for scene in cases:
        a = scene.arrays()
        mask = plane_free_lights(scene)
        P = _points_on_bounded(scene, rng, 500)
        for l, light in enumerate(a["lights"]):
            occ = _plane_occludes(P, light, a["planes"])
            if (mask >> l) & 1:
                assert not occ.any(), f"light {l} is flagged plane-free but a plane hides it from {int(occ.sum())} points"
            else:
                seen_clear = True
    assert seen_clear, "at least one light of the random scenes should have a plane in the way"


def test_primary_rects_fuzz():
    """Random balls (radius 3e-5 ... 300, any distance, also around or behind the eye), random cameras (5 ... 150 degrees,
    tilted), small odd images: no sample position may see a ball outside its rectangle."""
    rng = np.random.default_rng(2024)
    checked = 0
    for _ in range(150):
        s = Scene()
        m = s.add_material(Material())
        for _j in range(int(rng.integers(1, 6))):
            scale = 10.0 ** rng.uniform(-2, 2)
            s.add_sphere(tuple(rng.normal(size=3) * scale * 3), float(10.0 ** rng.uniform(-2.5, 0.5) * scale), m)
        s.add_light((0, 50, 0))
        cam = Camera(tuple(rng.normal(size=3) * 10.0 ** rng.uniform(-1, 1.5)), tuple(rng.normal(size=3) * 3),
                     up=tuple(rng.normal(size=3)), vfov_deg=float(rng.uniform(5, 150)))
        w, h = int(rng.integers(8, 90)), int(rng.integers(8, 70))
        rects = primary_rects(s, make_params(w, h, 4, 1, cam.resolve(w, h))).astype(np.int64)
        eye, D, px, py = _primary_dirs(cam, w, h, 2)
        for j, ball in enumerate(cull_tables(s)["bsph"]):
            touch = _line_touches_ball(eye, D, ball)
            x0, x1, y0, y1 = rects[j]
            inside = ((py >= y0) & (py <= y1))[:, None] & ((px >= x0) & (px <= x1))[None, :]
            assert not (touch & ~inside).any(), (j, rects[j], ball, w, h)
            checked += 1
    assert checked > 300


def test_plane_free_lights_fuzz():
    """Spheres resting on, hovering over or sunk into a floor (by 1e-13 ... 2), walls, lights on either side: a flagged
    light must never be hidden by a plane from any point of a sphere - checked with the exact rule at the smallest
    ray epsilon the proof covers (1e-7), densely around the points nearest to every plane."""
    rng = np.random.default_rng(77)
    eps = 1e-7
    flagged = 0
    for _ in range(250):
        s = Scene()
        m = s.add_material(Material())
        n = np.array([0.0, 1.0, 0.0])
        if rng.random() < 0.4:
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
        d0 = float(rng.uniform(-3, 3))
        s.add_plane(tuple(n), d0, m)
        if rng.random() < 0.5:
            s.add_plane((1, 0, 0), -8.0, m)
            s.add_plane((-1, 0, 0), -8.0, m)
        for _j in range(int(rng.integers(1, 5))):
            r = float(10 ** rng.uniform(-1, 0.5))
            sink = float(rng.choice([0, 0, 1e-13, 1e-10, 1e-8, 1e-6, 1e-3, -1e-9, -0.5, -2]))
            base = rng.normal(size=3) * 3
            base -= n * (base @ n - d0)
            s.add_sphere(tuple(base + n * (r - sink)), r, m)
        for _l in range(2):
            lp = rng.normal(size=3) * 6
            lp = lp - n * ((lp @ n) - d0) + n * (rng.uniform(0.5, 9) if rng.random() < 0.8 else rng.uniform(-3, 0.01))
            s.add_light(tuple(lp))
        a = s.arrays()
        mask = plane_free_lights(s)
        pts = []
        for sp in a["spheres"]:
            v = rng.normal(size=(200, 3))
            pts.append(sp[:3] + sp[3] * v / np.linalg.norm(v, axis=1, keepdims=True))
            for pl in a["planes"]:
                nn = pl[:3] / np.linalg.norm(pl[:3])
                for sg in (-1.0, 1.0):
                    vv = sg * nn[None, :] + rng.normal(size=(40, 3)) * 10 ** rng.uniform(-9, -1)
                    pts.append(sp[:3] + sp[3] * vv / np.linalg.norm(vv, axis=1, keepdims=True))
                    pts.append((sp[:3] + sg * sp[3] * nn)[None, :])
        P = np.concatenate(pts)
        for l, light in enumerate(a["lights"]):
            if not (mask >> l) & 1:
                continue
            flagged += 1
            Lv = light[:3] - P
            dist = np.sqrt((Lv * Lv).sum(-1))
            L = Lv / dist[:, None]
            for pl in a["planes"]:
                dn = (L * pl[:3]).sum(-1)
                num = pl[3] - (P * pl[:3]).sum(-1)
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = num / dn
                assert not ((dn != 0) & (t > eps) & (t < dist)).any(), (a["spheres"], pl, light)
    assert flagged >= 10


def test_diagnostic_entry_points_refuse_ineligible_scenes_and_bad_arguments():
    """nt_primary_rects / nt_plane_free_lights follow nt_cull_tables: NT_ERR_INVALID for scenes without culling tables
    (no bounded primitive, > 64 of them), for a bad image size and for NULL outputs - never a crash."""
    import ctypes as C
    from nettracer_b200 import abi
    from nettracer_b200.lib import NetTracerError, load
    scene, cam = scenes.cornell_box()
    p = make_params(64, 48, 1, 1, cam.resolve(64, 48))
    L = load()
    desc, keep = scene.to_desc()
    assert L.nt_primary_rects(C.byref(desc), C.byref(p), None) == abi.NT_ERR_INVALID
    assert L.nt_primary_rects(C.byref(desc), None, None) == abi.NT_ERR_INVALID
    assert L.nt_plane_free_lights(C.byref(desc), None) == abi.NT_ERR_INVALID
    bad = make_params(0, 48, 1, 1, cam.resolve(64, 48))
    with pytest.raises(NetTracerError):
        primary_rects(scene, bad)
    planes_only = Scene()
    m = planes_only.add_material(Material())
    planes_only.add_plane((0, 1, 0), 0.0, m)
    planes_only.add_light((0, 5, 0))
    big = scenes.random_mixed(60, 1, 30, seed=3)[0]  # 90 bounded primitives: a BVH scene
    for sc in (planes_only, big):
        with pytest.raises(NetTracerError):
            primary_rects(sc, p)
        with pytest.raises(NetTracerError):
            plane_free_lights(sc)


def _axis_plane_occludes_exact(P, light, planes, eps):
    """SPEC §3 / §4 in their exact operation order (binary64, no FMA): shadow ray from P towards the light; does any plane
    with a normal of exactly +-e_k give eps < t < dist?"""
    Lv = light[:3] - P
    d2 = (Lv[:, 0] * Lv[:, 0] + Lv[:, 1] * Lv[:, 1]) + Lv[:, 2] * Lv[:, 2]
    dist = np.sqrt(d2)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dist
        L = Lv * inv[:, None]
        occ = np.zeros(len(P), dtype=bool)
        for pl in planes:
            n = pl[:3]
            if not (np.abs(n).max() == 1.0 and np.count_nonzero(n) == 1):
                continue
            dn = (n[0] * L[:, 0] + n[1] * L[:, 1]) + n[2] * L[:, 2]
            num = pl[3] - ((n[0] * P[:, 0] + n[1] * P[:, 1]) + n[2] * P[:, 2])
            t = num / dn
            occ |= (dn != 0) & (t > eps) & (t < dist)
    return occ, dist


def _room_samples(room, light, rng, n):
    """Points that stress the room test: uniform inside, exactly on every face, within a few ulps and within the dilation
    of every face (both sides), far away along the faces, next to the light."""
    lo, hi = room[0:6:2].copy(), room[1:6:2].copy()
    span = np.where(np.isfinite(hi - lo), hi - lo, 50.0)
    lo_f = np.where(np.isfinite(lo), lo, light[:3] - 25.0)
    hi_f = np.where(np.isfinite(hi), hi, light[:3] + 25.0)
    P = lo_f + (hi_f - lo_f) * rng.random((n, 3))
    out = [P, light[:3] + rng.normal(size=(n // 4, 3)) * 1e-3]
    for k in range(3):
        for face in (lo[k], hi[k]):
            if not np.isfinite(face):
                continue
            for off in (0.0, 1e-16, -1e-16, 3e-15, -3e-15, 1e-12, -1e-12, 1e-10 * span[k], -1e-10 * span[k], 1e-9, -1e-9, 1e-7, -1e-7):
                Q = lo_f + (hi_f - lo_f) * rng.random((n // 8, 3))
                Q[:, k] = face + off
                out.append(Q)
                R_ = Q.copy()  # grazing: far along the face, so that the direction component towards the plane is tiny
                R_[:, (k + 1) % 3] = light[(k + 1) % 3] + rng.choice([-1, 1], n // 8) * 10.0 ** rng.uniform(0, 4, n // 8)
                out.append(R_)
    return np.concatenate(out)


@pytest.mark.parametrize("eps", [1e-6, 1e-7, 1e-3, 1e-9])
def test_light_rooms_are_conservative(eps):
    """nt_light_rooms: whenever the kernel's test passes (P inside the box, dist <= min(eps * cap_per_eps, cap_max)) the
    exact rule must not find an axis-aligned plane between P and the light - the Cornell box, rooms with lights near a
    wall, open rooms (missing walls), several planes per axis, lights outside the box and lights ON a plane."""
    rng = np.random.default_rng(5)
    cases = [scenes.cornell_box()[0]]
    for _ in range(40):
        s = Scene()
        m = s.add_material(Material())
        for k in range(3):
            for _j in range(int(rng.integers(0, 4))):
                n = [0.0, 0.0, 0.0]
                n[k] = float(rng.choice([-1.0, 1.0]))
                s.add_plane(tuple(n), float(np.round(rng.uniform(-20, 20), int(rng.integers(0, 4)))), m)
        if rng.random() < 0.3:
            v = rng.normal(size=3)
            s.add_plane(tuple(v / np.linalg.norm(v)), 30.0, m)  # a general plane: never part of a room
        s.add_sphere((0, 0, 0), 1.0, m)
        for _l in range(3):
            lp = rng.uniform(-25, 25, size=3)
            if rng.random() < 0.3 and len(s.planes):
                pl = s.planes[int(rng.integers(len(s.planes)))]
                k = int(np.argmax(np.abs(pl[:3])))
                if abs(pl[k]) == 1.0:
                    lp[k] = pl[k] * pl[3] + float(rng.choice([0.0, 1e-12, 1e-9, 1e-6, 1e-3, 0.5]))
            s.add_light(tuple(lp))
        cases.append(s)
    passed = 0
    for scene in cases:
        a = scene.arrays()
        rooms = light_rooms(scene)
        assert rooms.shape == (len(a["lights"]), 8)
        for l, light in enumerate(a["lights"]):
            room = rooms[l]
            P = _room_samples(room, light, rng, 4000)
            occ, dist = _axis_plane_occludes_exact(P, light, a["planes"], eps)
            cap = min(eps * room[6], room[7])
            inside = ((P[:, 0] >= room[0]) & (P[:, 0] <= room[1]) & (P[:, 1] >= room[2]) & (P[:, 1] <= room[3]) &
                      (P[:, 2] >= room[4]) & (P[:, 2] <= room[5]) & (dist <= cap))
            assert not (inside & occ).any(), (l, light, room, P[inside & occ][:3])
            passed += int(inside.sum())
    assert passed > 100000, "the rooms must accept the bulk of the inside points, or the test proves nothing"


def test_light_rooms_cornell_values():
    scene = scenes.cornell_box()[0]
    rooms = light_rooms(scene)
    d = 16.0 * 2.0 ** -33
    np.testing.assert_allclose(rooms[0, :6], [-6 - d, 6 + d, 0 - d, 10 + d, -8 - d, 16 + d], rtol=0, atol=1e-15)
    assert rooms[0, 6] == pytest.approx(0.5 * 0.8 / d, rel=1e-12) and rooms[0, 7] == pytest.approx(0.8e12, rel=1e-12)
    # hit points of a 1080p frame lie well inside the caps: the longest shadow ray of the box is ~18 units
    assert 1e-6 * rooms[:, 6].min() > 30.0
