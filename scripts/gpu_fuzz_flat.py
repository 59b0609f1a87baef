"""Ad-hoc parity fuzz of the flat kernels (development aid, not part of the suite): random axis-aligned rooms (0-3 planes per
axis, closed or open), occasional general planes and triangles, glass / mirror spheres, 1-4 lights inside, outside, on and near
walls, random cameras and ray epsilons - strict image and counters against the oracle, fast image within a loose tolerance.
   python scripts/gpu_fuzz_flat.py [n_scenes] [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from nettracer_b200 import abi
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import Camera, Material, Scene, make_params
from oracle import oracle

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
only = set(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else None  # scene indices to render (the others are only generated)
KEYS = ["rays_primary", "rays_secondary", "rays_shadow", "light_evals", "sphere_tests", "plane_tests", "triangle_tests"]
bad = soft = 0
for it in range(n_scenes):
    s = Scene(ambient=(1.0, 1.0, 1.0), background=(0.1, 0.15, 0.2))
    mats = [s.add_material(Material(tuple(rng.uniform(0.2, 1, 3)), ka=0.1, kd=0.7, ks=0.3, shininess=float(rng.choice([8.0, 20.0, 33.5])), kr=float(rng.choice([0, 0, 0.3])))),
            s.add_material(Material((0.9, 0.95, 1.0), ka=0.0, kd=0.1, ks=0.4, shininess=80.0, kr=0.1, kt=0.8, ior=float(rng.uniform(1.2, 1.7)))),
            s.add_material(Material((0.9, 0.9, 0.9), ka=0.05, kd=0.2, ks=0.5, shininess=60.0, kr=0.7))]
    ext = float(10 ** rng.uniform(0, 1.5))
    for k in range(3):
        npl = int(rng.choice([0, 1, 2, 2, 2, 3]))
        pos = sorted(rng.uniform(-ext, ext, npl))
        for j, p in enumerate(pos):
            n = [0.0, 0.0, 0.0]
            n[k] = 1.0 if (j == 0) == (rng.random() < 0.8) else -1.0   # mostly facing inward
            s.add_plane(tuple(n), float(n[k] * p), mats[int(rng.integers(0, 3)) if rng.random() < 0.3 else 0])
    if rng.random() < 0.3:
        v = rng.normal(size=3)
        s.add_plane(tuple(v / np.linalg.norm(v)), float(-rng.uniform(0.5, 1.0) * ext), mats[0])
    for _ in range(int(rng.integers(1, 9))):
        s.add_sphere(tuple(rng.uniform(-0.7 * ext, 0.7 * ext, 3)), float(rng.uniform(0.05, 0.3) * ext), mats[int(rng.integers(0, 3))])
    if rng.random() < 0.3:
        for _ in range(int(rng.integers(1, 5))):
            c = rng.uniform(-0.6 * ext, 0.6 * ext, 3)
            s.add_triangle(*[tuple(c + rng.normal(size=3) * 0.3 * ext) for _ in range(3)], mats[int(rng.integers(0, 3))])
    for _ in range(int(rng.integers(1, 5))):
        lp = rng.uniform(-1.3 * ext, 1.3 * ext, 3)
        if rng.random() < 0.3 and len(s.planes):
            pl = s.planes[int(rng.integers(len(s.planes)))]
            k = int(np.argmax(np.abs(pl[:3])))
            if abs(pl[k]) == 1.0:
                lp[k] = pl[k] * pl[3] + float(rng.choice([0.0, 1e-9, 1e-5, 1e-2, -1e-2])) * ext
        s.add_light(tuple(lp), tuple(rng.uniform(0.2, 0.6, 3)))
    eye = rng.uniform(-0.9 * ext, 0.9 * ext, 3)
    cam = Camera(eye=tuple(eye), at=tuple(rng.uniform(-0.3 * ext, 0.3 * ext, 3)), up=(0, 1, 0), vfov_deg=float(rng.uniform(30, 90)))
    w, h, spp, depth = 96, 64, int(rng.choice([1, 4])), int(rng.integers(1, 6))
    eps = float(rng.choice([0.0, 1e-6, 1e-8, 1e-4]))
    kw = {}
    if rng.random() < 0.3:  # a shard of the frame: interleaved row bands, compact layout
        sc_ = int(rng.integers(2, 6))
        kw = dict(shard_index=int(rng.integers(0, sc_)), shard_count=sc_, band_rows=int(rng.choice([1, 3, 8, 16])), layout=abi.NT_LAYOUT_COMPACT)
    p = make_params(w, h, spp, depth, cam.resolve(w, h), abi.NT_F64_STRICT, ray_epsilon=eps, **kw)
    if only is not None and it not in only:
        continue
    try:
        with Renderer(s) as r:
            info = r.info()
            img, st = r.render_params(p)
            fast, _ = r.render_params(make_params(w, h, spp, depth, cam.resolve(w, h), abi.NT_F32_FAST, ray_epsilon=eps, **kw))
    except Exception as e:  # noqa: BLE001
        print(f"scene {it}: {type(e).__name__}: {e}")
        bad += 1
        continue
    if img.size == 0:  # a shard that owns no row
        continue
    ref, rst = oracle.render(s, p, compact_rows=img.shape[0]) if kw else oracle.render(s, p)
    diff = np.abs(img.astype(int) - ref.astype(int))
    nbad = int((diff.max(axis=-1) > 0).sum())
    cnt = [k for k in KEYS if (not info["uses_bvh"] or not k.endswith("_tests")) and st[k] != rst[k]]
    fd = np.abs(fast.astype(int) - ref.astype(int))[..., :3].max(axis=-1)
    if only is not None:
        print(f"scene {it}: counters gpu {[st[k] for k in KEYS]} oracle {[rst[k] for k in KEYS]} kw {kw}")
        print(f"scene {it}: fast within 3 LSB {float((fd <= 3).mean()):.3f}, mean abs diff {float(np.abs(fast.astype(int) - ref.astype(int))[..., :3].mean()):.2f}, planes {s.planes}, lights {s.lights}, eye {eye}")
    if (fd <= 3).mean() < 0.9:
        soft += 1  # the fast mode has no bit contract: degenerate scenes (a light ON a wall, eps = 0 against its 1e-4) differ legitimately
    if nbad > 2 or diff.max() > 1 or cnt:
        bad += 1
        print(f"scene {it}: strict diff pixels {nbad} max {diff.max()} counters {cnt} fast within 3 LSB {float((fd <= 3).mean()):.3f} "
              f"(planes {len(s.planes)} spheres {len(s.spheres)} tris {len(s.triangles)} lights {len(s.lights)} ext {ext:.2f} spp {spp} depth {depth} eps {eps} bvh {info['uses_bvh']})")
print(f"{n_scenes} scenes, {bad} bad (strict image or counters differ from the oracle, or an error), {soft} with a fast image far from the strict one")
