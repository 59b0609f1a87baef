"""Ad-hoc parity fuzz of the BVH paths (development aid, not part of the suite): 70-400 small mirror / glass / matte spheres,
0-300 triangles, 0-3 planes, depth 1-6 (mirror chains: drifted directions, per-set passes, deferred cone walks and sweeps),
host- and GPU-built trees, wavefront and state machine, tiny workspaces - strict image and ray counters against the
brute-force oracle.   python scripts/gpu_fuzz_bvh.py [n_scenes] [seed]"""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
from nettracer_b200 import abi
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import Camera, Material, Scene, make_params
from oracle import oracle

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
KEYS = ["rays_primary", "rays_secondary", "rays_shadow", "light_evals"]
bad = 0
for it in range(n_scenes):
    s = Scene(ambient=(1.0, 1.0, 1.0), background=(0.2, 0.3, 0.5))
    mats = [s.add_material(Material(tuple(rng.uniform(0.2, 1, 3)), ka=0.1, kd=0.7, ks=0.3, shininess=20.0, kr=float(rng.choice([0, 0.3])))),
            s.add_material(Material((0.9, 0.95, 1.0), ka=0.0, kd=0.1, ks=0.4, shininess=80.0, kr=0.1, kt=0.8, ior=1.5)),
            s.add_material(Material((0.9, 0.9, 0.9), ka=0.05, kd=0.1, ks=0.5, shininess=60.0, kr=0.9))]
    ext = float(10 ** rng.uniform(0.5, 2))
    ns, nt = int(rng.integers(70, 400)), int(rng.choice([0, 0, 50, 300]))
    rmax = float(rng.choice([0.01, 0.03, 0.1])) * ext
    for _ in range(ns):
        s.add_sphere(tuple(rng.uniform(-ext, ext, 3)), float(rng.uniform(0.2, 1.0) * rmax), mats[int(rng.choice([0, 1, 2, 2]))])
    for _ in range(nt):
        c = rng.uniform(-ext, ext, 3)
        s.add_triangle(*[tuple(c + rng.normal(size=3) * 0.1 * ext) for _ in range(3)], mats[int(rng.integers(0, 3))])
    for _ in range(int(rng.integers(0, 4))):
        v = rng.normal(size=3) if rng.random() < 0.5 else np.eye(3)[int(rng.integers(0, 3))] * rng.choice([-1.0, 1.0])
        s.add_plane(tuple(v / np.linalg.norm(v)), float(-rng.uniform(1.0, 1.5) * ext), mats[0])
    for _ in range(int(rng.integers(1, 4))):
        s.add_light(tuple(rng.uniform(-2 * ext, 2 * ext, 3)), tuple(rng.uniform(0.2, 0.6, 3)))
    cam = Camera(eye=tuple(rng.uniform(-1.5 * ext, 1.5 * ext, 3)), at=tuple(rng.uniform(-0.3 * ext, 0.3 * ext, 3)), up=(0, 1, 0), vfov_deg=float(rng.uniform(30, 80)))
    w, h, spp, depth = 80, 56, int(rng.choice([1, 4])), int(rng.integers(1, 7))
    mode = int(rng.integers(0, 5))
    for k in ("NT_BVH_BUILD", "NT_WAVEFRONT", "NT_WF_MB"):
        os.environ.pop(k, None)
    if mode == 1:
        os.environ["NT_BVH_BUILD"] = "gpu"
    elif mode == 2:
        os.environ["NT_WAVEFRONT"] = "0"
    elif mode == 3:
        os.environ["NT_WF_MB"] = "1"
    p = make_params(w, h, spp, depth, cam.resolve(w, h), abi.NT_F64_STRICT)
    try:
        with Renderer(s) as r:
            info = r.info()
            img, st = r.render_params(p)
            r.render_params(make_params(w, h, spp, depth, cam.resolve(w, h), abi.NT_F32_FAST))
    except Exception as e:  # noqa: BLE001
        print(f"scene {it}: {type(e).__name__}: {e}", flush=True)
        bad += 1
        continue
    ref, rst = oracle.render(s, p, accel=0)
    diff = np.abs(img.astype(int) - ref.astype(int))
    nbad = int((diff.max(axis=-1) > 0).sum())
    cnt = [k for k in KEYS if st[k] != rst[k]]
    if nbad > 2 or diff.max() > 1 or cnt or not info["uses_bvh"]:
        bad += 1
        print(f"scene {it}: diff pixels {nbad} max {diff.max()} counters {cnt} (spheres {ns} tris {nt} planes {len(s.planes)} ext {ext:.1f} rmax {rmax:.3f} spp {spp} depth {depth} mode {mode} bvh {info['uses_bvh']})", flush=True)
print(f"{n_scenes} scenes, {bad} bad")
