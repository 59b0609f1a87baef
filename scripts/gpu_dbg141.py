import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
src = open("scripts/gpu_fuzz_flat.py").read()
gen = src[src.index("for it in range(n_scenes):"):src.index("    if only is not None and it not in only:")]
from nettracer_b200 import abi
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import Camera, Material, Scene, make_params
from oracle import oracle
rng = np.random.default_rng(31); n_scenes = 142
scenes_ = []
exec(gen + "    scenes_.append((s,cam,w,h,spp,depth,eps,kw))\n")
s, cam, w, h, spp, depth, eps, kw = scenes_[141]
KEYS = ["rays_primary", "rays_secondary", "rays_shadow", "light_evals", "sphere_tests", "plane_tests"]
with Renderer(s) as r:
    for y in range(h):
        p = make_params(w, h, spp, depth, cam.resolve(w, h), abi.NT_F64_STRICT, ray_epsilon=eps, shard_index=y, shard_count=h, band_rows=1, layout=abi.NT_LAYOUT_COMPACT)
        img, st = r.render_params(p)
        ref, rst = oracle.render(s, p, compact_rows=1)
        if any(st[k] != rst[k] for k in KEYS):
            print("row", y, [st[k] for k in KEYS], [rst[k] for k in KEYS])
            # narrow by depth and spp
            for d2 in range(1, depth + 1):
                p2 = make_params(w, h, spp, d2, cam.resolve(w, h), abi.NT_F64_STRICT, ray_epsilon=eps, shard_index=y, shard_count=h, band_rows=1, layout=abi.NT_LAYOUT_COMPACT)
                _, s2 = r.render_params(p2); _, o2 = oracle.render(s, p2, compact_rows=1)
                print("   depth", d2, [s2[k] for k in KEYS], [o2[k] for k in KEYS])
