"""Traversal statistics of configs[3] for the three tree builders (development aid): tests per ray and frame time.
  python scripts/gpu_bvh_stats.py [f64|f32]"""
import os
import sys
sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
prec = abi.NT_F32_FAST if "f32" in sys.argv else abi.NT_F64_STRICT
factory, w, h, spp, depth = scenes.CONFIGS["cfg4_mesh1m_4k_4spp_d3"]
scene, cam = factory()
for name, env in [("host SAH", {}), ("gpu PLOC", {"NT_BVH_BUILD": "gpu"}), ("gpu LBVH", {"NT_BVH_BUILD": "gpu", "NT_BVH_GPU_ALGO": "lbvh"})]:
    for k in ("NT_BVH_BUILD", "NT_BVH_GPU_ALGO"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with Renderer(scene) as r:
        info = r.info()
        best = None
        for _ in range(3):
            img, st = r.render(cam, w, h, spp, depth, prec)
            best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
        rays = best["rays"]
        print(f"{name}: {best['kernel_ms']:.2f} ms, nodes {info['bvh_nodes']}, build {info['bvh_build_ms']:.1f} ms; per ray: box {best['box_tests'] / rays:.1f} "
              f"tri {best['triangle_tests'] / rays:.2f} sph {best['sphere_tests'] / rays:.2f}", flush=True)
