"""Build time of the GPU BVH builder on configs[3]'s scene, second creation in the process (the first pays module loads):
   NT_BVH_BUILD=gpu NT_BVH_TIMING=1 python scripts/gpu_bvh_build_time.py"""
import sys, time
sys.path.insert(0, ".")
from nettracer_b200 import scenes
from nettracer_b200.renderer import Renderer
scene, cam = scenes.spheres_and_mesh()
for i in range(3):
    t0 = time.perf_counter()
    with Renderer(scene) as r:
        info = r.info()
    print(f"create {i}: {time.perf_counter() - t0:.3f} s, build {info.get('bvh_build_ms')} ms, nodes {info.get('bvh_nodes')}", flush=True)
