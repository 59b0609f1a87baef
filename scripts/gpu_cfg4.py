"""cfg4 timing of one library build: python scripts/gpu_cfg4.py [f64|f32] (best of 3 kernel_ms)"""
import sys
sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
factory, w, h, spp, depth = scenes.CONFIGS["cfg4_mesh1m_4k_4spp_d3"]
scene, cam = factory()
with Renderer(scene) as r:
    best = None
    for _ in range(3):
        img, st = r.render(cam, w, h, spp, depth, abi.NT_F64_STRICT if prec == "f64" else abi.NT_F32_FAST)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
print(f"{prec} kernel {best['kernel_ms']:.2f} ms rays {best['rays']} box {best['box_tests']} tri {best['triangle_tests']} sph {best['sphere_tests']} checksum {int(img.astype('uint64').sum())}")
