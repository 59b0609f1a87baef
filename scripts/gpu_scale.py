"""Timing vs image size for one scene (development aid)."""
import sys
sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
name = sys.argv[1]; prec = abi.NT_F64_STRICT if sys.argv[2] == "f64" else abi.NT_F32_FAST
factory, w, h, spp, depth = scenes.CONFIGS[name]
scene, cam = factory()
with Renderer(scene) as r:
    for (ww, hh, ss, dd) in [(960, 540, 1, 3), (1920, 1080, 1, 3), (3840, 2160, 1, 3), (1920, 1080, 4, 3), (3840, 2160, 4, 3), (3840, 2160, 4, 1), (3840, 2160, 1, 1)]:
        best = None
        for _ in range(2):
            img, st = r.render(cam, ww, hh, ss, dd, prec)
            if best is None or st["kernel_ms"] < best["kernel_ms"]: best = st
        print(f"{ww}x{hh} spp{ss} d{dd}: {best['kernel_ms']:.2f} ms rays {best['rays']/1e6:.1f}M -> {best['rays']/best['kernel_ms']/1e3:.0f} Mrays/s, box/ray {best['box_tests']/best['rays']:.1f} tri/ray {best['triangle_tests']/best['rays']:.2f} sph/ray {best['sphere_tests']/best['rays']:.2f}")
