"""Quick GPU probe: peaks + timings of the named configs (development aid, not the bench)."""
import json
import sys
import time

sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes  # noqa: E402
from nettracer_b200.renderer import Renderer, measure_peaks  # noqa: E402

which = sys.argv[1:] or ["cfg2_cornell_1080p_1spp_d1", "cfg3_cornell_1080p_4spp_d5"]
print("peaks", json.dumps(measure_peaks(0)))
for name in which:
    factory, w, h, spp, depth = scenes.CONFIGS[name]
    t0 = time.time()
    scene, cam = factory()
    t1 = time.time()
    with Renderer(scene) as r:
        t2 = time.time()
        print(name, "scene gen %.2fs create %.2fs" % (t1 - t0, t2 - t1), r.info())
        for prec, pn in ((abi.NT_F64_STRICT, "f64"), (abi.NT_F32_FAST, "f32")):
            best = None
            for it in range(4):
                img, st = r.render(cam, w, h, spp, depth, prec)
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            fl = abi.algorithmic_flops(best)
            print(f"  {pn}: kernel {best['kernel_ms']:.3f} ms total {best['total_ms']:.3f} ms rays {best['rays']} "
                  f"-> {best['rays'] / best['kernel_ms'] / 1e3:.1f} Mrays/s, {fl / best['kernel_ms'] / 1e9:.2f} TFLOP/s algorithmic",
                  {k: best[k] for k in ('sphere_tests', 'plane_tests', 'triangle_tests', 'box_tests', 'light_evals')})
