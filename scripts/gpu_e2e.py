"""e2e timing of nt_render with pageable / pinned host buffers (development aid)."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import make_params
for name in sys.argv[1:] or ["cfg3_cornell_1080p_4spp_d5", "cfg2_cornell_1080p_1spp_d1"]:
    factory, w, h, spp, depth = scenes.CONFIGS[name]
    scene, cam = factory()
    p = make_params(w, h, spp, depth, cam.resolve(w, h))
    with Renderer(scene) as r:
        pageable = np.zeros((h, w, 4), dtype=np.uint8)
        pinned_t = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
        pinned = pinned_t.numpy()
        for label, buf in (("pageable", pageable), ("pinned", pinned)):
            ts = []
            for i in range(30):
                t0 = time.perf_counter(); _, st = r.render_params(p, buf); ts.append(time.perf_counter() - t0)
            ts = sorted(ts[5:])
            print(f"{name} {label}: median {1e3*ts[len(ts)//2]:.3f} ms min {1e3*ts[0]:.3f} ms kernel {st['kernel_ms']:.3f} ms")
        assert np.array_equal(pageable, pinned)
