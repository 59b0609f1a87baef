"""Per-shard kernel time on ONE GPU (imbalance + fixed overhead of small launches; development aid)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import make_params
factory, w, h, spp, depth = scenes.CONFIGS["cfg3_cornell_1080p_4spp_d5"]
scene, cam = factory()
with Renderer(scene) as r:
    full = min(r.render(cam, w, h, spp, depth)[1]["kernel_ms"] for _ in range(3))
    print("full frame kernel", round(full, 4), "ms; /8 =", round(full / 8, 4))
    for n in (8,):
        for band in (1, 2, 4, 8, 16, 32):
            ts, rays = [], []
            for i in range(n):
                p = make_params(w, h, spp, depth, cam.resolve(w, h), shard_index=i, shard_count=n, band_rows=band, layout=abi.NT_LAYOUT_COMPACT)
                best = None
                for _ in range(3):
                    _, st = r.render_params(p)
                    best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
                ts.append(best["kernel_ms"]); rays.append(best["rays"])
            print(f"N={n} band={band}: kernel ms per shard min {min(ts):.4f} max {max(ts):.4f} mean {np.mean(ts):.4f}; rays max/mean {max(rays)/np.mean(rays):.3f}")
