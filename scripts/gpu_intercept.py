"""Kernel time vs fraction of the frame (fixed overhead / tail of the persistent kernel; development aid)."""
import sys
sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import make_params
factory, w, h, spp, _ = scenes.CONFIGS["cfg3_cornell_1080p_4spp_d5"]
scene, cam = factory()
with Renderer(scene) as r:
    for depth in (5, 1):
        for prec in (abi.NT_F64_STRICT, abi.NT_F32_FAST):
            row = []
            for n in (1, 2, 4, 8, 16, 32, 64):
                p = make_params(w, h, spp, depth, cam.resolve(w, h), prec, shard_index=0, shard_count=n, band_rows=1, layout=abi.NT_LAYOUT_COMPACT)
                best = min(r.render_params(p)[1]["kernel_ms"] for _ in range(4))
                row.append(f"1/{n}: {best:.4f}")
            print("depth", depth, "f64" if prec == 0 else "f32", " | ".join(row))
