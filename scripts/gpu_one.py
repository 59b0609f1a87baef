"""Render one config a few times (profiling target):
   python scripts/gpu_one.py <config> <f64|f32> [iters] [WxH] [spp] [depth]"""
import sys

sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes  # noqa: E402
from nettracer_b200.renderer import Renderer  # noqa: E402

name, prec = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
factory, w, h, spp, depth = scenes.CONFIGS[name]
if len(sys.argv) > 4:
    w, h = map(int, sys.argv[4].split("x"))
if len(sys.argv) > 5:
    spp = int(sys.argv[5])
if len(sys.argv) > 6:
    depth = int(sys.argv[6])
import os  # noqa: E402
shards = int(os.environ.get("NT_ONE_SHARDS", "1"))  # render shard 0 of this many (band 8) instead of the whole frame
scene, cam = factory()
with Renderer(scene) as r:
    for _ in range(iters):
        img, st = r.render(cam, w, h, spp, depth, abi.NT_F64_STRICT if prec == "f64" else abi.NT_F32_FAST,
                           shard_index=0, shard_count=shards, band_rows=8, layout=abi.NT_LAYOUT_COMPACT if shards > 1 else abi.NT_LAYOUT_FULL)
    print(name, prec, w, h, spp, depth, st)
