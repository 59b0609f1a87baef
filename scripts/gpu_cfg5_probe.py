"""cfg5 shard 0 of 8 on one GPU with and without NT_RULE_RENORMALIZE (development aid): how much of the strict frame is the
walk of drifted directions (query_arm, nt_bvh_trace.cuh)?   python scripts/gpu_cfg5_probe.py [WxH]"""
import sys
sys.path.insert(0, ".")
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import make_params
factory, w, h, spp, depth = scenes.CONFIGS["cfg5_mesh1m_8k_16spp_d5"]
if len(sys.argv) > 1:
    w, h = map(int, sys.argv[1].split("x"))
scene, cam = factory()
with Renderer(scene) as r:
    for name, prec, flags in (("f64", abi.NT_F64_STRICT, 0), ("f64 renormalised", abi.NT_F64_STRICT, abi.NT_RULE_RENORMALIZE),
                              ("f32", abi.NT_F32_FAST, 0)):
        p = make_params(w, h, spp, depth, cam.resolve(w, h), prec, shard_index=0, shard_count=8, band_rows=8,
                        layout=abi.NT_LAYOUT_COMPACT, flags=flags)
        best = min(r.render_params(p)[1]["kernel_ms"] for _ in range(2))
        st = r.render_params(p)[1]
        print(f"{name}: {best:.1f} ms rays {st['rays']} box {st['box_tests']} sph {st['sphere_tests']} tri {st['triangle_tests']}", flush=True)
