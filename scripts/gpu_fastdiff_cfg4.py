"""configs[3], fast mode against strict mode on the GPU (development aid): where do the two frames differ?
   python scripts/gpu_fastdiff_cfg4.py"""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
factory, w, h, spp, depth = scenes.CONFIGS["cfg4_mesh1m_4k_4spp_d3"]
s, cam = factory()
out = {}
with Renderer(s) as r:
    a, _ = r.render(cam, w, h, spp, depth, abi.NT_F64_STRICT)
    for name, env, kw in (("fast", {}, {}), ("fast_sm", {"NT_WAVEFRONT": "0"}, {}), ("fast_d1", {}, {"depth": 1}),
                          ("strict_eps1e-4", {}, {"prec": abi.NT_F64_STRICT, "ray_epsilon": 1e-4})):
        os.environ.update(env)
        d = kw.pop("depth", depth)
        prec = kw.pop("prec", abi.NT_F32_FAST)
        ref = a if d == depth else r.render(cam, w, h, spp, d, abi.NT_F64_STRICT)[0]
        b, _ = r.render(cam, w, h, spp, d, prec, **kw)
        for k in env:
            os.environ.pop(k)
        diff = np.abs(a.astype(int) - b.astype(int))[..., :3].max(axis=-1) if d == depth else np.abs(ref.astype(int) - b.astype(int))[..., :3].max(axis=-1)
        hist = np.bincount(np.minimum(diff, 16).ravel(), minlength=17)
        rows = (diff > 2).mean(axis=1)
        print(name, "frac > 2 LSB", float((diff > 2).mean()), "hist (0..15, >= 16)", hist.tolist(), flush=True)
        print("   by 270-row band", [round(float(rows[i:i + 270].mean()), 4) for i in range(0, h, 270)], flush=True)
        out[name] = (diff > 2)[::8, ::8]
np.savez_compressed("gpurun_out/fastdiff_cfg4.npz", **out)
