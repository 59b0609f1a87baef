"""A/B timing of library variants (development aid): python scripts/gpu_ab.py [name ...]
Each name is nettracer_b200/variants/libnt_<name>.so ("main" = the shipped library), optionally followed by
":ENV=VALUE,ENV=VALUE" (environment switches read at scene creation / launch, e.g. main:NT_SLAB=0).  Prints, per variant and
precision, the median / min kernel time of configs[2] over 40 frames, the 1/8-frame time, an image hash and the ray
count (variants must agree bit for bit)."""
import hashlib
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ".")
    import numpy as np
    from nettracer_b200 import abi, scenes
    from nettracer_b200.renderer import Renderer
    from nettracer_b200.scene import make_params
    cfg = os.environ.get("NT_AB_CFG", "cfg3_cornell_1080p_4spp_d5")
    factory, w, h, spp, depth = scenes.CONFIGS[cfg]
    n = int(os.environ.get("NT_AB_ITERS", "40"))
    scene, cam = factory()
    with Renderer(scene) as r:
        for prec, pn in ((abi.NT_F64_STRICT, "f64"), (abi.NT_F32_FAST, "f32")):
            ts = []
            for _ in range(n):
                img, st = r.render(cam, w, h, spp, depth, prec)
                ts.append(st["kernel_ms"])
            ts.sort()
            p = make_params(w, h, spp, depth, cam.resolve(w, h), prec, shard_index=0, shard_count=8, band_rows=8, layout=abi.NT_LAYOUT_COMPACT)
            t8 = min(r.render_params(p)[1]["kernel_ms"] for _ in range(8))
            print(f"  {pn}: median {ts[len(ts) // 2]:.4f} min {ts[0]:.4f} ms | 1/8 frame {t8:.4f} ms | rays {st['rays']} "
                  f"sha {hashlib.sha1(np.ascontiguousarray(img).tobytes()).hexdigest()[:12]}", flush=True)
    sys.exit(0)

for spec in sys.argv[1:] or ["main"]:
    name, _, envs = spec.partition(":")
    env = dict(os.environ)
    for kv in filter(None, envs.split(",")):
        key, _, val = kv.partition("=")
        env[key] = val
    if name != "main":
        env["NT_LIB_PATH"] = os.path.abspath(f"nettracer_b200/variants/libnt_{name}.so")
    print(spec, flush=True)
    subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
