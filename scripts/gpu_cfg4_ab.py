"""cfg4 (configs[3]) A/B of library variants: python scripts/gpu_cfg4_ab.py [name[:ENV=V,...]] ...  (see gpu_ab.py)"""
import os
import subprocess
import sys

for spec in sys.argv[1:] or ["main"]:
    name, _, envs = spec.partition(":")
    env = dict(os.environ)
    for kv in filter(None, envs.split(",")):
        key, _, val = kv.partition("=")
        env[key] = val
    if name != "main":
        env["NT_LIB_PATH"] = os.path.abspath(f"nettracer_b200/variants/libnt_{name}.so")
    for prec in ("f64", "f32"):
        out = subprocess.run([sys.executable, "scripts/gpu_cfg4.py", prec], env=env, capture_output=True, text=True)
        print(spec, (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1], flush=True)
