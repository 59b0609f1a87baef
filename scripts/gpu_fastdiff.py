import sys
sys.path.insert(0, ".")
import numpy as np
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
s, cam = scenes.cornell_box()
w, h = 480, 270
with Renderer(s) as r:
    a, _ = r.render(cam, w, h, 4, 5, abi.NT_F64_STRICT)
    b, _ = r.render(cam, w, h, 4, 5, abi.NT_F32_FAST)
    c, _ = r.render(cam, w, h, 4, 5, abi.NT_F64_STRICT, ray_epsilon=1e-4)
np.savez_compressed("gpurun_out/fastdiff.npz", a=a, b=b, c=c)
