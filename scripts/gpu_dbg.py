"""debug: render the BVH test scene, dump stats + image to an npy"""
import sys
sys.path.insert(0, ".")
import numpy as np
from nettracer_b200 import abi, scenes
from nettracer_b200.renderer import Renderer
from nettracer_b200.scene import make_params
s, cam = scenes.random_mixed(150, 2, 300, seed=4)
p = make_params(224, 160, 4, 4, cam.resolve(224, 160), abi.NT_F64_STRICT)
with Renderer(s) as r:
    print(r.info())
    img, st = r.render_params(p)
print(st)
np.save(sys.argv[1], img)
