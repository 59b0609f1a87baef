"""Generates tests/golden/*.npz from the CPU oracle (run: python tests/make_golden.py).

These fixtures do NOT come from NetTracer — no reference source, scene or image exists
(/root/reference/README:1-3) — so they pin the oracle and the CUDA path to each other and to
SPEC-PROVISIONAL.md across compilers and rounds, nothing more ("parity unpinned")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nettracer_b200 import scenes  # noqa: E402
from nettracer_b200.scene import make_params  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {
    # name: (factory, kwargs, width, height, spp, depth)
    "cornell_160x90_s4_d5": (scenes.cornell_box, {}, 160, 90, 4, 5),
    "cornell_96x54_s1_d1": (scenes.cornell_box, {}, 96, 54, 1, 1),
    "mixed7_128x96_s4_d6": (scenes.random_mixed, dict(n_spheres=10, n_planes=2, n_triangles=14, seed=7), 128, 96, 4, 6),
    "mixedbvh11_112x80_s4_d4": (scenes.random_mixed, dict(n_spheres=90, n_planes=2, n_triangles=160, seed=11), 112, 80, 4, 4),
    "mesh_128x72_s4_d3": (scenes.spheres_and_mesh, dict(n_spheres=500, mesh_n=48), 128, 72, 4, 3),
}
KEYS = ["rays_primary", "rays_secondary", "rays_shadow", "light_evals"]


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, (factory, kw, w, h, spp, depth) in CASES.items():
        scene, cam = factory(**kw)
        p = make_params(w, h, spp, depth, cam.resolve(w, h))
        img, st = oracle.render(scene, p, accel=0)
        np.savez_compressed(os.path.join(out, name + ".npz"), rgba=img,
                            counters=np.array([st[k] for k in KEYS], dtype=np.uint64))
        print(name, img.shape, {k: st[k] for k in KEYS})


if __name__ == "__main__":
    main()
