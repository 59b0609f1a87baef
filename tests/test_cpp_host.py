"""The C++ host above the C ABI (host/nt_host.hpp, host/render_demo.cpp): builds everywhere; without a
GPU it must fail loudly (exit code 3, NT_ERR_NO_DEVICE); on a B200 its image must equal the CPU oracle's
image of the same scene built through the Python host classes (so both host mirrors flatten alike)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nettracer_b200 import lib
from nettracer_b200.scene import Camera, Material, Scene, make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "host", "render_demo")


def build_demo():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.DEVNULL)
    return DEMO


def demo_scene():
    s = Scene(ambient=(1, 1, 1), background=(0.05, 0.07, 0.12))
    m0 = s.add_material(Material((0.7, 0.7, 0.72), ka=0.1, kd=0.8, kr=0.15))
    m1 = s.add_material(Material((1, 1, 1), ka=0.1, kd=0.05, ks=0.5, shininess=120, kr=0.1, kt=0.85, ior=1.5))
    m2 = s.add_material(Material((0.85, 0.2, 0.15), ka=0.1, kd=0.8, ks=0.4, shininess=40))
    m3 = s.add_material(Material((1, 1, 1), ka=0.1, kd=0.15, ks=0.6, shininess=100, kr=0.75))
    s.add_plane((0, 1, 0), 0.0, m0)
    s.add_plane((0.2, 0.1, 1), -9.0, m0)
    s.add_sphere((-1.6, 1.0, 0.0), 1.0, m1)
    s.add_sphere((1.2, 0.8, -0.8), 0.8, m2)
    s.add_sphere((0.2, 0.5, 1.6), 0.5, m3)
    s.add_triangle((-3.5, 0.0, -2.5), (-1.5, 0.0, -3.5), (-2.5, 2.4, -3.0), m2)
    s.add_light((-4, 7, 5), (0.7, 0.68, 0.65))
    s.add_light((5, 6, 2), (0.35, 0.38, 0.45))
    return s, Camera((0.3, 2.2, 7.5), (0, 0.8, 0), vfov_deg=42)


def _has_gpu():
    n = C.c_int(0)
    return lib.load().nt_device_count(C.byref(n)) == 0 and n.value > 0


def test_cpp_host_builds_and_refuses_without_gpu(tmp_path):
    exe = build_demo()
    if _has_gpu():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, str(tmp_path / "x.ppm")], capture_output=True, text=True)
    assert r.returncode == 3 and "error -2" in r.stderr
    assert not (tmp_path / "x.ppm").exists()


@pytest.mark.gpu
def test_cpp_host_image_equals_oracle(tmp_path):
    from oracle import oracle
    exe = build_demo()
    out = tmp_path / "demo.ppm"
    w, h, spp, depth = 160, 90, 4, 4
    r = subprocess.run([exe, str(out), str(w), str(h), str(spp), str(depth), "f64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    header = b"P6\n%d %d\n255\n" % (w, h)
    assert raw.startswith(header)
    img = np.frombuffer(raw[len(header):], dtype=np.uint8).reshape(h, w, 3)
    s, cam = demo_scene()
    ref, st = oracle.render(s, make_params(w, h, spp, depth, cam.resolve(w, h)))
    assert np.array_equal(img, ref[..., :3])
    assert f"{st['rays']} rays" in r.stdout


def cornell_fixed():
    """The scene render_demo builds for `cornell` (fixed sphere positions), through the Python host classes."""
    s = Scene(ambient=(1, 1, 1), background=(0.02, 0.02, 0.03))
    white = s.add_material(Material((0.75, 0.75, 0.75), ka=0.08, kd=0.85))
    red = s.add_material(Material((0.75, 0.15, 0.15), ka=0.08, kd=0.85))
    green = s.add_material(Material((0.15, 0.75, 0.15), ka=0.08, kd=0.85))
    floor = s.add_material(Material((0.6, 0.6, 0.65), ka=0.08, kd=0.7, ks=0.2, shininess=40, kr=0.2))
    mirror = s.add_material(Material((0.9, 0.9, 0.95), ka=0.02, kd=0.15, ks=0.6, shininess=120, kr=0.75))
    glass = s.add_material(Material((0.95, 0.98, 1.0), ka=0.0, kd=0.05, ks=0.5, shininess=200, kr=0.1, kt=0.85, ior=1.5))
    blue = s.add_material(Material((0.2, 0.35, 0.85), ka=0.1, kd=0.7, ks=0.4, shininess=30))
    for n, d, m in [((0, 1, 0), 0.0, floor), ((0, -1, 0), -10.0, white), ((1, 0, 0), -6.0, red),
                    ((-1, 0, 0), -6.0, green), ((0, 0, 1), -8.0, white), ((0, 0, -1), -16.0, white)]:
        s.add_plane(n, d, m)
    mats = [mirror, glass, blue, mirror, glass, blue, mirror, glass]
    for k in range(8):
        r = 0.95 + 0.05 * k
        s.add_sphere((-4.2 + (k % 4) * 2.8, r + (0.5 * k if k % 3 == 1 else 0.0), -3.5 + (k // 4) * 4.5), r, mats[k])
    s.add_light((-3.0, 9.2, 4.0), (0.65, 0.62, 0.6))
    s.add_light((3.5, 8.8, -2.0), (0.45, 0.47, 0.5))
    return s, Camera((0.0, 5.0, 15.0), (0.0, 3.2, 0.0), vfov_deg=42)


@pytest.mark.gpu
def test_cpp_host_multi_gpu_cornell_equals_oracle(tmp_path):
    """No Python in the rendering process: render_demo drives every visible GPU through nt_multi_render (n_gpus = 0)
    on the Cornell-style box at 960x540, 4 spp, depth 5; the PPM must equal the oracle's frame."""
    from oracle import oracle
    exe = build_demo()
    out = tmp_path / "cornell.ppm"
    w, h, spp, depth = 960, 540, 4, 5
    r = subprocess.run([exe, str(out), str(w), str(h), str(spp), str(depth), "f64", "0", "cornell"], capture_output=True, text=True)
    assert r.returncode == 0 and "multi-GPU:" in r.stdout, r.stdout + r.stderr
    raw = out.read_bytes()
    header = b"P6\n%d %d\n255\n" % (w, h)
    img = np.frombuffer(raw[len(header):], dtype=np.uint8).reshape(h, w, 3)
    s, cam = cornell_fixed()
    ref, st = oracle.render(s, make_params(w, h, spp, depth, cam.resolve(w, h)))
    assert np.array_equal(img, ref[..., :3])
    assert f"{st['rays']} rays" in r.stdout
