"""ctypes loader for oracle/liboracle.so — TEST INFRASTRUCTURE ONLY, PARITY UNPINNED.

Never imported by nettracer_b200/ (tests/test_boundary.py greps for that)."""
import ctypes as C
import os
import subprocess

import numpy as np

from nettracer_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("nt_oracle.c", "nt_oracle.h")] + \
           [os.path.join(_HERE, "..", "include", "nettracer_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.nto_render_radiance.restype = C.c_int
        L.nto_render_radiance.argtypes = [C.POINTER(abi.nt_scene_desc), C.POINTER(abi.nt_render_params),
                                          C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.POINTER(abi.nt_render_stats), C.c_int, C.c_int, C.c_uint32]
        L.nto_trace_rays.restype = C.c_int
        L.nto_trace_rays.argtypes = [C.POINTER(abi.nt_scene_desc), C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.nto_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads():
    return lib().nto_max_threads()


def render(scene, params, accel=0, n_threads=0, row_step=1, radiance=False, compact_rows=None):
    """Returns (rgba uint8 [rows,w,4], stats dict[, radiance float64 [h,w,3]])."""
    desc, keep = scene.to_desc()
    h, w = params.height, params.width
    rows = h if params.layout == abi.NT_LAYOUT_FULL else compact_rows
    img = np.zeros((rows, w, 4), dtype=np.uint8)
    rad = np.zeros((h, w, 3), dtype=np.float64) if radiance else None
    st = abi.nt_render_stats()
    rc = lib().nto_render_radiance(C.byref(desc), C.byref(params), img.ctypes.data, w * 4,
                                   rad.ctypes.data if radiance else None, C.byref(st),
                                   int(accel), int(n_threads), int(row_step))
    if rc != 0:
        raise RuntimeError(f"oracle render failed: {rc}")
    del keep
    return (img, st.as_dict(), rad) if radiance else (img, st.as_dict())


def sample_costs(scene, params, accel=0, n_threads=0):
    """Rays cast and ray-tree nodes visited by every sample: uint16 [h, w, spp, 2] (analysis aid)."""
    desc, keep = scene.to_desc()
    out = np.zeros((params.height, params.width, params.spp, 2), dtype=np.uint16)
    st = abi.nt_render_stats()
    L = lib()
    L.nto_sample_costs.restype = C.c_int
    L.nto_sample_costs.argtypes = [C.POINTER(abi.nt_scene_desc), C.POINTER(abi.nt_render_params), C.c_void_p,
                                   C.POINTER(abi.nt_render_stats), C.c_int, C.c_int]
    rc = L.nto_sample_costs(C.byref(desc), C.byref(params), out.ctypes.data, C.byref(st), int(accel), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle sample_costs failed: {rc}")
    del keep
    return out, st.as_dict()


def trace_rays(scene, origins, dirs, ray_epsilon=0.0, accel=0):
    desc, keep = scene.to_desc()
    o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)
    t = np.zeros(len(o), dtype=np.float64)
    prim = np.zeros(len(o), dtype=np.int32)
    rc = lib().nto_trace_rays(C.byref(desc), len(o), o.ctypes.data, d.ctypes.data, float(ray_epsilon),
                              int(accel), t.ctypes.data, prim.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle trace failed: {rc}")
    del keep
    return t, prim
