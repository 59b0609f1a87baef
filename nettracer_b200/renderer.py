"""Host-side renderer over the C ABI (include/nettracer_b200.h): the call a NetTracer host would
make in place of its per-pixel intersect-and-shade loop.  The reference's render API is unknown
(/root/reference/README:1-3); this mirrors SURVEY.md §8(b)'s proposed surface one to one.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .lib import check, load
from .scene import Camera, Scene, make_params, owned_rows, shard_rows


class Renderer:
    """One scene resident on one GPU.  `render()` = nt_render (host buffer, blocking);
    `render_device()` = nt_render_device (device buffer, asynchronous on a stream)."""

    def __init__(self, scene: Scene, device: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        desc, keep = scene.to_desc()
        check(self._lib.nt_scene_create(C.byref(desc), int(device), C.byref(self._h)))
        del keep
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.nt_scene_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self) -> dict:
        a = (C.c_uint64 * 4)()
        check(self._lib.nt_scene_info(self._h, a))
        return {"uses_bvh": bool(a[0] & 1), "bvh_on_gpu": bool(a[0] & 2), "culling": bool(a[0] & 4),
                "bvh_build_ms": (int(a[0]) >> 8) / 1000.0,
                "bvh_nodes": int(a[1]), "device_bytes": int(a[2]), "device": int(a[3]) & 0xffffffff,
                "last_launches": int(a[3]) >> 32}

    # -- host path --
    def render_params(self, params: abi.nt_render_params, out: np.ndarray | None = None):
        """out: uint8 [rows, width, 4] C-contiguous host array (rows = height for FULL, owned rows for
        COMPACT).  Returns (out, stats dict)."""
        rows = params.height if params.layout == abi.NT_LAYOUT_FULL else shard_rows(
            params.height, params.band_rows or 1, params.shard_index, params.shard_count or 1)
        if out is None:
            out = np.zeros((rows, params.width, 4), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (rows, params.width, 4)
        st = abi.nt_render_stats()
        check(self._lib.nt_render(self._h, C.byref(params), out.ctypes.data, params.width * 4, C.byref(st)))
        return out, st.as_dict()

    def render(self, camera: Camera, width, height, spp=1, max_depth=1, precision=abi.NT_F64_STRICT,
               out=None, **kw):
        p = make_params(width, height, spp, max_depth, camera.resolve(width, height), precision, **kw)
        return self.render_params(p, out)

    # -- device path --
    def render_device(self, params: abi.nt_render_params, dev_ptr: int, row_stride_bytes: int, stream: int = 0,
                      sync: abi.nt_frame_sync | None = None):
        """nt_render_device, or nt_render_device_sync when `sync` names frame-synchronisation flags."""
        check(self._lib.nt_render_device_sync(self._h, C.byref(params), C.c_void_p(dev_ptr), row_stride_bytes,
                                              C.c_void_p(stream), C.byref(sync) if sync is not None else None))

    def device_stats(self, stream: int = 0) -> dict:
        st = abi.nt_render_stats()
        check(self._lib.nt_render_device_stats(self._h, C.c_void_p(stream), C.byref(st)))
        return st.as_dict()

    # -- unit-level --
    def trace_rays(self, origins, dirs, precision=abi.NT_F64_STRICT, ray_epsilon=0.0):
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)
        t = np.zeros(len(o), dtype=np.float64)
        prim = np.zeros(len(o), dtype=np.int32)
        check(self._lib.nt_trace_rays(self._h, len(o), o.ctypes.data, d.ctypes.data, int(precision),
                                      float(ray_epsilon), t.ctypes.data, prim.ctypes.data))
        return t, prim


def cull_tables(scene: Scene) -> dict:
    """The conservative culling tables nt_scene_create builds for a flat scene (nt_cull_tables: host only,
    no GPU needed).  Returns k, lbuf [n_lights, 6, k, k] uint64, nbr [n_spheres] uint64, bsph [nb, 4]."""
    lib = load()
    desc, keep = scene.to_desc()
    k = C.c_uint32()
    check(lib.nt_cull_tables(C.byref(desc), C.byref(k), None, 0, None, None))
    nb = desc.n_spheres + desc.n_triangles
    lbuf = np.zeros((desc.n_lights, 6, k.value, k.value), dtype=np.uint64)
    nbr = np.zeros(desc.n_spheres, dtype=np.uint64)
    bsph = np.zeros((nb, 4), dtype=np.float64)
    check(lib.nt_cull_tables(C.byref(desc), C.byref(k), lbuf.ctypes.data, lbuf.size, nbr.ctypes.data, bsph.ctypes.data))
    del keep
    return {"k": int(k.value), "lbuf": lbuf, "nbr": nbr, "bsph": bsph}


def primary_rects(scene: Scene, params) -> np.ndarray:
    """Per bounded primitive the pixel rectangle [x0, x1, y0, y1] (inclusive; x0 > x1 = empty) outside which no
    primary ray of `params`' camera can touch it (nt_primary_rects: host only).  uint16 [n_spheres + n_triangles, 4]."""
    desc, keep = scene.to_desc()
    out = np.zeros((desc.n_spheres + desc.n_triangles, 4), dtype=np.uint16)
    check(load().nt_primary_rects(C.byref(desc), C.byref(params), out.ctypes.data))
    del keep
    return out


def plane_free_lights(scene: Scene) -> int:
    """Bit mask of the lights towards which a shadow query from a bounded primitive cannot be stopped by a plane
    (nt_plane_free_lights: host only)."""
    desc, keep = scene.to_desc()
    m = C.c_uint32()
    check(load().nt_plane_free_lights(C.byref(desc), C.byref(m)))
    del keep
    return int(m.value)


def shadow_grid(scene: Scene, light: int):
    """The shadow grid of one light of a BVH scene (nt_shadow_grid: host only): None when the light has none, else
    (params float32[16], K, off uint32[K*K + 1] rebased to 0, items uint32[...]) - see include/nettracer_b200.h."""
    desc, keep = scene.to_desc()
    lib = load()
    k, n = C.c_uint32(), C.c_size_t()
    check(lib.nt_shadow_grid(C.byref(desc), light, None, C.byref(k), None, 0, None, 0, C.byref(n)))
    if k.value == 0:
        del keep
        return None
    params = np.zeros(16, dtype=np.float32)
    off = np.zeros(k.value * k.value + 1, dtype=np.uint32)
    items = np.zeros(max(n.value, 1), dtype=np.uint32)
    check(lib.nt_shadow_grid(C.byref(desc), light, params.ctypes.data, C.byref(k), off.ctypes.data, off.size, items.ctypes.data, items.size, C.byref(n)))
    del keep
    return params, int(k.value), off - off[0], items[:n.value]


def light_rooms(scene: Scene) -> np.ndarray:
    """Per light the axis-aligned room [lo_x, hi_x, lo_y, hi_y, lo_z, hi_z, cap_per_eps, cap_max] inside which a shadow
    query skips the axis-aligned planes (nt_light_rooms: host only).  float64 [n_lights, 8]."""
    desc, keep = scene.to_desc()
    out = np.zeros((max(desc.n_lights, 1), 8), dtype=np.float64)
    check(load().nt_light_rooms(C.byref(desc), out.ctypes.data))
    del keep
    return out[:desc.n_lights]


def measure_peaks(device=0) -> dict:
    p = abi.nt_peaks()
    check(load().nt_measure_peaks(int(device), C.byref(p)))
    return p.as_dict()


def deinterleave_host(compact_shards, height, width, band_rows):
    """Host twin of nt_deinterleave_device (used by CPU tests of the sharding arithmetic)."""
    n = len(compact_shards)
    full = np.zeros((height, width, 4), dtype=np.uint8)
    for i, buf in enumerate(compact_shards):
        full[owned_rows(height, band_rows, i, n)] = buf
    return full
