"""Loader for the C-ABI library libnettracer_b200.so (built in-tree by __graft_entry__.build()).

Fails loudly when the library is missing — there is no CPU or PyTorch fallback for the hot path."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NT_LIB_PATH") or os.path.join(_HERE, "libnettracer_b200.so")  # override: experiments only
_LIB = None


class NetTracerError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"nettracer_b200 error {code}: {text}")
        self.code = code


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C nettracer_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64p = C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)
    sig = {
        "nt_abi_version": (C.c_int, []),
        "nt_last_error": (C.c_char_p, []),
        "nt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
        "nt_scene_create": (C.c_int, [C.POINTER(abi.nt_scene_desc), C.c_int, C.POINTER(vp)]),
        "nt_scene_destroy": (None, [vp]),
        "nt_scene_info": (C.c_int, [vp, u64p]),
        "nt_render": (C.c_int, [vp, C.POINTER(abi.nt_render_params), vp, C.c_size_t, C.POINTER(abi.nt_render_stats)]),
        "nt_render_device": (C.c_int, [vp, C.POINTER(abi.nt_render_params), vp, C.c_size_t, vp]),
        "nt_render_device_stats": (C.c_int, [vp, vp, C.POINTER(abi.nt_render_stats)]),
        "nt_trace_rays": (C.c_int, [vp, u32, vp, vp, u32, C.c_double, vp, vp]),
        "nt_shard_rows": (u32, [u32, u32, u32, u32]),
        "nt_deinterleave_device": (C.c_int, [vp, C.c_size_t, vp, C.c_size_t, u32, u32, u32, u32, C.c_int, vp]),
        "nt_device_malloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(vp)]),
        "nt_device_free": (C.c_int, [C.c_int, vp]),
        "nt_ipc_export": (C.c_int, [vp, C.c_int, vp]),
        "nt_ipc_open": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
        "nt_ipc_close": (C.c_int, [vp, C.c_int]),
        "nt_measure_peaks": (C.c_int, [C.c_int, C.POINTER(abi.nt_peaks)]),
        "nt_cull_tables": (C.c_int, [C.POINTER(abi.nt_scene_desc), C.POINTER(u32), vp, C.c_size_t, vp, vp]),
        "nt_primary_rects": (C.c_int, [C.POINTER(abi.nt_scene_desc), C.POINTER(abi.nt_render_params), vp]),
        "nt_plane_free_lights": (C.c_int, [C.POINTER(abi.nt_scene_desc), C.POINTER(u32)]),
        "nt_light_rooms": (C.c_int, [C.POINTER(abi.nt_scene_desc), vp]),
        "nt_shadow_grid": (C.c_int, [C.POINTER(abi.nt_scene_desc), u32, vp, C.POINTER(u32), vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "nt_render_device_sync": (C.c_int, [vp, C.POINTER(abi.nt_render_params), vp, C.c_size_t, vp, C.POINTER(abi.nt_frame_sync)]),
        "nt_flags_wait_device": (C.c_int, [vp, vp, u32, u32, vp]),
        "nt_multi_create": (C.c_int, [C.POINTER(abi.nt_scene_desc), C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "nt_multi_destroy": (None, [vp]),
        "nt_multi_device_count": (C.c_int, [vp]),
        "nt_multi_render": (C.c_int, [vp, C.POINTER(abi.nt_render_params), vp, C.c_size_t, C.POINTER(abi.nt_render_stats)]),
        "nt_host_frame_open": (C.c_int, [C.c_char_p, C.c_size_t, u32, C.c_int, C.c_int, C.POINTER(vp)]),
        "nt_host_frame_pixels": (vp, [vp]),
        "nt_host_frame_flag": (vp, [vp, u32]),
        "nt_host_frame_post": (C.c_int, [vp, u32, u32]),
        "nt_host_frame_wait_all": (C.c_int, [vp, u32, u32]),
        "nt_host_frame_ack": (C.c_int, [vp, u32]),
        "nt_host_frame_wait_ack": (C.c_int, [vp, u32, u32]),
        "nt_host_frame_close": (None, [vp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name, None)
        if fn is None:
            if os.environ.get("NT_LIB_PATH"):  # an older experiment build: tolerate a missing newer entry point
                continue
            raise ImportError(f"{LIB_PATH} does not export {name}: rebuild it (make -C nettracer_b200/csrc)")
        fn.restype, fn.argtypes = res, args
    if L.nt_abi_version() != abi.NT_ABI_VERSION:
        raise ImportError("libnettracer_b200.so ABI version mismatch")
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise NetTracerError(rc, load().nt_last_error().decode(errors="replace"))
