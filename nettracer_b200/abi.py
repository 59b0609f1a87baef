"""ctypes mirror of include/nettracer_b200.h (the C-ABI drop-in boundary).

No reference interface can be cited: /root/reference/README:1-3 is the whole reference.
The same structs are consumed by the CUDA library and by the test-only CPU oracle.
"""
import ctypes as C

NT_ABI_VERSION = 2
NT_OK, NT_ERR_INVALID, NT_ERR_NO_DEVICE, NT_ERR_CUDA, NT_ERR_NOMEM, NT_ERR_TIMEOUT, NT_ERR_SYSTEM = 0, -1, -2, -3, -4, -5, -6
NT_F64_STRICT, NT_F32_FAST = 0, 1
NT_LAYOUT_FULL, NT_LAYOUT_COMPACT = 0, 1
NT_MAX_DEPTH = 16
NT_RENDER_COUNT_EXECUTED = 1
NT_RULE_QUANTIZE_TRUNCATE, NT_RULE_ATTENUATE_INV_SQUARE, NT_RULE_SAMPLE_CORNER, NT_RULE_RENORMALIZE = 2, 4, 8, 16
NT_RULE_MASK = 30

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class nt_camera(C.Structure):
    _fields_ = [("eye", C.c_double * 3), ("p00", C.c_double * 3),
                ("dx", C.c_double * 3), ("dy", C.c_double * 3)]


class nt_scene_desc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32),
                ("n_spheres", C.c_uint32), ("n_planes", C.c_uint32), ("n_triangles", C.c_uint32),
                ("n_materials", C.c_uint32), ("n_lights", C.c_uint32),
                ("spheres", _pd), ("sphere_mat", _pi),
                ("planes", _pd), ("plane_mat", _pi),
                ("triangles", _pd), ("triangle_mat", _pi),
                ("materials", _pd), ("lights", _pd),
                ("ambient", C.c_double * 3), ("background", C.c_double * 3)]


class nt_render_params(C.Structure):
    _fields_ = [("struct_size", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32),
                ("spp", C.c_uint32), ("max_depth", C.c_uint32), ("precision", C.c_uint32),
                ("ray_epsilon", C.c_double),
                ("camera", nt_camera),
                ("shard_index", C.c_uint32), ("shard_count", C.c_uint32), ("band_rows", C.c_uint32),
                ("layout", C.c_uint32), ("flags", C.c_uint32)]


class nt_render_stats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_secondary", C.c_uint64),
                ("rays_shadow", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("plane_tests", C.c_uint64),
                ("triangle_tests", C.c_uint64), ("box_tests", C.c_uint64),
                ("light_evals", C.c_uint64),
                ("sphere_tests_executed", C.c_uint64), ("plane_tests_executed", C.c_uint64),
                ("triangle_tests_executed", C.c_uint64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["rays"] = self.rays_primary + self.rays_secondary + self.rays_shadow
        return d


class nt_frame_sync(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("post_at_start_value", C.c_uint32), ("wait_value", C.c_uint32),
                ("post_when_done_value", C.c_uint32),
                ("post_at_start", C.c_void_p), ("wait_before_store", C.c_void_p), ("post_when_done", C.c_void_p)]


class nt_peaks(C.Structure):
    _fields_ = [("f64_fma_gflops", C.c_double), ("f64_nofma_gflops", C.c_double),
                ("f32_fma_gflops", C.c_double), ("f32_nofma_gflops", C.c_double),
                ("sm_clock_mhz_est", C.c_double), ("sm_count", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# Every symbol include/nettracer_b200.h declares (tests check the .so exports each one).
EXPORTS = [
    "nt_abi_version", "nt_last_error", "nt_device_count",
    "nt_scene_create", "nt_scene_destroy", "nt_scene_info",
    "nt_render", "nt_render_device", "nt_render_device_stats", "nt_trace_rays",
    "nt_shard_rows", "nt_deinterleave_device",
    "nt_device_malloc", "nt_device_free", "nt_ipc_export", "nt_ipc_open", "nt_ipc_close",
    "nt_measure_peaks", "nt_cull_tables", "nt_primary_rects", "nt_plane_free_lights", "nt_light_rooms", "nt_shadow_grid",
    "nt_render_device_sync", "nt_flags_wait_device",
    "nt_multi_create", "nt_multi_destroy", "nt_multi_device_count", "nt_multi_render",
    "nt_host_frame_open", "nt_host_frame_pixels", "nt_host_frame_flag", "nt_host_frame_post", "nt_host_frame_wait_all",
    "nt_host_frame_ack", "nt_host_frame_wait_ack", "nt_host_frame_close",
]

# Flop-counting convention fixed in SURVEY.md §8(d) (FMA = 2, div/sqrt = 1).
FLOPS = {"sphere_tests": 17, "plane_tests": 11, "triangle_tests": 39, "box_tests": 18,
         "light_evals": 40}


def algorithmic_flops(stats: dict) -> int:
    return sum(stats[k] * v for k, v in FLOPS.items())


def executed_flops(stats: dict) -> int:
    """Same weights on the tests an instrumented flat-scene launch really started (NT_RENDER_COUNT_EXECUTED)."""
    return (stats["sphere_tests_executed"] * FLOPS["sphere_tests"] + stats["plane_tests_executed"] * FLOPS["plane_tests"]
            + stats["triangle_tests_executed"] * FLOPS["triangle_tests"] + stats["box_tests"] * FLOPS["box_tests"]
            + stats["light_evals"] * FLOPS["light_evals"])
