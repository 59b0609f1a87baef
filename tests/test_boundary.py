"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, its structs have the layout the ctypes mirror assumes, it refuses to compute without an
sm_100 device (no CPU fallback), and the product package never touches oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

from nettracer_b200 import abi, lib, scenes
from nettracer_b200.scene import shard_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nettracer_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    names = header_functions()
    assert set(names) == set(abi.EXPORTS), set(names) ^ set(abi.EXPORTS)
    L = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    nm = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (nt_[a-z0-9_]+)", nm))
    assert exported == set(names), exported ^ set(names)


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nettracer_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(nt_scene_desc),sizeof(nt_render_params),sizeof(nt_render_stats),sizeof(nt_camera),sizeof(nt_peaks),"
                   "offsetof(nt_render_params,camera),offsetof(nt_scene_desc,ambient),sizeof(nt_frame_sync),"
                   "offsetof(nt_frame_sync,post_when_done));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(abi.nt_scene_desc), C.sizeof(abi.nt_render_params), C.sizeof(abi.nt_render_stats),
            C.sizeof(abi.nt_camera), C.sizeof(abi.nt_peaks), abi.nt_render_params.camera.offset,
            abi.nt_scene_desc.ambient.offset, C.sizeof(abi.nt_frame_sync), abi.nt_frame_sync.post_when_done.offset]
    assert got == want


def test_library_loads_and_pure_host_entry_points():
    L = lib.load()
    assert L.nt_abi_version() == abi.NT_ABI_VERSION
    for h, band, n in [(1080, 16, 8), (1080, 8, 3), (17, 5, 4), (1, 1, 1), (100, 7, 7)]:
        rows = [L.nt_shard_rows(h, band, i, n) for i in range(n)]
        assert rows == [shard_rows(h, band, i, n) for i in range(n)]
        assert sum(rows) == h
    assert L.nt_shard_rows(10, 0, 0, 1) == 0 and L.nt_shard_rows(10, 4, 2, 2) == 0


def _has_gpu():
    n = C.c_int(0)
    return lib.load().nt_device_count(C.byref(n)) == 0 and n.value > 0


def test_no_device_means_error_not_fallback():
    if _has_gpu():
        pytest.skip("a GPU is present")
    from nettracer_b200.renderer import Renderer, measure_peaks
    s, cam = scenes.cornell_box()
    with pytest.raises(lib.NetTracerError) as e:
        Renderer(s)
    assert e.value.code == abi.NT_ERR_NO_DEVICE
    with pytest.raises(lib.NetTracerError):
        measure_peaks(0)


def test_multi_gpu_entry_points_refuse_without_gpu_and_validate():
    """nt_multi_* / nt_host_frame_* (ABI v2): argument errors are reported before any device work, and without a GPU
    they fail with NT_ERR_NO_DEVICE - no CPU rendering path hides behind them."""
    L = lib.load()
    s, cam = scenes.cornell_box()
    d, keep = s.to_desc()
    h = C.c_void_p()
    assert L.nt_multi_create(C.byref(d), None, 1, C.byref(h)) == abi.NT_ERR_INVALID
    devs = (C.c_int * 2)(0, 0)
    assert L.nt_multi_create(C.byref(d), devs, 2, C.byref(h)) == abi.NT_ERR_INVALID and b"twice" in L.nt_last_error()
    assert L.nt_multi_device_count(None) == 0
    assert L.nt_host_frame_open(b"no-slash", 64, 1, 1, 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert L.nt_host_frame_open(b"/nt_test_x", 0, 1, 1, 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert L.nt_host_frame_wait_all(None, 1, 1) == abi.NT_ERR_INVALID
    assert L.nt_flags_wait_device(None, None, 1, 1, None) == abi.NT_ERR_INVALID
    if _has_gpu():
        return
    devs = (C.c_int * 1)(0)
    assert L.nt_multi_create(C.byref(d), devs, 1, C.byref(h)) == abi.NT_ERR_NO_DEVICE and not h.value
    assert L.nt_host_frame_open(b"/nt_test_x", 4096, 1, 1, 0, C.byref(h)) == abi.NT_ERR_NO_DEVICE
    assert not os.path.exists("/dev/shm/nt_test_x")


def test_invalid_scene_rejected_before_device_work():
    L = lib.load()
    s, cam = scenes.cornell_box()
    d, keep = s.to_desc()
    d.struct_size = 3
    h = C.c_void_p()
    assert L.nt_scene_create(C.byref(d), 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert b"struct_size" in L.nt_last_error()
    d, keep = s.to_desc()
    d.n_materials = 0
    assert L.nt_scene_create(C.byref(d), 0, C.byref(h)) == abi.NT_ERR_INVALID
    s.spheres[0] = (0.0, 0.0, 0.0, -1.0)
    d, keep = s.to_desc()
    assert L.nt_scene_create(C.byref(d), 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert L.nt_scene_create(None, 0, C.byref(h)) == abi.NT_ERR_INVALID
    # non-finite lights / materials are refused too (they would poison the culling tables and every pixel)
    s, cam = scenes.cornell_box()
    s.lights[0] = (float("nan"), 9.0, 4.0, 1.0, 1.0, 1.0)
    d, keep = s.to_desc()
    assert L.nt_scene_create(C.byref(d), 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert b"light 0" in L.nt_last_error()
    s, cam = scenes.cornell_box()
    s.materials[1].kd = float("inf")
    d, keep = s.to_desc()
    assert L.nt_scene_create(C.byref(d), 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert b"material 1" in L.nt_last_error()
    s, cam = scenes.cornell_box()
    s.materials[2].ior = 0.0  # 1 / ior is a derived value of every material (SPEC-PROVISIONAL §1)
    d, keep = s.to_desc()
    assert L.nt_scene_create(C.byref(d), 0, C.byref(h)) == abi.NT_ERR_INVALID
    assert b"index of refraction" in L.nt_last_error()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under nettracer_b200/ may import, link or exec it."""
    pkg = os.path.join(ROOT, "nettracer_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|\bnto_[a-z_]+\s*\(|oracle/|oracle\.oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), os.path.join(dirpath, f)
    ldd = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_kernels_are_sm100a_and_strict_unit_forbids_fma():
    """The library carries sm_100a SASS only, and the strict binary64 translation unit is compiled with
    -fmad=false (bit-exact agreement with the oracle on the GPU is the functional proof)."""
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    mk = open(os.path.join(ROOT, "nettracer_b200", "csrc", "Makefile")).read()
    rule = "\n".join(mk[mk.index("nt_kernels_f64.o:"):].splitlines()[:2])
    assert "-fmad=false" in rule and "use_fast_math" not in rule
