#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the intersect-and-shade hot path (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f64|f32] [--workload NAME]
                  [--mode gather|p2p_store] [--impl reference]

A step = one frame of the workload (default: BASELINE.json configs[2], the configuration the metric
is quoted on: Cornell-style box, 1920x1080, 4 spp, depth 5).  N > 1 shards the SAME frame across the
ranks in interleaved row bands (strong scaling) and lands it on rank 0 with one exchange per frame.

value        whole-job Mrays/s, frame resident on the device: rays cast per frame (counted on the
             device, deterministic) / mean device time per step (CUDA events, max over ranks).
e2e          the same metric through the reference-facing call nt_render with a pinned HOST output
             buffer (N = 1) or sharded render + exchange + device->host copy on rank 0 (N > 1), timed
             by the host clock around blocking calls.
roofline     the render kernel against the measured FP64 (strict) / FP32 (fast) issue-rate peak
             (nt_measure_peaks, same process, same clocks): algorithmic flops by SURVEY.md §8(d)'s
             convention / kernel time.  The path is FP-pipe bound; DRAM traffic is ~0.
cpu_baseline the CPU oracle (a port of SPEC-PROVISIONAL.md — NOT NetTracer, whose sources do not exist
             here) on all host threads over a bounded sample of the same frame.
--impl reference  times that same oracle alone (there is no reference implementation to run:
             /root/reference holds one README and the image has no JVM).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s"
DEFAULT_WORKLOAD = "cfg3_cornell_1080p_4spp_d5"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--mode", default="p2p_store", choices=["gather", "p2p_store"])
    ap.add_argument("--band-rows", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bvh-build", default="host", choices=["host", "gpu"],
                    help="BVH scenes: binned-SAH build on the host (default) or LBVH build on the GPU")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU; only samples taken between
    begin() and end() (the timed regions) are summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.windows = index, [], []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sel = [s for t, s in self.samples if any(a <= t <= (b or t) for a, b in self.windows)]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(sel),
               "window": "device-resident and e2e timed regions"}
        try:
            sm = sorted(float(s[0]) for s in sel)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(sel[0][1])
                out["power_w_max"] = max(float(s[2]) for s in sel)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(s[3 + i].lower().startswith("active") for s in sel):
                    out["reasons"].append(n)
        except Exception:
            pass
        return out


_json_out = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner on the
    first communicator) is sent to stderr by pointing fd 1 at fd 2; the line itself goes to a private copy of fd 1."""
    global _json_out
    if _json_out is None:
        sys.stdout.flush()
        _json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _json_out or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload(name):
    from nettracer_b200 import scenes
    factory, w, h, spp, depth = scenes.CONFIGS[name]
    scene, cam = factory()
    return scene, cam, w, h, spp, depth


def cpu_oracle_sample(scene, cam, w, h, spp, depth, target_s):
    """All host threads over every k-th image row of the frame (k = 1 when a whole frame is cheap),
    repeated until ~target_s of CPU time has been spent."""
    from nettracer_b200.scene import make_params
    from oracle import oracle
    accel = 1 if len(scene.spheres) + len(scene.triangles) > 64 else 0
    p = make_params(w, h, spp, depth, cam.resolve(w, h))
    k = 64
    t0 = time.perf_counter()
    oracle.render(scene, p, accel=accel, row_step=k)
    dt = time.perf_counter() - t0
    k = max(1, min(64, int(k * dt / target_s * 1.1) or 1))  # densest row sample that fits the target
    reps, total, rays = 0, 0.0, 0
    while total < target_s and reps < 200:
        t0 = time.perf_counter()
        _, st = oracle.render(scene, p, accel=accel, row_step=k)
        total += time.perf_counter() - t0
        rays += st["rays"]
        reps += 1
    rows = len(range(0, h, k))
    return {"value": rays / total / 1e6, "unit": METRIC, "cores": oracle.max_threads(), "kind": "port",
            "sample": f"{reps} x every {k}th row ({rows} of {h} rows) of {w}x{h} {spp}spp depth{depth}: {rays} rays in "
                      f"{total:.1f} s, oracle accel={'bvh' if accel else 'brute force'}",
            "seconds": total, "rays": rays, "row_step": k, "ms_per_frame_extrapolated": 1e3 * total / reps * h / rows,
            "note": "CPU oracle = port of SPEC-PROVISIONAL.md, NOT NetTracer (no reference source, no JVM)"}


def run_reference(a):
    """--impl reference: the oracle alone (the only CPU implementation of this path that exists)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, cam, w, h, spp, depth = workload(a.workload)
    from nettracer_b200.scene import make_params
    from oracle import oracle
    accel = 1 if len(scene.spheres) + len(scene.triangles) > 64 else 0
    p = make_params(w, h, spp, depth, cam.resolve(w, h))
    # size the per-step sample so that warmup+steps stay within ~2 minutes
    k = 64
    t0 = time.perf_counter()
    _, st = oracle.render(scene, p, accel=accel, row_step=k)
    dt = time.perf_counter() - t0
    budget = 100.0 / max(1, a.steps + a.warmup)
    k = max(1, min(64, int(k * dt / budget) + 1))
    times, rays = [], 0
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        _, st = oracle.render(scene, p, accel=accel, row_step=k)
        dt = time.perf_counter() - t0
        if i >= a.warmup:
            times.append(dt)
            rays = st["rays"]
    ms = 1e3 * sum(times) / len(times)
    val = rays / (ms * 1e-3) / 1e6
    rows = len(range(0, h, k))
    frame_ms = ms * h / rows
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "ms_per_frame_extrapolated": frame_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": a.workload, "width": w, "height": h, "spp": spp, "max_depth": depth,
                       "sample": f"every {k}th row ({rows} of {h})"},
            "cpu_baseline": {"value": val, "unit": METRIC, "cores": oracle.max_threads(), "kind": "port",
                             "sample": f"every {k}th row ({rows} of {h} rows) per step",
                             "note": "no NetTracer source or JVM exists here; this is the SPEC-PROVISIONAL oracle"},
            "e2e": {"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    a = parse()
    claim_stdout()
    if a.impl == "reference":
        run_reference(a)
        return
    import ctypes as C

    import torch
    import torch.distributed as dist

    from nettracer_b200 import abi
    from nettracer_b200.renderer import measure_peaks
    from nettracer_b200.scene import make_params
    from nettracer_b200.sharded import CudaBackend, ShardedRenderer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the hot path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if a.bvh_build == "gpu":
        os.environ["NT_BVH_BUILD"] = "gpu"
    scene, cam, w, h, spp, depth = workload(a.workload)
    prec = abi.NT_F64_STRICT if a.precision == "f64" else abi.NT_F32_FAST
    t0 = time.perf_counter()
    backend = CudaBackend(scene, local)
    scene_create_s = time.perf_counter() - t0
    sr = ShardedRenderer(backend, rank, world, band_rows=a.band_rows, mode=a.mode)
    camera = cam.resolve(w, h)
    sp = sr.shard_params(w, h, spp, depth, camera, prec)
    info = backend.renderer.info()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    peaks = measure_peaks(local) if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---------- device-resident timing ----------
    def step():
        return sr.render(sp)

    for _ in range(a.warmup):
        step()
        flush.zero_()
    barrier()
    # per-frame work of this rank (deterministic): rays and algorithmic flops
    st = backend.stats()
    rays_total = sum_over_ranks(float(st["rays"]))
    flops_local = float(abi.algorithmic_flops(st))
    flops_total = sum_over_ranks(flops_local)

    if sampler:
        sampler.begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(a.steps)]
    barrier()
    wall0 = time.perf_counter()
    for e0, ek, e1 in ev:
        flush.zero_()  # L2 flush between timed steps, outside the step's event pair
        e0.record()
        sr.render(sp, kernel_done=ek)  # render kernel | ek | exchange (+ deinterleave)
        e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    if sampler:
        sampler.end()
    step_ms = [e0.elapsed_time(e1) for e0, ek, e1 in ev]
    kern_ms = [e0.elapsed_time(ek) for e0, ek, e1 in ev]
    ms_per_step = max_over_ranks(sum(step_ms) / len(step_ms))
    kernel_ms_local = sum(kern_ms) / len(kern_ms)
    kernel_ms = max_over_ranks(kernel_ms_local)
    value = rays_total / (ms_per_step * 1e-3) / 1e6

    # ---------- end-to-end through the host-buffer call ----------
    host = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
    e2e_times = []
    for i in range(a.warmup + a.steps):
        if i == a.warmup and sampler:
            sampler.begin()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            full_p = make_params(w, h, spp, depth, camera, prec)
            stats = abi.nt_render_stats()
            from nettracer_b200.lib import check, load
            check(load().nt_render(backend.renderer._h, C.byref(full_p), C.c_void_p(host.data_ptr()), w * 4, C.byref(stats)))
        else:
            full = sr.render(sp)
            if rank == 0:
                host.copy_(full, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= a.warmup:
            e2e_times.append(max_over_ranks(dt))
    if sampler:
        sampler.end()
    clocks = sampler.stop() if sampler else None
    e2e_ms = 1e3 * sum(e2e_times) / len(e2e_times)
    e2e_val = rays_total / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        strict = a.precision == "f64"
        per_frame_launches = backend.renderer.info()["last_launches"]
        peak_gf = peaks["f64_nofma_gflops"] if strict else peaks["f32_fma_gflops"]
        achieved_tf = flops_local / (kernel_ms_local * 1e-3) / 1e12
        roof = {"bound": "fp64" if strict else "fp32", "achieved": achieved_tf, "peak": peak_gf / 1e3, "unit": "TFLOP/s",
                "frac": achieved_tf / (peak_gf / 1e3),
                "traffic": ({"cfg3_cornell_1080p_4spp_d5": 7.423e7}.get(a.workload) if strict else None),
                "traffic_note": "dram read+write bytes of one render launch (ncu --set full, profiles/r01j_cfg3_f64_final.md): "
                                "2.3 MB read + 71.9 MB written, of which 8.3 MB is the frame and the rest evicted local-memory "
                                "(ray-tree stack and spill) lines; 1.5 % of HBM bandwidth - the kernel is FP-issue bound, not HBM bound",
                "kernel": ("wf_trace_kernel + wf_shade_kernel (wavefront pipeline, all kernels of a frame)" if info["uses_bvh"] and per_frame_launches > 2
                           else "render_%skernel<%s>" % ("bvh_" if info["uses_bvh"] else "", "double" if strict else "float")),
                "kernel_ms": kernel_ms_local, "algorithmic_flops_per_launch": flops_local,
                "peak_source": "nt_measure_peaks in this process: " + ("FP64 mul/add issue rate without FMA (strict mode may not fuse)"
                                                                        if strict else "FP32 FMA issue rate"),
                "peaks_measured_gflops": peaks,
                "note": "path is FP-pipe bound, not HBM/tensor (BASELINE.json north_star); flop convention SURVEY.md §8(d); "
                        "traffic: see profiles/ (DRAM bytes per launch are ~0.3 MB)"}
        line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "ms_per_frame": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
                "config": {"workload": a.workload, "width": w, "height": h, "spp": spp, "max_depth": depth,
                           "rays_per_frame": rays_total, "sharding": f"{world} x interleaved {a.band_rows}-row bands" if world > 1 else "none",
                           "exchange": (sr.mode if world > 1 else "none"), "uses_bvh": info["uses_bvh"],
                           "l2": "256 MiB memset between timed steps (outside each step's event pair)",
                           "scene_create_s": scene_create_s, "bvh_on_gpu": info.get("bvh_on_gpu", False),
                           "bvh_build_ms": info.get("bvh_build_ms", 0.0), "bvh_nodes": info["bvh_nodes"]},
                "e2e": {"value": e2e_val, "unit": METRIC, "ms_per_frame": e2e_ms, "h2d_bytes_per_step": C.sizeof(abi.nt_render_params),
                        "d2h_bytes_per_step": h * w * 4 + 8 * 8 * 32,
                        "api": "nt_render -> pinned host RGBA8" if world == 1 else "ShardedRenderer.render + D2H on rank 0"},
                "gpu_launches": a.steps * (per_frame_launches + (1 if sr.mode == "gather" and world > 1 else 0)),
                "gpu_launches_note": "per frame and rank, counted by the library (nt_scene_info): flat scenes 1 render kernel; BVH scenes "
                                     "the wavefront pipeline's trace/shadow/shade kernels per level and chunk + sum + resolve "
                                     "(or 2 with NT_WAVEFRONT=0); + deinterleave on rank 0 in gather mode",
                "kernel_ms": kernel_ms, "wall_s_timed_region": wall, "clocks": clocks, "roofline": roof}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_sample(scene, cam, w, h, spp, depth, a.cpu_seconds)
        emit(line)
    barrier()
    sr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
