#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the intersect-and-shade hot path (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f64|f32] [--workload NAME]
                  [--mode gather|p2p_store] [--impl reference] [--no-secondary]

A step = one frame of the workload (default: BASELINE.json configs[2], the configuration the metric
is quoted on: Cornell-style box, 1920x1080, 4 spp, depth 5).  N > 1 shards the SAME frame across the
ranks in interleaved row bands (strong scaling) and lands it on rank 0 with one exchange per frame.

value        whole-job Mrays/s, frame resident on the device of rank 0: rays cast per frame (counted on
             the device, deterministic) / mean device time per step (CUDA events on the launching stream,
             max over ranks; a step = render kernel + exchange: peer stores over NVLink + flag wait).
e2e          the same metric through the reference-facing call with a HOST frame: nt_render into pinned
             memory (N = 1); ShardedRenderer.render_host (N > 1) - every rank's kernel stores its bands
             into one shared page-locked host frame and posts its flag there - timed by rank 0's host clock.
frame_check  after the timed regions rank 0 compares the device frame AND the host frame with the CPU
             oracle's frame of the same parameters; a mismatch makes the run fail (exit code 1).
roofline     the render kernel against the measured FP64 (strict) / FP32 (fast) issue-rate peak
             (nt_measure_peaks, same process, same clocks): `frac` = algorithmic flops by SURVEY.md
             §8(d)'s convention (every query credited with every primitive) / kernel time / peak;
             `executed_frac` = the same weights on the tests the kernel really starts (instrumented twin
             of the kernel, one extra launch); `frac_vs_fma_peak` = against the FMA rate the strict mode
             may not use.  The path is FP-issue bound, not HBM or tensor bound.
secondary    bounded runs of the other BASELINE.json configs (outside the headline's timed region).
cpu_baseline the CPU oracle (a port of SPEC-PROVISIONAL.md — NOT NetTracer, whose sources do not exist
             here) on all host threads over a bounded sample of the same frame.
--impl reference  times that same oracle alone (there is no reference implementation to run:
             /root/reference holds one README and the image has no JVM).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s"
DEFAULT_WORKLOAD = "cfg3_cornell_1080p_4spp_d5"
L2_NOTE = "GPU arm: 256 MiB memset between timed steps, outside each step's event pair; CPU arm: not applicable"


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the bounded runs of the other configs")
    ap.add_argument("--no-frame-check", action="store_true")
    ap.add_argument("--bvh-build", default="host", choices=["host", "gpu"],
                    help="BVH scenes: binned-SAH build on the host (default) or LBVH build on the GPU")
    return ap.parse_args()


def config_of(name, w, h, spp, depth):
    """The SAME dictionary in both arms, so that the driver's config comparison sees one configuration."""
    return {"workload": name, "width": w, "height": h, "spp": spp, "max_depth": depth, "l2": L2_NOTE}


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons of one GPU through NVML, in this process, every ~2 ms; only samples
    taken between begin() and end() (the timed regions) are summarised.  (Round 1 polled nvidia-smi every 50 ms and
    never landed a sample inside a 15 ms timed region.)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.windows = index, [], []
        self._stop_flag = False
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_flag:
                t = time.perf_counter()
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                if any(a <= t and b is None for a, b in self.windows):  # the costlier queries only inside a window
                    self.samples.append((t, sm, reasons_fn(h), nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.002)
        except Exception as e:  # no NVML: the line says so instead of inventing clocks
            self.error = repr(e)

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def stop(self):
        self._stop_flag = True
        self.join(timeout=1.0)
        out = {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_sm", None), "reasons": [], "samples": len(self.samples),
               "source": "NVML in-process, 2 ms period", "window": "device-resident and e2e timed regions"}
        if self.error:
            out["error"] = self.error
        if self.samples:
            sm = sorted(s[1] for s in self.samples)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_mhz_min"] = sm[0]
            out["power_w_max"] = max(s[3] for s in self.samples)
            bits = 0
            for s in self.samples:
                bits |= int(s[2])
            # NVML reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal, 0x40 hw_thermal, 0x80 hw_power_brake
            for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                              (0x4, "sw_power_cap"), (0x80, "hw_power_brake_slowdown")):
                if bits & bit:
                    out["reasons"].append(name)
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


def uses_accel(scene):
    return 1 if len(scene.spheres) + len(scene.triangles) > 64 else 0


def cpu_oracle_sample(scene, cam, w, h, spp, depth, target_s):
    """All host threads over every k-th image row of the frame (k = 1 when a whole frame is cheap),
    repeated until ~target_s of CPU time has been spent."""
    from nettracer_b200.scene import make_params
    from oracle import oracle
    accel = uses_accel(scene)
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
    accel = uses_accel(scene)
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
    sample = f"every {k}th row ({rows} of {h} rows) per step" if k > 1 else "the whole frame per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "ms_per_frame_extrapolated": frame_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(a.workload, w, h, spp, depth),
            "cpu_baseline": {"value": val, "unit": METRIC, "cores": oracle.max_threads(), "kind": "port", "sample": sample,
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

    import numpy as np
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

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if world > 1 else x

    if a.bvh_build == "gpu":
        os.environ["NT_BVH_BUILD"] = "gpu"
    prec = abi.NT_F64_STRICT if a.precision == "f64" else abi.NT_F32_FAST
    strict = a.precision == "f64"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    peaks = measure_peaks(local) if rank == 0 else None
    peak_gf = (peaks["f64_nofma_gflops"] if strict else peaks["f32_fma_gflops"]) if peaks else None

    def measure(name, steps, warmup, want_e2e, sample_clocks):
        """One workload, sharded over the ranks: device-resident timing, optional end-to-end timing, frame check data.
        Returns a dict on every rank (aggregates are identical on all ranks)."""
        scene, cam, w, h, spp, depth = workload(name)
        t0 = time.perf_counter()
        backend = CudaBackend(scene, local)
        scene_create_s = time.perf_counter() - t0
        sr = ShardedRenderer(backend, rank, world, band_rows=a.band_rows, mode=a.mode)
        camera = cam.resolve(w, h)
        sp = sr.shard_params(w, h, spp, depth, camera, prec)
        info = backend.renderer.info()
        for _ in range(warmup):
            sr.render(sp)
            flush.zero_()
        barrier()
        # per-frame work of this rank (deterministic): rays and algorithmic flops
        st = backend.stats()
        rays_total = sum_over_ranks(float(st["rays"]))
        flops_local = float(abi.algorithmic_flops(st))
        launches = backend.renderer.info()["last_launches"] + (1 if world > 1 and rank == 0 and sr.mode == "p2p_store" else 0)
        # executed work: one launch of the instrumented twin (flat scenes; BVH counters are executed counts already)
        xp = sr.shard_params(w, h, spp, depth, camera, prec)
        xp.flags = abi.NT_RENDER_COUNT_EXECUTED
        sr.render(xp)
        barrier()
        xst = backend.stats()
        executed_local = float(abi.executed_flops(xst)) if not info["uses_bvh"] else flops_local
        sr.render(sp)   # the frame the check below reads is a production frame again
        barrier()

        if sampler and sample_clocks:
            sampler.begin()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        barrier()
        wall0 = time.perf_counter()
        full = None
        for e0, ek, e1 in ev:
            flush.zero_()  # L2 flush between timed steps, outside the step's event pair
            e0.record()
            full = sr.render(sp, kernel_done=ek)  # render kernel (+ flag posts) | ek | exchange: flag wait / gather
            e1.record()
        barrier()
        wall = time.perf_counter() - wall0
        if sampler and sample_clocks:
            sampler.end()
        ms_per_step = max_over_ranks(sum(e0.elapsed_time(e1) for e0, ek, e1 in ev) / steps)
        kernel_ms_local = sum(e0.elapsed_time(ek) for e0, ek, e1 in ev) / steps
        res = {"name": name, "w": w, "h": h, "spp": spp, "depth": depth, "info": info, "scene_create_s": scene_create_s,
               "rays_total": rays_total, "flops_local": flops_local, "executed_local": executed_local,
               "ms_per_step": ms_per_step, "kernel_ms_local": kernel_ms_local, "kernel_ms": max_over_ranks(kernel_ms_local),
               "launches": launches, "wall": wall, "exchange": sr.mode if world > 1 else "none", "scene": scene, "cam": cam}
        res["value"] = rays_total / (ms_per_step * 1e-3) / 1e6
        dev_frame = full.cpu().numpy() if rank == 0 else None

        # ---------- end to end: the frame in host memory ----------
        host_frame = None
        if want_e2e:
            hp = make_params(w, h, spp, depth, camera, prec, shard_index=rank, shard_count=world, band_rows=a.band_rows)
            host = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() if world == 1 else None
            from nettracer_b200.lib import check, load

            def e2e_frame():
                if world == 1:  # the image only (stats = NULL): rays per frame are known from the device-timed steps
                    check(load().nt_render(backend.renderer._h, C.byref(hp), C.c_void_p(host.data_ptr()), w * 4, None))
                    return host.numpy()
                return sr.render_host(hp)[0]
            for _ in range(warmup):
                e2e_frame()
            barrier()
            if sampler and sample_clocks:
                sampler.begin()
            t0 = time.perf_counter()
            for _ in range(steps):
                host_frame = e2e_frame()   # rank 0 returns when every rank's bands are in the host frame
            e2e_ms_local = 1e3 * (time.perf_counter() - t0) / steps
            if sampler and sample_clocks:
                sampler.end()
            if rank == 0 and host_frame is not None:
                host_frame = np.array(host_frame)
            barrier()
            res["e2e_ms"] = max_over_ranks(e2e_ms_local)
        res["dev_frame"], res["host_frame"] = dev_frame, host_frame

        # ---------- two frames in flight (extra; NOT the headline): consecutive frames on two streams ----------
        # A shard's launch ends on the latency of its deepest tile; an animation does not have to
        # wait for it - frame f + 1 can fill the SMs frame f's tail leaves idle.  Two independent ShardedRenderers (own
        # scene copy, own frame buffers and flags), even frames on one stream, odd frames on the other; the time is the
        # bracket over all frames / their number, no L2 flush inside the bracket (it would serialise the two streams).
        if want_e2e and not info["uses_bvh"]:
            backend_b = CudaBackend(scene, local)
            sr_b = ShardedRenderer(backend_b, rank, world, band_rows=a.band_rows, mode=a.mode)
            streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
            pair = [(sr, streams[0]), (sr_b, streams[1])]
            for i in range(4):
                with torch.cuda.stream(pair[i & 1][1]):
                    pair[i & 1][0].render(sp)
            barrier()
            n_pipe = max(steps, 20)
            t0 = time.perf_counter()
            e_beg = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e_end = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for i in range(2):
                e_beg[i].record(streams[i])
            for i in range(n_pipe):
                with torch.cuda.stream(pair[i & 1][1]):
                    pair[i & 1][0].render(sp)
            for i in range(2):
                e_end[i].record(streams[i])
            barrier()
            span = max(e_beg[0].elapsed_time(e_end[0]), e_beg[0].elapsed_time(e_end[1]), e_beg[1].elapsed_time(e_end[0]),
                       e_beg[1].elapsed_time(e_end[1]))
            ms_pipe = max_over_ranks(span / n_pipe)
            res["pipelined"] = {"frames_in_flight": 2, "frames": n_pipe, "ms_per_frame": ms_pipe,
                                "Mrays_per_s": rays_total / (ms_pipe * 1e-3) / 1e6,
                                "note": "throughput of consecutive frames on two streams (device-resident, no L2 flush inside "
                                        "the bracket); the headline `value` is one frame at a time"}
            sr_b.close()
        sr.close()
        return res

    def frame_check(res):
        """Rank 0: device frame and host frame against the oracle's frame (every k-th row when a whole frame would
        take the host too long)."""
        from oracle import oracle
        scene, cam, w, h, spp, depth = res["scene"], res["cam"], res["w"], res["h"], res["spp"], res["depth"]
        accel = uses_accel(scene)
        p = make_params(w, h, spp, depth, cam.resolve(w, h))
        k = 1 if not accel else max(1, -(-(w * h * spp) // 2_000_000))  # BVH scenes: ~2 M samples of oracle work
        t0 = time.perf_counter()
        ref, rst = oracle.render(scene, p, accel=accel, row_step=k)
        rows = slice(0, h, k)
        out = {"sha256_device_frame": hashlib.sha256(res["dev_frame"].tobytes()).hexdigest()[:16],
               "rows_checked": len(range(0, h, k)), "oracle_s": time.perf_counter() - t0}
        diff = np.abs(res["dev_frame"][rows].astype(np.int16) - ref[rows].astype(np.int16)).max(axis=-1)
        # strict mode: any difference counts; fast mode (no bit contract, SPEC section 7): pixels more than 1 LSB off (flat
        # scenes) / 2 LSB off (BVH scenes), the tolerances of tests/test_parity_gpu.py::test_fast_mode_tolerance*
        lsb = 0 if strict else (2 if accel else 1)
        out["diff_pixels_vs_oracle"] = int((diff > lsb).sum())
        out["max_abs_diff"] = int(diff.max())
        if k == 1 and strict:
            out["rays_equal_oracle"] = bool(rst["rays"] == int(res["rays_total"]))
        if res["host_frame"] is not None:
            out["host_frame_equals_device_frame"] = bool(np.array_equal(res["host_frame"], res["dev_frame"]))
        # strict: `pow` may move <= 2 pixels by 1 LSB; fast: >= 99.5 % of the pixels within 1 LSB (tests/test_parity_gpu.py)
        tol = 2 if strict else max(2, int((0.01 if accel else 0.005) * out["rows_checked"] * w))
        out["tolerance_pixels"] = tol
        out["ok"] = bool(out["diff_pixels_vs_oracle"] <= tol and out.get("host_frame_equals_device_frame", True)
                         and out.get("rays_equal_oracle", True))
        return out

    main_res = measure(a.workload, a.steps, a.warmup, True, True)
    clocks = sampler.stop() if sampler else None

    secondary = []
    if not a.no_secondary:
        # bounded runs of the other configs (BASELINE.json configs[1], [3], and [4] from 4 GPUs up), a few frames each
        plan = [("cfg2_cornell_1080p_1spp_d1", 20, 3), ("cfg4_mesh1m_4k_4spp_d3", 3, 2)]
        if world >= 4:
            plan.append(("cfg5_mesh1m_8k_16spp_d5", 2, 1))
        for name, steps, warmup in plan:
            if name == a.workload:
                continue
            r = measure(name, steps, warmup, False, False)
            entry = {"workload": name, "frames": steps, "ms_per_frame": r["ms_per_step"], "Mrays_per_s": r["value"],
                     "rays_per_frame": r["rays_total"], "kernel_ms": r["kernel_ms"], "launches_per_frame": r["launches"],
                     "uses_bvh": r["info"]["uses_bvh"], "scene_create_s": r["scene_create_s"], "dtype": a.precision}
            if rank == 0:
                entry["roofline_frac"] = r["flops_local"] / (r["kernel_ms_local"] * 1e-3) / 1e9 / peak_gf
                if not a.no_frame_check:
                    entry["frame_check"] = frame_check(r)
            secondary.append(entry)

    ok = True
    if rank == 0:
        r = main_res
        w, h, spp, depth, info = r["w"], r["h"], r["spp"], r["depth"], r["info"]
        achieved_tf = r["flops_local"] / (r["kernel_ms_local"] * 1e-3) / 1e12
        executed_tf = r["executed_local"] / (r["kernel_ms_local"] * 1e-3) / 1e12
        fma_peak_tf = (peaks["f64_fma_gflops"] if strict else peaks["f32_fma_gflops"]) / 1e3
        traffic = {"cfg3_cornell_1080p_4spp_d5": 5.380e7}.get(a.workload) if strict and world == 1 else None
        roof = {"bound": "fp64" if strict else "fp32", "achieved": achieved_tf, "peak": peak_gf / 1e3, "unit": "TFLOP/s",
                "frac": achieved_tf / (peak_gf / 1e3),
                "executed": executed_tf, "executed_frac": executed_tf / (peak_gf / 1e3),
                "frac_vs_fma_peak": achieved_tf / fma_peak_tf, "executed_frac_vs_fma_peak": executed_tf / fma_peak_tf,
                "traffic": traffic,
                "traffic_source": ("profile constant, not measured in this run: dram read + write bytes of one render launch from "
                                   "ncu --set full (profiles/r02w_cfg3_f64_lean.md, the current kernel): 2.2 MB read + 51.6 MB written, of which "
                                   "8.3 MB is the frame and the rest local-memory (ray-stack, spill) lines evicted from L2; "
                                   "1.0 % of HBM bandwidth") if traffic else None,
                "kernel": ("wf_trace_kernel + wf_shade_kernel (wavefront pipeline, all kernels of a frame)" if info["uses_bvh"] and r["launches"] > 3
                           else "render_%skernel<%s>" % ("bvh_" if info["uses_bvh"] else "", "double" if strict else "float")),
                "kernel_ms": r["kernel_ms_local"], "algorithmic_flops_per_launch": r["flops_local"],
                "executed_flops_per_launch": r["executed_local"],
                "peak_source": "nt_measure_peaks in this process: " + ("FP64 mul/add issue rate without FMA (strict mode may not fuse)"
                                                                        if strict else "FP32 FMA issue rate"),
                "peaks_measured_gflops": peaks,
                "note": "path is FP-issue bound, not HBM/tensor (BASELINE.json north_star); flop convention SURVEY.md §8(d); `frac` "
                        "credits every query with every primitive (algorithmic work), `executed_frac` counts the tests the kernel "
                        "starts after culling; rank 0's shard when n_gpus > 1"}
        e2e_val = r["rays_total"] / (r["e2e_ms"] * 1e-3) / 1e6
        line = {"metric": METRIC, "value": r["value"], "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": r["ms_per_step"], "ms_per_frame": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
                "config": config_of(a.workload, w, h, spp, depth),
                "run": {"rays_per_frame": r["rays_total"],
                        "sharding": f"{world} x interleaved {a.band_rows}-row bands" if world > 1 else "none",
                        "exchange": ("peer stores over NVLink into rank 0's double-buffered frame + device flags (no collective)"
                                     if r["exchange"] == "p2p_store" else r["exchange"]),
                        "uses_bvh": info["uses_bvh"], "scene_create_s": r["scene_create_s"], "bvh_on_gpu": info.get("bvh_on_gpu", False),
                        "bvh_build_ms": info.get("bvh_build_ms", 0.0), "bvh_nodes": info["bvh_nodes"]},
                "e2e": {"value": e2e_val, "unit": METRIC, "ms_per_frame": r["e2e_ms"], "h2d_bytes_per_step": C.sizeof(abi.nt_render_params) * world,
                        "d2h_bytes_per_step": h * w * 4 + 8 * abi_counter_bytes(abi) * world,
                        "api": ("nt_render -> pinned host RGBA8 (zero-copy stores from the kernel)" if world == 1 else
                                "ShardedRenderer.render_host: every rank's kernel stores its bands into ONE shared page-locked host "
                                "frame (nt_host_frame_*) and posts its completion flag there; rank 0's host spins on the flags; "
                                "rank 0's clock")},
                "gpu_launches": a.steps * r["launches"],
                "gpu_launches_note": "per frame on rank 0, counted by the library (nt_scene_info): flat scenes 1 render kernel (exchange flags "
                                     "are posted inside it); BVH scenes the 8 small kernels of the eye grid + the wavefront pipeline's kernels per level and chunk "
                                     "+ sum + resolve; + 1 flag-wait kernel on rank 0 when n_gpus > 1 (p2p_store) / + deinterleave (gather)",
                "kernel_ms": r["kernel_ms"], "wall_s_timed_region": r["wall"], "clocks": clocks, "roofline": roof}
        if not a.no_frame_check:
            line["frame_check"] = frame_check(r)
            ok = line["frame_check"]["ok"] and all(s.get("frame_check", {"ok": True})["ok"] for s in secondary)
        if "pipelined" in r:
            line["pipelined"] = r["pipelined"]
        if secondary:
            line["secondary"] = secondary
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_sample(r["scene"], r["cam"], w, h, spp, depth, a.cpu_seconds)
        emit(line)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: FRAME CHECK FAILED - the rendered frame differs from the oracle (see frame_check)\n")
        sys.exit(1)


def abi_counter_bytes(abi):
    return 11 * 32  # NT_NCOUNTERS x NT_COUNTER_SLOTS 64-bit counters read back per call


if __name__ == "__main__":
    main()
