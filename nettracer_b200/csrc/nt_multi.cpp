// nt_multi.cpp — several GPUs behind one C-ABI call (include/nettracer_b200.h nt_multi_*; SURVEY.md §8(b)'s "gpu list",
// §8(e)): one host process, the scene replicated per device, one persistent host thread per device that calls nt_render
// for its shard of interleaved row bands.  Every GPU's render kernel stores its RGBA8 words straight into ONE pinned host
// frame over its own PCIe link (the caller's buffer when that is page-locked, else a staging frame of the library), so
// there is no gather and no device-to-host copy.  Replaces the reference's thread / network tile distribution, which
// cannot be cited (/root/reference/README:1-3 holds no code).  Host logic only; no CPU rendering path exists.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nettracer_b200.h"

int nt_fail_public(int code, const char *fmt, ...); // nt_api.cu: sets nt_last_error() of the calling thread

struct nt_multi {
    int n = 0;
    std::vector<nt_scene *> scenes;
    std::vector<int> devices;
    std::vector<std::thread> workers;
    // job hand-over: the caller bumps `gen`; workers spin on it briefly (a frame of the metric's configuration takes
    // ~0.1 ms per GPU, a condition-variable wake-up alone is 20-50 us), then sleep on the condition variable
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::atomic<uint64_t> gen{ 0 };
    std::atomic<int> pending{ 0 };
    bool quit = false;
    nt_render_params params{};
    uint8_t *target = nullptr; // pinned frame the kernels store into
    size_t target_stride = 0;
    uint8_t *copy_out = nullptr; // pageable caller buffer: every worker copies its own bands out of the staging frame
    size_t copy_stride = 0;
    std::vector<int> rc;
    std::vector<nt_render_stats> stats;
    std::vector<std::string> err;
    uint8_t *staging = nullptr;
    size_t staging_bytes = 0;
    std::mutex call_mu; // one nt_multi_render at a time
};

static void worker_main(nt_multi *m, int i) {
    cudaSetDevice(m->devices[i]);
    uint64_t seen = 0;
    for (;;) {
        // wait for the next job
        const auto t0 = std::chrono::steady_clock::now();
        while (m->gen.load(std::memory_order_acquire) == seen) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                std::unique_lock<std::mutex> lk(m->mu);
                m->cv_go.wait(lk, [&] { return m->gen.load(std::memory_order_acquire) != seen || m->quit; });
                break;
            }
        }
        if (m->quit) return;
        seen = m->gen.load(std::memory_order_acquire);
        nt_render_params p = m->params;
        p.shard_index = (uint32_t)i; p.shard_count = (uint32_t)m->n; p.layout = NT_LAYOUT_FULL;
        m->rc[i] = nt_render(m->scenes[i], &p, m->target, m->target_stride, &m->stats[i]);
        if (m->rc[i]) m->err[i] = nt_last_error();
        else if (m->copy_out) { // staging frame -> the caller's pageable buffer, this shard's bands only
            const uint32_t band = p.band_rows, nb = (p.height + band - 1) / band;
            for (uint32_t b = (uint32_t)i; b < nb; b += (uint32_t)m->n) {
                const uint32_t y0 = b * band, y1 = y0 + band > p.height ? p.height : y0 + band;
                for (uint32_t y = y0; y < y1; ++y)
                    memcpy(m->copy_out + (size_t)y * m->copy_stride, m->target + (size_t)y * m->target_stride, (size_t)p.width * 4);
            }
        }
        if (m->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(m->mu);
            m->cv_done.notify_all();
        }
    }
}

extern "C" void nt_multi_destroy(nt_multi *m) {
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
        m->gen.fetch_add(1, std::memory_order_release);
    }
    m->cv_go.notify_all();
    for (auto &t : m->workers)
        if (t.joinable()) t.join();
    for (nt_scene *s : m->scenes) nt_scene_destroy(s);
    if (m->staging) cudaFreeHost(m->staging);
    delete m;
}

extern "C" int nt_multi_create(const nt_scene_desc *desc, const int *devices, int n_devices, nt_multi **out) {
    if (!out) return nt_fail_public(NT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return nt_fail_public(NT_ERR_INVALID, "need 1..64 devices");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return nt_fail_public(NT_ERR_INVALID, "device %d listed twice", devices[i]);
    nt_multi *m = new (std::nothrow) nt_multi;
    if (!m) return nt_fail_public(NT_ERR_NOMEM, "out of host memory");
    m->n = n_devices;
    m->devices.assign(devices, devices + n_devices);
    m->scenes.assign((size_t)n_devices, nullptr);
    m->rc.assign((size_t)n_devices, 0);
    m->stats.assign((size_t)n_devices, nt_render_stats{});
    m->err.assign((size_t)n_devices, std::string());
    // replicate the scene: one thread per device (BVH builds and uploads run side by side)
    {
        std::vector<std::thread> th;
        for (int i = 0; i < n_devices; ++i)
            th.emplace_back([m, desc, i] {
                m->rc[i] = nt_scene_create(desc, m->devices[i], &m->scenes[i]);
                if (m->rc[i]) m->err[i] = nt_last_error();
            });
        for (auto &t : th) t.join();
    }
    for (int i = 0; i < n_devices; ++i)
        if (m->rc[i]) {
            const int rc = m->rc[i];
            const std::string e = m->err[i];
            nt_multi_destroy(m);
            return nt_fail_public(rc, "device %d: %s", devices[i], e.c_str());
        }
    try {
        for (int i = 0; i < n_devices; ++i) m->workers.emplace_back(worker_main, m, i);
    } catch (...) {
        nt_multi_destroy(m);
        return nt_fail_public(NT_ERR_SYSTEM, "could not start the per-device host threads");
    }
    *out = m;
    return NT_OK;
}

extern "C" int nt_multi_device_count(const nt_multi *m) { return m ? m->n : 0; }

extern "C" int nt_multi_render(nt_multi *m, const nt_render_params *params, uint8_t *rgba_out, size_t stride, nt_render_stats *stats) {
    if (!m || !params || !rgba_out) return nt_fail_public(NT_ERR_INVALID, "NULL argument");
    if (params->struct_size != sizeof(nt_render_params)) return nt_fail_public(NT_ERR_INVALID, "nt_render_params.struct_size %u != %zu", params->struct_size, sizeof(nt_render_params));
    if (params->width == 0 || params->height == 0) return nt_fail_public(NT_ERR_INVALID, "bad image size");
    if (stride < (size_t)params->width * 4 || stride % 4 || ((uintptr_t)rgba_out) % 4) return nt_fail_public(NT_ERR_INVALID, "row stride must be >= 4*width, stride and pointer multiples of 4");
    std::lock_guard<std::mutex> call(m->call_mu);
    const auto t0 = std::chrono::steady_clock::now();
    m->params = *params;
    if (m->params.band_rows == 0) m->params.band_rows = 8;
    // page-locked caller buffer: the kernels of every device store into it directly (UVA: one address for all devices)
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, rgba_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer;
    cudaGetLastError();
    if (pinned) {
        m->target = rgba_out; m->target_stride = stride; m->copy_out = nullptr;
    } else {
        const size_t need = (size_t)params->width * 4 * params->height;
        if (need > m->staging_bytes) {
            if (m->staging) cudaFreeHost(m->staging);
            m->staging = nullptr; m->staging_bytes = 0;
            cudaSetDevice(m->devices[0]);
            cudaError_t e = cudaHostAlloc((void **)&m->staging, need, cudaHostAllocPortable | cudaHostAllocMapped);
            if (e != cudaSuccess) { cudaGetLastError(); return nt_fail_public(NT_ERR_NOMEM, "staging frame cudaHostAlloc(%zu): %s", need, cudaGetErrorString(e)); }
            m->staging_bytes = need;
        }
        m->target = m->staging; m->target_stride = (size_t)params->width * 4;
        m->copy_out = rgba_out; m->copy_stride = stride;
    }
    m->pending.store(m->n, std::memory_order_release);
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->gen.fetch_add(1, std::memory_order_release);
    }
    m->cv_go.notify_all();
    {
        const auto s0 = std::chrono::steady_clock::now();
        while (m->pending.load(std::memory_order_acquire) != 0) {
            if (std::chrono::steady_clock::now() - s0 > std::chrono::milliseconds(2)) {
                std::unique_lock<std::mutex> lk(m->mu);
                m->cv_done.wait(lk, [&] { return m->pending.load(std::memory_order_acquire) == 0; });
                break;
            }
        }
    }
    for (int i = 0; i < m->n; ++i)
        if (m->rc[i]) return nt_fail_public(m->rc[i], "device %d: %s", m->devices[i], m->err[i].c_str());
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int i = 0; i < m->n; ++i) {
            const nt_render_stats &s = m->stats[i];
            stats->rays_primary += s.rays_primary; stats->rays_secondary += s.rays_secondary; stats->rays_shadow += s.rays_shadow;
            stats->sphere_tests += s.sphere_tests; stats->plane_tests += s.plane_tests; stats->triangle_tests += s.triangle_tests;
            stats->box_tests += s.box_tests; stats->light_evals += s.light_evals;
            if (s.kernel_ms > stats->kernel_ms) stats->kernel_ms = s.kernel_ms;
        }
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return NT_OK;
}
