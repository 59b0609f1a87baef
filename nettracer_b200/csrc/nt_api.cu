// nt_api.cu — the C ABI of include/nettracer_b200.h: validation, scene flattening + upload,
// launch configuration, host/device render entry points, sharding helpers, CUDA IPC wrappers.
// Host logic only; the arithmetic lives in nt_trace.cuh.  No CPU fallback anywhere: every compute
// entry point needs an sm_100 device and fails with NT_ERR_NO_DEVICE / NT_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/nettracer_b200.h"
#include "nt_bvh.h"
#include "nt_cull.h"
#include "nt_shadowgrid.h"
#include "nt_device.h"
#include "nt_sync.cuh"

static_assert(NT_MAX_DEPTH == NT_MAX_DEPTH_DEV, "depth limits must agree");
static_assert(NT_RULE_QUANTIZE_TRUNCATE == NT_DEV_RULE_TRUNCATE && NT_RULE_ATTENUATE_INV_SQUARE == NT_DEV_RULE_ATTENUATE &&
              NT_RULE_RENORMALIZE == NT_DEV_RULE_RENORMALIZE, "rule bits must agree");

// ---------------- errors ----------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
// for the host-only translation units of the library (nt_multi.cpp, nt_hostframe.cpp)
int nt_fail_public(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(NT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

extern "C" int nt_abi_version(void) { return NT_ABI_VERSION; }
extern "C" const char *nt_last_error(void) { return g_err; }

extern "C" int nt_device_count(int *count) {
    if (!count) return fail(NT_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); *count = 0; return fail(NT_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = n;
    return NT_OK;
}

static int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(NT_ERR_NO_DEVICE, "no CUDA device (%s)", cudaGetErrorString(e)); }
    if (device < 0 || device >= n) return fail(NT_ERR_NO_DEVICE, "device %d out of range (%d devices)", device, n);
    int major = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(NT_ERR_NO_DEVICE, "device %d is sm_%d0, this library is built for sm_100a only", device, major);
    return NT_OK;
}

// ---------------- scene ----------------
struct nt_scene {
    int device = 0;
    NtDevScene ds{};
    std::vector<void *> allocs;
    size_t device_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Work counters: one block per call in flight, taken round-robin from a small ring, so that flat-scene renders on
    // different streams do not share the tile counter (each block remembers the stream and an event of its last use)
    struct CounterBlock { unsigned long long *d = nullptr; cudaEvent_t ev = nullptr; cudaStream_t st = nullptr; bool used = false; };
    static const int kRing = 4;
    CounterBlock ring[kRing];
    int ring_next = 0, ring_last = -1;
    unsigned long long *h_counters = nullptr;
    unsigned *h_done = nullptr, *d_done = nullptr; // nt_render without stats: completion flag the kernel posts into pinned host memory
    unsigned done_seq = 0;
    uint8_t *d_fb = nullptr;
    size_t fb_bytes = 0;
    bool bvh_on_gpu = false;
    double bvh_build_ms = 0;
    void *d_samples = nullptr; // BVH scenes: per-sample radiance scratch
    size_t samples_bytes = 0;
    unsigned last_launches = 0; // kernels launched by the last render call
    unsigned long long heavy_bits = 0; // flat scenes: bounded primitives whose material both reflects and transmits (branching ray trees)
    uint32_t lfree = 0;         // flat scenes with culling: lights no plane can hide from a bounded primitive (nt_cull.h)
    std::vector<double> h_bsph; // flat scenes with culling: bounding spheres (host copy, for the per-camera pixel rectangles)
    void *d_wf = nullptr;      // BVH scenes: wavefront workspace (level records of one chunk of samples)
    size_t wf_bytes = 0, wf_limit = (size_t)-1;
    std::mutex mu;
};

template <typename T>
static int upload(nt_scene *sc, const std::vector<T> &v, const T **out) {
    *out = nullptr;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? NT_ERR_NOMEM : NT_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    sc->allocs.push_back(p);
    sc->device_bytes += bytes;
    if (!v.empty()) CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)p;
    return NT_OK;
}

static int validate_desc(const nt_scene_desc *d) {
    if (!d) return fail(NT_ERR_INVALID, "desc is NULL");
    if (d->struct_size != sizeof(nt_scene_desc)) return fail(NT_ERR_INVALID, "nt_scene_desc.struct_size %u != %zu", d->struct_size, sizeof(nt_scene_desc));
    if (d->n_materials == 0 || !d->materials) return fail(NT_ERR_INVALID, "scene needs at least one material");
    if ((d->n_spheres && (!d->spheres || !d->sphere_mat)) || (d->n_planes && (!d->planes || !d->plane_mat)) ||
        (d->n_triangles && (!d->triangles || !d->triangle_mat)) || (d->n_lights && !d->lights))
        return fail(NT_ERR_INVALID, "array pointer is NULL for a non-zero count");
    for (uint32_t i = 0; i < d->n_spheres; ++i) {
        if (d->sphere_mat[i] < 0 || (uint32_t)d->sphere_mat[i] >= d->n_materials) return fail(NT_ERR_INVALID, "sphere %u: material %d out of range", i, d->sphere_mat[i]);
        const double *s = d->spheres + 4 * (size_t)i;
        if (!(s[3] > 0) || !std::isfinite(s[0] + s[1] + s[2] + s[3])) return fail(NT_ERR_INVALID, "sphere %u: radius must be > 0 and all values finite", i);
    }
    for (uint32_t i = 0; i < d->n_planes; ++i) {
        if (d->plane_mat[i] < 0 || (uint32_t)d->plane_mat[i] >= d->n_materials) return fail(NT_ERR_INVALID, "plane %u: material %d out of range", i, d->plane_mat[i]);
        const double *p = d->planes + 4 * (size_t)i;
        if (!std::isfinite(p[0] + p[1] + p[2] + p[3])) return fail(NT_ERR_INVALID, "plane %u: non-finite value", i);
    }
    for (uint32_t i = 0; i < d->n_triangles; ++i) {
        if (d->triangle_mat[i] < 0 || (uint32_t)d->triangle_mat[i] >= d->n_materials) return fail(NT_ERR_INVALID, "triangle %u: material %d out of range", i, d->triangle_mat[i]);
        const double *t = d->triangles + 9 * (size_t)i;
        double sum = 0;
        for (int k = 0; k < 9; ++k) sum += t[k];
        if (!std::isfinite(sum)) return fail(NT_ERR_INVALID, "triangle %u: non-finite value", i);
    }
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        double sum = 0;
        for (int k = 0; k < 6; ++k) sum += d->lights[6 * (size_t)i + k];
        if (!std::isfinite(sum)) return fail(NT_ERR_INVALID, "light %u: non-finite value", i);
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        double sum = 0;
        for (int k = 0; k < 10; ++k) sum += d->materials[10 * (size_t)i + k];
        if (!std::isfinite(sum)) return fail(NT_ERR_INVALID, "material %u: non-finite value", i);
        if (!(d->materials[10 * (size_t)i + 9] > 0)) return fail(NT_ERR_INVALID, "material %u: index of refraction must be > 0", i);
    }
    return NT_OK;
}

// nt_cull.h: lights towards which a strict-mode shadow query from a bounded primitive cannot be stopped by any
// plane, for ray epsilons >= kPlaneFreeEps (launch() clears the bits for smaller ones)
static const double kPlaneFreeEps = 1e-7;
static uint32_t plane_free_lights(const nt_scene_desc *d) {
    return nt_cull_plane_free_lights(d->spheres, d->n_spheres, d->triangles, d->n_triangles, d->planes, d->n_planes, d->lights, d->n_lights, kPlaneFreeEps);
}

static const size_t kSmemBudget = 38 * 1024; // + 8.3 KB of static shared memory + list padding + light rooms (<= 1.1 KB) stays under the 48 KB default limit
static const size_t kCounterBytes = sizeof(unsigned long long) * (NT_COUNTER_SLOTS * NT_NCOUNTERS + NT_COUNTER_EXTRA);
static const uint32_t kFlatMaxBounded = 64;

extern "C" void nt_scene_destroy(nt_scene *sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    for (void *p : sc->allocs) cudaFree(p);
    if (sc->d_fb) cudaFree(sc->d_fb);
    if (sc->d_samples) cudaFree(sc->d_samples);
    if (sc->d_wf) cudaFree(sc->d_wf);
    for (auto &cb : sc->ring) {
        if (cb.d) cudaFree(cb.d);
        if (cb.ev) cudaEventDestroy(cb.ev);
    }
    if (sc->h_counters) cudaFreeHost(sc->h_counters);
    if (sc->h_done) cudaFreeHost(sc->h_done);
    if (sc->ev0) cudaEventDestroy(sc->ev0);
    if (sc->ev1) cudaEventDestroy(sc->ev1);
    if (sc->stream) cudaStreamDestroy(sc->stream);
    delete sc;
}

static int scene_create_impl(const nt_scene_desc *d, int device, nt_scene *sc) {
    CU(cudaSetDevice(device));
    sc->device = device;
    const uint32_t ns = d->n_spheres, np = d->n_planes, nt = d->n_triangles, nm = d->n_materials, nl = d->n_lights;

    // flat (everything staged in shared memory) or BVH (bounded primitives in HBM behind a tree)
    bool use_bvh = ns + nt > kFlatMaxBounded ||
                   ((size_t)ns * 4 + (size_t)np * 7 + (size_t)nt * NT_TRI_STRIDE) * sizeof(double) > kSmemBudget;
    if (const char *e = getenv("NT_BVH")) {
        if (e[0] == '1') use_bvh = true;
        else if (e[0] == '0' && ns + nt <= kFlatMaxBounded && ((size_t)ns * 4 + (size_t)np * 7 + (size_t)nt * NT_TRI_STRIDE) * sizeof(double) <= kSmemBudget) use_bvh = false;
    }
    if (ns >= (1u << 26) || nt >= (1u << 26)) return fail(NT_ERR_INVALID, "more than 2^26 spheres or triangles");
    if ((size_t)np * 4 * sizeof(double) > kSmemBudget) return fail(NT_ERR_INVALID, "too many planes (%u): planes are staged in shared memory, limit %zu", np, kSmemBudget / 32);

    NtBvhBuild bvh;
    std::vector<int> sph_order(ns), tri_order(nt);
    for (uint32_t i = 0; i < ns; ++i) sph_order[i] = (int)i;
    for (uint32_t i = 0; i < nt; ++i) tri_order[i] = (int)i;
    const NtBvhNode4 *gpu_nodes = nullptr; // set when the tree was built on the device
    uint32_t gpu_n_nodes = 0;
    if (use_bvh) {
        int leaf_max = 4;
        if (const char *e = getenv("NT_BVH_LEAF")) leaf_max = std::min(std::max(atoi(e), 1), NT_LEAF_MAX);
        const char *bm = getenv("NT_BVH_BUILD"); // "gpu": LBVH on the device (row f3); default: binned SAH on the host
        if (bm && !strcmp(bm, "gpu")) {
            const auto tb0 = std::chrono::steady_clock::now();
            double *d_s = nullptr, *d_t = nullptr;
            cudaError_t e = cudaSuccess;
            if (ns && (e = cudaMalloc((void **)&d_s, sizeof(double) * 4 * (size_t)ns)) == cudaSuccess)
                e = cudaMemcpy(d_s, d->spheres, sizeof(double) * 4 * (size_t)ns, cudaMemcpyHostToDevice);
            if (e == cudaSuccess && nt && (e = cudaMalloc((void **)&d_t, sizeof(double) * 9 * (size_t)nt)) == cudaSuccess)
                e = cudaMemcpy(d_t, d->triangles, sizeof(double) * 9 * (size_t)nt, cudaMemcpyHostToDevice);
            NtBvhNode4 *nodes = nullptr;
            int depth4 = 0, brc = (int)e;
            if (e == cudaSuccess)
                brc = nt_bvh_build_gpu(d_s, ns, d_t, nt, leaf_max, nullptr, &nodes, &gpu_n_nodes, sph_order, tri_order, bvh.blo, bvh.bhi,
                                       &bvh.max_abs, &depth4);
            cudaFree(d_s); cudaFree(d_t);
            if (brc) return fail(NT_ERR_CUDA, "GPU BVH build: %s", cudaGetErrorString((cudaError_t)brc));
            if (3 * depth4 + 4 > NT_BVH_STACK) { // pathological Morton tree: fall back to the host builder
                cudaFree(nodes);
                for (uint32_t i = 0; i < ns; ++i) sph_order[i] = (int)i;
                for (uint32_t i = 0; i < nt; ++i) tri_order[i] = (int)i;
            } else {
                gpu_nodes = nodes;
                sc->allocs.push_back(nodes);
                sc->device_bytes += sizeof(NtBvhNode4) * (size_t)gpu_n_nodes;
                sc->bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count();
            }
        }
        if (!gpu_nodes) {
            const auto tb0 = std::chrono::steady_clock::now();
            nt_bvh_build(d->spheres, ns, d->triangles, nt, leaf_max, bvh);
            if (3 * bvh.depth4 + 4 > NT_BVH_STACK) return fail(NT_ERR_INVALID, "BVH too deep (%d levels) for the %d-entry traversal stack", bvh.depth4, NT_BVH_STACK);
            sph_order = bvh.sph_order;
            tri_order = bvh.tri_order;
            sc->bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count();
        }
    }

    // SPEC-PROVISIONAL §1 derived quantities, in binary64; the float view is the rounded double view
    std::vector<double> sph(4 * (size_t)ns), sph_invr(ns), pln(4 * (size_t)np), tri(NT_TRI_STRIDE * (size_t)nt),
        mat(NT_MAT_STRIDE * (size_t)nm), lights(6 * (size_t)nl), globals(8, 0.0);
    std::vector<int> sph_mat(ns), sph_gid(ns), pln_mat(np), tri_mat(nt), tri_gid(nt);
    std::vector<unsigned> pln_code((np + 15) / 16 + 1, 0u);
    for (uint32_t k = 0; k < ns; ++k) {
        const int i = sph_order[k];
        const double *s = d->spheres + 4 * (size_t)i;
        sph[4 * (size_t)k] = s[0]; sph[4 * (size_t)k + 1] = s[1]; sph[4 * (size_t)k + 2] = s[2];
        sph[4 * (size_t)k + 3] = s[3] * s[3];
        sph_invr[k] = 1.0 / s[3];
        sph_mat[k] = d->sphere_mat[i];
        sph_gid[k] = i;
    }
    for (uint32_t i = 0; i < np; ++i) {
        for (int k = 0; k < 4; ++k) pln[4 * (size_t)i + k] = d->planes[4 * (size_t)i + k];
        pln_mat[i] = d->plane_mat[i];
        const double *n = d->planes + 4 * (size_t)i;
        unsigned code = 3; // exact axis-aligned unit normals take the one-product fast path (bit-identical)
        for (int k = 0; k < 3; ++k)
            if (std::fabs(n[k]) == 1.0 && n[(k + 1) % 3] == 0.0 && n[(k + 2) % 3] == 0.0) code = (unsigned)k;
        pln_code[i / 16] |= code << (2 * (i % 16));
    }
    // flat kernels: axis-aligned planes by axis as (p = n_k * d, index) pairs, everything else in a general list
    std::vector<double> axl64;
    std::vector<float> axl32;
    std::vector<int> pgen;
    uint32_t nax[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; ++k) {
        for (uint32_t i = 0; i < np; ++i)
            if (((pln_code[i / 16] >> (2 * (i % 16))) & 3u) == (unsigned)k) {
                const double *n = d->planes + 4 * (size_t)i;
                const double pk = n[k] * n[3]; // exact: n[k] = +-1
                double bits64 = 0;
                float bits32 = 0;
                const long long i64 = (long long)i;
                const int i32 = (int)i;
                memcpy(&bits64, &i64, 8);
                memcpy(&bits32, &i32, 4);
                axl64.push_back(pk); axl64.push_back(bits64);
                axl32.push_back((float)pk); axl32.push_back(bits32);
                ++nax[k];
            }
    }
    for (uint32_t i = 0; i < np; ++i)
        if (((pln_code[i / 16] >> (2 * (i % 16))) & 3u) == 3u) pgen.push_back((int)i);
    // at most two axis-aligned planes per axis (a room): the same entries as [3][2] slots padded with (NaN, -1), for the
    // straight-line nearest-plane routine; NT_SLAB=0 keeps the list loops (A/B, tests)
    std::vector<double> axs64(12, std::nan(""));
    std::vector<float> axs32(12, std::nanf(""));
    bool slab = !use_bvh && nax[0] + nax[1] + nax[2] > 0 && nax[0] <= 2 && nax[1] <= 2 && nax[2] <= 2;
    if (const char *e = getenv("NT_SLAB")) if (e[0] == '0') slab = false;
    {
        const long long m64 = -1;
        const int m32 = -1;
        for (int j = 0; j < 6; ++j) { memcpy(&axs64[2 * j + 1], &m64, 8); memcpy(&axs32[4 * (j / 2) + 2 + (j % 2)], &m32, 4); } // binary32: p0 p1 i0 i1 per axis
        size_t src = 0;
        for (int k = 0; k < 3 && slab; ++k)
            for (uint32_t j = 0; j < nax[k]; ++j, ++src)
                for (int e = 0; e < 2; ++e) { axs64[4 * k + 2 * j + e] = axl64[2 * src + e]; axs32[4 * k + 2 * e + j] = axl32[2 * src + e]; }
    }
    // light rooms (nt_cull.h): shadow queries from inside skip the axis-aligned planes; NT_LIGHT_ROOMS=0 disables (A/B, tests)
    std::vector<double> room64(8 * (size_t)nl + 8, 0.0);
    std::vector<float> room32(8 * (size_t)nl + 8, 0.0f);
    bool rooms = !use_bvh && nl > 0 && nl <= 16 && nax[0] + nax[1] + nax[2] > 0;
    if (const char *e = getenv("NT_LIGHT_ROOMS")) if (e[0] == '0') rooms = false;
    if (rooms) nt_cull_light_rooms(d->planes, np, d->lights, nl, room64.data(), room32.data());
    for (uint32_t k = 0; k < nt; ++k) {
        const int i = tri_order[k];
        const double *t = d->triangles + 9 * (size_t)i;
        double *o = tri.data() + NT_TRI_STRIDE * (size_t)k;
        const double e1[3] = { t[3] - t[0], t[4] - t[1], t[5] - t[2] }, e2[3] = { t[6] - t[0], t[7] - t[1], t[8] - t[2] };
        const double c[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
        const double inv = 1.0 / std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
        for (int a = 0; a < 3; ++a) { o[a] = t[a]; o[3 + a] = e1[a]; o[6 + a] = e2[a]; o[9 + a] = c[a] * inv; }
        tri_mat[k] = d->triangle_mat[i];
        tri_gid[k] = (int)(ns + np) + i;
    }
    for (uint32_t i = 0; i < nm; ++i) {
        const double *m = d->materials + 10 * (size_t)i;
        double *o = mat.data() + NT_MAT_STRIDE * (size_t)i;
        for (int k = 0; k < 10; ++k) o[k] = m[k];
        o[10] = 1.0 / m[9];
        o[11] = 0;
    }
    for (size_t i = 0; i < 6 * (size_t)nl; ++i) lights[i] = d->lights[i];
    for (int k = 0; k < 3; ++k) { globals[k] = d->ambient[k]; globals[3 + k] = d->background[k]; }

    auto to_f = [](const std::vector<double> &v) { std::vector<float> f(v.size()); for (size_t i = 0; i < v.size(); ++i) f[i] = (float)v[i]; return f; };
    NtDevScene &ds = sc->ds;
    ds.ns = ns; ds.np = np; ds.nt = nt; ds.nm = nm; ds.nl = nl;
    ds.use_bvh = use_bvh ? 1 : 0;
    ds.n_nodes = use_bvh && ns + nt > 0 ? (gpu_nodes ? gpu_n_nodes : (uint32_t)bvh.nodes4.size()) : 0;
    sc->bvh_on_gpu = gpu_nodes != nullptr;
    ds.max_abs = bvh.max_abs;
    if (!use_bvh) { // flat scenes: extent of the bounded primitives, for the binary32 filter margin
        double mx = 0;
        for (uint32_t i = 0; i < ns; ++i)
            for (int a = 0; a < 3; ++a) mx = std::max(mx, std::fabs(d->spheres[4 * (size_t)i + a]) + d->spheres[4 * (size_t)i + 3]);
        for (size_t i = 0; i < 9 * (size_t)nt; ++i) mx = std::max(mx, std::fabs(d->triangles[i]));
        ds.max_abs = (float)mx * 1.000001f;
    }
    for (int a = 0; a < 3; ++a) { ds.blo[a] = bvh.blo[a]; ds.bhi[a] = bvh.bhi[a]; }
    int rc;
#define UP(vec, dst) if ((rc = upload(sc, vec, &(dst))) != NT_OK) return rc
    UP(sph, ds.v64.sph); UP(sph_invr, ds.v64.sph_invr); UP(pln, ds.v64.pln); UP(tri, ds.v64.tri);
    UP(mat, ds.v64.mat); UP(lights, ds.v64.lights); UP(globals, ds.v64.globals);
    const std::vector<float> fsph = to_f(sph), fir = to_f(sph_invr), fpln = to_f(pln), ftri = to_f(tri),
                             fmat = to_f(mat), fl = to_f(lights), fg = to_f(globals);
    UP(fsph, ds.v32.sph); UP(fir, ds.v32.sph_invr); UP(fpln, ds.v32.pln); UP(ftri, ds.v32.tri);
    UP(fmat, ds.v32.mat); UP(fl, ds.v32.lights); UP(fg, ds.v32.globals);
    UP(sph_mat, ds.sph_mat); UP(sph_gid, ds.sph_gid); UP(pln_mat, ds.pln_mat); UP(pln_code, ds.pln_code); UP(tri_mat, ds.tri_mat); UP(tri_gid, ds.tri_gid);
    if (gpu_nodes) ds.nodes = gpu_nodes; else UP(bvh.nodes4, ds.nodes);
    for (int k = 0; k < 3; ++k) ds.nax[k] = nax[k];
    ds.ngen = (uint32_t)pgen.size();
    UP(axl64, ds.axl64); UP(axl32, ds.axl32); UP(pgen, ds.pgen);
    ds.slab = slab ? 1 : 0; ds.rooms = rooms ? 1 : 0;
    for (int w = 0; w < 2; ++w) { // the layout of nt_trace.cuh stage_scene, R = float / double
        const size_t rs = w ? 8 : 4;
        const size_t r_units = (size_t)ns * 4 + (size_t)np * 4 + (size_t)nt * NT_TRI_STRIDE + ((2 * ((size_t)nax[0] + nax[1] + nax[2]) + 3) & ~(size_t)3);
        ds.room_off[w] = (uint32_t)(r_units * rs + ((pgen.size() + 3) & ~(size_t)3) * 4);
        ds.axs_off[w] = ds.room_off[w] + (uint32_t)((rooms ? 8 * (size_t)nl : 0) * rs);
    }
    UP(axs64, ds.axs64); UP(axs32, ds.axs32); UP(room64, ds.room64); UP(room32, ds.room32);
    // flat scenes: conservative culling tables (nt_cull.h); NT_CULL=0 renders by brute force (A/B, tests)
    ds.cull = 0; ds.lbuf_k = NT_LBUF_K;
    ds.sph_bits = ns >= 64 ? ~0ull : (1ull << ns) - 1ull;
    ds.all_bits = ns + nt >= 64 ? ~0ull : (1ull << (ns + nt)) - 1ull;
    NtCullTables ct;
    const char *ce = getenv("NT_CULL");
    if (!use_bvh && !(ce && ce[0] == '0') && nt_cull_build(d->spheres, ns, d->triangles, nt, d->lights, nl, ct)) {
        ds.cull = 1; ds.lbuf_k = ct.k;
        {   // light buffers are trusted up to 1e5 x the smallest sphere radius from the light (nt_trace.cuh lbuf_mask); a scene
            // whose own extent exceeds that gets no culling at all
            double rmin = HUGE_VAL, ext = 0;
            for (uint32_t j = 0; j < ns; ++j) {
                rmin = std::min(rmin, d->spheres[4 * (size_t)j + 3]);
                for (int a = 0; a < 3; ++a) ext = std::max(ext, std::fabs(d->spheres[4 * (size_t)j + a]) + d->spheres[4 * (size_t)j + 3]);
            }
            for (uint32_t l = 0; l < nl; ++l) for (int a = 0; a < 3; ++a) ext = std::max(ext, std::fabs(d->lights[6 * (size_t)l + a]));
            ds.cull_far = ns ? (float)std::min(1e5 * rmin, 1e30) : 1e30f;
            if (ns && !(2.0 * ext < 1e5 * rmin)) ds.cull = 0; // tiny spheres in a huge scene: brute force
        }
        sc->h_bsph = ct.bsph;
        for (uint32_t j = 0; j < ns; ++j) { const double *m = d->materials + 10 * (size_t)d->sphere_mat[j]; if (m[7] > 0 && m[8] > 0) sc->heavy_bits |= 1ull << j; }
        for (uint32_t j = 0; j < nt; ++j) { const double *m = d->materials + 10 * (size_t)d->triangle_mat[j]; if (m[7] > 0 && m[8] > 0) sc->heavy_bits |= 1ull << (ns + j); }
        const char *pe = getenv("NT_PLANE_FREE"); // 0: always run the plane loops of a shadow query (A/B, tests)
        if (!(pe && pe[0] == '0')) sc->lfree = plane_free_lights(d);
    }
    UP(ct.lbuf, ds.lbuf); UP(ct.nbr, ds.nbr); // 16-byte placeholders when culling is off
    // BVH scenes: shadow grids (nt_shadowgrid.h); NT_SHADOW_GRID=0 keeps every occlusion query on the whole tree (A/B, tests)
    {
        std::vector<NtShadowGrid> grids(nl);
        std::vector<uint32_t> sg_off, sg_items;
        const char *ge = getenv("NT_SHADOW_GRID");
        ds.sg_on = 0;
        if (use_bvh && ns > 0 && nl > 0 && !(ge && ge[0] == '0'))
            ds.sg_on = nt_shadow_grids_build(sph.data(), ns, lights.data(), nl, (double)ds.max_abs, grids, sg_off, sg_items) > 0 ? 1 : 0;
        if (!ds.sg_on) { grids.assign(std::max<uint32_t>(nl, 1), NtShadowGrid{}); sg_off.clear(); sg_items.clear(); }
        // ... and room for the grid of the primary rays, rebuilt on the device for every call's eye (nt_eyegrid.cuh):
        // grid nl, K0 x K0 + 2 offsets (+ the scan's block totals), 48 entries per sphere, the spheres' rectangles
        const char *ee = getenv("NT_EYE_GRID");
        ds.eg_on = use_bvh && ns > 0 && !(ee && ee[0] == '0') ? 1 : 0;
        grids.resize((size_t)nl + 1, NtShadowGrid{});
        ds.eg_off = nullptr; ds.eg_items = nullptr; ds.eg_boxes = nullptr; ds.eg_acc = nullptr; ds.eg_items_cap = 0; ds.eg_k0 = 0;
        if (ds.eg_on) {
            const uint32_t k0 = nt_shadow_grid_k0(ns);
            ds.eg_k0 = k0;
            ds.eg_items_cap = (uint32_t)std::min<size_t>(std::max<size_t>(48 * (size_t)ns, 65536), (size_t)1 << 30);
            const size_t off_bytes = sizeof(uint32_t) * ((size_t)k0 * k0 + 2 + 1024), item_bytes = sizeof(uint32_t) * (size_t)ds.eg_items_cap;
            void *p_off = nullptr, *p_items = nullptr; // device-only: written by the kernels of nt_eyegrid.cuh
            cudaError_t ge2 = cudaMalloc(&p_off, off_bytes);
            if (ge2 == cudaSuccess) { sc->allocs.push_back(p_off); ge2 = cudaMalloc(&p_items, item_bytes); }
            if (ge2 == cudaSuccess) { sc->allocs.push_back(p_items); ge2 = cudaMemset(p_off, 0, off_bytes); }
            if (ge2 != cudaSuccess) return fail(ge2 == cudaErrorMemoryAllocation ? NT_ERR_NOMEM : NT_ERR_CUDA, "eye grid cudaMalloc(%zu + %zu): %s", off_bytes, item_bytes, cudaGetErrorString(ge2));
            sc->device_bytes += off_bytes + item_bytes;
            ds.eg_off = (uint32_t *)p_off; ds.eg_items = (uint32_t *)p_items;
            double clo[3] = { HUGE_VAL, HUGE_VAL, HUGE_VAL }, chi[3] = { -HUGE_VAL, -HUGE_VAL, -HUGE_VAL };
            for (uint32_t i = 0; i < ns; ++i)
                for (int a = 0; a < 3; ++a) { clo[a] = std::min(clo[a], sph[4 * (size_t)i + a]); chi[a] = std::max(chi[a], sph[4 * (size_t)i + a]); }
            for (int a = 0; a < 3; ++a) { ds.sph_lo[a] = (float)clo[a]; ds.sph_hi[a] = (float)chi[a]; }
            std::vector<double> boxes(4 * (size_t)ns, 0.0);
            std::vector<unsigned long long> acc(8, 0ull);
            const double *d_boxes = nullptr;
            const unsigned long long *d_acc = nullptr;
            UP(boxes, d_boxes); UP(acc, d_acc);
            ds.eg_boxes = const_cast<double *>(d_boxes); ds.eg_acc = const_cast<unsigned long long *>(d_acc);
        }
        UP(grids, ds.sgrid); UP(sg_off, ds.sg_off); UP(sg_items, ds.sg_items);
    }
#undef UP
    CU(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&sc->ev0));
    CU(cudaEventCreate(&sc->ev1));
    for (auto &cb : sc->ring) {
        CU(cudaMalloc(&cb.d, kCounterBytes));
        CU(cudaEventCreateWithFlags(&cb.ev, cudaEventDisableTiming));
    }
    CU(cudaMallocHost(&sc->h_counters, kCounterBytes));
    if (cudaHostAlloc((void **)&sc->h_done, 64, cudaHostAllocMapped) == cudaSuccess) {
        *sc->h_done = 0;
        if (cudaHostGetDevicePointer((void **)&sc->d_done, sc->h_done, 0) != cudaSuccess) sc->d_done = nullptr;
    }
    cudaGetLastError();
    return NT_OK;
}

extern "C" int nt_scene_create(const nt_scene_desc *desc, int device, nt_scene **out) {
    if (!out) return fail(NT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc) return rc;
    if ((rc = check_device(device)) != NT_OK) return rc;
    nt_scene *sc = new (std::nothrow) nt_scene;
    if (!sc) return fail(NT_ERR_NOMEM, "out of host memory");
    rc = scene_create_impl(desc, device, sc);
    if (rc) { nt_scene_destroy(sc); return rc; }
    *out = sc;
    return NT_OK;
}

extern "C" int nt_cull_tables(const nt_scene_desc *desc, uint32_t *k_out, uint64_t *lbuf_out, size_t lbuf_capacity,
                              uint64_t *nbr_out, double *bsph_out) {
    int rc = validate_desc(desc);
    if (rc) return rc;
    NtCullTables ct;
    if (!nt_cull_build(desc->spheres, desc->n_spheres, desc->triangles, desc->n_triangles, desc->lights, desc->n_lights, ct))
        return fail(NT_ERR_INVALID, "scene is not eligible for the flat culling tables (0 or > 64 bounded primitives, or > %d lights)", NT_CULL_MAX_LIGHTS);
    if (k_out) *k_out = ct.k;
    if (lbuf_out) {
        if (lbuf_capacity < ct.lbuf.size()) return fail(NT_ERR_INVALID, "lbuf_out holds %zu masks, %zu needed", lbuf_capacity, ct.lbuf.size());
        for (size_t i = 0; i < ct.lbuf.size(); ++i) lbuf_out[i] = ct.lbuf[i];
    }
    if (nbr_out) for (size_t i = 0; i < ct.nbr.size(); ++i) nbr_out[i] = ct.nbr[i];
    if (bsph_out) for (size_t i = 0; i < ct.bsph.size(); ++i) bsph_out[i] = ct.bsph[i];
    return NT_OK;
}

extern "C" int nt_primary_rects(const nt_scene_desc *desc, const nt_render_params *p, uint16_t *rects_out) {
    int rc = validate_desc(desc);
    if (rc) return rc;
    if (!p || !rects_out) return fail(NT_ERR_INVALID, "NULL argument");
    if (p->width == 0 || p->height == 0 || p->width > 65536 || p->height > 65536) return fail(NT_ERR_INVALID, "bad image size %ux%u", p->width, p->height);
    NtCullTables ct;
    if (!nt_cull_build(desc->spheres, desc->n_spheres, desc->triangles, desc->n_triangles, desc->lights, desc->n_lights, ct))
        return fail(NT_ERR_INVALID, "scene is not eligible for the flat culling tables");
    double cam[12], eye_inf = 0, mx = 0;
    for (int k = 0; k < 3; ++k) {
        cam[k] = p->camera.eye[k]; cam[3 + k] = p->camera.p00[k]; cam[6 + k] = p->camera.dx[k]; cam[9 + k] = p->camera.dy[k];
        eye_inf = std::max(eye_inf, std::fabs(cam[k]));
    }
    for (uint32_t i = 0; i < desc->n_spheres; ++i)
        for (int a = 0; a < 3; ++a) mx = std::max(mx, std::fabs(desc->spheres[4 * (size_t)i + a]) + desc->spheres[4 * (size_t)i + 3]);
    for (size_t i = 0; i < 9 * (size_t)desc->n_triangles; ++i) mx = std::max(mx, std::fabs(desc->triangles[i]));
    nt_cull_primary_rects(ct.bsph.data(), desc->n_spheres + desc->n_triangles, cam, p->width, p->height, 1e-5 * (eye_inf + mx), rects_out);
    return NT_OK;
}

extern "C" int nt_shadow_grid(const nt_scene_desc *desc, uint32_t light, float *params_out, uint32_t *k_out, uint32_t *off_out,
                              size_t off_capacity, uint32_t *items_out, size_t items_capacity, size_t *n_items_out) {
    int rc = validate_desc(desc);
    if (rc) return rc;
    if (!k_out || !n_items_out) return fail(NT_ERR_INVALID, "NULL argument");
    if (light >= desc->n_lights) return fail(NT_ERR_INVALID, "light %u >= n_lights %u", light, desc->n_lights);
    const uint32_t ns = desc->n_spheres;
    std::vector<double> sph(4 * (size_t)ns);
    double mx = 0; // as nt_scene_create: the largest |coordinate| of a bounded primitive
    for (uint32_t i = 0; i < ns; ++i) {
        const double *s = desc->spheres + 4 * (size_t)i;
        sph[4 * (size_t)i] = s[0]; sph[4 * (size_t)i + 1] = s[1]; sph[4 * (size_t)i + 2] = s[2]; sph[4 * (size_t)i + 3] = s[3] * s[3];
        for (int a = 0; a < 3; ++a) mx = std::max(mx, std::fabs(s[a]) + s[3]);
    }
    for (size_t i = 0; i < 9 * (size_t)desc->n_triangles; ++i) mx = std::max(mx, std::fabs(desc->triangles[i]));
    std::vector<NtShadowGrid> grids;
    std::vector<uint32_t> off, items;
    nt_shadow_grids_build(sph.data(), ns, desc->lights + 6 * (size_t)light, 1, mx, grids, off, items);
    *k_out = 0; *n_items_out = 0;
    if (grids.empty() || !grids[0].valid) return NT_OK;
    const NtShadowGrid &g = grids[0];
    *k_out = g.K; *n_items_out = items.size();
    if (params_out) {
        for (int k = 0; k < 3; ++k) { params_out[k] = g.L[k]; params_out[3 + k] = g.axis[k]; params_out[6 + k] = g.U[k]; params_out[9 + k] = g.V[k]; }
        params_out[12] = g.u0; params_out[13] = g.v0; params_out[14] = g.su; params_out[15] = g.sv;
    }
    if (off_out && off_capacity >= off.size()) std::copy(off.begin(), off.end(), off_out);
    if (items_out && items_capacity >= items.size()) std::copy(items.begin(), items.end(), items_out);
    return NT_OK;
}

extern "C" int nt_light_rooms(const nt_scene_desc *desc, double *rooms_out) {
    int rc = validate_desc(desc);
    if (rc) return rc;
    if (!rooms_out) return fail(NT_ERR_INVALID, "rooms_out is NULL");
    std::vector<float> f32(8 * (size_t)desc->n_lights + 8);
    nt_cull_light_rooms(desc->planes, desc->n_planes, desc->lights, desc->n_lights, rooms_out, f32.data());
    return NT_OK;
}

extern "C" int nt_plane_free_lights(const nt_scene_desc *desc, uint32_t *mask_out) {
    int rc = validate_desc(desc);
    if (rc) return rc;
    if (!mask_out) return fail(NT_ERR_INVALID, "NULL argument");
    NtCullTables ct;
    if (!nt_cull_build(desc->spheres, desc->n_spheres, desc->triangles, desc->n_triangles, desc->lights, desc->n_lights, ct))
        return fail(NT_ERR_INVALID, "scene is not eligible for the flat culling tables");
    *mask_out = plane_free_lights(desc);
    return NT_OK;
}

extern "C" int nt_scene_info(const nt_scene *sc, uint64_t info[4]) {
    if (!sc || !info) return fail(NT_ERR_INVALID, "NULL argument");
    info[0] = (uint64_t)sc->ds.use_bvh | ((uint64_t)sc->bvh_on_gpu << 1) | ((uint64_t)sc->ds.cull << 2) | ((uint64_t)(sc->bvh_build_ms * 1000.0) << 8);
    info[1] = sc->ds.n_nodes; info[2] = sc->device_bytes; info[3] = (uint64_t)sc->device | ((uint64_t)sc->last_launches << 32);
    return NT_OK;
}

// ---------------- sharding arithmetic ----------------
extern "C" uint32_t nt_shard_rows(uint32_t height, uint32_t band_rows, uint32_t shard_index, uint32_t shard_count) {
    if (band_rows == 0 || shard_count == 0 || shard_index >= shard_count) return 0;
    const uint32_t nb = (height + band_rows - 1) / band_rows;
    uint32_t rows = 0;
    for (uint32_t b = shard_index; b < nb; b += shard_count) {
        const uint32_t y0 = b * band_rows, y1 = y0 + band_rows > height ? height : y0 + band_rows;
        rows += y1 - y0;
    }
    return rows;
}

// ---------------- render ----------------
// nt_trace.cuh plane_below_eps: eps * (1 - 2^-50), rounded; 0 switches the shortcut off for absurdly small epsilons
static double eps_low_bound(double eps) { return eps >= 1e-290 ? eps * (1.0 - 8.8817841970012523e-16) : 0.0; }
// Warp tile = twx x (32 / lanes / twx) pixels; everything derived from the shape (a->lanes, vrows, width must be set)
static void set_tile_shape(NtRenderArgs *a, uint32_t twx) {
    a->twx = twx; a->twy = 32u / a->lanes / twx;
    a->tiles_x = (a->width + a->twx - 1) / a->twx;
    a->tiles_y = (a->vrows + a->twy - 1) / a->twy;
    a->n_tiles = a->tiles_x * a->tiles_y;
    a->tile_rot = 0;
    for (a->log2_twx = 0; (1u << a->log2_twx) < a->twx; ++a->log2_twx) {}
    for (a->log2_twy = 0; (1u << a->log2_twy) < a->twy; ++a->log2_twy) {}
    a->inv_tiles_x = 1.0f / (float)a->tiles_x;
}

static int make_args(const nt_render_params *p, size_t stride, NtRenderArgs *a) {
    if (!p) return fail(NT_ERR_INVALID, "params is NULL");
    if (p->struct_size != sizeof(nt_render_params)) return fail(NT_ERR_INVALID, "nt_render_params.struct_size %u != %zu", p->struct_size, sizeof(nt_render_params));
    if (p->width == 0 || p->height == 0 || p->width > 65536 || p->height > 65535) return fail(NT_ERR_INVALID, "bad image size %ux%u (at most 65536 x 65535)", p->width, p->height);
    uint32_t n = 0;
    for (uint32_t k = 1; k <= 8; ++k) if (k * k == p->spp) n = k;
    if (!n) return fail(NT_ERR_INVALID, "spp %u is not a perfect square in 1..64", p->spp);
    if (p->max_depth < 1 || p->max_depth > NT_MAX_DEPTH) return fail(NT_ERR_INVALID, "max_depth %u not in 1..%d", p->max_depth, NT_MAX_DEPTH);
    if ((uint64_t)(p->width + 31) * (uint64_t)(p->height + 31) * p->spp >= (1ull << 31))
        return fail(NT_ERR_INVALID, "%ux%u at %u spp has 2^31 samples or more: sample and tile indices are 32-bit; render it in shards", p->width, p->height, p->spp);
    if (p->precision != NT_F64_STRICT && p->precision != NT_F32_FAST) return fail(NT_ERR_INVALID, "unknown precision %u", p->precision);
    if (p->layout != NT_LAYOUT_FULL && p->layout != NT_LAYOUT_COMPACT) return fail(NT_ERR_INVALID, "unknown layout %u", p->layout);
    if (p->flags & ~(NT_RENDER_COUNT_EXECUTED | NT_RULE_MASK)) return fail(NT_ERR_INVALID, "unknown flags 0x%x", p->flags);
    const uint32_t scount = p->shard_count ? p->shard_count : 1;
    const uint32_t band = p->band_rows ? (p->band_rows > 65536 ? 65536 : p->band_rows) : 1; // height <= 65536: same partition
    if (p->shard_index >= scount) return fail(NT_ERR_INVALID, "shard_index %u >= shard_count %u", p->shard_index, scount);
    if (stride < (size_t)p->width * 4 || stride % 4) return fail(NT_ERR_INVALID, "row stride %zu must be >= 4*width and a multiple of 4", stride);
    memset(a, 0, sizeof *a);
    a->width = p->width; a->height = p->height; a->spp = p->spp; a->n = n; a->max_depth = p->max_depth;
    a->shard_index = p->shard_index; a->shard_count = scount; a->band_rows = band; a->layout = p->layout;
    a->vrows = nt_shard_rows(p->height, band, p->shard_index, scount);
    uint32_t L = 1;
    while (L < 32 && p->spp % (L * 2) == 0) L *= 2;
    a->lanes = L;
    int li = 0;
    while ((1u << li) < L) ++li;
    a->log2_lanes = (uint32_t)li;
    static const uint32_t tw[6] = { 8, 4, 4, 2, 2, 1 }; // L = 1,2,4,8,16,32: near-square warp tiles (best culling / coherence)
    uint32_t twx = tw[li];
    if (const char *e = getenv("NT_TILE_W")) { // experiment: warp tile width in pixels (power of two <= 32 / lanes)
        const uint32_t wreq = (uint32_t)atoi(e);
        if (wreq >= 1 && wreq <= 32u / L && (wreq & (wreq - 1)) == 0) twx = wreq;
    }
    set_tile_shape(a, twx);
    a->band_magic = band > 1 ? (uint32_t)(((1ull << 32) + band - 1) / band) : 0u;
    a->n_mul = (65536u + n - 1) / n;
    a->eps = p->ray_epsilon > 0 ? p->ray_epsilon : 1e-6;
    if (p->precision == NT_F32_FAST && a->eps < 1e-4) a->eps = 1e-4; // SPEC-PROVISIONAL §7
    a->eps_lo = eps_low_bound(a->eps);
    for (int k = 0; k < 3; ++k) {
        a->cam[k] = p->camera.eye[k]; a->cam[3 + k] = p->camera.p00[k];
        a->cam[6 + k] = p->camera.dx[k]; a->cam[9 + k] = p->camera.dy[k];
    }
    a->stride = stride;
    a->count_executed = p->flags & NT_RENDER_COUNT_EXECUTED;
    a->rules = p->flags & NT_RULE_MASK;
    const double half = (p->flags & NT_RULE_SAMPLE_CORNER) ? 0.0 : 0.5; // SPEC-PROVISIONAL section 8
    for (uint32_t i = 0; i < n; ++i) a->samp_off[i] = ((double)i + half) / (double)n;
    a->inv_spp = 1.0 / (double)p->spp;
    for (int k = 0; k < 12; ++k) a->camf[k] = (float)a->cam[k];
    for (int k = 0; k < 8; ++k) a->samp_off_f[k] = (float)a->samp_off[k];
    a->inv_spp_f = (float)a->inv_spp;
    return NT_OK;
}

// Flags of the multi-GPU exchange for a shard that owns no row: nothing to render, but the protocol goes on.
static int sync_only(const NtRenderArgs &a, cudaStream_t st);

static int launch(nt_scene *sc, NtRenderArgs &a, uint32_t precision, cudaStream_t st, const nt_frame_sync *sync = nullptr) {
    a.sync_post_ptr = nullptr; a.sync_wait_ptr = nullptr; a.sync_done_ptr = nullptr;
    if (sync) {
        if (sync->struct_size != sizeof(nt_frame_sync)) return fail(NT_ERR_INVALID, "nt_frame_sync.struct_size %u != %zu", sync->struct_size, sizeof(nt_frame_sync));
        if (((uintptr_t)sync->post_at_start | (uintptr_t)sync->wait_before_store | (uintptr_t)sync->post_when_done) % 4) return fail(NT_ERR_INVALID, "flag pointers must be 4-byte aligned");
        a.sync_post_ptr = sync->post_at_start; a.sync_post_val = sync->post_at_start_value;
        a.sync_wait_ptr = sync->wait_before_store; a.sync_wait_val = sync->wait_value;
        a.sync_done_ptr = sync->post_when_done; a.sync_done_val = sync->post_when_done_value;
    }
    // this call's block of work counters; calls on other streams that used it (or, for BVH scenes, the scene's shared
    // scratch buffers) are ordered before this one
    nt_scene::CounterBlock &cb = sc->ring[sc->ring_next];
    if (cb.used && cb.st != st) CU(cudaStreamWaitEvent(st, cb.ev, 0));
    if (sc->ds.use_bvh && sc->ring_last >= 0 && sc->ring[sc->ring_last].st != st) CU(cudaStreamWaitEvent(st, sc->ring[sc->ring_last].ev, 0));
    CU(cudaMemsetAsync(cb.d, 0, kCounterBytes, st));
    a.counters = cb.d;
    sc->ring_last = sc->ring_next;
    sc->ring_next = (sc->ring_next + 1) % nt_scene::kRing;
    cb.st = st; cb.used = true;
    struct Mark { nt_scene::CounterBlock &cb; cudaStream_t st; ~Mark() { cudaEventRecord(cb.ev, st); } } mark{ cb, st };
    sc->last_launches = 0;
    if (a.vrows == 0) return sync_only(a, st);
    const size_t need = nt_sample_buffer_bytes(sc->ds, a, (int)precision);
    if (need > sc->samples_bytes) { // grows on demand; cudaFree waits for kernels still using the old one
        if (sc->d_samples) cudaFree(sc->d_samples);
        sc->d_samples = nullptr; sc->samples_bytes = 0;
        cudaError_t e = cudaMalloc(&sc->d_samples, need);
        if (e != cudaSuccess) return fail(NT_ERR_NOMEM, "sample buffer cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
        sc->samples_bytes = need;
    }
    a.samples = sc->d_samples;
    a.n_launches = &sc->last_launches;
    // BVH scenes render through the wavefront pipeline (nt_wavefront.cuh) unless NT_WAVEFRONT=0, the trees are deeper
    // than NT_WF_MAX_DEPTH or the scene has more than 32 lights; the frame is cut into chunks that fit the workspace
    // (NT_WF_MB megabytes at most, default 16384)
    a.wf = nullptr; a.wf_bytes = 0;
    {
        const char *e_on = getenv("NT_WAVEFRONT"), *e_mb = getenv("NT_WF_MB"); // read per launch: tests toggle them
        const bool on = !(e_on && e_on[0] == '0');
        const size_t budget = (size_t)(e_mb && atoll(e_mb) > 0 ? atoll(e_mb) : 16384) << 20;
        size_t want = on ? nt_wavefront_bytes(sc->ds, a, (int)precision) : 0;
        if (want > budget) want = budget;
        if (want > sc->wf_limit) want = sc->wf_limit; // what the device could give last time
        if (want && want < nt_wavefront_min_bytes(a, (int)precision)) want = 0; // not even one warp of samples fits: state machine
        if (want) {
            if (want > sc->wf_bytes) {
                if (sc->d_wf) cudaFree(sc->d_wf);
                sc->d_wf = nullptr; sc->wf_bytes = 0;
                cudaError_t e = cudaMalloc(&sc->d_wf, want);
                while (e == cudaErrorMemoryAllocation && want / 2 >= nt_wavefront_min_bytes(a, (int)precision)) { // smaller chunks, more launches
                    cudaGetLastError();
                    want /= 2;
                    e = cudaMalloc(&sc->d_wf, want);
                }
                if (e != cudaSuccess) { cudaGetLastError(); sc->d_wf = nullptr; want = 0; } // not even that: the per-lane state machine needs no workspace
                if (e != cudaSuccess || want < sc->wf_limit) sc->wf_limit = want;
                sc->wf_bytes = want;
            }
            if (want) { a.wf = sc->d_wf; a.wf_bytes = want; }
        }
    }
    if (!sc->ds.use_bvh && sc->ds.cull) { // primary rays: per-primitive pixel rectangles for this camera (nt_cull.h)
        double eye_inf = 0;
        for (int k = 0; k < 3; ++k) eye_inf = std::max(eye_inf, std::fabs(a.cam[k]));
        nt_cull_primary_rects(sc->h_bsph.data(), sc->ds.ns + sc->ds.nt, a.cam, a.width, a.height, 1e-5 * (eye_inf + (double)sc->ds.max_abs), &a.prect[0][0]);
        // start the tile sequence at the first row that shows a branching primitive (any primitive when there is none)
        const char *re = getenv("NT_TILE_ROT");
        uint32_t y0 = a.height;
        for (int pass = 0; pass < 2 && y0 == a.height; ++pass)
            for (uint32_t j = 0; j < sc->ds.ns + sc->ds.nt; ++j)
                if ((pass == 1 || ((sc->heavy_bits >> j) & 1ull)) && a.prect[j][0] <= a.prect[j][1]) y0 = std::min<uint32_t>(y0, a.prect[j][2]);
        if (y0 < a.height && a.max_depth > 1 && !(re && re[0] == '0')) {
            const uint32_t vr = (uint32_t)((uint64_t)y0 * a.vrows / a.height); // owned rows are spread evenly over the image
            a.tile_rot = std::min(vr / a.twy, a.tiles_y - 1) * a.tiles_x;
        }
    }
    sc->ds.lfree = a.eps >= kPlaneFreeEps ? sc->lfree : 0u; // the proof behind the bits assumes this epsilon at least
    const int e = precision == NT_F64_STRICT ? nt_launch_render_f64(sc->ds, a, st) : nt_launch_render_f32(sc->ds, a, st);
    if (e) return fail(NT_ERR_CUDA, "render kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    return NT_OK;
}

static void sum_counters(const unsigned long long *h, nt_render_stats *s) {
    unsigned long long t[NT_NCOUNTERS] = { 0 };
    for (int slot = 0; slot < NT_COUNTER_SLOTS; ++slot)
        for (int i = 0; i < NT_NCOUNTERS; ++i) t[i] += h[slot * NT_NCOUNTERS + i];
    s->rays_primary = t[0]; s->rays_secondary = t[1]; s->rays_shadow = t[2];
    s->sphere_tests = t[3]; s->plane_tests = t[4]; s->triangle_tests = t[5]; s->box_tests = t[6]; s->light_evals = t[7];
    s->sphere_tests_executed = t[8]; s->plane_tests_executed = t[9]; s->triangle_tests_executed = t[10];
}

extern "C" int nt_render(nt_scene *sc, const nt_render_params *p, uint8_t *rgba_out, size_t stride,
                         nt_render_stats *stats) {
    if (!sc || !rgba_out) return fail(NT_ERR_INVALID, "NULL argument");
    const auto t0 = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> lock(sc->mu);
    NtRenderArgs a;
    int rc = make_args(p, (size_t)(p ? p->width : 0) * 4, &a); // device frame is tightly packed
    if (rc) return rc;
    if (stride < (size_t)p->width * 4) return fail(NT_ERR_INVALID, "row stride %zu < 4*width", stride);
    CU(cudaSetDevice(sc->device));
    const size_t pitch = (size_t)p->width * 4;
    const uint32_t out_rows = p->layout == NT_LAYOUT_COMPACT ? a.vrows : p->height;
    // Pinned (page-locked) caller buffer: the kernel stores the RGBA8 words straight into it over PCIe
    // (UVA maps pinned host memory into the device address space), so the device->host transfer overlaps
    // the tracing instead of following it.  Pageable buffers go through a device frame + cudaMemcpy2D.
    uint8_t *mapped = nullptr;
    {
        static const bool allow = [] { const char *e = getenv("NT_ZEROCOPY"); return !(e && e[0] == '0'); }();
        cudaPointerAttributes at;
        if (allow && stride % 4 == 0 && ((uintptr_t)rgba_out) % 4 == 0 &&
            cudaPointerGetAttributes(&at, rgba_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            mapped = (uint8_t *)at.devicePointer;
        cudaGetLastError();
    }
    if (mapped) {
        a.out = mapped;
        a.stride = stride;
        // stores cross PCIe: one row of 32 / lanes pixels per warp tile makes the widest contiguous write (configs[2]:
        // 4x2 tiles 0.852 ms per frame end to end, 8x1 0.783; configs[1]: 8x4 0.372, 32x1 0.206)
        if (!getenv("NT_TILE_W")) set_tile_shape(&a, 32u / a.lanes);
    } else {
        const size_t need = pitch * out_rows;
        if (need > sc->fb_bytes) {
            if (sc->d_fb) cudaFree(sc->d_fb);
            sc->d_fb = nullptr; sc->fb_bytes = 0;
            cudaError_t e = cudaMalloc((void **)&sc->d_fb, need);
            if (e != cudaSuccess) return fail(NT_ERR_NOMEM, "framebuffer cudaMalloc(%zu): %s", need, cudaGetErrorString(e));
            sc->fb_bytes = need;
        }
        a.out = sc->d_fb;
    }
    const size_t cbytes = kCounterBytes;
    // A caller that wants the image only (stats == NULL) and hands a pinned frame to a flat scene: no events, no copy of the
    // work counters - the kernel's last block posts a sequence number into pinned host memory after its last pixel store
    // (nt_sync.cuh frame_sync_end: posted PCIe writes stay ordered) and this thread spins on it; every 1024 spins it asks
    // the stream, so that a failed launch cannot hang the caller.  (~20 us of a 0.7 ms frame.)
    if (!stats && mapped && !sc->ds.use_bvh && sc->d_done && a.vrows) {
        nt_frame_sync fs;
        memset(&fs, 0, sizeof fs);
        fs.struct_size = sizeof fs;
        fs.post_when_done = sc->d_done;
        fs.post_when_done_value = ++sc->done_seq;
        if ((rc = launch(sc, a, p->precision, sc->stream, &fs)) != NT_OK) return rc;
        volatile unsigned *flag = sc->h_done;
        for (unsigned spins = 1;; ++spins) {
            if ((int)(*flag - sc->done_seq) >= 0) return NT_OK;
            if ((spins & 1023u) == 0) {
                const cudaError_t e = cudaStreamQuery(sc->stream);
                if (e != cudaErrorNotReady) { // finished (the flag is there by now) or failed
                    CU(cudaStreamSynchronize(sc->stream));
                    return NT_OK;
                }
            }
        }
    }
    CU(cudaEventRecord(sc->ev0, sc->stream));
    if ((rc = launch(sc, a, p->precision, sc->stream)) != NT_OK) return rc;
    CU(cudaEventRecord(sc->ev1, sc->stream));
    if (a.vrows && !mapped) {
        if (p->layout == NT_LAYOUT_COMPACT || a.shard_count == 1) {
            CU(cudaMemcpy2DAsync(rgba_out, stride, sc->d_fb, pitch, pitch, out_rows, cudaMemcpyDeviceToHost, sc->stream));
        } else { // full layout, sharded: only the owned bands reach the caller's buffer
            const uint32_t nb = (p->height + a.band_rows - 1) / a.band_rows;
            for (uint32_t b = a.shard_index; b < nb; b += a.shard_count) {
                const uint32_t y0 = b * a.band_rows, y1 = std::min(p->height, y0 + a.band_rows);
                CU(cudaMemcpy2DAsync(rgba_out + (size_t)y0 * stride, stride, sc->d_fb + (size_t)y0 * pitch, pitch, pitch, y1 - y0, cudaMemcpyDeviceToHost, sc->stream));
            }
        }
    }
    CU(cudaMemcpyAsync(sc->h_counters, a.counters, cbytes, cudaMemcpyDeviceToHost, sc->stream));
    CU(cudaStreamSynchronize(sc->stream));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        sum_counters(sc->h_counters, stats);
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
        stats->kernel_ms = ms;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return NT_OK;
}

extern "C" int nt_render_device_sync(nt_scene *sc, const nt_render_params *p, void *rgba_out_dev, size_t stride,
                                     void *cuda_stream, const nt_frame_sync *sync) {
    if (!sc || !rgba_out_dev) return fail(NT_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lock(sc->mu);
    NtRenderArgs a;
    int rc = make_args(p, stride, &a);
    if (rc) return rc;
    if (((uintptr_t)rgba_out_dev) % 4) return fail(NT_ERR_INVALID, "output pointer must be 4-byte aligned");
    CU(cudaSetDevice(sc->device));
    a.out = (uint8_t *)rgba_out_dev;
    {   // a page-locked HOST frame as the target (nt_host_frame_pixels, cudaHostAlloc): the stores cross PCIe, and one row of
        // 32 / lanes pixels per warp tile makes the widest contiguous write (see nt_render)
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, rgba_out_dev) == cudaSuccess && at.type == cudaMemoryTypeHost && !getenv("NT_TILE_W"))
            set_tile_shape(&a, 32u / a.lanes);
        cudaGetLastError();
    }
    return launch(sc, a, p->precision, (cudaStream_t)cuda_stream, sync);
}

extern "C" int nt_render_device(nt_scene *sc, const nt_render_params *p, void *rgba_out_dev, size_t stride,
                                void *cuda_stream) {
    return nt_render_device_sync(sc, p, rgba_out_dev, stride, cuda_stream, nullptr);
}

extern "C" int nt_render_device_stats(nt_scene *sc, void *cuda_stream, nt_render_stats *stats) {
    if (!sc || !stats) return fail(NT_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lock(sc->mu);
    CU(cudaSetDevice(sc->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    memset(stats, 0, sizeof *stats);
    if (sc->ring_last < 0) return NT_OK; // nothing rendered yet
    const nt_scene::CounterBlock &cb = sc->ring[sc->ring_last];
    if (cb.st != st) CU(cudaStreamWaitEvent(st, cb.ev, 0));
    CU(cudaMemcpyAsync(sc->h_counters, cb.d, kCounterBytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    sum_counters(sc->h_counters, stats);
    if (sc->h_counters[NT_COUNTER_SLOTS * NT_NCOUNTERS + 2])
        return fail(NT_ERR_TIMEOUT, "the frame gave up waiting for a synchronisation flag (%llu waits timed out)", sc->h_counters[NT_COUNTER_SLOTS * NT_NCOUNTERS + 2]);
    return NT_OK;
}

// ---------------- unit-level nearest-hit queries ----------------
extern "C" int nt_trace_rays(nt_scene *sc, uint32_t n, const double *origins, const double *dirs, uint32_t precision,
                             double ray_epsilon, double *t_out, int32_t *prim_out) {
    if (!sc || (n && (!origins || !dirs || !t_out || !prim_out))) return fail(NT_ERR_INVALID, "NULL argument");
    if (precision != NT_F64_STRICT && precision != NT_F32_FAST) return fail(NT_ERR_INVALID, "unknown precision %u", precision);
    if (n == 0) return NT_OK;
    std::lock_guard<std::mutex> lock(sc->mu);
    CU(cudaSetDevice(sc->device));
    double *d_o = nullptr, *d_d = nullptr, *d_t = nullptr;
    int *d_p = nullptr;
    const size_t vb = sizeof(double) * 3 * (size_t)n;
    int rc = NT_OK;
    do {
        cudaError_t e;
        if ((e = cudaMalloc((void **)&d_o, vb)) || (e = cudaMalloc((void **)&d_d, vb)) || (e = cudaMalloc((void **)&d_t, sizeof(double) * n)) ||
            (e = cudaMalloc((void **)&d_p, sizeof(int) * n))) { rc = fail(NT_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); break; }
        if ((e = cudaMemcpyAsync(d_o, origins, vb, cudaMemcpyHostToDevice, sc->stream)) || (e = cudaMemcpyAsync(d_d, dirs, vb, cudaMemcpyHostToDevice, sc->stream))) {
            rc = fail(NT_ERR_CUDA, "H2D: %s", cudaGetErrorString(e)); break;
        }
        NtTraceArgs a;
        a.n = n;
        a.eps = ray_epsilon > 0 ? ray_epsilon : 1e-6;
        if (precision == NT_F32_FAST && a.eps < 1e-4) a.eps = 1e-4;
        a.eps_lo = eps_low_bound(a.eps);
        a.origins = d_o; a.dirs = d_d; a.t_out = d_t; a.prim_out = d_p;
        const int le = precision == NT_F64_STRICT ? nt_launch_trace_f64(sc->ds, a, sc->stream) : nt_launch_trace_f32(sc->ds, a, sc->stream);
        if (le) { rc = fail(NT_ERR_CUDA, "trace kernel launch: %s", cudaGetErrorString((cudaError_t)le)); break; }
        if ((e = cudaMemcpyAsync(t_out, d_t, sizeof(double) * n, cudaMemcpyDeviceToHost, sc->stream)) ||
            (e = cudaMemcpyAsync(prim_out, d_p, sizeof(int) * n, cudaMemcpyDeviceToHost, sc->stream)) || (e = cudaStreamSynchronize(sc->stream))) {
            rc = fail(NT_ERR_CUDA, "D2H: %s", cudaGetErrorString(e)); break;
        }
    } while (0);
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_p);
    return rc;
}

// ---------------- gather helper: compact shard buffers -> full frame ----------------
__global__ void deinterleave_kernel(const uint32_t *__restrict__ compact, size_t shard_stride_words, uint32_t *__restrict__ full,
                                    size_t row_stride_words, uint32_t width, uint32_t height, uint32_t band, uint32_t shards) {
    const uint32_t y = blockIdx.y;
    const uint32_t b = y / band, shard = b % shards, vr = (b / shards) * band + y % band;
    const uint32_t *src = compact + shard * shard_stride_words + (size_t)vr * width;
    uint32_t *dst = full + (size_t)y * row_stride_words;
    for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < width; x += gridDim.x * blockDim.x) dst[x] = __ldg(src + x);
}

extern "C" int nt_deinterleave_device(const void *compact_all, size_t shard_stride_bytes, void *full_out, size_t row_stride_bytes,
                                      uint32_t width, uint32_t height, uint32_t band_rows, uint32_t shard_count, int device,
                                      void *cuda_stream) {
    if (!compact_all || !full_out || !width || !height || !band_rows || !shard_count) return fail(NT_ERR_INVALID, "bad argument");
    if (shard_stride_bytes % 4 || row_stride_bytes % 4 || row_stride_bytes < (size_t)width * 4) return fail(NT_ERR_INVALID, "strides must be multiples of 4 and >= 4*width");
    if (height > 65535) return fail(NT_ERR_INVALID, "height %u too large", height);
    int rc = check_device(device);
    if (rc) return rc;
    CU(cudaSetDevice(device));
    dim3 grid((width + 255) / 256 > 8 ? 8 : (width + 255) / 256, height), block(256);
    deinterleave_kernel<<<grid, block, 0, (cudaStream_t)cuda_stream>>>((const uint32_t *)compact_all, shard_stride_bytes / 4, (uint32_t *)full_out,
                                                                       row_stride_bytes / 4, width, height, band_rows, shard_count);
    CU(cudaGetLastError());
    return NT_OK;
}

// ---------------- frame synchronisation flags (nt_sync.cuh) ----------------
static int sync_only(const NtRenderArgs &a, cudaStream_t st) {
    if (!a.sync_post_ptr && !a.sync_wait_ptr && !a.sync_done_ptr) return NT_OK;
    nt::sync_kernel<<<1, 32, 0, st>>>(a.sync_post_ptr, a.sync_post_val, a.sync_wait_ptr, 1, a.sync_wait_val, a.sync_done_ptr, a.sync_done_val,
                                      a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS + 2);
    CU(cudaGetLastError());
    return NT_OK;
}

extern "C" int nt_flags_wait_device(nt_scene *sc, const uint32_t *flags, uint32_t n, uint32_t value, void *cuda_stream) {
    if (!sc || !flags || n == 0 || n > 1024 || ((uintptr_t)flags) % 4) return fail(NT_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lock(sc->mu);
    CU(cudaSetDevice(sc->device));
    // a wait that gives up is recorded with the scene's last frame (reported by nt_render_device_stats)
    unsigned long long *timeouts = sc->ring_last >= 0 ? sc->ring[sc->ring_last].d + NT_COUNTER_SLOTS * NT_NCOUNTERS + 2 : nullptr;
    nt::sync_kernel<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(nullptr, 0, flags, n, value, nullptr, 0, timeouts);
    CU(cudaGetLastError());
    return NT_OK;
}

// ---------------- CUDA IPC (peer framebuffer over NVLink) ----------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");

extern "C" int nt_device_malloc(int device, size_t bytes, void **dev_ptr_out) {
    if (!dev_ptr_out || bytes == 0) return fail(NT_ERR_INVALID, "bad argument");
    *dev_ptr_out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    CU(cudaSetDevice(device));
    cudaError_t e = cudaMalloc(dev_ptr_out, bytes);
    if (e != cudaSuccess) return fail(NT_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return NT_OK;
}

extern "C" int nt_device_free(int device, void *dev_ptr) {
    if (!dev_ptr) return NT_OK;
    CU(cudaSetDevice(device));
    CU(cudaFree(dev_ptr));
    return NT_OK;
}

extern "C" int nt_ipc_export(const void *dev_ptr, int device, uint8_t handle_out[64]) {
    if (!dev_ptr || !handle_out) return fail(NT_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, (void *)dev_ptr));
    memcpy(handle_out, &h, 64);
    return NT_OK;
}

extern "C" int nt_ipc_open(const uint8_t handle[64], int device, void **dev_ptr_out) {
    if (!handle || !dev_ptr_out) return fail(NT_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return NT_OK;
}

extern "C" int nt_ipc_close(void *dev_ptr, int device) {
    if (!dev_ptr) return fail(NT_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return NT_OK;
}
