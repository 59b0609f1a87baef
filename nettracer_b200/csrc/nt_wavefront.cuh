// nt_wavefront.cuh — wavefront renderer for BVH scenes (BASELINE.json configs[3], [4]): the render loop as
// per-bounce ray queues.  Same arithmetic as nt_bvh_trace.cuh / nt_trace.cuh (SPEC-PROVISIONAL §2-§5, same
// operation order, so the strict mode stays bit-exact against the oracle), different schedule:
//
//   A frame is cut into chunks of S samples.  Tree level l = 1..max_depth of a chunk is a record array in heap
//   layout: the node of sample i reached by the reflection/transmission choices `path` (l-1 bits, first choice
//   in the top bit) is record (i << (l-1)) | path; its reflection child is record 2r, its transmission
//   child 2r+1 of level l+1.  So no pointers are stored and the final sum can walk the tree in pre-order.
//
//   per level:  wf_trace<nearest>  persistent warps pull ray records from the level's task list (one atomicAdd
//                                  per warp, ballot + popc prefix), traverse, write (t, primitive).  A lane that
//                                  finishes pulls the next ray: no lane waits for another lane's shading.
//               wf_trace<shadow>   tasks = (record, light): the lane rebuilds the hit point and the shadow ray,
//                                  runs the any-hit query and sets the light's bit in the record when visible.
//               wf_shade           one thread per record: Phong with the visibility bits, term = W * local,
//                                  children written to level l+1 and appended to its task list (compacted with
//                                  ballot/popc, one atomicAdd per warp).
//   per chunk:  wf_sum             one thread per sample: adds the terms of its tree in depth-first pre-order
//                                  (SPEC §4), writes the per-sample radiance the resolve kernel consumes.
//
// In the per-lane state machine of nt_bvh_trace.cuh a warp's traversal runs at ~13 of 32 lanes (lanes park
// while others finish, and shading interleaves with traversal); here traversal, shadow and shading are
// separate coherent passes.  The price is the record traffic (~100 B per tree node), a few percent of HBM
// bandwidth at the ray rates this path reaches.
#pragma once
#include "nt_bvh_trace.cuh"

#define NT_WF_MAX_DEPTH 6  // deeper trees use the per-lane state machine (heap layout: 2^depth - 1 records per sample)
#ifndef NT_WF_REFILL
#define NT_WF_REFILL 20    // idle lanes before a warp pulls new tasks (configs[3] f64: 4 -> 70.0 ms, 8 -> 69.8, 16 -> 62.5, 24 -> 62.7;
                           // with sparse pulls repeated and NT_DESCEND_MIN 12: 12 -> 57.0, 16 -> 55.4, 20 -> 54.8, 24 -> 54.9, 28 -> 56.1)
#endif

#ifndef NT_WF_SORT_DEFAULT
#define NT_WF_SORT_DEFAULT 0   // see wf_sort_mode: configs[3] f64 57.7 ms unsorted, 55.8 with 2, 56.0 with 6, 60.4 with 3 BEFORE the shadow grids
                               // (profiles/r04_wf_sort.txt); with them the shadow passes are half as long and the sort no longer pays: 38.4 / 39.3
#endif

#ifndef NT_WF_REFILL_SHADOW
#define NT_WF_REFILL_SHADOW NT_WF_REFILL // the occlusion passes' own threshold (their walks are short since the shadow grids): 12 / 20 / 26 -> 39.3 / 38.2 / 38.2 ms
#endif
#ifndef NT_WF_EARLY_RETIRE
#define NT_WF_EARLY_RETIRE 1
#endif
#ifndef NT_WF_REPULL
#define NT_WF_REPULL 1
#endif
struct NtWfLevel {
    void *ray;            // R[7][cap]: ox oy oz dx dy dz W   (levels >= 2; level 1 rays come from the camera)
    void *hit_t;          // R[cap]
    int *prim;            // kind << 28 | idx, -1 miss, -2 no ray
    unsigned *vis;        // bit l: light l faces the surface and is not occluded
    void *term;           // R[3][cap]: W * local (or W * background)
    unsigned char *kids;  // bit 0 reflection child, bit 1 transmission child
    unsigned *tasks;      // record indices with a ray at this level (level >= 2)
    unsigned cap;
};
struct NtWfArgs {
    NtWfLevel lv[NT_WF_MAX_DEPTH];
    unsigned *counts;             // [NT_WF_MAX_DEPTH + 1] tasks per level (level 1: samples of the chunk)
    unsigned long long *fetch;    // [2 * NT_WF_MAX_DEPTH] task-fetch cursors (nearest, shadow) per level
    unsigned sid0, n_samples;     // this chunk: first sample id, sample count
    unsigned level;               // 1-based
    // nearest-hit queries whose direction has drifted so far that their sphere pass would test every sphere
    // (nt_bvh_trace.cuh query_second_pass): record indices, swept by wf_sweep_kernel after the level's nearest pass
    unsigned *sweep_list, *sweep_count, sweep_cap;
    // ... and those with a moderate drift, whose sphere pass walks the sphere tree under the cone margin: the same list
    // mechanism, walked by the CONEPASS instantiation of wf_trace_kernel (the main kernel carries no cone code)
    unsigned *cone_list, *cone_count;
    unsigned long long *cone_fetch;
    // the task list the level's kernels read: NULL = record i is task i (level 1), else lv[level - 1].tasks or a sorted
    // copy of it (wf_sort_* below)
    const unsigned *cur;
    unsigned *sort_keys, *sort_out, *sort_hist; // scratch of the task sort: key per task, sorted list, bucket counters
};

namespace nt {

// Work counters of the traversal kernels in per-thread shared-memory slots: eight live registers less in a loop that
// spilled its binary32 box parameters (64 registers, 4 blocks per SM).  Only the box-test count, bumped at every node,
// stays in a register; the others change once per ray or per leaf.
__shared__ unsigned nt_wf_cnt[7][NT_BLOCK_THREADS];
template <int I> struct SmCounter {
    __device__ __forceinline__ void operator++(int) { nt_wf_cnt[I][threadIdx.x] += 1u; }
    __device__ __forceinline__ void operator+=(unsigned v) { nt_wf_cnt[I][threadIdx.x] += v; }
    __device__ __forceinline__ unsigned get() const { return nt_wf_cnt[I][threadIdx.x]; }
};
struct CountersSm {
    SmCounter<0> prim; SmCounter<1> sec; SmCounter<2> shadow; SmCounter<3> sph; SmCounter<4> pln; SmCounter<5> tri;
    unsigned box;
    SmCounter<6> light;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 7; ++i) nt_wf_cnt[i][threadIdx.x] = 0u;
        box = 0u;
    }
};
__device__ __forceinline__ void flush_counters(const CountersSm &k, unsigned long long *counters, unsigned long long *s_cnt) {
    const unsigned vals[NT_NCOUNTERS] = { k.prim.get(), k.sec.get(), k.shadow.get(), k.sph.get(), k.pln.get(), k.tri.get(), k.box, k.light.get(), 0u, 0u, 0u };
    flush_counter_values(vals, counters, s_cnt);
}

template <typename R> struct WfHit {
    V3<R> P, Ng;
    int mat;
};
// Hit point and geometric normal of a record (SPEC §4), exactly as lane_advance computes them.
template <typename R>
__device__ __forceinline__ WfHit<R> wf_hit(const Ctx<R, true> &c, const V3<R> &o, const V3<R> &d, R t, int prim) {
    const NtDevScene &s = *c.s;
    const NtSceneView<R> &v = *c.v;
    WfHit<R> h;
    h.P = { o.x + d.x * t, o.y + d.y * t, o.z + d.z * t };
    const int kind = prim >> 28, idx = prim & 0x0fffffff;
    if (kind == 0) {
        R p[4];
        c.ld_sph(idx, p);
        const R ir = __ldg(v.sph_invr + idx);
        h.Ng = { (h.P.x - p[0]) * ir, (h.P.y - p[1]) * ir, (h.P.z - p[2]) * ir };
        h.mat = __ldg(s.sph_mat + idx);
    } else if (kind == 1) {
        R p[4];
        c.ld_pln(idx, p);
        h.Ng = { p[0], p[1], p[2] };
        h.mat = __ldg(s.pln_mat + idx);
    } else {
        const R *tp = v.tri + (size_t)idx * NT_TRI_STRIDE + 9;
        h.Ng = { __ldg(tp), __ldg(tp + 1), __ldg(tp + 2) };
        h.mat = __ldg(s.tri_mat + idx);
    }
    return h;
}

// The ray of a record: level 1 from the camera (SPEC §2), deeper levels from the record array.
template <typename R>
__device__ __forceinline__ bool wf_ray(const NtRenderArgs &a, const NtWfArgs &w, unsigned rec, V3<R> &o, V3<R> &d, R &W) {
    const NtWfLevel &L = w.lv[w.level - 1];
    if (w.level == 1) {
        const SampleMap m = map_sample(a, w.sid0 + rec);
        if (!m.live) return false;
        const unsigned si = m.sidx % a.n, sj = m.sidx / a.n;
        const R ox = ArgsView<R>::samp_off(a, si), oy = ArgsView<R>::samp_off(a, sj); // (i + 0.5) / n (or i / n, SPEC §8), divided on the host
        const R fx = (R)m.px + ox, fy = (R)m.y + oy;
        const V3<R> D = { (ArgsView<R>::cam(a, 3) + ArgsView<R>::cam(a, 6) * fx) + ArgsView<R>::cam(a, 9) * fy,
                          (ArgsView<R>::cam(a, 4) + ArgsView<R>::cam(a, 7) * fx) + ArgsView<R>::cam(a, 10) * fy,
                          (ArgsView<R>::cam(a, 5) + ArgsView<R>::cam(a, 8) * fx) + ArgsView<R>::cam(a, 11) * fy };
        o = { ArgsView<R>::cam(a, 0), ArgsView<R>::cam(a, 1), ArgsView<R>::cam(a, 2) };
        d = scale(D, Math<R>::rcp(Math<R>::sqrt_(dot(D, D))));
        W = R(1);
        return true;
    }
    const R *ray = (const R *)L.ray;
    const size_t cap = L.cap;
    o = { ray[rec], ray[cap + rec], ray[2 * cap + rec] };
    d = { ray[3 * cap + rec], ray[4 * cap + rec], ray[5 * cap + rec] };
    W = ray[6 * cap + rec];
    return true;
}

// Persistent traversal kernel.  SHADOW == false: one task = one record, nearest hit.  SHADOW == true: one task =
// (record, light), any hit between the hit point and the light.
// CONEPASS: the second pass of the nearest-hit queries the main kernel deferred (NtWfArgs::cone_list): one task = one
// listed record, whose nearest hit among planes and triangles is already in the record; the sphere set is walked under
// the cone margin (query_arm pass 2) and the record keeps the nearer hit.
template <typename R, bool SHADOW, bool CONEPASS = false>
__global__ void __launch_bounds__(NT_BLOCK_THREADS, NT_MIN_BLOCKS_BVH)
wf_trace_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a, const __grid_constant__ NtWfArgs w) {
    __shared__ unsigned long long s_cnt[NT_NCOUNTERS];
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, true> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.eps_lo = (R)a.eps_lo; c.max_depth = a.max_depth; c.rules = a.rules;
    stage_scene<R, true>(s, v, c);
    const unsigned lane = threadIdx.x & 31;
    constexpr int REFILL = SHADOW ? NT_WF_REFILL_SHADOW : NT_WF_REFILL;
    const NtWfLevel &L = w.lv[w.level - 1];
    const unsigned n_rec = CONEPASS ? min(*w.cone_count, w.sweep_cap) : w.level == 1 ? w.n_samples : w.counts[w.level];
    const unsigned long long n_tasks = SHADOW ? (((unsigned long long)n_rec + 31) / 32) * 32 * s.nl : n_rec;
    unsigned long long *cursor = CONEPASS ? w.cone_fetch : w.fetch + 2 * (w.level - 1) + (SHADOW ? 1 : 0);
    CountersSm k;
    k.clear();

    BvhQuery<R> q;
    int2 bstack[NT_BVH_STACK];
    // Strict mode: the exact ray (12 registers) is needed at the leaves only, and with it live the inner step spilled
    // the binary32 box parameters and reloaded them for every node (12 local loads per visit): it waits in shared memory.
    constexpr bool RAY_IN_SMEM = sizeof(R) == 8;
    __shared__ R s_ray[RAY_IN_SMEM ? 6 : 1][RAY_IN_SMEM ? NT_BLOCK_THREADS : 1];
    q.done = true; q.cur = NT_REF_EMPTY; q.sp = 0; q.found = false; q.any = SHADOW;
    bool active = false, exhausted = false;
    unsigned rec = 0, light = 0;

    for (;;) {
        // ---- refill: idle lanes pull tasks, one atomic per warp ----
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (__popc(idle) >= REFILL || idle == 0xffffffffu)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(idle) >= n_tasks) exhausted = true;
            if (!active) {
                const unsigned long long ti = base + __popc(idle & ((1u << lane) - 1));
                if (ti < n_tasks) {
                    if constexpr (CONEPASS) {
                        rec = w.cone_list[ti];
                        V3<R> o, d;
                        R W;
                        wf_ray<R>(a, w, rec, o, d, W);
                        const int prim = L.prim[rec];
                        q.o = o; q.d = d; q.tb = ((const R *)L.hit_t)[rec]; q.any = false; q.found = prim >= 0; q.sp = 0;
                        q.best.kind = prim >= 0 ? prim >> 28 : -1; q.best.idx = prim >= 0 ? (prim & 0x0fffffff) : -1;
                        q.best.gid = prim < 0 ? 0x7fffffff : (prim >> 28) == 1 ? (int)s.ns + (prim & 0x0fffffff)
                                                           : __ldg(((prim >> 28) == 2 ? s.tri_gid : s.sph_gid) + (prim & 0x0fffffff));
                        query_arm<R>(c, q, o, d, 2);
                        if constexpr (RAY_IN_SMEM) {
                            s_ray[0][threadIdx.x] = o.x; s_ray[1][threadIdx.x] = o.y; s_ray[2][threadIdx.x] = o.z;
                            s_ray[3][threadIdx.x] = d.x; s_ray[4][threadIdx.x] = d.y; s_ray[5][threadIdx.x] = d.z;
                        }
                        active = true;
                    } else if constexpr (!SHADOW) {
                        rec = w.cur ? w.cur[ti] : (unsigned)ti;
                        V3<R> o, d;
                        R W;
                        if (wf_ray<R>(a, w, rec, o, d, W)) {
                            if (w.level == 1) { k.prim++; primary_query_start<R>(c, q, o, d, k); }
                            else query_start<R>(c, q, o, d, Math<R>::inf(), false, k);
                            if constexpr (RAY_IN_SMEM) {
                                s_ray[0][threadIdx.x] = o.x; s_ray[1][threadIdx.x] = o.y; s_ray[2][threadIdx.x] = o.z;
                                s_ray[3][threadIdx.x] = d.x; s_ray[4][threadIdx.x] = d.y; s_ray[5][threadIdx.x] = d.z;
                            }
                            active = true;
                        } else {
                            L.prim[rec] = -2;
                        }
                    } else {
                        bool continue_task = true;
                        // light-major inside groups of 32 records: the lanes of a warp get consecutive records and
                        // the same light, i.e. coherent shadow rays
                        const unsigned long long grp = ti / (32ull * s.nl);
                        const unsigned within = (unsigned)(ti - grp * 32ull * s.nl);
                        light = within >> 5;
                        const unsigned long long ri64 = grp * 32ull + (within & 31u);
                        if (ri64 >= n_rec) continue_task = false;
                        const unsigned ri = (unsigned)(ri64 < n_rec ? ri64 : 0);
                        rec = w.cur ? w.cur[ri] : ri;
                        const int prim = continue_task ? L.prim[rec] : -2;
                        if (prim >= 0) {
                            V3<R> o, d;
                            R W;
                            wf_ray<R>(a, w, rec, o, d, W);
                            const WfHit<R> h = wf_hit<R>(c, o, d, ((const R *)L.hit_t)[rec], prim);
                            const bool entering = dot(d, h.Ng) < R(0);
                            const V3<R> N = entering ? h.Ng : V3<R>{ -h.Ng.x, -h.Ng.y, -h.Ng.z };
                            const R *lp = v.lights + 6 * light;
                            const V3<R> Lv = { __ldg(lp) - h.P.x, __ldg(lp + 1) - h.P.y, __ldg(lp + 2) - h.P.z };
                            const R d2 = dot(Lv, Lv);
                            const R dist = Math<R>::sqrt_(d2);
                            const V3<R> Ld = scale(Lv, Math<R>::rcp(dist));
                            const R ndl = dot(N, Ld);
                            if (ndl > R(0)) {
                                k.shadow++;
                                shadow_query_start<R>(c, q, h.P, Ld, dist, light, k);
                                if constexpr (RAY_IN_SMEM) {
                                    s_ray[0][threadIdx.x] = h.P.x; s_ray[1][threadIdx.x] = h.P.y; s_ray[2][threadIdx.x] = h.P.z;
                                    s_ray[3][threadIdx.x] = Ld.x; s_ray[4][threadIdx.x] = Ld.y; s_ray[5][threadIdx.x] = Ld.z;
                                }
                                active = true;
                            }
                        }
                    }
                }
            }
        }
        // an occlusion query can end where it starts - a listed sphere of the light's shadow grid occludes, the ray leaves the
        // scene's bounds at once: such a lane retires here and counts as idle, or the warp would walk with the few lanes whose
        // queries are real (the level-2 shadow pass of configs[3] ran at 11 of 32 lanes)
        if constexpr (SHADOW && NT_WF_EARLY_RETIRE) {
            if (active && q.done) {
                if (!q.found) atomicOr(L.vis + rec, 1u << light);
                active = false;
            }
        }
        const unsigned started = __ballot_sync(0xffffffffu, active);
        if (started == 0) {
            if (exhausted) break;
            continue;
        }
        // sparse tasks - records that hit nothing, lights behind the surface, samples of the last ragged tile start no query -:
        // pull again before walking, or the warp walks with whatever share of its lanes the list happened to fill
        if (NT_WF_REPULL && !exhausted && __popc(~started) >= REFILL) continue;
        // ---- traversal rounds (as in render_bvh_kernel) until enough lanes have finished ----
        for (;;) {
            for (;;) {
                const bool inner = active && !q.done && q.cur >= 0;
                const unsigned im = __ballot_sync(0xffffffffu, inner);
                if (im == 0) break;
                if (inner) query_inner_step_t<CONEPASS, R>(c, q, bstack, k);
                if (__popc(im) < NT_DESCEND_MIN &&
                    __ballot_sync(0xffffffffu, active && !q.done && ref_is_leaf(q.cur)) != 0) break;
            }
            const bool leaf = active && !q.done && ref_is_leaf(q.cur);
            if (leaf) {
                if constexpr (RAY_IN_SMEM) {
                    const V3<R> o = { s_ray[0][threadIdx.x], s_ray[1][threadIdx.x], s_ray[2][threadIdx.x] };
                    const V3<R> d = { s_ray[3][threadIdx.x], s_ray[4][threadIdx.x], s_ray[5][threadIdx.x] };
                    query_leaf_step<R>(c, q, o, d, bstack, k);
                } else {
                    query_leaf_step<R>(c, q, q.o, q.d, bstack, k);
                }
            }
            // drifted directions (query_arm): the sphere set after the triangle set - a walk under the cone margin, or the
            // whole warp sweeping every sphere for a ray whose cone holds the scene anyway
            if constexpr (!CONEPASS && sizeof(R) == 8) {
                // a drifted direction (query_arm) has walked the triangle set only: its sphere pass is deferred - to the
                // sweep list when the cone would hold the scene anyway, else to the cone list.  Shadow rays are normalised
                // and never get here; a full list leaves the lane to walk the sphere set itself (constant margin).
                if (active && q.done && q.pass == 1) {
                    bool deferred = false;
                    if constexpr (!SHADOW) {
                        V3<R> d = q.d;
                        if constexpr (RAY_IN_SMEM) d = { s_ray[3][threadIdx.x], s_ray[4][threadIdx.x], s_ray[5][threadIdx.x] };
                        const bool wide = query_wants_sweep<R>(d);
                        const unsigned slot = atomicAdd(wide ? w.sweep_count : w.cone_count, 1u);
                        if (slot < w.sweep_cap) { (wide ? w.sweep_list : w.cone_list)[slot] = rec; q.pass = 2; deferred = true; }
                    }
                    if (!deferred && !(q.any && q.found)) {
                        V3<R> o = q.o, d = q.d;
                        if constexpr (RAY_IN_SMEM) {
                            o = { s_ray[0][threadIdx.x], s_ray[1][threadIdx.x], s_ray[2][threadIdx.x] };
                            d = { s_ray[3][threadIdx.x], s_ray[4][threadIdx.x], s_ray[5][threadIdx.x] };
                        }
                        query_arm<R>(c, q, o, d, 3);
                    }
                }
            }
            const unsigned parked = __ballot_sync(0xffffffffu, !active || q.done);
            if (parked == 0xffffffffu || (!exhausted && __popc(parked) >= REFILL)) break;
        }
        // ---- retire finished queries ----
        if (active && q.done) {
            if constexpr (!SHADOW) {
                ((R *)L.hit_t)[rec] = q.tb;
                L.prim[rec] = q.best.kind >= 0 ? (q.best.kind << 28) | q.best.idx : -1;
                if constexpr (!CONEPASS) L.vis[rec] = 0u;
            } else {
                if (!q.found) atomicOr(L.vis + rec, 1u << light);
            }
            active = false;
        }
    }
    flush_counters(k, a.counters, s_cnt);
}

// Deferred sphere sweeps of one level (see NtWfArgs::sweep_list): one warp per listed record.  The record holds the
// nearest hit among planes and triangles (the query's pass 1); the warp tests EVERY sphere with the exact rule and the
// record keeps the nearer of the two, ties to the smaller global id - what the walk of the grown sphere tree would have
// returned.
template <typename R>
__global__ void __launch_bounds__(NT_BLOCK_THREADS)
wf_sweep_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a, const __grid_constant__ NtWfArgs w) {
    __shared__ unsigned long long s_cnt[NT_NCOUNTERS];
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    const NtWfLevel &L = w.lv[w.level - 1];
    const unsigned n = min(*w.sweep_count, w.sweep_cap), lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (unsigned e = warp; e < n; e += n_warps) {
        const unsigned rec = w.sweep_list[e];
        V3<R> o, d;
        R W;
        wf_ray<R>(a, w, rec, o, d, W);
        const SweepHit<R> h = sweep_spheres<R>(v.sph, s.sph_gid, s.ns, (R)a.eps, o, d);
        k.sph += (s.ns + 31u - lane) / 32u;
        if (lane == 0 && h.idx >= 0) {
            const R tb = ((const R *)L.hit_t)[rec];
            const int prim = L.prim[rec];
            int gid = 0x7fffffff;
            if (prim >= 0) gid = (prim >> 28) == 1 ? (int)s.ns + (prim & 0x0fffffff) : __ldg(((prim >> 28) == 2 ? s.tri_gid : s.sph_gid) + (prim & 0x0fffffff));
            if (h.t < tb || (h.t == tb && h.gid < gid)) {
                ((R *)L.hit_t)[rec] = h.t;
                L.prim[rec] = h.idx; // kind 0
            }
        }
    }
    flush_counters(k, a.counters, s_cnt);
}

// Shade one level: term = W * local (SPEC §4) and the children of every record that has a ray.
template <typename R>
__global__ void __launch_bounds__(NT_BLOCK_THREADS)
wf_shade_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a, const __grid_constant__ NtWfArgs w) {
    __shared__ unsigned long long s_cnt[NT_NCOUNTERS];
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, true> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.eps_lo = (R)a.eps_lo; c.max_depth = a.max_depth; c.rules = a.rules;
    stage_scene<R, true>(s, v, c);
    const NtWfLevel &L = w.lv[w.level - 1];
    const unsigned n_rec = w.level == 1 ? w.n_samples : w.counts[w.level];
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };
    const unsigned lane = threadIdx.x & 31;
    const unsigned stride = gridDim.x * blockDim.x;
    // whole warps iterate together (the child append uses warp ballots)
    for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_rec; base += stride) {
        const unsigned i = base + lane;
        bool refl = false, trans = false;
        V3<R> P = { R(0), R(0), R(0) }, Rd = P, T = P;
        R Wr = R(0), Wt = R(0);
        unsigned rec = 0;
        if (i < n_rec) {
            rec = w.cur ? w.cur[i] : i;
            const int prim = L.prim[rec];
            R *term = (R *)L.term;
            const size_t cap = L.cap;
            unsigned char kids = 0;
            if (prim != -2) {
                V3<R> o, d;
                R W;
                wf_ray<R>(a, w, rec, o, d, W);
                if (prim < 0) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) term[ch * cap + rec] = W * __ldg(v.globals + 3 + ch);
                } else {
                    const WfHit<R> h = wf_hit<R>(c, o, d, ((const R *)L.hit_t)[rec], prim);
                    P = h.P;
                    const R *mp = v.mat + (size_t)h.mat * NT_MAT_STRIDE;
                    R m0[4], m1[4];
                    Ld<R>::g4(mp, m0);     // r g b ka
                    Ld<R>::g4(mp + 4, m1); // kd ks shininess kr
                    const bool entering = dot(d, h.Ng) < R(0);
                    const V3<R> N = entering ? h.Ng : V3<R>{ -h.Ng.x, -h.Ng.y, -h.Ng.z };
                    R local[3];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) local[ch] = __ldg(v.globals + ch) * (m0[3] * m0[ch]);
                    const unsigned vis = L.vis[rec];
                    for (unsigned l = 0; l < s.nl; ++l) {
                        if (!((vis >> l) & 1u)) continue;
                        k.light++;
                        const R *lp = v.lights + 6 * l;
                        const V3<R> Lv = { __ldg(lp) - P.x, __ldg(lp + 1) - P.y, __ldg(lp + 2) - P.z };
                        const R d2 = dot(Lv, Lv);
                        const R dist = Math<R>::sqrt_(d2);
                        const V3<R> Ld_ = scale(Lv, Math<R>::rcp(dist));
                        const R ndl = dot(N, Ld_);
                        R lc[3] = { __ldg(lp + 3), __ldg(lp + 4), __ldg(lp + 5) };
                        if (a.rules & NT_DEV_RULE_ATTENUATE) { // SPEC §8
                            const R att = Math<R>::rcp(d2);
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) lc[ch] = lc[ch] * att;
                        }
                        const R kdn = m1[0] * ndl;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) local[ch] = local[ch] + lc[ch] * (m0[ch] * kdn);
                        const R two = R(2) * ndl;
                        const V3<R> Rv = { N.x * two - Ld_.x, N.y * two - Ld_.y, N.z * two - Ld_.z };
                        const R rv = -dot(Rv, d);
                        if (m1[1] > R(0) && rv > R(0)) {
                            const R sterm = m1[1] * Math<R>::pow_(rv, m1[2]);
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) local[ch] = local[ch] + lc[ch] * sterm;
                        }
                    }
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) term[ch * cap + rec] = W * local[ch];
                    if (w.level < a.max_depth) {
                        R m2[4];
                        Ld<R>::g4(mp + 8, m2); // kt ior inv_ior pad
                        const R kr = m1[3], kt = m2[0];
                        const R cosi = -dot(d, N);
                        R wr = kr, wt = R(0);
                        if (kt > R(0)) {
                            const R eta = entering ? m2[2] : m2[1];
                            const R kk = R(1) - (eta * eta) * (R(1) - cosi * cosi);
                            if (kk < R(0)) wr = kr + kt;
                            else {
                                wt = kt;
                                const R sterm = eta * cosi - Math<R>::sqrt_(kk);
                                T = { d.x * eta + N.x * sterm, d.y * eta + N.y * sterm, d.z * eta + N.z * sterm };
                                if (fast_renormalises<R>::value || (a.rules & NT_DEV_RULE_RENORMALIZE)) T = scale(T, Math<R>::rcp(Math<R>::sqrt_(dot(T, T)))); // SPEC §8
                            }
                        }
                        if (wt > R(0)) { k.sec++; trans = true; Wt = W * wt; }
                        if (wr > R(0)) {
                            k.sec++;
                            refl = true; Wr = W * wr;
                            const R two = R(2) * cosi;
                            Rd = { d.x + N.x * two, d.y + N.y * two, d.z + N.z * two };
                            if (fast_renormalises<R>::value || (a.rules & NT_DEV_RULE_RENORMALIZE)) Rd = scale(Rd, Math<R>::rcp(Math<R>::sqrt_(dot(Rd, Rd))));
                        }
                        kids = (unsigned char)((refl ? 1 : 0) | (trans ? 2 : 0));
                    }
                }
            }
            L.kids[rec] = kids;
        }
        // ---- children -> level + 1 records (heap layout) and its compacted task list ----
        if (w.level < a.max_depth) { // uniform
            const NtWfLevel &C = w.lv[w.level];
            const unsigned mr = __ballot_sync(0xffffffffu, refl), mt = __ballot_sync(0xffffffffu, trans);
            const unsigned total = __popc(mr) + __popc(mt);
            if (total) {
                unsigned pos = 0;
                if (lane == 0) pos = atomicAdd(w.counts + w.level + 1, total);
                pos = __shfl_sync(0xffffffffu, pos, 0);
                R *ray = (R *)C.ray;
                const size_t cap = C.cap;
                const unsigned below = (1u << lane) - 1;
                if (refl) {
                    const unsigned cr = 2 * rec;
                    ray[cr] = P.x; ray[cap + cr] = P.y; ray[2 * cap + cr] = P.z;
                    ray[3 * cap + cr] = Rd.x; ray[4 * cap + cr] = Rd.y; ray[5 * cap + cr] = Rd.z;
                    ray[6 * cap + cr] = Wr;
                    C.tasks[pos + __popc(mr & below)] = cr;
                }
                if (trans) {
                    const unsigned cr = 2 * rec + 1;
                    ray[cr] = P.x; ray[cap + cr] = P.y; ray[2 * cap + cr] = P.z;
                    ray[3 * cap + cr] = T.x; ray[4 * cap + cr] = T.y; ray[5 * cap + cr] = T.z;
                    ray[6 * cap + cr] = Wt;
                    C.tasks[pos + __popc(mr) + __popc(mt & below)] = cr;
                }
            }
        }
    }
    flush_counters(k, a.counters, s_cnt);
}

// Per sample: the terms of its ray tree added in depth-first pre-order, reflection subtree before transmission
// subtree (SPEC §4), into a running sum that starts at 0.
template <typename R>
__global__ void __launch_bounds__(256)
wf_sum_kernel(const __grid_constant__ NtRenderArgs a, const __grid_constant__ NtWfArgs w) {
    R *samples = (R *)a.samples;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < w.n_samples; i += gridDim.x * blockDim.x) {
        R acc[3] = { R(0), R(0), R(0) };
        unsigned level = 1, path = 0;
        unsigned pend[NT_WF_MAX_DEPTH]; // deferred transmission children: level << 16 | path  (path < 2^5)
        int sp = 0;
        if (w.lv[0].prim[i] != -2) {
            for (;;) {
                const NtWfLevel &L = w.lv[level - 1];
                const unsigned rec = (i << (level - 1)) | path;
                const R *term = (const R *)L.term;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) acc[ch] = acc[ch] + term[(size_t)ch * L.cap + rec];
                const unsigned kids = L.kids[rec];
                if ((kids & 3u) == 3u) pend[sp++] = (level + 1) << 16 | (2 * path + 1);
                if (kids & 1u) { ++level; path = 2 * path; continue; }
                if (kids & 2u) { ++level; path = 2 * path + 1; continue; }
                if (sp == 0) break;
                const unsigned e = pend[--sp];
                level = e >> 16; path = e & 0xffffu;
            }
        }
        R *dst = samples + 3 * (size_t)(w.sid0 + i);
        dst[0] = acc[0]; dst[1] = acc[1]; dst[2] = acc[2];
    }
}

// ---- task sort (NT_WF_SORT) ---------------------------------------------------------------------------------------------
// A level's task list is in creation order: the children of neighbouring samples, reflection and transmission rays
// alternating.  Before a traversal pass the list can be bucket-sorted by where its rays start - a Morton code of the
// origin's cell in the scene bounds, then the direction's octant (nearest-hit pass), or the cell of the hit point the
// shadow rays leave from (shadow pass) - so that the lanes of a warp walk the same nodes.  Three small kernels: keys +
// bucket histogram, exclusive scan, scatter; atomics are warp-aggregated with match.any, equal keys of one warp keep their
// order.  The order of a task list never changes a result (every task writes its own record), only the schedule.
#define NT_WF_SORT_BITS 18
#define NT_WF_SORT_BUCKETS (1u << NT_WF_SORT_BITS)

__device__ __forceinline__ unsigned wf_spread3(unsigned v) { // bits of v (<= 6 of them) to every third position
    v &= 0x3fu;
    v = (v | (v << 8)) & 0x0000300fu;
    v = (v | (v << 4)) & 0x000030c3u;
    v = (v | (v << 2)) & 0x00009249u;
    return v;
}
template <int BITS>
__device__ __forceinline__ unsigned wf_cell_code(const NtDevScene &s, float x, float y, float z) {
    const float n = (float)(1 << BITS);
    const float fx = (x - s.blo[0]) * (n / fmaxf(s.bhi[0] - s.blo[0], 1e-30f));
    const float fy = (y - s.blo[1]) * (n / fmaxf(s.bhi[1] - s.blo[1], 1e-30f));
    const float fz = (z - s.blo[2]) * (n / fmaxf(s.bhi[2] - s.blo[2], 1e-30f));
    const unsigned ix = (unsigned)fminf(fmaxf(fx, 0.f), n - 1.f), iy = (unsigned)fminf(fmaxf(fy, 0.f), n - 1.f), iz = (unsigned)fminf(fmaxf(fz, 0.f), n - 1.f);
    return wf_spread3(ix) | wf_spread3(iy) << 1 | wf_spread3(iz) << 2;
}

// HITPOINT == false: key of the record's ray (5 bits per axis + octant); true: key of what it hit (records without a hit go
// to the last bucket: their shadow tasks end at once).
template <typename R, bool HITPOINT>
__global__ void __launch_bounds__(256)
wf_sort_hist_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a, const __grid_constant__ NtWfArgs w) {
    const NtWfLevel &L = w.lv[w.level - 1];
    const unsigned n = w.level == 1 ? w.n_samples : w.counts[w.level];
    const unsigned stride = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned i = base + lane;
        unsigned key = 0xffffffffu;
        if (i < n) {
            const unsigned rec = w.cur ? w.cur[i] : i;
            V3<R> o, d;
            R W;
            key = NT_WF_SORT_BUCKETS - 1;
            if constexpr (HITPOINT) {
                // where the shadow rays start = what the record hit.  The primitive arrays of a BVH scene are in the
                // depth-first order of the tree, so the index itself is a spatial code at every scale, and it costs one
                // 4-byte gather (the hit point's cell needs the ray and t: 7 gathers of R, 2.2 ms per frame of configs[3]
                // against 3.2 ms saved in the shadow passes).  Planes are unbounded: those few records use the cell.
                const int prim = L.prim[rec];
                if (prim >= 0) {
                    const unsigned kind = (unsigned)prim >> 28, idx = (unsigned)prim & 0x0fffffffu;
                    if (kind == 2) key = (unsigned)(((unsigned long long)idx << (NT_WF_SORT_BITS - 1)) / s.nt);
                    else if (kind == 0) key = (1u << (NT_WF_SORT_BITS - 1)) + (unsigned)(((unsigned long long)idx << (NT_WF_SORT_BITS - 2)) / s.ns);
                    else if (wf_ray<R>(a, w, rec, o, d, W)) {
                        const R t = ((const R *)L.hit_t)[rec];
                        key = (3u << (NT_WF_SORT_BITS - 2)) + wf_cell_code<5>(s, (float)(o.x + d.x * t), (float)(o.y + d.y * t), (float)(o.z + d.z * t));
                    }
                }
            } else if (wf_ray<R>(a, w, rec, o, d, W)) {
                key = wf_cell_code<5>(s, (float)o.x, (float)o.y, (float)o.z) << 3 | (d.x < R(0) ? 1u : 0u) | (d.y < R(0) ? 2u : 0u) | (d.z < R(0) ? 4u : 0u);
            }
            w.sort_keys[i] = key;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (i < n && lane == (unsigned)__ffs(peers) - 1) atomicAdd(w.sort_hist + key, (unsigned)__popc(peers));
    }
}

// Exclusive scan of the bucket counters in place, two launches: every block scans 1024 counters and posts its total, then
// every block adds the totals of the blocks before it.  (One block walking all 2^18 counters took 0.45 ms.)
static __device__ __forceinline__ unsigned wf_block_scan_1024(unsigned v, unsigned *s_warp, unsigned &total) { // inclusive
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += u; }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        unsigned t = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, t, o); if ((int)lane >= o) t += u; }
        s_warp[lane] = t;
    }
    __syncthreads();
    total = s_warp[31];
    return v + (wid ? s_warp[wid - 1] : 0u);
}
static __global__ void __launch_bounds__(1024) wf_sort_scan1_kernel(unsigned *hist, unsigned *sums) {
    __shared__ unsigned s_warp[32];
    const unsigned i = blockIdx.x * 1024 + threadIdx.x, c = hist[i];
    unsigned total;
    const unsigned inc = wf_block_scan_1024(c, s_warp, total);
    hist[i] = inc - c;
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(1024) wf_sort_scan2_kernel(unsigned *hist, const unsigned *sums) {
    __shared__ unsigned s_warp[32];
    static_assert(NT_WF_SORT_BUCKETS / 1024 <= 1024, "one block adds the block totals up");
    unsigned total;
    wf_block_scan_1024(threadIdx.x < blockIdx.x ? sums[threadIdx.x] : 0u, s_warp, total);
    hist[blockIdx.x * 1024 + threadIdx.x] += total;
}

static __global__ void __launch_bounds__(256)
wf_sort_scatter_kernel(const __grid_constant__ NtWfArgs w) {
    const unsigned n = w.level == 1 ? w.n_samples : w.counts[w.level];
    const unsigned stride = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned i = base + lane;
        const unsigned key = i < n ? w.sort_keys[i] : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const unsigned leader = (unsigned)__ffs(peers) - 1;
        unsigned pos = 0;
        if (i < n && lane == leader) pos = atomicAdd(w.sort_hist + key, (unsigned)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (i < n) w.sort_out[pos + __popc(peers & ((1u << lane) - 1))] = w.cur ? w.cur[i] : i;
    }
}

// Scratch of the task sort for chunks of S samples: keys and the sorted list at the capacity of the deepest level, plus the
// bucket counters.  Only workspaces that dwarf the counters sort at all (tests run the pipeline in 1 MB).
inline size_t wf_sort_fixed_bytes() { return 4 * (size_t)NT_WF_SORT_BUCKETS + 4096 + 1024; } // + the alignment of three arrays
inline size_t wf_sort_bytes_per_sample(unsigned depth) { return (size_t)8 << (depth - 1); }
inline unsigned wf_sort_mode() { // bit 0: rays of levels >= 2 before the nearest-hit pass; bit 1: hit points of levels >= 2 before the shadow pass; bit 2: hit points of level 1
    const char *e = getenv("NT_WF_SORT");
    return e ? (unsigned)atoi(e) & 7u : NT_WF_SORT_DEFAULT;
}

// Bytes of workspace per sample for trees of depth D, and the layout of one chunk inside the workspace.
template <typename R>
inline size_t wf_bytes_per_sample(unsigned depth) {
    size_t b = 0;
    for (unsigned l = 1; l <= depth; ++l) {
        const size_t per = (l >= 2 ? 7 * sizeof(R) : 0) + sizeof(R) + 4 + 4 + 3 * sizeof(R) + 1 + (l >= 2 ? 4 : 0);
        b += per << (l - 1);
    }
    return b + 16; // slack for alignment
}

// Room for the deferred-sweep list of a chunk of n samples (strict mode only): one entry per 16 samples, 64 K to 4 M entries.
template <typename R>
inline size_t wf_sweep_bytes(size_t n_samples) {
    if (sizeof(R) != 8) return 0;
    size_t e = n_samples / 16;
    e = e < (64u << 10) ? (64u << 10) : e > (4u << 20) ? (4u << 20) : e;
    return 4 * e;
}

template <typename R>
inline int launch_wavefront(const NtDevScene &s, const NtRenderArgs &a, cudaStream_t st, int sms, int blocks_per_sm) {
    const unsigned depth = a.max_depth;
    const unsigned n_sids = a.tiles_x * a.tiles_y * (a.spp / a.lanes) * 32;
    const size_t header = 512;
    if (a.wf_bytes <= header + 4096) return (int)cudaErrorInvalidValue;
    const size_t sweep_bytes = a.wf_bytes > 64 * wf_sweep_bytes<R>(n_sids) ? wf_sweep_bytes<R>(n_sids) : 0; // a tiny workspace keeps everything for records
    unsigned sort_mode = wf_sort_mode();
    if (a.wf_bytes < 4 * wf_sort_fixed_bytes() || s.nl == 0) sort_mode = 0; // tiny workspaces keep everything for records
    const size_t sort_fixed = sort_mode ? wf_sort_fixed_bytes() : 0;
    size_t S = (a.wf_bytes - header - 256 * 8 * depth - sweep_bytes - sort_fixed) / (wf_bytes_per_sample<R>(depth) + (sort_mode ? wf_sort_bytes_per_sample(depth) : 0));
    S &= ~(size_t)31;
    if (S > n_sids) S = n_sids;
    const size_t max_s = ((size_t)1 << 31) >> (depth - 1); // record indices must fit 32 bits
    if (S > max_s) S = max_s & ~(size_t)31;
    if (S < 32) return (int)cudaErrorInvalidValue;

    NtWfArgs w;
    memset(&w, 0, sizeof w);
    unsigned char *p = (unsigned char *)a.wf;
    w.counts = (unsigned *)p;                                  // [NT_WF_MAX_DEPTH + 2]
    w.fetch = (unsigned long long *)(p + 64);                  // [2 * NT_WF_MAX_DEPTH]
    w.sweep_count = (unsigned *)(p + 192);                     // [NT_WF_MAX_DEPTH]: one counter per level (the header is zeroed per chunk)
    w.cone_count = (unsigned *)(p + 224);                      // [NT_WF_MAX_DEPTH]
    w.cone_fetch = (unsigned long long *)(p + 256);            // [NT_WF_MAX_DEPTH]
    p += header;
    auto take = [&](size_t bytes) { void *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    for (unsigned l = 1; l <= depth; ++l) {
        NtWfLevel &L = w.lv[l - 1];
        const size_t cap = S << (l - 1);
        L.cap = (unsigned)cap;
        L.ray = l >= 2 ? take(7 * sizeof(R) * cap) : nullptr;
        L.hit_t = take(sizeof(R) * cap);
        L.term = take(3 * sizeof(R) * cap);
        L.prim = (int *)take(4 * cap);
        L.vis = (unsigned *)take(4 * cap);
        L.tasks = l >= 2 ? (unsigned *)take(4 * cap) : nullptr;
        L.kids = (unsigned char *)take(cap);
    }
    if (sort_mode) {
        const size_t cap = S << (depth - 1);
        w.sort_keys = (unsigned *)take(4 * cap);
        w.sort_out = (unsigned *)take(4 * cap);
        w.sort_hist = (unsigned *)take(4 * (size_t)NT_WF_SORT_BUCKETS + 4096); // + the scan's block totals
    }
    // deferred sphere sweeps: whatever is left of the workspace, at most 4 M entries (a full list only means that the
    // remaining rays walk the sphere tree themselves)
    {
        const size_t used = (size_t)(p - (unsigned char *)a.wf);
        if (used > a.wf_bytes) return (int)cudaErrorInvalidValue;
        size_t cap = (a.wf_bytes - used) / 8; // two lists
        if (cap > sweep_bytes / 8) cap = sweep_bytes / 8;
        w.sweep_list = (unsigned *)p;
        w.cone_list = w.sweep_list + cap;
        w.sweep_cap = (unsigned)cap;
    }

    unsigned *const sort_scratch = w.sort_out;
    const size_t smem = flat_smem_bytes<R>(s, true);
    const unsigned grid_full = (unsigned)(sms * blocks_per_sm);
    for (unsigned sid0 = 0; sid0 < n_sids; sid0 += (unsigned)S) {
        w.sid0 = sid0;
        w.n_samples = (unsigned)(n_sids - sid0 < S ? n_sids - sid0 : S);
        cudaMemsetAsync(a.wf, 0, header, st);
        for (unsigned l = 1; l <= depth; ++l) {
            w.level = l;
            // an upper bound of the level's records (the exact count lives on the device)
            const size_t bound = (size_t)w.n_samples << (l - 1);
            unsigned grid = grid_full;
            const size_t warps_needed = (bound + 31) / 32, wpb = NT_BLOCK_THREADS / 32;
            if (grid > (warps_needed + wpb - 1) / wpb) grid = (unsigned)((warps_needed + wpb - 1) / wpb);
            unsigned *sweep_counts = w.sweep_count, *cone_counts = w.cone_count;
            unsigned long long *cone_fetches = w.cone_fetch;
            w.sweep_count = sweep_counts + (l - 1); w.cone_count = cone_counts + (l - 1); w.cone_fetch = cone_fetches + (l - 1);
            // the level's task list, and (NT_WF_SORT) its sorted copies: the level's own buffer and the scratch take turns
            w.cur = l >= 2 ? w.lv[l - 1].tasks : nullptr;
            unsigned *spare = sort_scratch;
            unsigned sortgrid = (unsigned)(sms * 8);
            if (sortgrid > (bound + 255) / 256) sortgrid = (unsigned)((bound + 255) / 256);
            auto sort_tasks = [&](bool hitpoint) {
                cudaMemsetAsync(w.sort_hist, 0, 4 * (size_t)NT_WF_SORT_BUCKETS, st);
                if (hitpoint) wf_sort_hist_kernel<R, true><<<sortgrid, 256, 0, st>>>(s, a, w);
                else wf_sort_hist_kernel<R, false><<<sortgrid, 256, 0, st>>>(s, a, w);
                wf_sort_scan1_kernel<<<NT_WF_SORT_BUCKETS / 1024, 1024, 0, st>>>(w.sort_hist, w.sort_hist + NT_WF_SORT_BUCKETS);
                wf_sort_scan2_kernel<<<NT_WF_SORT_BUCKETS / 1024, 1024, 0, st>>>(w.sort_hist, w.sort_hist + NT_WF_SORT_BUCKETS);
                unsigned *dst = spare;
                w.sort_out = dst;
                wf_sort_scatter_kernel<<<sortgrid, 256, 0, st>>>(w);
                spare = dst == sort_scratch ? w.lv[l - 1].tasks : sort_scratch; // a second sort of this level goes back into its own buffer (level 1 has none)
                w.cur = dst;
                if (a.n_launches) *a.n_launches += 4;
            };
            if (l >= 2 && (sort_mode & 1u)) sort_tasks(false);
            wf_trace_kernel<R, false><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a, w);
            // second passes of drifted directions: a direction leaves unit length at a sphere bounce, so from level 2 on
            const bool second = sizeof(R) == 8 && l >= 2 && s.ns > 0 && w.sweep_cap > 0;
            if constexpr (sizeof(R) == 8) {
                if (second) {
                    wf_sweep_kernel<R><<<sms, NT_BLOCK_THREADS, 0, st>>>(s, a, w);
                    wf_trace_kernel<R, false, true><<<sms, NT_BLOCK_THREADS, smem, st>>>(s, a, w);
                }
            }
            if (s.nl && (sort_mode & (l >= 2 ? 2u : 4u)) && spare) sort_tasks(true);
            if (s.nl) wf_trace_kernel<R, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a, w);
            w.sweep_count = sweep_counts; w.cone_count = cone_counts; w.cone_fetch = cone_fetches;
            if (a.n_launches) *a.n_launches += (s.nl ? 3 : 2) + (second ? 2 : 0);
            unsigned sgrid = (unsigned)(sms * 8);
            if (sgrid > (bound + NT_BLOCK_THREADS - 1) / NT_BLOCK_THREADS) sgrid = (unsigned)((bound + NT_BLOCK_THREADS - 1) / NT_BLOCK_THREADS);
            // shading reads and writes whole records: the creation-order list (coalesced) while it still exists
            if (l == 1) w.cur = nullptr;
            else if (w.cur == sort_scratch && !(sort_mode & 1u)) w.cur = w.lv[l - 1].tasks;
            wf_shade_kernel<R><<<sgrid, NT_BLOCK_THREADS, smem, st>>>(s, a, w);
        }
        unsigned ggrid = (unsigned)(sms * 8);
        if (ggrid > (w.n_samples + 255) / 256) ggrid = (w.n_samples + 255) / 256;
        wf_sum_kernel<R><<<ggrid, 256, 0, st>>>(a, w);
        if (a.n_launches) *a.n_launches += 1;
    }
    resolve_kernel<R><<<dim3((a.width + 255) / 256, a.vrows), 256, 0, st>>>(a);
    if (a.n_launches) *a.n_launches += 1;
    return (int)cudaGetLastError();
}

} // namespace nt
