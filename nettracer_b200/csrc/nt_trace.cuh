// nt_trace.cuh — the intersect-and-shade hot path, templated on the arithmetic
// type R: double = strict mode (this header is then compiled with -fmad=false so no multiply-add
// is fused and every operation happens in the order SPEC-PROVISIONAL.md writes it), float = fast
// mode (FMA contraction, approximate div/sqrt/pow).  No reference file can be cited — the spec is
// this repository's own (SPEC-PROVISIONAL.md §n is quoted beside each routine).
//
// One thread traces one image sample: its whole ray tree, depth-first in pre-order with an explicit
// stack (SPEC §4).  `lanes` adjacent lanes own the samples of one pixel and combine them with
// shuffles in sample order (SPEC §5), so supersampling never leaves registers.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include <type_traits>

#include "nt_device.h"

// what-if timing hooks (experiments only; the shipped build defines none of them)
#ifdef NT_EXP_NOPLANES
#define NT_EXP_NP(x) 0u
#elif defined(NT_EXP_FIXED_NP)
#define NT_EXP_NP(x) ((unsigned)NT_EXP_FIXED_NP) // what-if: compile-time primitive counts (full unrolling)
#else
#define NT_EXP_NP(x) (x)
#endif
#ifdef NT_EXP_NOSPHERES
#define NT_EXP_NS(x) 0u
#elif defined(NT_EXP_FIXED_NS)
#define NT_EXP_NS(x) ((unsigned)NT_EXP_FIXED_NS)
#else
#define NT_EXP_NS(x) (x)
#endif

namespace nt {

template <typename R> struct V3 { R x, y, z; };

template <typename R> __device__ __forceinline__ R dot(const V3<R> &a, const V3<R> &b) {
    return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
template <typename R> __device__ __forceinline__ V3<R> cross(const V3<R> &a, const V3<R> &b) {
    return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
}
template <typename R> __device__ __forceinline__ V3<R> sub(const V3<R> &a, const V3<R> &b) {
    return { a.x - b.x, a.y - b.y, a.z - b.z };
}
template <typename R> __device__ __forceinline__ V3<R> scale(const V3<R> &a, R s) {
    return { a.x * s, a.y * s, a.z * s };
}

template <typename R> struct Math;
template <> struct Math<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
#ifdef NT_EXP_NOPOW
    static __device__ __forceinline__ double pow_(double a, double b) { return a * b; }
#else
    static __device__ __forceinline__ double pow_(double a, double b) { return pow(a, b); }
#endif
    static __device__ __forceinline__ float up(double x) { return __double2float_ru(x); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};
template <> struct Math<float> {
    static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float pow_(float a, float b) { return __powf(a, b); }
    static __device__ __forceinline__ float up(float x) { return x; }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};

// ---- loads: 128-bit; read-only path for global data, ld.shared by 32-bit shared address for the
// staged data (an explicit state space: a generic pointer made the compiler rebuild the shared
// window base inside the loops) ----
template <typename R> struct Ld;
template <> struct Ld<double> {
    static __device__ __forceinline__ void g4(const double *p, double *o) {
        const double2 a = __ldg((const double2 *)p), b = __ldg((const double2 *)p + 1);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
    static __device__ __forceinline__ void s4(unsigned addr, double *o) {
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(o[2]), "=d"(o[3]) : "r"(addr));
    }
    static __device__ __forceinline__ void g9(const double *p, double *o) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double2 a = __ldg((const double2 *)p + i);
            o[2 * i] = a.x;
            if (i < 4) o[2 * i + 1] = a.y;
        }
    }
    static __device__ __forceinline__ void s9(unsigned addr, double *o) {
        double pad;
        (void)&pad;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(o[2]), "=d"(o[3]) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+32];" : "=d"(o[4]), "=d"(o[5]) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+48];" : "=d"(o[6]), "=d"(o[7]) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+64];" : "=d"(o[8]), "=d"(pad) : "r"(addr));
    }
};
template <> struct Ld<float> {
    static __device__ __forceinline__ void g4(const float *p, float *o) {
        const float4 a = __ldg((const float4 *)p);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    }
    static __device__ __forceinline__ void s4(unsigned addr, float *o) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(addr));
    }
    static __device__ __forceinline__ void g9(const float *p, float *o) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float4 a = __ldg((const float4 *)p + i);
            o[4 * i] = a.x;
            if (i < 2) { o[4 * i + 1] = a.y; o[4 * i + 2] = a.z; o[4 * i + 3] = a.w; }
        }
    }
    static __device__ __forceinline__ void s9(unsigned addr, float *o) {
        float p0, p1, p2;
        (void)&p0; (void)&p1; (void)&p2;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(addr));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "r"(addr));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+32];" : "=f"(o[8]), "=f"(p0), "=f"(p1), "=f"(p2) : "r"(addr));
    }
};

struct Counters { // per-thread work counters in registers (BVH and unit kernels)
    unsigned prim, sec, shadow, sph, pln, tri, box, light;
};
// Same fields backed by per-thread shared-memory slots (flat render kernel): an update costs
// LDS + IADD + STS off the critical path instead of 8 live registers in a register-bound kernel.
struct CountersRef {
    unsigned &prim, &sec, &shadow, &sph, &pln, &tri, &box, &light;
};

// ---- SPEC §3 intersections ----
// Sphere, part 1 (branch-free, so two spheres can be interleaved): b and the discriminant.
template <typename R>
__device__ __forceinline__ void sphere_eval(const R q[4], const V3<R> &o, const V3<R> &d, R &b, R &disc) {
    V3<R> oc = { o.x - q[0], o.y - q[1], o.z - q[2] };
    b = dot(oc, d);
    if constexpr (sizeof(R) == 8) {
        R cc = dot(oc, oc) - q[3];
        disc = b * b - cc;
    } else {
        // fast mode (SPEC §7): b*b - (|oc|^2 - r^2) cancels catastrophically in binary32 for distant
        // spheres (hit points land ~1e-4 off the surface and refracted rays re-hit it), so take the
        // discriminant from the component of oc perpendicular to the (unit) direction instead
        V3<R> l = { oc.x - b * d.x, oc.y - b * d.y, oc.z - b * d.z };
        disc = q[3] - dot(l, l);
    }
}
// Sphere, part 2: roots (only reached when disc >= 0).  Deliberately NOT inlined: the binary64 sqrt
// expands to ~30 instructions and this is reached from six places; one shared copy keeps the kernel
// inside the instruction cache (the profile showed no_instruction stalls at 60 KB of code).
template <typename R>
__device__ __noinline__ bool sphere_finish(R b, R disc, R eps, R &t_out) {
    R sq = Math<R>::sqrt_(disc);
    R t = -b - sq;
    if (!(t > eps)) t = -b + sq;
    if (!(t > eps)) return false;
    t_out = t;
    return true;
}
// The plane quotient, shared for the same reason (binary64 division is ~30 instructions).
template <typename R>
__device__ __noinline__ R plane_quotient(R num, R dn) { return Math<R>::div(num, dn); }
template <typename R>
__device__ __forceinline__ bool hit_sphere(const R q[4], const V3<R> &o, const V3<R> &d, R eps, R &t_out) {
    R b, disc;
    sphere_eval<R>(q, o, d, b, disc);
    if (disc < R(0)) return false;
    return sphere_finish<R>(b, disc, eps, t_out);
}

// Plane, split in three so that two planes can be evaluated together (independent chains interleave):
//   plane_eval   dn = dot(n, d), num = d_plane - dot(n, o)        (branch-free apart from the uniform switch)
//   plane_reject strict mode: true when the quotient provably cannot lie in (0, tmax)  — see hit_plane
//   plane_finish the quotient and the t > eps test
template <typename R>
__device__ __forceinline__ void plane_eval(const R q[4], int code, const V3<R> &o, const V3<R> &d, R &dn, R &num) {
    // code (warp-uniform: every lane tests the same plane): 0..2 = the normal is +-e_k, 3 = general.
    // For an axis-aligned unit normal the two dot products of SPEC §3 reduce EXACTLY to one product:
    // (0*a + n_k*b) + 0*c == n_k*b in IEEE arithmetic for finite inputs and n_k = +-1, so 8 of the 10
    // multiply/adds are skipped with bit-identical dn and num.  Uniform switch: no divergence.
    R dno;
    switch (sizeof(R) == 8 ? code : 3) { // strict mode only: in binary32 the products are cheaper than the switch
    case 0: dn = q[0] * d.x; dno = q[0] * o.x; break;
    case 1: dn = q[1] * d.y; dno = q[1] * o.y; break;
    case 2: dn = q[2] * d.z; dno = q[2] * o.z; break;
    default: {
        V3<R> n = { q[0], q[1], q[2] };
        dn = dot(n, d);
        dno = dot(n, o);
    }
    }
    num = q[3] - dno;
}
// `tm1` = tmax * (1 + 2e-15), computed once per query (and again when the bound shrinks) so that a plane
// costs one product here.  If |num| >= fl(|dn| * tm1) then |num|/|dn| >= tmax*(1+2e-15)(1-2^-53) > tmax*(1+1.8e-15),
// hence the correctly rounded quotient is >= tmax: the plane cannot be nearer than the bound.
template <typename R>
__device__ __forceinline__ R plane_bound(R tmax) {
    if constexpr (sizeof(R) == 8) return tmax * (1.0 + 2e-15); else return tmax;
}
template <typename R>
__device__ __forceinline__ bool plane_reject(R dn, R num, R tm1) {
    if constexpr (sizeof(R) == 8)
        return ((__double2hiint(num) ^ __double2hiint(dn)) < 0) | (fabs(num) >= fabs(dn) * tm1);
    else
        return false;
}
template <typename R>
__device__ __forceinline__ bool plane_finish(R dn, R num, R eps, R &t_out) {
    if (dn == R(0)) return false;
    R t;
    if constexpr (sizeof(R) == 8) t = plane_quotient<R>(num, dn); else t = Math<R>::div(num, dn);
    if (!(t > eps)) return false;
    t_out = t;
    return true;
}
// Returns true with t when the plane is hit (t > eps) AND t could be < tmax; when it returns false the
// plane is either missed or certainly not nearer than tmax.  The strict mode avoids the binary64
// division (~15 FP64-pipe instructions) unless the quotient can matter:
//   * num and dn of different sign            -> t < 0, a miss
//   * |num| >= |dn| * (tmax * (1 + 2e-15))     -> the correctly rounded quotient is >= tmax (see plane_bound)
// Both tests are implied by the exact rule, so the result is bit-identical to dividing always.
template <typename R>
__device__ __forceinline__ bool hit_plane(const R q[4], int code, const V3<R> &o, const V3<R> &d, R eps, R tm1,
                                          R &t_out) {
    R dn, num;
    plane_eval<R>(q, code, o, d, dn, num);
    if (plane_reject<R>(dn, num, tm1)) return false;
    return plane_finish<R>(dn, num, eps, t_out);
}

template <typename R>
__device__ __forceinline__ bool hit_triangle(const R q[9], const V3<R> &o, const V3<R> &d, R eps, R &t_out) {
    V3<R> v0 = { q[0], q[1], q[2] }, e1 = { q[3], q[4], q[5] }, e2 = { q[6], q[7], q[8] };
    V3<R> p = cross(d, e2);
    R det = dot(e1, p);
    if (det == R(0)) return false;
    R inv = Math<R>::rcp(det);
    V3<R> tv = sub(o, v0);
    R u = dot(tv, p) * inv;
    if (u < R(0) || u > R(1)) return false;
    V3<R> qv = cross(tv, e1);
    R v = dot(d, qv) * inv;
    if (v < R(0) || u + v > R(1)) return false;
    R t = dot(e2, qv) * inv;
    if (!(t > eps)) return false;
    t_out = t;
    return true;
}

// ---- binary32 pre-filter for the strict sphere test (flat scenes) ----
// The strict mode must return exactly what the binary64 rule of SPEC §3 returns, but it does not have
// to evaluate that rule for a sphere that provably cannot matter.  The filter works on a binary32 copy
// of the ray and of the sphere (centre rounded to nearest, radius rounded up) and answers one question
// conservatively: could this sphere have a hit with t in (0, tmax]?  It reports "no" only if
//   * the centre is farther from the ray's line than r + m            (clear miss), or
//   * the whole chord lies behind the origin, t_far + m < 0            (both roots negative), or
//   * the chord starts beyond the bound, t_near - m > tmax.
// m = 4e-5 * (|origin|_inf + scene extent) is ~20x the worst-case binary32 error of the quantities
// involved (conversions 2^-24 relative, then <= ~12 roundings on values bounded by 2*(|o|+extent)); the
// comparisons are written so that NaN/overflow answers "maybe".  Every "maybe" gets the exact test, in
// the original order, so results are bit-identical (tests/test_parity_gpu.py compares against the
// oracle, which has no filter).  Counters still count every sphere as tested.
struct SphFilterRay {
    float ox, oy, oz, dx, dy, dz, m;
};
__device__ __forceinline__ bool sphere_maybe(const SphFilterRay &r, float4 s /*cx cy cz r_up*/, float tmaxf) {
    // branch-free on purpose: NT_FILTER_BATCH of these are evaluated back to back so that their
    // (4-cycle-latency) dependency chains interleave; the kernel is latency-, not throughput-bound
    const float x = r.ox - s.x, y = r.oy - s.y, z = r.oz - s.z;
    const float b = __fmaf_rn(z, r.dz, __fmaf_rn(y, r.dy, x * r.dx));
    const float lx = __fmaf_rn(-b, r.dx, x), ly = __fmaf_rn(-b, r.dy, y), lz = __fmaf_rn(-b, r.dz, z);
    const float l2 = __fmaf_rn(lz, lz, __fmaf_rn(ly, ly, lx * lx));
    const float rm = s.w + r.m, rm2 = rm * rm * 1.000001f;
    const float h = sqrtf(fmaxf(rm2 - l2, 0.0f)) * 1.000001f;
    const bool no = (l2 > rm2) | (-b + h + r.m < 0.0f) | (-b - h - r.m > tmaxf);
    return !no;
}

template <typename R>
__device__ __forceinline__ SphFilterRay make_filter_ray(const V3<R> &o, const V3<R> &d, float scene_max_abs) {
    SphFilterRay r;
    r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
    r.dx = (float)d.x; r.dy = (float)d.y; r.dz = (float)d.z;
    r.m = 4e-5f * (fmaxf(fmaxf(fabsf(r.ox), fabsf(r.oy)), fabsf(r.oz)) + scene_max_abs);
    return r;
}

// ---- scene context of one block ----
// Dynamic shared memory: the flat intersection data staged by stage_scene().  Addressed through this
// symbol (plus element offsets kept in Ctx) so that every access is a plain LDS with an immediate
// base; pointers kept in a struct made the compiler rebuild the shared window address per iteration.
extern __shared__ __align__(16) unsigned char nt_smem[];

template <typename R, bool BVH> struct Ctx {
    const NtDevScene *s;
    const NtSceneView<R> *v;
    unsigned sph_addr, pln_addr, tri_addr, code_addr, fsph_addr; // 32-bit shared-memory byte addresses of the staged arrays
    R eps;
    unsigned max_depth;
    __device__ __forceinline__ void ld_sph(unsigned i, R *q) const {
        if constexpr (BVH) Ld<R>::g4(v->sph + 4 * (size_t)i, q); else Ld<R>::s4(sph_addr + i * (4 * (unsigned)sizeof(R)), q);
    }
    __device__ __forceinline__ void ld_pln(unsigned i, R *q) const { Ld<R>::s4(pln_addr + i * (4 * (unsigned)sizeof(R)), q); }
    __device__ __forceinline__ float4 ld_fsph(unsigned i) const {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(fsph_addr + 16 * i));
        return v;
    }
    // plane classes, 2 bits per plane, 16 planes per word
    __device__ __forceinline__ unsigned pln_codes(unsigned word) const {
        unsigned v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(code_addr + 4 * word));
        return v;
    }
    __device__ __forceinline__ void ld_tri(unsigned i, R *q) const {
        if constexpr (BVH) Ld<R>::g9(v->tri + NT_TRI_STRIDE * (size_t)i, q);
        else Ld<R>::s9(tri_addr + i * (NT_TRI_STRIDE * (unsigned)sizeof(R)), q);
    }
};

struct Hit {
    int kind; // 0 sphere, 1 plane, 2 triangle, -1 none
    int idx;  // index into the device array of that kind
    int gid;  // global primitive id (tie-break)
};

__device__ __forceinline__ void slab(float lo, float hi, float o, float inv, float m, float &tn,
                                     float &tf) {
    float a = (lo - m - o) * inv, b = (hi + m - o) * inv;
    tn = fmaxf(tn, fminf(a, b));
    tf = fminf(tf, fmaxf(a, b));
}

// SPEC §3 nearest hit: smallest t; equal t -> smallest global primitive id.
template <typename R, bool BVH, typename K>
__device__ __forceinline__ bool nearest_hit(const Ctx<R, BVH> &c, const V3<R> &o, const V3<R> &d,
                                            R &tb, Hit &best, K &k) {
    const NtDevScene &s = *c.s;
    tb = Math<R>::inf();
    best.kind = -1; best.idx = -1; best.gid = 0x7fffffff;
    R t;
    if constexpr (!BVH && sizeof(R) == 8 && NT_SPHERE_FILTER) {
        const SphFilterRay fr = make_filter_ray(o, d, s.max_abs);
        float tbf = CUDART_INF_F;
        for (unsigned i0 = 0; i0 < NT_EXP_NS(s.ns); i0 += NT_FILTER_BATCH) {
            unsigned mask = 0;
#pragma unroll
            for (int u = 0; u < NT_FILTER_BATCH; ++u)
                if (i0 + u < s.ns && sphere_maybe(fr, c.ld_fsph(i0 + u), tbf)) mask |= 1u << u;
            while (mask) { // exact tests of the survivors, in index order
                const unsigned i = i0 + (unsigned)__ffs((int)mask) - 1;
                mask &= mask - 1;
                R q[4];
                c.ld_sph(i, q);
                if (hit_sphere<R>(q, o, d, c.eps, t) && t < tb) { tb = t; best.kind = 0; best.idx = (int)i; tbf = Math<R>::up(tb); }
            }
        }
        k.sph += s.ns;
    } else if constexpr (!BVH) {
        // two spheres per iteration: the two discriminant chains are independent and interleave
        unsigned i = 0;
        for (; i + 2 <= NT_EXP_NS(s.ns); i += 2) {
            R q0[4], q1[4], b0, b1, d0, d1;
            c.ld_sph(i, q0);
            c.ld_sph(i + 1, q1);
            sphere_eval<R>(q0, o, d, b0, d0);
            sphere_eval<R>(q1, o, d, b1, d1);
            if (!(d0 < R(0)) && sphere_finish<R>(b0, d0, c.eps, t) && t < tb) { tb = t; best.kind = 0; best.idx = (int)i; }
            if (!(d1 < R(0)) && sphere_finish<R>(b1, d1, c.eps, t) && t < tb) { tb = t; best.kind = 0; best.idx = (int)i + 1; }
        }
        if (i < NT_EXP_NS(s.ns)) {
            R q[4];
            c.ld_sph(i, q);
            if (hit_sphere<R>(q, o, d, c.eps, t) && t < tb) { tb = t; best.kind = 0; best.idx = (int)i; }
        }
        k.sph += s.ns;
    }
    {
        // two planes per iteration (their chains interleave); a plane rejected against the bound that was
        // current when the pair started is rejected a fortiori against a smaller one
        const unsigned np = NT_EXP_NP(s.np);
        unsigned codes = 0, i = 0;
        R tm1 = plane_bound<R>(tb);
        // (pairs pay in binary32: 0.727 -> 0.700 ms; in binary64 the extra live values spill: 1.295 -> 1.372 ms)
        for (; sizeof(R) == 4 && i + 2 <= np; i += 2) {
            R q0[4], q1[4], dn0, num0, dn1, num1;
            if ((i & 15) == 0) codes = c.pln_codes(i >> 4);
            c.ld_pln(i, q0);
            c.ld_pln(i + 1, q1);
            plane_eval<R>(q0, (int)(codes & 3u), o, d, dn0, num0);
            plane_eval<R>(q1, (int)((codes >> 2) & 3u), o, d, dn1, num1);
            codes >>= 4;
            const bool r0 = plane_reject<R>(dn0, num0, tm1), r1 = plane_reject<R>(dn1, num1, tm1);
            if (!r0 && plane_finish<R>(dn0, num0, c.eps, t) && t < tb) { tb = t; best.kind = 1; best.idx = (int)i; best.gid = (int)(s.ns + i); }
            if (!r1 && plane_finish<R>(dn1, num1, c.eps, t) && t < tb) { tb = t; best.kind = 1; best.idx = (int)i + 1; best.gid = (int)(s.ns + i + 1); }
            tm1 = plane_bound<R>(tb);
        }
        for (; i < np; ++i) {
            R q[4];
            if ((i & 15) == 0) codes = c.pln_codes(i >> 4);
            c.ld_pln(i, q);
            const int code = (int)(codes & 3u);
            codes >>= 2;
            if (hit_plane<R>(q, code, o, d, c.eps, tm1, t) && t < tb) { tb = t; best.kind = 1; best.idx = (int)i; best.gid = (int)(s.ns + i); tm1 = plane_bound<R>(tb); }
        }
    }
    k.pln += s.np;
    if constexpr (!BVH) {
        for (unsigned i = 0; i < s.nt; ++i) {
            R q[9];
            c.ld_tri(i, q);
            if (hit_triangle<R>(q, o, d, c.eps, t) && t < tb) { tb = t; best.kind = 2; best.idx = (int)i; }
        }
        k.tri += s.nt;
    }
    return best.kind >= 0;
}

// SPEC §3 occlusion: any primitive hit (t > eps) with t < dist; first found ends the query.
// Counters follow the sequential rule (tests up to and including the first occluder).
template <typename R, bool BVH, typename K>
__device__ __forceinline__ bool occluded(const Ctx<R, BVH> &c, const V3<R> &o, const V3<R> &d, R dist,
                                         K &k) {
    const NtDevScene &s = *c.s;
    R t;
    if constexpr (!BVH && sizeof(R) == 8 && NT_SPHERE_FILTER) {
        const SphFilterRay fr = make_filter_ray(o, d, s.max_abs);
        const float distf = Math<R>::up(dist);
        for (unsigned i0 = 0; i0 < NT_EXP_NS(s.ns); i0 += NT_FILTER_BATCH) {
            unsigned mask = 0;
#pragma unroll
            for (int u = 0; u < NT_FILTER_BATCH; ++u)
                if (i0 + u < s.ns && sphere_maybe(fr, c.ld_fsph(i0 + u), distf)) mask |= 1u << u;
            while (mask) {
                const unsigned i = i0 + (unsigned)__ffs((int)mask) - 1;
                mask &= mask - 1;
                R q[4];
                c.ld_sph(i, q);
                if (hit_sphere<R>(q, o, d, c.eps, t) && t < dist) { k.sph += i + 1; return true; }
            }
        }
        k.sph += s.ns;
    } else if constexpr (!BVH) {
        unsigned i = 0;
        for (; i + 2 <= NT_EXP_NS(s.ns); i += 2) {
            R q0[4], q1[4], b0, b1, d0, d1;
            c.ld_sph(i, q0);
            c.ld_sph(i + 1, q1);
            sphere_eval<R>(q0, o, d, b0, d0);
            sphere_eval<R>(q1, o, d, b1, d1);
            if (!(d0 < R(0)) && sphere_finish<R>(b0, d0, c.eps, t) && t < dist) { k.sph += i + 1; return true; }
            if (!(d1 < R(0)) && sphere_finish<R>(b1, d1, c.eps, t) && t < dist) { k.sph += i + 2; return true; }
        }
        if (i < NT_EXP_NS(s.ns)) {
            R q[4];
            c.ld_sph(i, q);
            if (hit_sphere<R>(q, o, d, c.eps, t) && t < dist) { k.sph += s.ns; return true; }
        }
        k.sph += s.ns;
    }
    {
        const unsigned np = NT_EXP_NP(s.np);
        unsigned codes = 0, i = 0;
        const R dm1 = plane_bound<R>(dist);
        for (; sizeof(R) == 4 && i + 2 <= np; i += 2) {
            R q0[4], q1[4], dn0, num0, dn1, num1;
            if ((i & 15) == 0) codes = c.pln_codes(i >> 4);
            c.ld_pln(i, q0);
            c.ld_pln(i + 1, q1);
            plane_eval<R>(q0, (int)(codes & 3u), o, d, dn0, num0);
            plane_eval<R>(q1, (int)((codes >> 2) & 3u), o, d, dn1, num1);
            codes >>= 4;
            const bool r0 = plane_reject<R>(dn0, num0, dm1), r1 = plane_reject<R>(dn1, num1, dm1);
            if (!r0 && plane_finish<R>(dn0, num0, c.eps, t) && t < dist) { k.pln += i + 1; return true; }
            if (!r1 && plane_finish<R>(dn1, num1, c.eps, t) && t < dist) { k.pln += i + 2; return true; }
        }
        for (; i < np; ++i) {
            R q[4];
            if ((i & 15) == 0) codes = c.pln_codes(i >> 4);
            c.ld_pln(i, q);
            const int code = (int)(codes & 3u);
            codes >>= 2;
            if (hit_plane<R>(q, code, o, d, c.eps, dm1, t) && t < dist) { k.pln += i + 1; return true; }
        }
    }
    k.pln += s.np;
    if constexpr (!BVH) {
        for (unsigned i = 0; i < s.nt; ++i) {
            R q[9];
            c.ld_tri(i, q);
            if (hit_triangle<R>(q, o, d, c.eps, t) && t < dist) { k.tri += i + 1; return true; }
        }
        k.tri += s.nt;
    }
    return false;
}

// SPEC §4: radiance of one sample = sum over its ray tree in depth-first pre-order of W * local.
// `accp` / `Wp`: the sample's running radiance sum (3 values, stride NT_BLOCK_THREADS) and path weight,
// kept in per-thread shared-memory slots: touched once per tree node, not worth 8 registers.
#define NT_ACC(ch) accp[(ch) * NT_BLOCK_THREADS]
template <typename R, bool BVH, typename K>
__device__ __forceinline__ void trace_sample(const Ctx<R, BVH> &c, V3<R> o, V3<R> d, R *accp, R *Wp, K &k) {
    const NtDevScene &s = *c.s;
    const NtSceneView<R> &v = *c.v;
    // deferred transmission children (reflection children are followed immediately)
    R st[NT_MAX_DEPTH_DEV][7];
    unsigned st_depth[NT_MAX_DEPTH_DEV];
    int sp = 0;
    *Wp = R(1);
    unsigned depth = 1;
    for (;;) {
        R t;
        Hit h;
        bool descend = false;
        if (!nearest_hit<R, BVH, K>(c, o, d, t, h, k)) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) NT_ACC(ch) = NT_ACC(ch) + *Wp * __ldg(v.globals + 3 + ch);
        } else {
            const V3<R> P = { o.x + d.x * t, o.y + d.y * t, o.z + d.z * t };
            V3<R> Ng;
            int mat;
            if (h.kind == 0) {
                R q[4];
                c.ld_sph(h.idx, q);
                const R ir = __ldg(v.sph_invr + h.idx);
                Ng = { (P.x - q[0]) * ir, (P.y - q[1]) * ir, (P.z - q[2]) * ir };
                mat = __ldg(s.sph_mat + h.idx);
            } else if (h.kind == 1) {
                R q[4];
                c.ld_pln(h.idx, q);
                Ng = { q[0], q[1], q[2] };
                mat = __ldg(s.pln_mat + h.idx);
            } else {
                const R *tp = v.tri + (size_t)h.idx * NT_TRI_STRIDE + 9;
                Ng = { __ldg(tp), __ldg(tp + 1), __ldg(tp + 2) };
                mat = __ldg(s.tri_mat + h.idx);
            }
            // material rows are re-read where they are used (128-bit __ldg, L1 hits) instead of being kept
            // live across the occlusion queries: the kernel is register-bound
            const R *mp = v.mat + (size_t)mat * NT_MAT_STRIDE;
            const R cosd = dot(d, Ng);
            const bool entering = cosd < R(0);
            V3<R> N = Ng;
            if (!entering) { N.x = -Ng.x; N.y = -Ng.y; N.z = -Ng.z; }
            R local[3];
            {
                R m0[4];
                Ld<R>::g4(mp, m0); // r g b ka
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) local[ch] = __ldg(v.globals + ch) * (m0[3] * m0[ch]);
            }
            for (unsigned l = 0; l < s.nl; ++l) {
                const R *lp = v.lights + 6 * l;
                const V3<R> Lv = { __ldg(lp) - P.x, __ldg(lp + 1) - P.y, __ldg(lp + 2) - P.z };
                const R d2 = dot(Lv, Lv);
                const R dist = Math<R>::sqrt_(d2);
                const V3<R> L = scale(Lv, Math<R>::rcp(dist));
                const R ndl = dot(N, L);
                if (!(ndl > R(0))) continue;
                k.shadow++;
                if (occluded<R, BVH, K>(c, P, L, dist, k)) continue;
                k.light++;
                R m0[4], m1[4];
                Ld<R>::g4(mp, m0);     // r g b ka
                Ld<R>::g4(mp + 4, m1); // kd ks shininess kr
                const R kdn = m1[0] * ndl;
                const R lc[3] = { __ldg(lp + 3), __ldg(lp + 4), __ldg(lp + 5) };
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) local[ch] = local[ch] + lc[ch] * (m0[ch] * kdn);
                const R two = R(2) * ndl;
                const V3<R> Rv = { N.x * two - L.x, N.y * two - L.y, N.z * two - L.z };
                const R rv = -dot(Rv, d);
                if (m1[1] > R(0) && rv > R(0)) {
                    const R sterm = m1[1] * Math<R>::pow_(rv, m1[2]);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) local[ch] = local[ch] + lc[ch] * sterm;
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) NT_ACC(ch) = NT_ACC(ch) + *Wp * local[ch];

            if (depth < c.max_depth) {
                R m2[4];
                Ld<R>::g4(mp + 8, m2); // kt ior inv_ior pad
                const R kr = __ldg(mp + 7), kt = m2[0];
                const R cosi = -dot(d, N);
                R wr = kr, wt = R(0);
                V3<R> T = { R(0), R(0), R(0) };
                if (kt > R(0)) {
                    const R eta = entering ? m2[2] : m2[1];
                    const R kk = R(1) - (eta * eta) * (R(1) - cosi * cosi);
                    if (kk < R(0)) wr = kr + kt;
                    else {
                        wt = kt;
                        const R sterm = eta * cosi - Math<R>::sqrt_(kk);
                        T = { d.x * eta + N.x * sterm, d.y * eta + N.y * sterm, d.z * eta + N.z * sterm };
                    }
                }
                if (wt > R(0)) {
                    k.sec++;
                    if (wr > R(0)) { // defer: reflection subtree comes first in pre-order
                        st[sp][0] = P.x; st[sp][1] = P.y; st[sp][2] = P.z;
                        st[sp][3] = T.x; st[sp][4] = T.y; st[sp][5] = T.z;
                        st[sp][6] = *Wp * wt; st_depth[sp] = depth + 1;
                        ++sp;
                    }
                }
                if (wr > R(0)) {
                    k.sec++;
                    const R two = R(2) * cosi;
                    const V3<R> Rd = { d.x + N.x * two, d.y + N.y * two, d.z + N.z * two };
                    o = P; d = Rd; *Wp = *Wp * wr; depth = depth + 1;
                    descend = true;
                } else if (wt > R(0)) {
                    o = P; d = T; *Wp = *Wp * wt; depth = depth + 1;
                    descend = true;
                }
            }
        }
        if (descend) continue;
        if (sp == 0) break;
        --sp;
        o = { st[sp][0], st[sp][1], st[sp][2] };
        d = { st[sp][3], st[sp][4], st[sp][5] };
        *Wp = st[sp][6];
        depth = st_depth[sp];
    }
}

// ---- block-level plumbing ----

// Stage the flat intersection data in shared memory with 128-bit loads.
template <typename R, bool BVH>
__device__ __forceinline__ void stage_scene(const NtDevScene &s, const NtSceneView<R> &v, Ctx<R, BVH> &c) {
    const unsigned n_sph = BVH ? 0u : s.ns * 4, n_pln = s.np * 4, n_tri = BVH ? 0u : s.nt * NT_TRI_STRIDE;
    R *smem = (R *)nt_smem;
    constexpr int VEC = 16 / sizeof(R);
    typedef typename std::conditional<sizeof(R) == 8, double2, float4>::type VT;
    if constexpr (!BVH)
#pragma unroll 1
        for (unsigned i = threadIdx.x; i < n_sph / VEC; i += blockDim.x) ((VT *)smem)[i] = __ldg((const VT *)v.sph + i);
#pragma unroll 1
    for (unsigned i = threadIdx.x; i < n_pln / VEC; i += blockDim.x) ((VT *)(smem + n_sph))[i] = __ldg((const VT *)v.pln + i);
    if constexpr (!BVH)
#pragma unroll 1
        for (unsigned i = threadIdx.x; i < n_tri / VEC; i += blockDim.x) ((VT *)(smem + n_sph + n_pln))[i] = __ldg((const VT *)v.tri + i);
    unsigned *codes = (unsigned *)(smem + n_sph + n_pln + n_tri);
    for (unsigned i = threadIdx.x; i < (s.np + 15) / 16; i += blockDim.x) codes[i] = __ldg(s.pln_code + i);
    const unsigned code_words = (s.np + 15) / 16, fs_off = (code_words + 3) & ~3u; // float4-aligned
    if constexpr (!BVH && sizeof(R) == 8) {
        float4 *fs = (float4 *)(codes + fs_off);
        for (unsigned i = threadIdx.x; i < s.ns; i += blockDim.x) {
            const double2 a = __ldg((const double2 *)v.sph + 2 * i), b2 = __ldg((const double2 *)v.sph + 2 * i + 1);
            fs[i] = make_float4((float)a.x, (float)a.y, (float)b2.x, __double2float_ru(sqrt(b2.y)));
        }
    }
    __syncthreads();
    unsigned base = (unsigned)__cvta_generic_to_shared(nt_smem);
    asm volatile("" : "+r"(base)); // opaque: otherwise ptxas re-derives the window base (S2R + 5 ops) per use
    c.sph_addr = base;
    c.pln_addr = base + n_sph * (unsigned)sizeof(R);
    c.tri_addr = base + (n_sph + n_pln) * (unsigned)sizeof(R);
    c.code_addr = base + (n_sph + n_pln + n_tri) * (unsigned)sizeof(R);
    c.fsph_addr = c.code_addr + 4 * fs_off;
}

// Per-thread counters -> one atomic per counter per block, spread over NT_COUNTER_SLOTS slots.
__device__ __forceinline__ void flush_counter_values(const unsigned vals[NT_NCOUNTERS], unsigned long long *counters,
                                                     unsigned long long *s_cnt) {
    if (threadIdx.x < NT_NCOUNTERS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NT_NCOUNTERS; ++i) {
        const unsigned w = __reduce_add_sync(0xffffffffu, vals[i]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_cnt[i], (unsigned long long)w);
    }
    __syncthreads();
    if (threadIdx.x < NT_NCOUNTERS && s_cnt[threadIdx.x]) {
        const unsigned slot = blockIdx.x % NT_COUNTER_SLOTS;
        atomicAdd(&counters[slot * NT_NCOUNTERS + threadIdx.x], s_cnt[threadIdx.x]);
    }
}
__device__ __forceinline__ void flush_counters(const Counters &k, unsigned long long *counters,
                                               unsigned long long *s_cnt) {
    const unsigned vals[NT_NCOUNTERS] = { k.prim, k.sec, k.shadow, k.sph, k.pln, k.tri, k.box, k.light };
    flush_counter_values(vals, counters, s_cnt);
}

// Persistent warps: the grid is sized to fill the machine once (SM count x resident blocks); every
// warp pulls warp-tiles (twx x twy pixels x `lanes` samples = 32 samples) from one atomic counter until
// the image is exhausted, so no block waits at a barrier for its slowest tile and an expensive region
// (glass, mirrors) is spread over all SMs.  The only block barriers are the scene staging at the start
// and the counter flush at the end.
//
// Register diet (the kernel is register- and latency-bound): the work counters, the
// sample's running sum and its path weight live in per-thread shared-memory slots; material rows are
// re-read where used; pixel coordinates are recomputed after the trace instead of being kept live;
// SINGLE = (spp / lanes == 1) drops the cross-round pixel sum.  With that 4 blocks/SM fit in 64
// registers with ~220 bytes of spills (configs[2] f64: 1.39 -> 1.30 ms; f32 prefers 3 blocks: 0.80 -> 0.72 ms).
// Putting the work counters into shared memory as well (NT_COUNTERS_SMEM) removes the remaining spills
// but costs more instructions than it saves (1.34 ms).
template <typename R, bool BVH, bool SINGLE>
__global__ void __launch_bounds__(NT_BLOCK_THREADS, sizeof(R) == 8 ? NT_MIN_BLOCKS_F64 : NT_MIN_BLOCKS_F32)
render_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a) {
    __shared__ unsigned long long s_cnt[NT_NCOUNTERS];
#if NT_COUNTERS_SMEM
    __shared__ unsigned s_k[NT_NCOUNTERS][NT_BLOCK_THREADS];
#endif
    __shared__ R s_state[4][NT_BLOCK_THREADS]; // acc r g b, W
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, BVH> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.max_depth = a.max_depth;
    stage_scene<R, BVH>(s, v, c);

    const unsigned tid = threadIdx.x, lane = tid & 31;
#if NT_COUNTERS_SMEM
#pragma unroll
    for (int i = 0; i < NT_NCOUNTERS; ++i) s_k[i][tid] = 0;
    CountersRef k = { s_k[0][tid], s_k[1][tid], s_k[2][tid], s_k[3][tid], s_k[4][tid], s_k[5][tid], s_k[6][tid], s_k[7][tid] };
    typedef CountersRef KT;
#else
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };
    typedef Counters KT;
#endif
    R *accp = &s_state[0][tid], *Wp = &s_state[3][tid];
    const unsigned warps_per_block = NT_BLOCK_THREADS / 32, total_warps = gridDim.x * warps_per_block;
    const unsigned n_tiles = a.tiles_x * a.tiles_y;
    unsigned long long *next_tile = a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS;

    unsigned tile = blockIdx.x * warps_per_block + (tid >> 5); // first tile: no atomic needed
    while (tile < n_tiles) {
        R sum[3] = { R(0), R(0), R(0) };
        const unsigned rounds = SINGLE ? 1u : a.spp / a.lanes;
        for (unsigned r = 0; r < rounds; ++r) {
            {
                const unsigned L = a.lanes, j = lane & (L - 1), pw = lane / L;
                const unsigned px = (tile % a.tiles_x) * a.twx + pw % a.twx;
                const unsigned vr = (tile / a.tiles_x) * a.twy + pw / a.twx;
                NT_ACC(0) = R(0); NT_ACC(1) = R(0); NT_ACC(2) = R(0);
                if (px < a.width && vr < a.vrows) {
                    const unsigned y = ((vr / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + vr % a.band_rows;
                    // SPEC §2: regular n x n grid, sample s = r*L + j
                    const unsigned sidx = r * L + j;
                    const unsigned si = sidx % a.n, sj = sidx / a.n;
                    const R rn = (R)a.n;
                    const R ox = Math<R>::div((R)si + R(0.5), rn), oy = Math<R>::div((R)sj + R(0.5), rn);
                    const R fx = (R)px + ox, fy = (R)y + oy;
                    const V3<R> D = { ((R)a.cam[3] + (R)a.cam[6] * fx) + (R)a.cam[9] * fy,
                                      ((R)a.cam[4] + (R)a.cam[7] * fx) + (R)a.cam[10] * fy,
                                      ((R)a.cam[5] + (R)a.cam[8] * fx) + (R)a.cam[11] * fy };
                    const V3<R> dir = scale(D, Math<R>::rcp(Math<R>::sqrt_(dot(D, D))));
                    const V3<R> eye = { (R)a.cam[0], (R)a.cam[1], (R)a.cam[2] };
                    k.prim++;
                    trace_sample<R, BVH, KT>(c, eye, dir, accp, Wp, k);
                }
            }
            asm volatile("" : "+r"(tile)); // pixel coordinates are recomputed below, not carried across the trace
            // SPEC §5: samples are added in sample order; the lanes of one pixel are adjacent
            const unsigned L = a.lanes;
            if (L == 1) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) sum[ch] = sum[ch] + NT_ACC(ch);
            } else {
                const unsigned base = lane & ~(L - 1);
                const R a0 = NT_ACC(0), a1 = NT_ACC(1), a2 = NT_ACC(2);
#pragma unroll 1
                for (unsigned jj = 0; jj < L; ++jj) {
                    sum[0] = sum[0] + __shfl_sync(0xffffffffu, a0, base + jj);
                    sum[1] = sum[1] + __shfl_sync(0xffffffffu, a1, base + jj);
                    sum[2] = sum[2] + __shfl_sync(0xffffffffu, a2, base + jj);
                }
            }
        }
        {
            const unsigned L = a.lanes, j = lane & (L - 1), pw = lane / L;
            const unsigned px = (tile % a.tiles_x) * a.twx + pw % a.twx;
            const unsigned vr = (tile / a.tiles_x) * a.twy + pw / a.twx;
            if (px < a.width && vr < a.vrows && j == 0) {
                const unsigned y = ((vr / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + vr % a.band_rows;
                const R inv_spp = Math<R>::rcp((R)a.spp);
                unsigned rgba = 0xff000000u;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const R cv = sum[ch] * inv_spp;
                    const unsigned q = cv <= R(0) ? 0u : cv >= R(1) ? 255u : (unsigned)(int)(cv * R(255) + R(0.5));
                    rgba |= q << (8 * ch);
                }
                const size_t row = a.layout == 1 ? vr : y;
                *(unsigned *)(a.out + row * a.stride + 4 * (size_t)px) = rgba;
            }
        }
        unsigned long long nt = 0;
        if (lane == 0) nt = atomicAdd(next_tile, 1ull) + total_warps;
        tile = (unsigned)__shfl_sync(0xffffffffu, nt, 0);
    }
#if NT_COUNTERS_SMEM
    const unsigned vals[NT_NCOUNTERS] = { s_k[0][tid], s_k[1][tid], s_k[2][tid], s_k[3][tid], s_k[4][tid], s_k[5][tid], s_k[6][tid], s_k[7][tid] };
    flush_counter_values(vals, a.counters, s_cnt);
#else
    flush_counters(k, a.counters, s_cnt);
#endif
}

// Unit-level entry: nearest hit of arbitrary rays (nt_trace_rays).
template <typename R, bool BVH>
__global__ void __launch_bounds__(NT_BLOCK_THREADS)
trace_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtTraceArgs a) {
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, BVH> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.max_depth = 1;
    stage_scene<R, BVH>(s, v, c);
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const V3<R> o = { (R)a.origins[3 * i], (R)a.origins[3 * i + 1], (R)a.origins[3 * i + 2] };
    const V3<R> d = { (R)a.dirs[3 * i], (R)a.dirs[3 * i + 1], (R)a.dirs[3 * i + 2] };
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };
    R t;
    Hit h;
    if (nearest_hit<R, BVH, Counters>(c, o, d, t, h, k)) {
        a.t_out[i] = (double)t;
        a.prim_out[i] = h.kind == 1 ? h.gid : (h.kind == 0 ? h.idx : (int)(s.ns + s.np) + h.idx);
    } else {
        a.t_out[i] = -1.0;
        a.prim_out[i] = -1;
    }
}

template <typename R>
inline size_t flat_smem_bytes(const NtDevScene &s, bool bvh) {
    size_t n = (size_t)s.np * 4;
    if (!bvh) n += (size_t)s.ns * 4 + (size_t)s.nt * NT_TRI_STRIDE;
    size_t bytes = n * sizeof(R) + (size_t)((((s.np + 15) / 16) + 3) & ~3u) * sizeof(unsigned);
    if (!bvh && sizeof(R) == 8) bytes += (size_t)s.ns * sizeof(float4); // binary32 filter spheres
    return bytes;
}

} // namespace nt

#include "nt_bvh_trace.cuh"

namespace nt {

template <typename R, bool BVH>
inline int launch_render_t(const NtDevScene &s, const NtRenderArgs &a, cudaStream_t st) {
    static int blocks_per_sm[64] = { 0 }, sms[64] = { 0 }; // per device, resolved once
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = flat_smem_bytes<R>(s, BVH);
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!sms[dev]) {
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (BVH) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], render_bvh_kernel<R>, NT_BLOCK_THREADS, 4096);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], render_kernel<R, false, true>, NT_BLOCK_THREADS, 4096);
        if (blocks_per_sm[dev] < 1) blocks_per_sm[dev] = 1;
    }
    const unsigned n_tiles = a.tiles_x * a.tiles_y, wpb = NT_BLOCK_THREADS / 32;
    unsigned grid = (unsigned)(sms[dev] * blocks_per_sm[dev]);
    if (grid > (n_tiles + wpb - 1) / wpb) grid = (n_tiles + wpb - 1) / wpb;
    if (BVH) {
        render_bvh_kernel<R><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
        resolve_kernel<R><<<dim3((a.width + 255) / 256, a.vrows), 256, 0, st>>>(a);
    } else if (a.spp == a.lanes) render_kernel<R, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
    else render_kernel<R, false, false><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
    return (int)cudaGetLastError();
}

template <typename R>
inline int launch_render(const NtDevScene &s, const NtRenderArgs &a, cudaStream_t st) {
    return s.use_bvh ? launch_render_t<R, true>(s, a, st) : launch_render_t<R, false>(s, a, st);
}

template <typename R>
inline int launch_trace(const NtDevScene &s, const NtTraceArgs &a, cudaStream_t st) {
    dim3 grid((a.n + NT_BLOCK_THREADS - 1) / NT_BLOCK_THREADS), block(NT_BLOCK_THREADS);
    const size_t smem = flat_smem_bytes<R>(s, s.use_bvh != 0);
    if (s.use_bvh) trace_bvh_kernel<R><<<grid, block, smem, st>>>(s, a);
    else trace_kernel<R, false><<<grid, block, smem, st>>>(s, a);
    return (int)cudaGetLastError();
}

} // namespace nt
