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
#include "nt_sync.cuh"

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

// libdevice pow (non-integer shininess): ~2.4 KB of code that the benchmark scenes never run - out of line, so that it
// does not sit between the hot blocks of the flat kernels (instruction cache)
static __device__ __noinline__ double pow_slow(double a, double b) { return pow(a, b); }
template <typename R> struct Math;
template <> struct Math<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
#ifdef NT_EXP_NOPOW
    static __device__ __forceinline__ double pow_(double a, double b) { return a * b; }
#else
    // SPEC §0: `pow` is the one operation that may differ from the oracle's libm in the last bits.  Integer
    // exponents (every shininess of the benchmark scenes) go through square-and-multiply: <= 20 roundings,
    // relative error < 3e-15, ~30 instructions executed instead of libdevice pow's ~150 (which also was 2.4 KB of
    // the strict kernel's hot code); anything else takes libdevice's pow.
    static __device__ __forceinline__ double pow_(double a, double b) {
        const int n = (int)b;
        if ((double)n == b && n >= 1 && n <= 1024) {
            double r = (n & 1) ? a : 1.0, x = a;
#pragma unroll 1
            for (int e = n >> 1; e; e >>= 1) {
                x = x * x;
                if (e & 1) r = r * x;
            }
            return r;
        }
        return pow_slow(a, b);
    }
#endif
    static __device__ __forceinline__ float up(double x) { return __double2float_ru(x); }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
    // SPEC §2 / §4: len = sqrt(x), inv = 1 / len - two correctly rounded operations, in this order
    static __device__ __forceinline__ void len_inv(double x, double &len, double &inv) { len = sqrt(x); inv = 1.0 / len; }
};
template <> struct Math<float> {
    static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float pow_(float a, float b) { return __powf(a, b); }
    static __device__ __forceinline__ float up(float x) { return x; }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    // fast mode: one MUFU.RSQ instead of MUFU.SQRT + MUFU.RCP
    static __device__ __forceinline__ void len_inv(float x, float &len, float &inv) { inv = rsqrtf(x); len = x * inv; }
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
    static __device__ __forceinline__ void s2(unsigned addr, double &p, int &idx) { // axis-plane entry: position, index bits
        double b;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p), "=d"(b) : "r"(addr));
        idx = __double2loint(b);
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
    static __device__ __forceinline__ void s2(unsigned addr, float &p, int &idx) {
        float b;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(p), "=f"(b) : "r"(addr));
        idx = __float_as_int(b);
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

struct Counters { // per-thread work counters in registers
    unsigned prim, sec, shadow, sph, pln, tri, box, light;
};
// Instrumented flat kernel (NT_RENDER_COUNT_EXECUTED): additionally the primitive tests the kernel really STARTS - the
// culling tables and the chord rule skip most of what the algorithmic counters above credit (those keep the brute-force
// meaning).  A separate instantiation, so the product kernel carries no extra registers.
struct CountersX : Counters {
    unsigned xsph, xpln, xtri;
};
template <typename K> struct counts_executed { static constexpr bool value = false; };
template <> struct counts_executed<CountersX> { static constexpr bool value = true; };
#define NT_X(k, field, n) do { if constexpr (counts_executed<K>::value) (k).field += (n); } while (0)
// The fast mode re-normalises reflected and refracted directions (what SPEC §8's NT_RULE_RENORMALIZE switches on in the
// strict mode).  SPEC §4 does not, and a bounce off a sphere of radius r seen from distance t multiplies the error of
// |d|^2 by ~4 (t / r)^2: harmless 1e-11 after one bounce in binary64, but 5e-4 in binary32 - and §3's sphere rule, written
// for unit directions, then sees every sphere sqrt(5e-4) t ~ a unit larger at t = 50.  configs[3] (10 000 spheres of
// radius 0.1 - 0.6 seen from 100 units): 4.7 % of the pixels - every mirror or glass sphere - more than 2 LSB from the
// strict frame without it (found by `bench.py --precision f32`'s frame check, scripts/gpu_fastdiff_cfg4.py).
template <typename R> struct fast_renormalises { static constexpr bool value = sizeof(R) == 4; };
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
// Strict mode: is the correctly rounded quotient num / dn certainly <= eps (a miss by the exact rule)?
//   eps_lo = fl(eps * k), k = 1 - 2^-50 (made on the host; 0 when eps < 1e-290), lo = fl(|dn| * eps_lo).  When lo is a
//   normal number, |num| <= lo implies |num| / |dn| <= eps * k * (1 + 2^-53)^2 < eps, and rounding is monotonic,
//   so |fl(num / dn)| <= eps.
// This is every ray that starts ON a plane (a shadow, reflection or refraction ray leaving a wall: num is 0 or
// a few ulps): 44 % of all plane quotients of configs[2], and a zero numerator takes the ~100-instruction
// special-operand path of the binary64 division (ncu: 4 % of the kernel's instructions and most of its
// instruction-cache misses, profiles/r01g_cfg3_f64_before_lowreject.md).
template <typename R>
__device__ __forceinline__ bool plane_below_eps(R dn, R num, R eps_lo) {
    if constexpr (sizeof(R) == 8) {
        const double lo = fabs(dn) * eps_lo;
        return fabs(num) <= lo && lo >= 2.2250738585072014e-308;
    } else {
        return false;
    }
}
template <typename R>
__device__ __forceinline__ bool plane_finish(R dn, R num, R eps, R eps_lo, R &t_out) {
    if (dn == R(0)) return false;
    if (plane_below_eps<R>(dn, num, eps_lo)) return false;
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
__device__ __forceinline__ bool hit_plane(const R q[4], int code, const V3<R> &o, const V3<R> &d, R eps, R eps_lo, R tm1,
                                          R &t_out) {
    R dn, num;
    plane_eval<R>(q, code, o, d, dn, num);
    if (plane_reject<R>(dn, num, tm1)) return false;
    return plane_finish<R>(dn, num, eps, eps_lo, t_out);
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

// ---- conservative culling for flat scenes (nt_cull.h; SPEC §3 "conservative culling only") ----
// Bit j of a mask = bounded primitive j (spheres 0..ns-1, then triangles).  A primitive that is absent from
// the mask a query uses provably cannot be hit by that query, so skipping it leaves every result bit-identical
// to the brute-force rule (the oracle has no culling; tests compare images, counters and t bit for bit).
// Work counters keep the brute-force meaning: they count the tests the sequential rule would have done.
__device__ __forceinline__ unsigned long long low_bits(unsigned n) { return n >= 64 ? ~0ull : (1ull << n) - 1ull; }

// Light buffer look-up for a shadow query.  Lv = light - P, so the direction light -> P is -Lv.  Cube face =
// major axis and sign; (a, b) = the two other components in increasing axis order, divided by |major|.
// Approximate binary32 arithmetic is fine: the host dilated every footprint by 2e-3 rad (nt_cull.cpp), the cell
// index is clamped, and a NaN direction reads cell 0 (such a query has no defined answer anyway).
template <typename R>
__device__ __forceinline__ unsigned long long lbuf_mask(const NtDevScene &s, unsigned l, const V3<R> &Lv) {
    const float x = -(float)Lv.x, y = -(float)Lv.y, z = -(float)Lv.z;
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    const bool fx = ax >= ay && ax >= az, fy = !fx && ay >= az;
    const float m = fx ? ax : (fy ? ay : az), comp = fx ? x : (fy ? y : z);
    const float a = fx ? y : x, b = (fx || fy) ? z : y;
    const unsigned face = (fx ? 0u : (fy ? 2u : 4u)) + (comp < 0.0f ? 1u : 0u);
    const int K = (int)s.lbuf_k;
    const float half = 0.5f * (float)K, sc = __fdividef(half, m);
    int iu = (int)(a * sc + half), iv = (int)(b * sc + half);
    iu = min(max(iu, 0), K - 1);
    iv = min(max(iv, 0), K - 1);
    const unsigned cell = ((l * 6u + face) * (unsigned)K + (unsigned)iv) * (unsigned)K + (unsigned)iu; // < 2^32: <= 32 lights (nt_cull.h)
    // A query from very far away - a hit on an unbounded plane near the horizon, millions of units out - is beyond what the
    // tables can promise: the sphere rule of SPEC section 3 cancels catastrophically there (its discriminant carries an
    // error of ~1e-15 |o - c|^2, the size of r^2 from |o - c| ~ 1e6 r on), so it "hits" spheres the ray misses by a wide
    // margin, and a culled GPU query would disagree with it.  Found by scripts/gpu_fuzz_flat.py (an exact image, one
    // work counter off by two: a sphere the rule hit from 2.2e7 units away was not in its cell).  Beyond cull_far = 1e5 x the
    // smallest radius every primitive is tested, as the rule does; m is the max norm of light - P.
    return m > s.cull_far ? s.all_bits : __ldg(s.lbuf + cell);
}

// ---- warp-tile coordinates without integer division (they are recomputed around every trace instead of
// being kept live, and a 32-bit division is ~20 instructions) ----
// tile -> (first pixel column, first owned row).  tile / tiles_x from a binary32 estimate + one fix-up step:
// the quotient is a tile row < 2^16, so the estimate is off by at most one.
// The tile sequence starts at the first tile row that shows a bounded primitive (a branching one - glass - when the
// scene has any) and wraps around: the rows above it, which show walls, sky or background only, are the cheapest
// and come last, so the launch does not end on a row of expensive tiles.  (configs[2]: 0.817 -> 0.77 ms; a full
// heavy-tiles-first order through a classification pass was no faster than the plain order.)
__device__ __forceinline__ void tile_origin(const NtRenderArgs &a, unsigned tile, unsigned &px0, unsigned &vr0) {
    tile += a.tile_rot;
    if (tile >= a.n_tiles) tile -= a.n_tiles;
    unsigned ty = __float2uint_rz(__uint2float_rz(tile) * a.inv_tiles_x);
    int r = (int)(tile - ty * a.tiles_x);
    if (r < 0) { --ty; r += (int)a.tiles_x; }
    else if (r >= (int)a.tiles_x) { ++ty; r -= (int)a.tiles_x; }
    px0 = (unsigned)r << a.log2_twx;
    vr0 = ty << a.log2_twy;
}
// owned (virtual) row -> image row.  vr < 2^16, so __umulhi(vr, ceil(2^32 / band_rows)) == vr / band_rows exactly.
__device__ __forceinline__ unsigned row_to_y(const NtRenderArgs &a, unsigned vr) {
    if (a.shard_count == 1) return vr;
    const unsigned b = a.band_rows == 1 ? vr : __umulhi(vr, a.band_magic);
    return (b * a.shard_count + a.shard_index) * a.band_rows + (vr - b * a.band_rows);
}

// Primary rays: the bounded primitives whose pixel rectangle (made on the host for this camera from the dilated
// bounding spheres, nt_cull.h nt_cull_primary_rects) overlaps the warp tile's pixel area x in [px0, px0 + twx),
// y in [ymin, ymax1).  Every lane tests primitive `lane` (and `lane + 32`); a ballot makes the warp-uniform mask.
// (The first version tested the ball against a cone around the tile in binary32 inside the kernel: 130 warp
// instructions per tile, 4.8 % of the kernel; the rectangles cost ~15.)
static __device__ __forceinline__ unsigned long long tile_mask(const NtDevScene &s, const NtRenderArgs &a, unsigned px0, unsigned ymin,
                                                               unsigned ymax1, unsigned lane) {
    const unsigned nb = s.ns + s.nt;
    const unsigned x1 = px0 + a.twx - 1u, y1 = ymax1 - 1u;
    unsigned long long mask = 0;
#pragma unroll 1
    for (unsigned base = 0; base < nb; base += 32) {
        const unsigned j = base + lane;
        bool in = false;
        if (j < nb) {
            const uint2 r = *(const uint2 *)a.prect[j]; // x0 | x1 << 16, y0 | y1 << 16
            in = (r.x & 0xffffu) <= x1 && (r.x >> 16) >= px0 && (r.y & 0xffffu) <= y1 && (r.y >> 16) >= ymin;
        }
        mask |= (unsigned long long)__ballot_sync(0xffffffffu, in) << base;
    }
    return mask;
}

// ---- scene context of one block ----
// Dynamic shared memory: the flat intersection data staged by stage_scene().  Addressed through this
// symbol (plus element offsets kept in Ctx) so that every access is a plain LDS with an immediate
// base; pointers kept in a struct made the compiler rebuild the shared window address per iteration.
extern __shared__ __align__(16) unsigned char nt_smem[];

// LEAN: the launch has neither triangles nor general (non-axis-aligned) planes - their loops are compiled out.  The flat
// kernels' hot code must stay inside the instruction cache (no_instruction stalls), and loops that
// never run still sit between the hot blocks.
template <typename R, bool BVH, bool LEAN = false> struct Ctx {
    static constexpr bool lean = LEAN;
    const NtDevScene *s;
    const NtSceneView<R> *v;
    unsigned sph_addr, pln_addr, tri_addr, code_addr; // 32-bit shared-memory byte addresses of the staged arrays
    unsigned axl_addr, gen_addr;                      // flat scenes: axis-aligned plane lists, general-plane index list
    R eps, eps_lo; // eps_lo: see plane_below_eps
    unsigned max_depth;
    unsigned rules; // NT_DEV_RULE_* bits (SPEC §8); the flat kernel reads them only in its RULES instantiation
    __device__ __forceinline__ void ld_sph(unsigned i, R *q) const {
        if constexpr (BVH) Ld<R>::g4(v->sph + 4 * (size_t)i, q); else Ld<R>::s4(sph_addr + i * (4 * (unsigned)sizeof(R)), q);
    }
    __device__ __forceinline__ void ld_pln(unsigned i, R *q) const { Ld<R>::s4(pln_addr + i * (4 * (unsigned)sizeof(R)), q); }
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

// ---- planes of a flat scene ----
// The host splits the planes into three lists of axis-aligned ones (normal exactly +-e_k) and a list of general
// ones (NtDevScene::nax / axl / pgen).  For n = s*e_k, s = +-1, SPEC §3 reduces EXACTLY (products with +-1 and 0
// are exact, negation commutes with rounding) to
//     t = fl( fl(p - o_k) / d_k ),   p = s * d_plane      (entry of the axis list: p and the plane index)
// i.e. one subtraction per plane, no per-plane switch.  The lists are not in index order, so ties are broken
// explicitly (smallest global id, SPEC §3); the conservative bound test only rejects quotients that are
// STRICTLY larger than the bound (see plane_bound), so a tie is never rejected.
// Strict mode: can the quotient num / dk lie in (0, bound)?  bk = |dk| * plane_bound(bound).  "No" when the
// signs differ (t < 0) or |num| >= bk (t > bound, strictly); NaN anywhere answers "no", and the exact rule then
// misses as well (a NaN quotient fails t > eps; dk == 0 is a miss).
template <typename R>
__device__ __forceinline__ bool axis_candidate(R dk, R num, R bk) {
    if constexpr (sizeof(R) == 8) return (__double2hiint(num) ^ __double2hiint(dk)) >= 0 && fabs(num) < bk;
    else return true;
}
template <typename R, typename K, bool LEAN>
__device__ __forceinline__ void planes_nearest(const Ctx<R, false, LEAN> &c, const V3<R> &o, const V3<R> &d, R &tb, Hit &best, K &k) {
    const NtDevScene &s = *c.s;
    NT_X(k, xpln, s.nax[0] + s.nax[1] + s.nax[2] + s.ngen);
    R tm1 = plane_bound<R>(tb);
    unsigned addr = c.axl_addr;
    const R oo[3] = { o.x, o.y, o.z }, dd[3] = { d.x, d.y, d.z };
#pragma unroll
    for (int k = 0; k < 3; ++k) { // unrolled over the axes (no selects, list lengths straight from the constant bank)
        const unsigned n = s.nax[k];
        const R ok = oo[k], dk = dd[k];
        R bk = fabs(dk) * tm1;
#pragma unroll 1
        for (unsigned j = 0; j < n; ++j, addr += 2u * (unsigned)sizeof(R)) {
            R p, t;
            int idx;
            Ld<R>::s2(addr, p, idx);
            const R num = p - ok;
            if (!axis_candidate<R>(dk, num, bk)) continue;
            if (!plane_finish<R>(dk, num, c.eps, c.eps_lo, t)) continue;
            const int gid = (int)s.ns + idx;
            if (t < tb || (t == tb && gid < best.gid)) {
                tb = t; best.kind = 1; best.idx = idx; best.gid = gid;
                tm1 = plane_bound<R>(tb); bk = fabs(dk) * tm1;
            }
        }
    }
    if constexpr (!LEAN) {
#pragma unroll 1
        for (unsigned j = 0; j < s.ngen; ++j) {
            int idx;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(idx) : "r"(c.gen_addr + 4u * j));
            R q[4], dn, num, t;
            c.ld_pln((unsigned)idx, q);
            plane_eval<R>(q, 3, o, d, dn, num);
            if (plane_reject<R>(dn, num, tm1)) continue;
            if (!plane_finish<R>(dn, num, c.eps, c.eps_lo, t)) continue;
            const int gid = (int)s.ns + idx;
            if (t < tb || (t == tb && gid < best.gid)) { tb = t; best.kind = 1; best.idx = idx; best.gid = gid; tm1 = plane_bound<R>(tb); }
        }
    }
}
// Any plane with eps < t < dist?
// `axis` false: the query's origin lies in its light's room (in_light_room), no axis-aligned plane can stop it.
template <typename R, typename K, bool LEAN>
__device__ __forceinline__ bool planes_occluded(const Ctx<R, false, LEAN> &c, const V3<R> &o, const V3<R> &d, R dist, bool axis, K &k) {
    const NtDevScene &s = *c.s;
    NT_X(k, xpln, (axis ? s.nax[0] + s.nax[1] + s.nax[2] : 0u) + s.ngen); // an upper bound when an occluder ends the loops early
    const R dm1 = plane_bound<R>(dist);
    unsigned addr = c.axl_addr;
    const R oo[3] = { o.x, o.y, o.z }, dd[3] = { d.x, d.y, d.z };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned n = axis ? s.nax[k] : 0u;
        const R ok = oo[k], dk = dd[k], bk = fabs(dk) * dm1;
#pragma unroll 1
        for (unsigned j = 0; j < n; ++j, addr += 2u * (unsigned)sizeof(R)) {
            R p, t;
            int idx;
            Ld<R>::s2(addr, p, idx);
            const R num = p - ok;
            if (!axis_candidate<R>(dk, num, bk)) continue;
            if (plane_finish<R>(dk, num, c.eps, c.eps_lo, t) && t < dist) return true;
        }
    }
    if constexpr (!LEAN) {
#pragma unroll 1
        for (unsigned j = 0; j < s.ngen; ++j) {
            int idx;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(idx) : "r"(c.gen_addr + 4u * j));
            R q[4], dn, num, t;
            c.ld_pln((unsigned)idx, q);
            plane_eval<R>(q, 3, o, d, dn, num);
            if (plane_reject<R>(dn, num, dm1)) continue;
            if (plane_finish<R>(dn, num, c.eps, c.eps_lo, t) && t < dist) return true;
        }
    }
    return false;
}
// The specialised kernels (LEAN) run for scenes whose shadow queries pass the room test nearly always: there the axis
// lists above are cold code, and inlined they sat in the middle of the hot loop (2 KB).  Arguments by value, a context
// rebuilt from them: nothing of the caller's state has to live in memory for the call.
template <typename R>
__device__ __noinline__ bool planes_occluded_cold(const NtDevScene *s, unsigned axl_addr, R eps, R eps_lo, R ox, R oy, R oz, R dx, R dy, R dz, R dist) {
    Ctx<R, false, true> c;
    c.s = s; c.v = nullptr; c.axl_addr = axl_addr; c.gen_addr = 0; c.sph_addr = c.pln_addr = c.tri_addr = c.code_addr = 0;
    c.eps = eps; c.eps_lo = eps_lo; c.max_depth = 0; c.rules = 0;
    Counters k{};
    const V3<R> o = { ox, oy, oz }, d = { dx, dy, dz };
    return planes_occluded<R, Counters, true>(c, o, d, dist, true, k);
}
// Rare path (a plane occludes the light): index of the FIRST occluding plane in index order, for the work
// counters' sequential rule.  Plain SPEC §3 formula; gives the same t as the list forms above, bit for bit.
template <typename R, bool LEAN>
__device__ __noinline__ unsigned first_occluding_plane(const Ctx<R, false, LEAN> &c, const V3<R> &o, const V3<R> &d, R dist) {
    const NtDevScene &s = *c.s;
#pragma unroll 1
    for (unsigned i = 0; i < s.np; ++i) {
        R q[4], dn, num, t;
        c.ld_pln(i, q);
        plane_eval<R>(q, 3, o, d, dn, num);
        if (plane_finish<R>(dn, num, c.eps, c.eps_lo, t) && t < dist) return i;
    }
    return s.np - 1;
}

// Fast mode, scenes with at most two axis-aligned planes per axis (NtDevScene::slab): t = (p - o_k) / d_k for both planes
// of an axis with ONE packed subtract and ONE packed multiply (sm_100 add.f32x2 / mul.f32x2), the reciprocal of d_k
// shared; staged per axis as p0 p1 i0 i1 so that a 128-bit shared load delivers the position pair in an aligned
// register pair.  Padding entries hold p = NaN (never > eps); d_k = 0 gives +-inf or NaN, never a hit.
template <bool LEAN>
__device__ __forceinline__ void planes_nearest_slab32(const Ctx<float, false, LEAN> &c, const V3<float> &o, const V3<float> &d, float &tb, Hit &best) {
    const float oo[3] = { o.x, o.y, o.z }, dd[3] = { d.x, d.y, d.z };
    int wi = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float q[4];
        Ld<float>::s4(c.sph_addr + c.s->axs_off[0] + 16u * k, q); // p0 p1 i0 i1
        const float inv = __fdividef(1.0f, dd[k]), nok = -oo[k];
        const float2 t = __fmul2_rn(__fadd2_rn(make_float2(q[0], q[1]), make_float2(nok, nok)), make_float2(inv, inv));
        const float t0 = t.x > c.eps ? t.x : CUDART_INF_F, t1 = t.y > c.eps ? t.y : CUDART_INF_F;
        if (t0 < tb) { tb = t0; wi = __float_as_int(q[2]); }
        if (t1 < tb) { tb = t1; wi = __float_as_int(q[3]); }
    }
    if (wi >= 0) { best.kind = 1; best.idx = wi; best.gid = (int)c.s->ns + wi; }
}

// SPEC §3 nearest hit: smallest t; equal t -> smallest global primitive id.
// Flat scenes: `mask` = the bounded primitives this ray can possibly hit (per lane).  `own` >= 0: the ray
// starts on sphere `own` (a reflection / refraction child): that sphere is tested first, and when it is hit
// the segment up to that hit is a chord of its ball, so only the balls touching it (nbr[own]) can stop the
// ray earlier.  Bounded primitives come in index order; the only out-of-order test is `own`, hence the
// explicit tie rule in the sphere loop.
template <typename R, bool BVH, typename K, bool LEAN>
__device__ __forceinline__ bool nearest_hit(const Ctx<R, BVH, LEAN> &c, const V3<R> &o, const V3<R> &d, unsigned long long mask,
                                            int own, R &tb, Hit &best, K &k) {
    const NtDevScene &s = *c.s;
    tb = Math<R>::inf();
    best.kind = -1; best.idx = -1; best.gid = 0x7fffffff;
    R t;
    if constexpr (!BVH) {
        if (own >= 0) {
            R q[4];
            NT_X(k, xsph, 1u);
            c.ld_sph((unsigned)own, q);
            if (hit_sphere<R>(q, o, d, c.eps, t)) {
                tb = t; best.kind = 0; best.idx = own; best.gid = own;
                // the chord argument needs a unit direction: SPEC §4 does not re-normalise, and with |d|^2 = 1 + e the
                // sphere rule accepts points up to sqrt(e) t outside a ball (nt_bvh_trace.cuh query_start).  The
                // neighbour table is dilated by 1e-3; beyond e = 1e-8 (a long mirror chain) every primitive is tested.
                if (fabs(dot(d, d) - R(1)) <= R(sizeof(R) == 8 ? 1e-8 : 1e-4)) mask = __ldg(s.nbr + own);
            } else mask &= ~(1ull << own);
        }
        unsigned long long m = mask & s.sph_bits;
        NT_X(k, xsph, (unsigned)__popcll(m));
        while (m) {
            const unsigned i = (unsigned)__ffsll((long long)m) - 1u;
            m &= m - 1;
            R q[4];
            c.ld_sph(i, q);
            if (hit_sphere<R>(q, o, d, c.eps, t) && (t < tb || (t == tb && (int)i < best.gid))) { tb = t; best.kind = 0; best.idx = (int)i; best.gid = (int)i; }
        }
    }
    if constexpr (!BVH && sizeof(R) == 8) planes_nearest<R, K>(c, o, d, tb, best, k);
    if constexpr (!BVH && sizeof(R) == 4) {
        // fast mode: the quotient is two instructions, so the plain SPEC §3 form in index order, two planes per
        // iteration and without branches around the division, beats the axis lists
        NT_X(k, xpln, s.np);
        unsigned i = 0;
        if (s.slab) { // a room: the slab form below covers every axis-aligned plane, the loop the general ones
            planes_nearest_slab32(c, o, d, tb, best);
#pragma unroll 1
            for (unsigned j = 0; !LEAN && j < s.ngen; ++j) {
                int idx;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(idx) : "r"(c.gen_addr + 4u * j));
                R q[4], dn, num;
                c.ld_pln((unsigned)idx, q);
                plane_eval<R>(q, 3, o, d, dn, num);
                if (plane_finish<R>(dn, num, c.eps, c.eps_lo, t) && t < tb) { tb = t; best.kind = 1; best.idx = idx; best.gid = (int)s.ns + idx; }
            }
            i = s.np;
        }
        for (; i + 2 <= s.np; i += 2) {
            R q0[4], q1[4], dn0, num0, dn1, num1;
            c.ld_pln(i, q0);
            c.ld_pln(i + 1, q1);
            plane_eval<R>(q0, 3, o, d, dn0, num0);
            plane_eval<R>(q1, 3, o, d, dn1, num1);
            if (plane_finish<R>(dn0, num0, c.eps, c.eps_lo, t) && t < tb) { tb = t; best.kind = 1; best.idx = (int)i; best.gid = (int)(s.ns + i); }
            if (plane_finish<R>(dn1, num1, c.eps, c.eps_lo, t) && t < tb) { tb = t; best.kind = 1; best.idx = (int)i + 1; best.gid = (int)(s.ns + i + 1); }
        }
        if (i < s.np) {
            R q[4], dn, num;
            c.ld_pln(i, q);
            plane_eval<R>(q, 3, o, d, dn, num);
            if (plane_finish<R>(dn, num, c.eps, c.eps_lo, t) && t < tb) { tb = t; best.kind = 1; best.idx = (int)i; best.gid = (int)(s.ns + i); }
        }
    }
    if constexpr (BVH) k.pln += s.np;
    if constexpr (!BVH && !LEAN) {
        if (s.nt) { // uniform
            unsigned long long m = s.ns >= 64 ? 0ull : mask >> s.ns;
            NT_X(k, xtri, (unsigned)__popcll(m));
            while (m) {
                const unsigned i = (unsigned)__ffsll((long long)m) - 1u;
                m &= m - 1;
                R q[9];
                c.ld_tri(i, q);
                if (hit_triangle<R>(q, o, d, c.eps, t) && t < tb) { tb = t; best.kind = 2; best.idx = (int)i; best.gid = (int)(s.ns + s.np + i); }
            }
        }
    }
    return best.kind >= 0;
}

// Is the origin P of a shadow query towards light l inside the light's room (nt_cull.h nt_cull_light_rooms: no axis-aligned
// plane can stop the query then)?  Six comparisons against the staged box and one against the distance cap.
template <typename R, bool LEAN>
__device__ __forceinline__ bool in_light_room(const Ctx<R, false, LEAN> &c, unsigned l, const V3<R> &P, R dist) {
    R a[4], b[4];
    const unsigned addr = c.sph_addr + c.s->room_off[sizeof(R) == 8] + l * (8u * (unsigned)sizeof(R));
    Ld<R>::s4(addr, a);
    Ld<R>::s4(addr + 4u * (unsigned)sizeof(R), b);
    return (P.x >= a[0]) & (P.x <= a[1]) & (P.y >= a[2]) & (P.y <= a[3]) & (P.z >= b[0]) & (P.z <= b[1]) & (dist <= b[2]);
}

// SPEC §3 occlusion: any primitive hit (t > eps) with t < dist; first found ends the query.
// Counters follow the sequential rule (tests up to and including the first occluder); a culled primitive is
// a certain miss, so the first occluder found in mask order is the first one in index order.
// Flat scenes count the tests NOT made: k.sph / k.pln / k.tri are deficits against "every query tests every
// primitive" and are touched only when a query ends early (the kernel's flush turns them into test counts:
// queries * n - deficit); an update per query was a spilled load-add-store on the common path.
// `planes`: 0 when the host proved that no plane can lie between this query's origin and its light (the
// origin is on a bounded primitive and the light's bit of NtDevScene::lfree is set, nt_cull.h); 2 when the origin
// lies in the light's room (only general planes are tested); 1 = every plane.
template <typename R, bool BVH, typename K, bool LEAN>
__device__ __forceinline__ bool occluded(const Ctx<R, BVH, LEAN> &c, const V3<R> &o, const V3<R> &d, R dist, unsigned long long mask,
                                         int planes, K &k) {
    const NtDevScene &s = *c.s;
    R t;
    if constexpr (!BVH) {
        unsigned long long m = mask & s.sph_bits;
        while (m) {
            const unsigned i = (unsigned)__ffsll((long long)m) - 1u;
            m &= m - 1;
            R q[4];
            NT_X(k, xsph, 1u);
            c.ld_sph(i, q);
            if (hit_sphere<R>(q, o, d, c.eps, t) && t < dist) { k.sph += s.ns - (i + 1); k.pln += s.np; k.tri += s.nt; return true; }
        }
    }
    if constexpr (!BVH && sizeof(R) == 8) {
        bool pocc = false;
        if constexpr (LEAN && !counts_executed<K>::value) { // no general planes: planes == 2 has nothing to test
            if (planes == 1) pocc = planes_occluded_cold<R>(c.s, c.axl_addr, c.eps, c.eps_lo, o.x, o.y, o.z, d.x, d.y, d.z, dist);
        } else {
            pocc = planes && planes_occluded<R, K>(c, o, d, dist, planes == 1, k);
        }
        if (pocc) { k.pln += s.np - (first_occluding_plane<R>(c, o, d, dist) + 1); k.tri += s.nt; return true; }
    }
    if constexpr (!BVH && sizeof(R) == 4) {
        if (planes == 2) { // origin in the light's room: general planes only
            NT_X(k, xpln, s.ngen);
#pragma unroll 1
            for (unsigned j = 0; !LEAN && j < s.ngen; ++j) {
                int idx;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(idx) : "r"(c.gen_addr + 4u * j));
                R q[4], dn, num;
                c.ld_pln((unsigned)idx, q);
                plane_eval<R>(q, 3, o, d, dn, num);
                if (plane_finish<R>(dn, num, c.eps, c.eps_lo, t) && t < dist) { k.pln += s.np - ((unsigned)idx + 1); k.tri += s.nt; return true; }
            }
        } else {
            NT_X(k, xpln, s.np); // an upper bound when an occluder ends the loop early
            unsigned i = 0;
            for (; i + 2 <= s.np; i += 2) {
                R q0[4], q1[4], dn0, num0, dn1, num1;
                c.ld_pln(i, q0);
                c.ld_pln(i + 1, q1);
                plane_eval<R>(q0, 3, o, d, dn0, num0);
                plane_eval<R>(q1, 3, o, d, dn1, num1);
                if (plane_finish<R>(dn0, num0, c.eps, c.eps_lo, t) && t < dist) { k.pln += s.np - (i + 1); k.tri += s.nt; return true; }
                if (plane_finish<R>(dn1, num1, c.eps, c.eps_lo, t) && t < dist) { k.pln += s.np - (i + 2); k.tri += s.nt; return true; }
            }
            if (i < s.np) {
                R q[4], dn, num;
                c.ld_pln(i, q);
                plane_eval<R>(q, 3, o, d, dn, num);
                if (plane_finish<R>(dn, num, c.eps, c.eps_lo, t) && t < dist) { k.pln += s.np - (i + 1); k.tri += s.nt; return true; }
            }
        }
    }
    if constexpr (BVH) k.pln += s.np;
    if constexpr (!BVH && !LEAN) {
        if (s.nt) { // uniform
            unsigned long long m = s.ns >= 64 ? 0ull : mask >> s.ns;
            while (m) {
                const unsigned i = (unsigned)__ffsll((long long)m) - 1u;
                m &= m - 1;
                R q[9];
                NT_X(k, xtri, 1u);
                c.ld_tri(i, q);
                if (hit_triangle<R>(q, o, d, c.eps, t) && t < dist) { k.tri += s.nt - (i + 1); return true; }
            }
        }
    }
    return false;
}

// SPEC §4: radiance of one sample = sum over its ray tree in depth-first pre-order of W * local.
// `accp` / `Wp`: the sample's running radiance sum (3 values, stride NT_BLOCK_THREADS) and path weight,
// kept in per-thread shared-memory slots: touched once per tree node, not worth 8 registers.
#define NT_ACC(ch) accp[(ch) * NT_BLOCK_THREADS]
// `pmask`: address of the warp's primary-ray mask (tile_mask) in shared memory.  `own` of the current ray:
// -2 primary ray (mask = *pmask), -1 no culling information (all bounded primitives), >= 0 the ray starts on
// that sphere (see nearest_hit).  Kept as ONE register instead of a live 64-bit mask: the kernel is register-bound.
// RULES: the instantiation that honours the rule switches of SPEC §8 (c.rules); the default kernel compiles them out.
template <typename R, bool BVH, typename K, bool RULES = false, bool LEAN = false>
__device__ __forceinline__ void trace_sample(const Ctx<R, BVH, LEAN> &c, V3<R> o, V3<R> d, R *accp, R *Wp, const unsigned long long *pmask,
                                             K &k) {
    const NtDevScene &s = *c.s;
    const NtSceneView<R> &v = *c.v;
    // deferred transmission children (reflection children are followed immediately)
    R st[NT_MAX_DEPTH_DEV][7];
    unsigned st_depth[NT_MAX_DEPTH_DEV]; // depth | (own + 1) << 8
    int sp = 0;
    *Wp = R(1);
    unsigned depth = 1;
    int own = -2;
    for (;;) {
        R t;
        Hit h;
        bool descend = false;
        const unsigned long long qmask = own == -2 ? *pmask : s.all_bits;
        if (!nearest_hit<R, BVH, K>(c, o, d, qmask, own, t, h, k)) {
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
            } else if constexpr (!LEAN) {
                const R *tp = v.tri + (size_t)h.idx * NT_TRI_STRIDE + 9;
                Ng = { __ldg(tp), __ldg(tp + 1), __ldg(tp + 2) };
                mat = __ldg(s.tri_mat + h.idx);
            } else {
                Ng = { R(0), R(0), R(0) }; mat = 0; // not reachable: a LEAN launch has no triangles
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
                R dist, inv_dist;
                Math<R>::len_inv(d2, dist, inv_dist);
                const V3<R> L = scale(Lv, inv_dist);
                const R ndl = dot(N, L);
                if (!(ndl > R(0))) continue;
                k.shadow++;
                const unsigned long long lmask = s.cull ? lbuf_mask<R>(s, l, Lv) : s.all_bits;
                int planes = 1; // strict mode only: 0 = plane-free light (nt_cull.h); both modes: 2 = origin in the light's room
                if constexpr (sizeof(R) == 8) planes = h.kind == 1 || !((s.lfree >> l) & 1u);
                if constexpr (!BVH) {
                    if (planes && s.rooms && in_light_room<R>(c, l, P, dist)) planes = 2;
                }
                if (occluded<R, BVH, K>(c, P, L, dist, lmask, planes, k)) continue;
                k.light++;
                R m0[4], m1[4];
                Ld<R>::g4(mp, m0);     // r g b ka
                Ld<R>::g4(mp + 4, m1); // kd ks shininess kr
                const R kdn = m1[0] * ndl;
                R lc[3] = { __ldg(lp + 3), __ldg(lp + 4), __ldg(lp + 5) };
                if constexpr (RULES) {
                    if (c.rules & NT_DEV_RULE_ATTENUATE) { // SPEC §8: light colour scaled by 1 / d2
                        const R att = Math<R>::rcp(d2);
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) lc[ch] = lc[ch] * att;
                    }
                }
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
                        // SPEC §8: re-normalised secondary directions - a rule switch in the strict mode, always in the fast mode
                        // (fast_renormalises below)
                        if constexpr (RULES || fast_renormalises<R>::value)
                            if (fast_renormalises<R>::value || (c.rules & NT_DEV_RULE_RENORMALIZE)) T = scale(T, Math<R>::rcp(Math<R>::sqrt_(dot(T, T))));
                    }
                }
                if (wt > R(0)) {
                    k.sec++;
                    if (wr > R(0)) { // defer: reflection subtree comes first in pre-order
                        st[sp][0] = P.x; st[sp][1] = P.y; st[sp][2] = P.z;
                        st[sp][3] = T.x; st[sp][4] = T.y; st[sp][5] = T.z;
                        st[sp][6] = *Wp * wt; st_depth[sp] = (depth + 1) | (unsigned)((s.cull && h.kind == 0 ? h.idx : -1) + 1) << 8;
                        ++sp;
                    }
                }
                if (wr > R(0)) {
                    k.sec++;
                    const R two = R(2) * cosi;
                    V3<R> Rd = { d.x + N.x * two, d.y + N.y * two, d.z + N.z * two };
                    if constexpr (RULES || fast_renormalises<R>::value)
                        if (fast_renormalises<R>::value || (c.rules & NT_DEV_RULE_RENORMALIZE)) Rd = scale(Rd, Math<R>::rcp(Math<R>::sqrt_(dot(Rd, Rd))));
                    o = P; d = Rd; *Wp = *Wp * wr; depth = depth + 1;
                    descend = true;
                } else if (wt > R(0)) {
                    o = P; d = T; *Wp = *Wp * wt; depth = depth + 1;
                    descend = true;
                }
                own = s.cull && h.kind == 0 ? h.idx : -1;
            }
        }
        if (descend) continue;
        if (sp == 0) break;
        --sp;
        o = { st[sp][0], st[sp][1], st[sp][2] };
        d = { st[sp][3], st[sp][4], st[sp][5] };
        *Wp = st[sp][6];
        depth = st_depth[sp] & 0xffu;
        own = (int)(st_depth[sp] >> 8) - 1;
    }
}

// ---- block-level plumbing ----

// Stage the flat intersection data in shared memory with 128-bit loads.
template <typename R, bool BVH, bool LEAN>
__device__ __forceinline__ void stage_scene(const NtDevScene &s, const NtSceneView<R> &v, Ctx<R, BVH, LEAN> &c) {
    const unsigned n_sph = BVH ? 0u : s.ns * 4, n_pln = s.np * 4, n_tri = BVH ? 0u : s.nt * NT_TRI_STRIDE;
    R *smem = (R *)nt_smem;
    constexpr int VEC = 16 / sizeof(R);
    typedef typename std::conditional<sizeof(R) == 8, double2, float4>::type VT;
    if constexpr (!BVH) {
#pragma unroll 1
        for (unsigned i = threadIdx.x; i < n_sph / VEC; i += blockDim.x) { ((VT *)smem)[i] = __ldg((const VT *)v.sph + i); }
    }
#pragma unroll 1
    for (unsigned i = threadIdx.x; i < n_pln / VEC; i += blockDim.x) { ((VT *)(smem + n_sph))[i] = __ldg((const VT *)v.pln + i); }
    if constexpr (!BVH) {
#pragma unroll 1
        for (unsigned i = threadIdx.x; i < n_tri / VEC; i += blockDim.x) { ((VT *)(smem + n_sph + n_pln))[i] = __ldg((const VT *)v.tri + i); }
    }
    const unsigned n_ax = BVH ? 0u : 2u * (s.nax[0] + s.nax[1] + s.nax[2]); // R units: (position, index bits) pairs
    // what follows the lists is read with 128-bit shared loads (rooms, slab entries): the lists are padded to 16 bytes (an
    // odd number of axis-aligned planes left the binary32 rooms 8-byte aligned - a misaligned-address fault, found by
    // scripts/gpu_fuzz_flat.py)
    const unsigned n_axp = (n_ax + 3u) & ~3u;
    if constexpr (BVH) {
        unsigned *codes = (unsigned *)(smem + n_sph + n_pln + n_tri);
        for (unsigned i = threadIdx.x; i < (s.np + 15) / 16; i += blockDim.x) codes[i] = __ldg(s.pln_code + i);
    } else {
        const R *axl = (const R *)(sizeof(R) == 8 ? (const void *)s.axl64 : (const void *)s.axl32);
        for (unsigned i = threadIdx.x; i < n_ax; i += blockDim.x) smem[n_sph + n_pln + n_tri + i] = __ldg(axl + i);
        int *gen = (int *)(smem + n_sph + n_pln + n_tri + n_axp);
        for (unsigned i = threadIdx.x; i < s.ngen; i += blockDim.x) gen[i] = __ldg(s.pgen + i);
        // light rooms [nl][8] (slot 6 becomes this launch's distance cap min(eps * cap_per_eps, cap_max)), slab entries [3][2][2]
        R *room = (R *)(gen + ((s.ngen + 3u) & ~3u));
        const R *rsrc = (const R *)(sizeof(R) == 8 ? (const void *)s.room64 : (const void *)s.room32);
        const R *asrc = (const R *)(sizeof(R) == 8 ? (const void *)s.axs64 : (const void *)s.axs32);
        const unsigned n_room = s.rooms ? 8u * s.nl : 0u;
        for (unsigned i = threadIdx.x; i < n_room; i += blockDim.x) {
            R v = __ldg(rsrc + i);
            if ((i & 7u) == 6u) { const R cap = c.eps * v, cmax = __ldg(rsrc + i + 1); v = cap < cmax ? cap : cmax; }
            room[i] = v;
        }
        for (unsigned i = threadIdx.x; i < 12u; i += blockDim.x) room[n_room + i] = __ldg(asrc + i);
    }
    __syncthreads();
    unsigned base = (unsigned)__cvta_generic_to_shared(nt_smem);
    asm volatile("" : "+r"(base)); // opaque: otherwise ptxas re-derives the window base (S2R + 5 ops) per use
    c.sph_addr = base;
    c.pln_addr = base + n_sph * (unsigned)sizeof(R);
    c.tri_addr = base + (n_sph + n_pln) * (unsigned)sizeof(R);
    c.code_addr = base + (n_sph + n_pln + n_tri) * (unsigned)sizeof(R);
    c.axl_addr = c.code_addr;
    c.gen_addr = c.axl_addr + n_axp * (unsigned)sizeof(R);
}

// Per-thread counters -> one atomic per counter per block, spread over NT_COUNTER_SLOTS slots.
__device__ __forceinline__ void flush_counter_values(const unsigned vals[NT_NCOUNTERS], unsigned long long *counters,
                                                     unsigned long long *s_cnt) {
    if (threadIdx.x < NT_NCOUNTERS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll 1
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
    const unsigned vals[NT_NCOUNTERS] = { k.prim, k.sec, k.shadow, k.sph, k.pln, k.tri, k.box, k.light, 0u, 0u, 0u };
    flush_counter_values(vals, counters, s_cnt);
}
__device__ __forceinline__ void flush_counters(const CountersX &k, unsigned long long *counters,
                                               unsigned long long *s_cnt) {
    const unsigned vals[NT_NCOUNTERS] = { k.prim, k.sec, k.shadow, k.sph, k.pln, k.tri, k.box, k.light, k.xsph, k.xpln, k.xtri };
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
// (Work counters in shared memory as well removed the remaining spills but cost more instructions than it
// saved: 1.34 ms against 1.30.)
// MINB != 0: the same kernel compiled for another number of resident blocks per SM.  The strict kernel exists twice:
// 4 blocks (64 registers) for whole frames, NT_MIN_BLOCKS_SMALL = 3 (80 registers, fewer spills, less contention per
// scheduler) for small launches - a 1/8-frame shard of an 8-GPU render is bound by the latency of its deepest tiles, not
// by throughput, and measured 0.153 ms against 0.166 (a whole frame: 0.831 against 0.810; profiles/r02a_ab_patches.txt).
template <typename R, bool BVH, bool SINGLE, bool EXEC = false, int MINB = 0, bool RULES = false, bool LEAN = false>
__global__ void __launch_bounds__(NT_BLOCK_THREADS, MINB ? MINB : (sizeof(R) == 8 ? NT_MIN_BLOCKS_F64 : NT_MIN_BLOCKS_F32))
render_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a) {
    __shared__ unsigned long long s_cnt[NT_NCOUNTERS];
    __shared__ R s_state[4][NT_BLOCK_THREADS]; // acc r g b, W
    __shared__ unsigned long long s_pmask[NT_BLOCK_THREADS / 32]; // per warp: primary-ray candidates of its tile
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, BVH, LEAN> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.eps_lo = (R)a.eps_lo; c.max_depth = a.max_depth; c.rules = a.rules;
    frame_sync_begin(a); // multi-GPU exchange: acknowledge / wait before the first pixel store (the barrier in stage_scene orders it)
    stage_scene<R, BVH>(s, v, c);

    const unsigned tid = threadIdx.x, lane = tid & 31;
    typedef typename std::conditional<EXEC, CountersX, Counters>::type KT; // EXEC: also the tests really started
    KT k{};
    R *accp = &s_state[0][tid], *Wp = &s_state[3][tid];
    const unsigned warps_per_block = NT_BLOCK_THREADS / 32, total_warps = gridDim.x * warps_per_block;
    const unsigned n_tiles = a.tiles_x * a.tiles_y;
    unsigned long long *next_tile = a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS;

    unsigned tile = blockIdx.x * warps_per_block + (tid >> 5); // first tile: no atomic needed
    while (tile < n_tiles) {
        R sum[3] = { R(0), R(0), R(0) };
        const unsigned rounds = SINGLE ? 1u : a.spp / a.lanes;
        if constexpr (!BVH) {
            unsigned long long pm = s.all_bits;
            if (s.cull) {
                unsigned px0, vr0;
                tile_origin(a, tile, px0, vr0);
                const unsigned vr1 = min(vr0 + a.twy, a.vrows) - 1; // first / last owned row of the tile -> image rows
                pm = tile_mask(s, a, px0, row_to_y(a, vr0), row_to_y(a, vr1) + 1, lane);
            }
            __syncwarp();
            if (lane == 0) s_pmask[tid >> 5] = pm;
            __syncwarp();
        }
        for (unsigned r = 0; r < rounds; ++r) {
            {
                const unsigned L = a.lanes, j = lane & (L - 1), pw = lane >> a.log2_lanes;
                unsigned px, vr;
                tile_origin(a, tile, px, vr);
                px += pw & (a.twx - 1);
                vr += pw >> a.log2_twx;
                NT_ACC(0) = R(0); NT_ACC(1) = R(0); NT_ACC(2) = R(0);
                if (px < a.width && vr < a.vrows) {
                    const unsigned y = row_to_y(a, vr);
                    // SPEC §2: regular n x n grid, sample s = r*L + j
                    const unsigned sidx = r * L + j;
                    const unsigned sj = (sidx * a.n_mul) >> 16, si = sidx - sj * a.n; // sidx / n, sidx % n (sidx < 64, n <= 8)
                    const R ox = ArgsView<R>::samp_off(a, si), oy = ArgsView<R>::samp_off(a, sj); // (i + 0.5) / n, divided on the host (IEEE, same value)
                    const R fx = (R)px + ox, fy = (R)y + oy;
                    const V3<R> D = { (ArgsView<R>::cam(a, 3) + ArgsView<R>::cam(a, 6) * fx) + ArgsView<R>::cam(a, 9) * fy,
                                      (ArgsView<R>::cam(a, 4) + ArgsView<R>::cam(a, 7) * fx) + ArgsView<R>::cam(a, 10) * fy,
                                      (ArgsView<R>::cam(a, 5) + ArgsView<R>::cam(a, 8) * fx) + ArgsView<R>::cam(a, 11) * fy };
                    R dlen, dinv;
                    Math<R>::len_inv(dot(D, D), dlen, dinv);
                    const V3<R> dir = scale(D, dinv);
                    const V3<R> eye = { ArgsView<R>::cam(a, 0), ArgsView<R>::cam(a, 1), ArgsView<R>::cam(a, 2) };
                    k.prim++;
                    trace_sample<R, BVH, KT, RULES, LEAN>(c, eye, dir, accp, Wp, &s_pmask[tid >> 5], k);
                }
            }
            asm volatile("" : "+r"(tile)); // pixel coordinates are recomputed below, not carried across the trace
            // SPEC §5: samples are added in sample order; the lanes of one pixel are adjacent
            const unsigned L = a.lanes;
            if (L == 1) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) sum[ch] = sum[ch] + NT_ACC(ch);
            } else {
                // only the pixel's first lane (which stores the pixel) needs the sum: its own sample, then the
                // samples of the L - 1 lanes above it in sample order; the other lanes' sums are never used
                const R a0 = NT_ACC(0), a1 = NT_ACC(1), a2 = NT_ACC(2);
                sum[0] = sum[0] + a0; sum[1] = sum[1] + a1; sum[2] = sum[2] + a2;
#pragma unroll 1
                for (unsigned jj = 1; jj < L; ++jj) {
                    sum[0] = sum[0] + __shfl_down_sync(0xffffffffu, a0, jj);
                    sum[1] = sum[1] + __shfl_down_sync(0xffffffffu, a1, jj);
                    sum[2] = sum[2] + __shfl_down_sync(0xffffffffu, a2, jj);
                }
            }
        }
        {
            const unsigned L = a.lanes, j = lane & (L - 1), pw = lane >> a.log2_lanes;
            unsigned px, vr;
            tile_origin(a, tile, px, vr);
            px += pw & (a.twx - 1);
            vr += pw >> a.log2_twx;
            if (px < a.width && vr < a.vrows && j == 0) {
                const unsigned y = row_to_y(a, vr);
                const R inv_spp = ArgsView<R>::inv_spp(a); // 1 / spp, divided on the host
                unsigned rgba = 0xff000000u;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const R cv = sum[ch] * inv_spp;
                    unsigned q = cv <= R(0) ? 0u : cv >= R(1) ? 255u : (unsigned)(int)(cv * R(255) + R(0.5));
                    if constexpr (RULES) // SPEC §8: truncation instead of rounding
                        if ((a.rules & NT_DEV_RULE_TRUNCATE) && cv > R(0) && cv < R(1)) q = (unsigned)(int)(cv * R(255));
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
    if constexpr (!BVH) { // deficits -> test counts (see occluded())
        const unsigned queries = k.prim + k.sec + k.shadow;
        k.sph = queries * s.ns - k.sph;
        k.pln = queries * s.np - k.pln;
        k.tri = queries * s.nt - k.tri;
    }
    flush_counters(k, a.counters, s_cnt);
    frame_sync_end(a); // multi-GPU exchange: "this shard is written" once the last block is through
}

// Unit-level entry: nearest hit of arbitrary rays (nt_trace_rays).
template <typename R, bool BVH>
__global__ void __launch_bounds__(NT_BLOCK_THREADS)
trace_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtTraceArgs a) {
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, BVH> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.eps_lo = (R)a.eps_lo; c.max_depth = 1; c.rules = 0;
    stage_scene<R, BVH>(s, v, c);
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const V3<R> o = { (R)a.origins[3 * i], (R)a.origins[3 * i + 1], (R)a.origins[3 * i + 2] };
    const V3<R> d = { (R)a.dirs[3 * i], (R)a.dirs[3 * i + 1], (R)a.dirs[3 * i + 2] };
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };
    R t;
    Hit h;
    if (nearest_hit<R, BVH, Counters>(c, o, d, s.all_bits, -1, t, h, k)) {
        a.t_out[i] = (double)t;
        a.prim_out[i] = h.gid;
    } else {
        a.t_out[i] = -1.0;
        a.prim_out[i] = -1;
    }
}

template <typename R>
inline size_t flat_smem_bytes(const NtDevScene &s, bool bvh) {
    size_t n = (size_t)s.np * 4;
    if (bvh) return n * sizeof(R) + (size_t)((((s.np + 15) / 16) + 3) & ~3u) * sizeof(unsigned);
    n += (size_t)s.ns * 4 + (size_t)s.nt * NT_TRI_STRIDE;
    n += (2 * ((size_t)s.nax[0] + s.nax[1] + s.nax[2]) + 3) & ~(size_t)3; // axis-aligned plane lists, padded to 16 bytes
    n += (s.rooms ? 8 * (size_t)s.nl : 0) + 12;                  // light rooms, slab entries
    return n * sizeof(R) + (((size_t)s.ngen + 3) & ~(size_t)3) * sizeof(int);
}

} // namespace nt

#include "nt_bvh_trace.cuh"
#include "nt_eyegrid.cuh"
#include "nt_wavefront.cuh"

namespace nt {

template <typename R, bool BVH>
inline int launch_render_t(const NtDevScene &s, const NtRenderArgs &a, cudaStream_t st) {
    static int blocks_per_sm[64] = { 0 }, blocks_small[64] = { 0 }, sms[64] = { 0 }; // per device, resolved once
    constexpr int MINB_SMALL = sizeof(R) == 8 && !BVH ? NT_MIN_BLOCKS_SMALL : 0; // strict flat kernel only
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = flat_smem_bytes<R>(s, BVH);
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!sms[dev]) {
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (BVH) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], render_bvh_kernel<R>, NT_BLOCK_THREADS, 4096);
        else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], render_kernel<R, false, true>, NT_BLOCK_THREADS, 4096);
            if constexpr (MINB_SMALL != 0)
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_small[dev], render_kernel<R, false, true, false, MINB_SMALL>, NT_BLOCK_THREADS, 4096);
            if (blocks_small[dev] < 1) blocks_small[dev] = 1;
        }
        if (blocks_per_sm[dev] < 1) blocks_per_sm[dev] = 1;
    }
    const unsigned n_tiles = a.tiles_x * a.tiles_y, wpb = NT_BLOCK_THREADS / 32;
    unsigned grid = (unsigned)(sms[dev] * blocks_per_sm[dev]);
    if (grid > (n_tiles + wpb - 1) / wpb) grid = (n_tiles + wpb - 1) / wpb;
    if (BVH) {
        // many kernels per frame: the exchange flags (nt_sync.cuh) are handled before and after the pipeline
        unsigned long long *timeouts = a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS + 2;
        if (a.sync_post_ptr || a.sync_wait_ptr) {
            sync_kernel<<<1, 32, 0, st>>>(a.sync_post_ptr, a.sync_post_val, a.sync_wait_ptr, 1, a.sync_wait_val, nullptr, 0, timeouts);
            if (a.n_launches) *a.n_launches += 1;
        }
        int rc = 0;
        if (s.eg_on) rc = launch_eye_grid(s, a.cam, st, a.n_launches); // this call's eye: the primary rays' sphere lists (nt_eyegrid.cuh)
        if (rc) return rc;
        if (a.wf) rc = launch_wavefront<R>(s, a, st, sms[dev], blocks_per_sm[dev]);
        else {
            render_bvh_kernel<R><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
            resolve_kernel<R><<<dim3((a.width + 255) / 256, a.vrows), 256, 0, st>>>(a);
            if (a.n_launches) *a.n_launches += 2;
            rc = (int)cudaGetLastError();
        }
        if (!rc && a.sync_done_ptr) {
            sync_kernel<<<1, 32, 0, st>>>(nullptr, 0, nullptr, 0, 0, a.sync_done_ptr, a.sync_done_val, timeouts);
            if (a.n_launches) *a.n_launches += 1;
            rc = (int)cudaGetLastError();
        }
        return rc;
    }
    if constexpr (!BVH) {
        // fewer than NT_SMALL_TILES_PER_WARP warp tiles per resident warp: the small-launch variant (NT_SMALL_LAUNCH=0 / 1
        // forces the choice, A/B)
        // ... and whose shadow queries skip the axis lists nearly always (light rooms) or have no plane at all
        bool lean = s.nt == 0 && s.ngen == 0 && (s.rooms || s.np == 0);
        if (const char *e = getenv("NT_LEAN")) if (e[0] == '0') lean = false; // A/B, tests
        bool small = MINB_SMALL != 0 && !a.count_executed && !a.rules && n_tiles < (unsigned)NT_SMALL_TILES_PER_WARP * (unsigned)(sms[dev] * blocks_per_sm[dev]) * wpb;
        if (const char *e = getenv("NT_SMALL_LAUNCH")) small = MINB_SMALL != 0 && !a.count_executed && !a.rules && e[0] == '1';
        if constexpr (MINB_SMALL != 0) {
            if (small) {
                grid = (unsigned)(sms[dev] * blocks_small[dev]);
                if (grid > (n_tiles + wpb - 1) / wpb) grid = (n_tiles + wpb - 1) / wpb;
                if (lean) {
                    if (a.spp == a.lanes) render_kernel<R, false, true, false, MINB_SMALL, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
                    else render_kernel<R, false, false, false, MINB_SMALL, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
                } else if (a.spp == a.lanes) render_kernel<R, false, true, false, MINB_SMALL><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
                else render_kernel<R, false, false, false, MINB_SMALL><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
                if (a.n_launches) *a.n_launches += 1;
                return (int)cudaGetLastError();
            }
        }
        if (a.rules) { // SPEC §8 rule switches: the instantiation that reads them (a little slower than the default kernel)
            if (a.spp == a.lanes) render_kernel<R, false, true, false, 0, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
            else render_kernel<R, false, false, false, 0, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
        } else if (a.count_executed) { // instrumented twin (nt_render_params.flags & NT_RENDER_COUNT_EXECUTED): measurement only
            if (a.spp == a.lanes) render_kernel<R, false, true, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
            else render_kernel<R, false, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
        } else if (lean) { // neither triangles nor general planes: the specialisation without their loops
            if (a.spp == a.lanes) render_kernel<R, false, true, false, 0, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
            else render_kernel<R, false, false, false, 0, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
        } else if (a.spp == a.lanes) render_kernel<R, false, true><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
        else render_kernel<R, false, false><<<grid, NT_BLOCK_THREADS, smem, st>>>(s, a);
        if (a.n_launches) *a.n_launches += 1;
    }
    return (int)cudaGetLastError();
}

// Whole-frame wavefront workspace; 0 when the wavefront path does not apply (flat scene, deep trees, > 32 lights).
template <typename R>
inline size_t wavefront_bytes(const NtDevScene &s, const NtRenderArgs &a) {
    if (!s.use_bvh || a.max_depth > NT_WF_MAX_DEPTH || s.nl > 32) return 0;
    const size_t n = (size_t)a.tiles_x * a.tiles_y * (a.spp / a.lanes) * 32;
    return 512 + 256 * 8 * (size_t)a.max_depth + n * wf_bytes_per_sample<R>(a.max_depth) + wf_sweep_bytes<R>(n) +
           (wf_sort_mode() ? 4 * wf_sort_fixed_bytes() + n * wf_sort_bytes_per_sample(a.max_depth) + 1024 : 0);
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
