// nt_bvh_trace.cuh — the intersect-and-shade path for scenes whose bounded primitives live in HBM
// behind a BVH (BASELINE.json configs[3], [4]).  Same arithmetic as nt_trace.cuh (SPEC-PROVISIONAL
// §2-§5, same operation order, so strict mode stays bit-exact), different control structure:
//
//   Every lane runs a small state machine over its sample's ray tree.  At any moment a lane has at
//   most ONE ray query in flight (nearest hit of a tree node, or the occlusion query of one light), and
//   ALL kinds of query run through the same traversal code.  The warp alternates between
//     (A) traversal rounds — one inner-node step and/or one leaf step per lane per round — and
//     (B) advance — lanes whose query has finished shade, pick the next light, spawn children or
//         finish — which is short.
//   With a recursive-looking per-lane structure (nt_trace.cuh's) the occlusion traversal and the
//   nearest-hit traversal are separate instruction streams and lanes that drift apart serialise: ncu
//   measured 6.1 of 32 threads active per instruction on the 1M-triangle scene (profiles/r01_*).
//   Here lanes drift in *state*, not in program counter.
//
//   Lanes are also refilled: a lane whose sample is complete stores the sample's radiance to the
//   per-sample buffer and, once enough lanes of the warp are idle, the warp claims that many new
//   samples with ONE atomicAdd (ballot + popc prefix gives each idle lane its sample id).  Ray trees
//   differ wildly on these scenes (a terrain hit is 3 queries, a glass sphere up to 21 at depth 3), so
//   with a static lane -> sample binding most lanes of a warp idle behind the deepest tree.  A small
//   resolve kernel then adds the samples of each pixel in sample order (SPEC §5) and quantises.
#pragma once
#include "nt_trace.cuh"

namespace nt {

#define NT_REF_EMPTY (-1)
// leaf ref = -2 - (first | (count-1) << 26 | type << 28); inner ref = node index >= 0
__device__ __forceinline__ bool ref_is_leaf(int r) { return r <= -2; }

template <typename R> struct BvhQuery { // scalars only: lives in registers (the stack is a separate array)
    V3<R> o, d;
    R tb;          // nearest: best t so far; any: the distance bound
    Hit best;
    // binary32 copy of the ray for the box tests (origin shifted by tshift): t = plane*i + c, with the
    // margin folded into c (cn*: near planes moved outward, cf*: far planes moved outward; both stored negated)
    float ix, iy, iz, cnx, cny, cnz, cfx, cfy, cfz, tmaxf;
    int nearx, neary, nearz; // float4 index of the near plane of each axis inside a node (lo: a, hi: 3 + a)
    R tshift;      // box-test frame: t' = t - tshift (0 unless the origin lies outside the scene bounds)
    int cur;       // ref being visited, NT_REF_EMPTY when the query needs a pop
    int sp;
    int pass;      // 0: one walk of the whole tree; 1: triangle set, the sphere set follows; 2: sphere set (query_arm)
    float fix, fiy, fiz; // pass 2 only: reciprocals for the FAR planes (the margin is a cone there, see query_arm)
    bool any, done, found;
};

// Arm the tree walk of a query: the binary32 copy of the ray, the box margin, the entry slab against the scene bounds.
//   pass 0  decide: one walk of the whole tree (from node 0) - or, when the direction has drifted from unit length so far
//           that the sphere boxes must grow by more than the rounding margin, pass 1;
//   pass 1  the triangle set only (root n_nodes - 2), plain margin; query_second_pass then arms
//   pass 2  the sphere set only (root n_nodes - 1), grown margin: a cone (below), walked by query_inner_step_t<true>;
//   pass 3  the same with the constant bound, for a caller whose loop has only the plain inner step.
//   pass 4  the triangle set only and nothing after it: an occlusion query whose spheres came from the light's shadow grid
//           (shadow_query_start).
// Why two passes: SPEC §3's sphere test takes the direction as a unit vector, but §4 does not re-normalise reflected /
// refracted directions: |d|^2 = L2 drifts from 1 along a mirror chain (every bounce off a small distant sphere amplifies
// the drift by ~ 4 (t / r)^2), and the rule then accepts roots t whose point o + d t lies at distance sqrt(r^2 + (L2 - 1)
// t^2) <= r + sqrt(L2 - 1) t from the centre - OUTSIDE the sphere and possibly outside its box.  Boxes may only cull what
// the rule cannot hit, so sphere boxes grow by that bound with t <= the largest distance from the origin to a scene point
// (found by bench.py's frame check against the brute-force oracle).  The triangle test does not depend on |d| (Moeller -
// Trumbore is exact for any scaling of the direction), so triangle boxes never grow: after three sphere bounces the growth
// is the whole scene, and a handful of such rays that walked the grown TRIANGLE tree tested every one of a million
// triangles each (configs[4], depth 5: 15 s per frame instead of 0.19 - measured in round 2, the reason for the passes).
// (Strict mode only.  The fast mode has no bit contract, its sphere test is a different formula (SPEC §7), and binary32
// normals of small distant spheres are unit vectors to 1e-4 at best: growing boxes for that would swallow the tree.)
template <typename R>
__device__ __forceinline__ void query_arm(const Ctx<R, true> &c, BvhQuery<R> &q, const V3<R> &o, const V3<R> &d, int pass) {
    const NtDevScene &s = *c.s;
    float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    const float dx = (float)d.x, dy = (float)d.y, dz = (float)d.z;
    // reciprocal direction, magnitude clamped so that plane*i - c never evaluates inf - inf
    // ... and never 0 (an infinite direction component): an empty slot's far plane is -inf * i, which must stay -inf - the
    // inner step tells empty slots by their inverted boxes alone
    const float ix = copysignf(fmaxf(fminf(1.0f / fabsf(dx), 1e18f), 1e-30f), dx), iy = copysignf(fmaxf(fminf(1.0f / fabsf(dy), 1e18f), 1e-30f), dy),
                iz = copysignf(fmaxf(fminf(1.0f / fabsf(dz), 1e18f), 1e-30f), dz);
    q.ix = ix; q.iy = iy; q.iz = iz;
    // Box margin: covers rounding the ray to binary32 (origin, direction, reciprocal, slab products).
    // It grows with |origin|, so a far origin (a camera outside the scene, a hit on an unbounded plane
    // kilometres away) is first slid along the exact ray to where it enters the scene bounds: only the
    // box tests use the shifted copy, the primitive tests keep the original ray.
    float m = 4e-6f * (fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz)) + s.max_abs);
    float grow = [&] {
        if constexpr (sizeof(R) == 4) return 0.0f;
        const R l2 = dot(d, d);
        if (!(l2 > R(1))) return l2 == l2 ? 0.0f : CUDART_INF_F;
        const float far_ = 1.7320508f * (fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz)) + s.max_abs);
        return __fsqrt_ru(Math<R>::up(l2 - R(1))) * far_ * 1.00001f;
    }();
    if constexpr (sizeof(R) == 8) { // the fast mode never grows boxes: it has one pass, and none of the code of the others
        if (pass == 0 && !(grow <= 4.0f * m)) pass = 1; // also when grow is NaN
    }
    if (pass == 1 || pass == 4) grow = 0.0f;
    q.pass = pass;
    const float grow_all = grow; // pass 2: the entry slab below keeps the constant bound, the node tests use the cone
    m += grow;
    float tn = 0.0f, tf = CUDART_INF_F;
    slab(s.blo[0], s.bhi[0], ox, ix, m, tn, tf);
    slab(s.blo[1], s.bhi[1], oy, iy, m, tn, tf);
    slab(s.blo[2], s.bhi[2], oz, iz, m, tn, tf);
    q.sp = 0;
    if (!(tn <= tf) || tn > Math<R>::up(q.tb)) { q.done = true; q.cur = NT_REF_EMPTY; return; } // misses every bounded primitive (of this pass)
    q.tshift = R(0);
    if (tn > 0.0f) {
        const R ts = (R)tn;
        const V3<R> os = { o.x + d.x * ts, o.y + d.y * ts, o.z + d.z * ts }; // a point of the exact ray
        ox = (float)os.x; oy = (float)os.y; oz = (float)os.z;
        m = 4e-6f * (fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz)) + s.max_abs) + grow; // grow: from the ORIGINAL origin, where t is measured
        q.tshift = ts;
    }
    // near plane of an axis: lo when the ray runs in +axis, hi otherwise; each moved outward by m
    const bool px = !(dx < 0.0f), py = !(dy < 0.0f), pz = !(dz < 0.0f);
    q.nearx = px ? 0 : 3; q.neary = py ? 1 : 4; q.nearz = pz ? 2 : 5;
    if (pass == 3) pass = 2; // (q.pass already says 3: the caller walks with the plain inner step, see below)
    else if (pass == 2 && grow_all < CUDART_INF_F) {
        // Pass 2, the margin as a CONE: a root t of the sphere rule lies within alpha t of its sphere, alpha = sqrt(L2 - 1),
        // so a box can hold a hit only if lo - alpha t <= o + d t <= hi + alpha t on every axis for some t: per axis
        // t (|d| + alpha) >= (near plane - o) and t (|d| - alpha) <= (far plane - o) - the slab test with one reciprocal
        // for the near planes and another for the far planes; an axis with |d| <= alpha (or nearly) bounds nothing above.
        // t is measured from the ORIGINAL origin: the slid origin adds the constant alpha * tshift.  alpha is taken 0.1 %
        // too large, which also covers the rounding of the two reciprocals.  (With the constant bound alpha * reach a ray
        // after two sphere bounces walked the whole sphere tree: a few thousand such rays of configs[4] held every
        // traversal kernel's tail for milliseconds - 394 ms per 1/8 frame against 176 ms with re-normalised directions.)
        const float alpha = __fsqrt_ru(Math<R>::up(dot(d, d) - R(1))) * 1.001f;
        const float mc = 4e-6f * (fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz)) + s.max_abs) + alpha * Math<R>::up(q.tshift) * 1.0001f;
        const float adx = fabsf(dx), ady = fabsf(dy), adz = fabsf(dz);
        const float inx = copysignf(1.0f / (adx + alpha), dx), iny = copysignf(1.0f / (ady + alpha), dy), inz = copysignf(1.0f / (adz + alpha), dz);
        const bool bx = adx - alpha > 1e-3f * adx, by = ady - alpha > 1e-3f * ady, bz = adz - alpha > 1e-3f * adz;
        q.ix = inx; q.iy = iny; q.iz = inz;
        q.fix = bx ? copysignf(1.0f / (adx - alpha), dx) : 0.0f;
        q.fiy = by ? copysignf(1.0f / (ady - alpha), dy) : 0.0f;
        q.fiz = bz ? copysignf(1.0f / (adz - alpha), dz) : 0.0f;
        q.cnx = -((px ? ox + mc : ox - mc) * inx); q.cfx = bx ? -((px ? ox - mc : ox + mc) * q.fix) : CUDART_INF_F;
        q.cny = -((py ? oy + mc : oy - mc) * iny); q.cfy = by ? -((py ? oy - mc : oy + mc) * q.fiy) : CUDART_INF_F;
        q.cnz = -((pz ? oz + mc : oz - mc) * inz); q.cfz = bz ? -((pz ? oz - mc : oz + mc) * q.fiz) : CUDART_INF_F;
        q.tmaxf = Math<R>::up(q.tb - q.tshift);
        q.done = false;
        q.cur = (int)s.n_nodes - 1;
        return;
    }
    q.fix = ix; q.fiy = iy; q.fiz = iz;
    // stored negated: they are the addends of the slab FMAs
    q.cnx = -((px ? ox + m : ox - m) * ix); q.cfx = -((px ? ox - m : ox + m) * ix);
    q.cny = -((py ? oy + m : oy - m) * iy); q.cfy = -((py ? oy - m : oy + m) * iy);
    q.cnz = -((pz ? oz + m : oz - m) * iz); q.cfz = -((pz ? oz - m : oz + m) * iz);
    q.tmaxf = Math<R>::up(q.tb - q.tshift);
    q.done = false;
    q.cur = pass == 0 ? 0 : pass == 4 ? (int)s.n_nodes - 2 : (int)s.n_nodes - 3 + pass; // pass 1 / 4: root n - 2 (triangles), pass 2 / 3: root n - 1 (spheres)
}
// A query that walked the triangle set alone (pass 1) goes on with the sphere set unless an occlusion query has already
// found its occluder.  Called by the traversal loops for lanes whose walk has just ended:
//   * moderate drift: pass 2 of query_arm, the sphere tree under the cone margin;
//   * alpha = sqrt(|d|^2 - 1) >= NT_SWEEP_ALPHA (a third sphere bounce and beyond): the cone holds most of the scene and
//     the rule itself accepts roots far off the spheres, so the walk degenerates into testing every sphere - one lane, ten
//     thousand exact tests, milliseconds during which its traversal kernel cannot end.  The wavefront pipeline DEFERS such
//     a ray to a list (query_wants_sweep) and a small kernel takes each listed ray with a whole warp: every lane tests the
//     spheres i = lane (mod 32) with the exact rule, a butterfly picks the nearest.  The same answer as the walk, 1/32 of
//     its latency, and none of its code in the traversal loop (inlined there, or called, it cost the common path 25 %).
//     Only nearest-hit queries can drift: a shadow ray's direction is normalised (SPEC section 4).
#ifndef NT_SWEEP_ALPHA
#define NT_SWEEP_ALPHA 0.05f
#endif
template <typename R>
__device__ __forceinline__ bool query_wants_sweep(const V3<R> &d) {
    const R e = dot(d, d) - R(1);
    return !(e < R(NT_SWEEP_ALPHA) * R(NT_SWEEP_ALPHA)); // also NaN
}
template <typename R>
__device__ __forceinline__ void query_second_pass(const Ctx<R, true> &c, BvhQuery<R> &q, const V3<R> &o, const V3<R> &d) {
    if constexpr (sizeof(R) == 8) {
        if (q.done && q.pass == 1 && !(q.any && q.found)) query_arm<R>(c, q, o, d, 2);
    }
}
template <typename R> struct SweepHit { R t; int idx, gid; }; // idx < 0: nothing
// The sweep (nt_wavefront.cuh wf_sweep_kernel): called by a whole warp with the same ray; returns the warp's nearest sphere
// hit by the exact rule, ties to the smallest global id (SPEC section 3).
template <typename R>
__device__ __forceinline__ SweepHit<R> sweep_spheres(const R *sph, const int *sph_gid, unsigned ns, R eps, const V3<R> &so, const V3<R> &sd) {
    const unsigned lane = threadIdx.x & 31;
    R bt = Math<R>::inf();
    int bidx = -1, bgid = 0x7fffffff;
    for (unsigned i = lane; i < ns; i += 32) {
        R p[4], t;
        Ld<R>::g4(sph + 4 * (size_t)i, p);
        if (!hit_sphere<R>(p, so, sd, eps, t)) continue;
        const int gid = __ldg(sph_gid + i);
        if (t < bt || (t == bt && gid < bgid)) { bt = t; bidx = (int)i; bgid = gid; }
    }
#pragma unroll 1
    for (int off = 16; off; off >>= 1) {
        const R t2 = __shfl_xor_sync(0xffffffffu, bt, off);
        const int g2 = __shfl_xor_sync(0xffffffffu, bgid, off), i2 = __shfl_xor_sync(0xffffffffu, bidx, off);
        if (t2 < bt || (t2 == bt && g2 < bgid)) { bt = t2; bgid = g2; bidx = i2; }
    }
    return { bt, bidx, bgid };
}

// Start a query: planes (unbounded, staged in shared memory) are tested here, then the tree is armed.
template <typename R, typename K>
__device__ __forceinline__ void query_start(const Ctx<R, true> &c, BvhQuery<R> &q, const V3<R> &o, const V3<R> &d,
                                            R tmax, bool any, K &k, int pass0 = 0) {
    const NtDevScene &s = *c.s;
    q.o = o; q.d = d; q.tb = tmax; q.any = any; q.done = false; q.found = false; q.pass = 0;
    q.best.kind = -1; q.best.idx = -1; q.best.gid = 0x7fffffff;
    q.cur = NT_REF_EMPTY; q.sp = 0;
    R t;
    unsigned codes = 0;
    for (unsigned i = 0; i < s.np; ++i) {
        R pq[4];
        if ((i & 15) == 0) codes = c.pln_codes(i >> 4);
        c.ld_pln(i, pq);
        const int code = (int)(codes & 3u);
        codes >>= 2;
        k.pln++;
        if (hit_plane<R>(pq, code, o, d, c.eps, c.eps_lo, plane_bound<R>(q.tb), t) && t < q.tb) {
            if (any) { q.done = true; q.found = true; return; }
            q.tb = t; q.best.kind = 1; q.best.idx = (int)i; q.best.gid = (int)(s.ns + i); q.found = true;
        }
    }
    if (s.n_nodes == 0) { q.done = true; return; }
    query_arm<R>(c, q, o, d, pass0);
}

// Occlusion query of light l from P (direction Ld, unit; the light at distance dist).  With a shadow grid for that light
// (NtShadowGrid, nt_shadowgrid.h) the spheres that can lie between P and the light are the few listed in P's cell: they are
// tested here with the exact rule, and the tree walk that follows covers the triangle set only (query_arm pass 4) - or is
// not needed at all.  Same boolean as the walk of the whole tree: the lists are conservative, the test is the leaf's.
template <typename R, typename K>
__device__ __forceinline__ void shadow_query_start(const Ctx<R, true> &c, BvhQuery<R> &q, const V3<R> &P, const V3<R> &Ld, R dist,
                                                   unsigned l, K &k) {
    const NtDevScene &s = *c.s;
    int pass0 = 0;
    if (s.sg_on) {
        const NtShadowGrid *g = s.sgrid + l;
        const uint4 kb = __ldg((const uint4 *)&g->K); // K base valid pad
        if (kb.z) {
            const float4 g0 = __ldg((const float4 *)g), g1 = __ldg((const float4 *)g + 1), g2 = __ldg((const float4 *)g + 2), g3 = __ldg((const float4 *)g + 3);
            // g0 = L.xyz axis.x | g1 = axis.yz U.xy | g2 = U.z V.xyz | g3 = u0 v0 su sv
            const float dx = (float)P.x - g0.x, dy = (float)P.y - g0.y, dz = (float)P.z - g0.z;
            const float w = dx * g0.w + dy * g1.x + dz * g1.y;
            pass0 = 4;
            if (w > 0.0f) {
                const float iw = 1.0f / w;
                const float fu = ((dx * g1.z + dy * g1.w + dz * g2.x) * iw - g3.x) * g3.z, fv = ((dx * g2.y + dy * g2.z + dz * g2.w) * iw - g3.y) * g3.w;
                const float kf = (float)kb.x;
                if (fu >= 0.0f && fv >= 0.0f && fu < kf && fv < kf) { // (NaN: outside)
                    const unsigned cell = kb.y + (unsigned)fv * kb.x + (unsigned)fu;
                    const uint32_t i0 = __ldg(s.sg_off + cell), i1 = __ldg(s.sg_off + cell + 1);
                    for (uint32_t i = i0; i < i1; ++i) {
                        const uint32_t idx = __ldg(s.sg_items + i);
                        R p[4], t;
                        c.ld_sph(idx, p);
                        k.sph++;
                        if (hit_sphere<R>(p, P, Ld, c.eps, t) && t < dist) {
                            q.o = P; q.d = Ld; q.tb = dist; q.any = true; q.found = true; q.done = true; q.pass = 4;
                            q.cur = NT_REF_EMPTY; q.sp = 0;
                            return;
                        }
                    }
                }
            }
        }
    }
    query_start<R>(c, q, P, Ld, dist, true, k, pass0);
}

// Nearest-hit query of a PRIMARY ray (origin = the eye of this render call).  With an eye grid (nt_eyegrid.cuh: grid nl, built
// at the start of the call) the spheres the ray can hit are those its cell lists: the planes first as always, then the walk
// of the triangle set is armed (pass 4), then the listed spheres are tested with the exact rule - nearest wins, ties to the
// smallest global id, exactly the leaf step's rule - and their bound goes into the walk.
template <typename R, typename K>
__device__ __forceinline__ void primary_query_start(const Ctx<R, true> &c, BvhQuery<R> &q, const V3<R> &o, const V3<R> &d, K &k) {
    const NtDevScene &s = *c.s;
    if (s.eg_on) {
        const NtShadowGrid *g = s.sgrid + s.nl;
        const uint4 kb = __ldg((const uint4 *)&g->K); // K base valid pad
        if (kb.z) {
            query_start<R>(c, q, o, d, Math<R>::inf(), false, k, 4);
            if (q.done) return; // past every bounded primitive, or a plane nearer than the scene's bounds
            const float4 g0 = __ldg((const float4 *)g), g1 = __ldg((const float4 *)g + 1), g2 = __ldg((const float4 *)g + 2), g3 = __ldg((const float4 *)g + 3);
            const float dx = (float)d.x, dy = (float)d.y, dz = (float)d.z;
            const float w = dx * g0.w + dy * g1.x + dz * g1.y;
            if (!(w > 0.0f)) return; // pointing away from every sphere
            const float iw = 1.0f / w;
            const float fu = ((dx * g1.z + dy * g1.w + dz * g2.x) * iw - g3.x) * g3.z, fv = ((dx * g2.y + dy * g2.z + dz * g2.w) * iw - g3.y) * g3.w;
            const float kf = (float)kb.x;
            if (!(fu >= 0.0f && fv >= 0.0f && fu < kf && fv < kf)) return;
            const unsigned cell = (unsigned)fv * kb.x + (unsigned)fu;
            const uint32_t i0 = __ldg(s.eg_off + cell), i1 = __ldg(s.eg_off + cell + 1);
            for (uint32_t i = i0; i < i1; ++i) {
                const int idx = (int)__ldg(s.eg_items + i);
                R p[4], t;
                c.ld_sph((unsigned)idx, p);
                k.sph++;
                if (!hit_sphere<R>(p, o, d, c.eps, t)) continue;
                if (t < q.tb) {
                    q.tb = t; q.best.kind = 0; q.best.idx = idx; q.best.gid = __ldg(s.sph_gid + idx); q.found = true;
                } else if (t == q.tb) {
                    const int gid = __ldg(s.sph_gid + idx);
                    if (gid < q.best.gid) { q.best.kind = 0; q.best.idx = idx; q.best.gid = gid; q.found = true; }
                }
            }
            q.tmaxf = Math<R>::up(q.tb - q.tshift);
            return;
        }
    }
    query_start<R>(c, q, o, d, Math<R>::inf(), false, k);
}

template <typename R> __device__ __forceinline__ void query_pop(BvhQuery<R> &q, const int2 *stack) {
    while (q.sp > 0) {
        const int2 e = stack[--q.sp];
        if (__int_as_float(e.y) <= q.tmaxf) { q.cur = e.x; return; }
    }
    q.cur = NT_REF_EMPTY;
    q.done = true;
}

// One inner node of the 4-wide tree: 7 x 128-bit loads (near planes, far planes, refs), 4 slab tests as
// 6 FMAs + max3/min3 each, then the nearest hit child (min over four (t_near bits | slot) keys) is visited next
// and the other hit children are pushed with their entry distance.
template <bool CONE, typename R, typename K>
__device__ __forceinline__ void query_inner_step_t(const Ctx<R, true> &c, BvhQuery<R> &q, int2 *stack, K &k) {
    const float4 *n = (const float4 *)(c.s->nodes + q.cur);
    const float4 nx = __ldg(n + q.nearx), ny = __ldg(n + q.neary), nz = __ldg(n + q.nearz);
    const float4 fx = __ldg(n + (3 - q.nearx)), fy = __ldg(n + (5 - q.neary)), fz = __ldg(n + (7 - q.nearz));
    const int4 rf = __ldg((const int4 *)(n + 6));
    k.box += 4;
    // The 24 slab products as 12 packed binary32 FMAs (sm_100 fma.rn.f32x2): children (0, 1) and (2, 3) of one plane
    // sit in an aligned register pair of its 128-bit load, the ray's reciprocal direction and plane offset are scalar
    // operands broadcast to both halves (SASS: FFMA2 R, R.F32x2.HI_LO, R.F32, -R.F32).  Entry = max3 + max, exit =
    // min3 + min with the query's bound folded in, so that one comparison decides; an empty slot holds an inverted
    // infinite box (both builders) and fails it without looking at its ref.
    const float2 ix = make_float2(q.ix, q.ix), iy = make_float2(q.iy, q.iy), iz = make_float2(q.iz, q.iz);
    const float2 cnx = make_float2(q.cnx, q.cnx), cny = make_float2(q.cny, q.cny), cnz = make_float2(q.cnz, q.cnz);
    const float2 cfx = make_float2(q.cfx, q.cfx), cfy = make_float2(q.cfy, q.cfy), cfz = make_float2(q.cfz, q.cfz);
    const float2 ax[2] = { __ffma2_rn(make_float2(nx.x, nx.y), ix, cnx), __ffma2_rn(make_float2(nx.z, nx.w), ix, cnx) };
    const float2 ay[2] = { __ffma2_rn(make_float2(ny.x, ny.y), iy, cny), __ffma2_rn(make_float2(ny.z, ny.w), iy, cny) };
    const float2 az[2] = { __ffma2_rn(make_float2(nz.x, nz.y), iz, cnz), __ffma2_rn(make_float2(nz.z, nz.w), iz, cnz) };
    // far planes: the same reciprocals, except in the cone pass of a drifted direction (query_arm)
    const float2 jx = CONE ? make_float2(q.fix, q.fix) : ix, jy = CONE ? make_float2(q.fiy, q.fiy) : iy, jz = CONE ? make_float2(q.fiz, q.fiz) : iz;
    const float2 bx[2] = { __ffma2_rn(make_float2(fx.x, fx.y), jx, cfx), __ffma2_rn(make_float2(fx.z, fx.w), jx, cfx) };
    const float2 by[2] = { __ffma2_rn(make_float2(fy.x, fy.y), jy, cfy), __ffma2_rn(make_float2(fy.z, fy.w), jy, cfy) };
    const float2 bz[2] = { __ffma2_rn(make_float2(fz.x, fz.y), jz, cfz), __ffma2_rn(make_float2(fz.z, fz.w), jz, cfz) };
    const int ra[4] = { rf.x, rf.y, rf.z, rf.w };
    int key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j >> 1;
        const float tn = fmaxf(fmaxf(j & 1 ? ax[h].y : ax[h].x, j & 1 ? ay[h].y : ay[h].x), fmaxf(j & 1 ? az[h].y : az[h].x, 0.0f));
        const float tf = fminf(fminf(j & 1 ? bx[h].y : bx[h].x, j & 1 ? by[h].y : by[h].x), fminf(j & 1 ? bz[h].y : bz[h].x, q.tmaxf));
        // (cone pass: an axis without a far bound multiplies the far plane by 0 - an empty slot's -inf becomes NaN, min
        // ignores it and the inverted box could pass with an unbounded query: there the ref is looked at)
        const bool hit = CONE ? (tn <= tf && ra[j] != NT_REF_EMPTY) : tn <= tf;
        key[j] = hit ? ((__float_as_int(tn) & ~3) | j) : 0x7fffffff; // tn >= 0: integer order == float order
    }
    // Next = the nearest hit child; the other hit children are pushed in slot order with their entry distance
    // (a pop discards entries that start beyond the current bound).  A full sort of the four keys (the first
    // version, a 5-exchange network) cost a third of this routine and saved only 1.4 % of the node visits on
    // configs[3] (66.6 -> 63.1 ms without it); occlusion queries do not care about the order at all.
    const int best = min(min(key[0], key[1]), min(key[2], key[3]));
    if (best == 0x7fffffff) { query_pop(q, stack); return; }
    const int slot = best & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (key[j] != 0x7fffffff && j != slot) stack[q.sp++] = make_int2(ra[j], key[j] & ~3);
    q.cur = slot == 0 ? rf.x : slot == 1 ? rf.y : slot == 2 ? rf.z : rf.w;
}
template <typename R, typename K>
__device__ __forceinline__ void query_inner_step(const Ctx<R, true> &c, BvhQuery<R> &q, int2 *stack, K &k) {
    if (sizeof(R) == 8 && q.pass == 2) query_inner_step_t<true, R, K>(c, q, stack, k); // rare: cold copy of the step
    else query_inner_step_t<false, R, K>(c, q, stack, k);
}

// One leaf: up to 4 primitives of one kind, exact tests in R.
// `o`, `d`: the query's exact ray.  The per-lane state machine passes q.o / q.d; the wavefront kernels keep the strict
// mode's ray in shared memory while the lane walks inner nodes (nt_wavefront.cuh) and hand it in from there.
template <typename R, typename K>
__device__ __forceinline__ void query_leaf_step(const Ctx<R, true> &c, BvhQuery<R> &q, const V3<R> &o, const V3<R> &d, const int2 *stack,
                                                K &k) {
    const int code = -2 - q.cur;
    const int first = code & 0x3ffffff, count = ((code >> 26) & 3) + 1;
    const bool is_tri = (code >> 28) & 1;
    for (int j = 0; j < count; ++j) {
        const int idx = first + j;
        R t;
        bool hit;
        if (is_tri) { R p[9]; c.ld_tri(idx, p); k.tri++; hit = hit_triangle<R>(p, o, d, c.eps, t); }
        else { R p[4]; c.ld_sph(idx, p); k.sph++; hit = hit_sphere<R>(p, o, d, c.eps, t); }
        if (!hit) continue;
        if (q.any) {
            if (t < q.tb) { q.found = true; q.done = true; q.cur = NT_REF_EMPTY; return; }
        } else {
            const int kind = is_tri ? 2 : 0;
            if (t < q.tb) {
                q.tb = t; q.best.kind = kind; q.best.idx = idx;
                q.best.gid = __ldg((is_tri ? c.s->tri_gid : c.s->sph_gid) + idx);
                q.tmaxf = Math<R>::up(q.tb - q.tshift); q.found = true;
            } else if (t == q.tb) { // SPEC §3 tie-break: the smallest global primitive id wins
                const int gid = __ldg((is_tri ? c.s->tri_gid : c.s->sph_gid) + idx);
                if (gid < q.best.gid) { q.best.kind = kind; q.best.idx = idx; q.best.gid = gid; q.found = true; }
            }
        }
    }
    query_pop(q, stack);
}

// ---- per-lane ray-tree state machine (scalars only; the deferred-children stack is a separate array) ----
template <typename R> struct Lane {
    V3<R> d;         // direction of the tree node being shaded (its origin and hit point live in the query)
    V3<R> N;
    R W, ndl, local[3], acc[3];
    unsigned depth, sid;
    int phase;       // 0: nearest-hit query in flight; 1 + l: occlusion query of light l in flight
    int mat, sp;
    bool active, entering;
};
template <typename R> struct ChildStack {
    R v[NT_MAX_DEPTH_DEV][7];
    unsigned depth[NT_MAX_DEPTH_DEV];
};

// Sample finished: hand its radiance to the resolve pass.
template <typename R>
__device__ __forceinline__ void lane_finish(Lane<R> &ln, R *samples) {
    R *dst = samples + 3 * (size_t)ln.sid;
    dst[0] = ln.acc[0]; dst[1] = ln.acc[1]; dst[2] = ln.acc[2];
    ln.active = false;
}

// Node finished: next ray of the tree (reflection child first, deferred transmission children after),
// or the sample is complete.  Starts the nearest-hit query of that ray.
template <typename R>
__device__ __forceinline__ void lane_next_ray(const Ctx<R, true> &c, Lane<R> &ln, BvhQuery<R> &q, ChildStack<R> &cs,
                                              bool descend, const V3<R> &o, R *samples, Counters &k) {
    V3<R> ro = o;
    if (!descend) {
        if (ln.sp == 0) { lane_finish<R>(ln, samples); q.done = true; return; }
        const int sp = --ln.sp;
        ro = { cs.v[sp][0], cs.v[sp][1], cs.v[sp][2] };
        ln.d = { cs.v[sp][3], cs.v[sp][4], cs.v[sp][5] };
        ln.W = cs.v[sp][6];
        ln.depth = cs.depth[sp];
    }
    ln.phase = 0;
    query_start<R>(c, q, ro, ln.d, Math<R>::inf(), false, k);
}

// SPEC §4, light loop from light `l` on: arm the occlusion query of the next light that faces the
// surface; when none is left, add the node to the sample, spawn its children and move on.
// P (the hit point) is the origin of the shadow queries, so it is passed around instead of stored.
template <typename R>
__device__ __forceinline__ void lane_lights_from(const Ctx<R, true> &c, Lane<R> &ln, BvhQuery<R> &q, ChildStack<R> &cs,
                                                 const V3<R> &P, unsigned l, R *samples, Counters &k) {
    const NtDevScene &s = *c.s;
    const NtSceneView<R> &v = *c.v;
    for (; l < s.nl; ++l) {
        const R *lp = v.lights + 6 * l;
        const V3<R> Lv = { __ldg(lp) - P.x, __ldg(lp + 1) - P.y, __ldg(lp + 2) - P.z };
        const R d2 = dot(Lv, Lv);
        const R dist = Math<R>::sqrt_(d2);
        const V3<R> L = scale(Lv, Math<R>::rcp(dist));
        const R ndl = dot(ln.N, L);
        if (!(ndl > R(0))) continue;
        k.shadow++;
        ln.ndl = ndl; ln.phase = 1 + (int)l;
        shadow_query_start<R>(c, q, P, L, dist, l, k);
        return;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) ln.acc[ch] = ln.acc[ch] + ln.W * ln.local[ch];
    bool descend = false;
    if (ln.depth < c.max_depth) {
        const R *mp = v.mat + (size_t)ln.mat * NT_MAT_STRIDE;
        R m2[4];
        Ld<R>::g4(mp + 8, m2); // kt ior inv_ior pad
        const R kr = __ldg(mp + 7), kt = m2[0];
        const V3<R> d = ln.d, N = ln.N;
        const R cosi = -dot(d, N);
        R wr = kr, wt = R(0);
        V3<R> T = { R(0), R(0), R(0) };
        if (kt > R(0)) {
            const R eta = ln.entering ? m2[2] : m2[1];
            const R kk = R(1) - (eta * eta) * (R(1) - cosi * cosi);
            if (kk < R(0)) wr = kr + kt;
            else {
                wt = kt;
                const R sterm = eta * cosi - Math<R>::sqrt_(kk);
                T = { d.x * eta + N.x * sterm, d.y * eta + N.y * sterm, d.z * eta + N.z * sterm };
                if (fast_renormalises<R>::value || (c.rules & NT_DEV_RULE_RENORMALIZE)) T = scale(T, Math<R>::rcp(Math<R>::sqrt_(dot(T, T)))); // SPEC §8
            }
        }
        if (wt > R(0)) {
            k.sec++;
            if (wr > R(0)) { // defer: the reflection subtree comes first in pre-order
                const int sp = ln.sp;
                cs.v[sp][0] = P.x; cs.v[sp][1] = P.y; cs.v[sp][2] = P.z;
                cs.v[sp][3] = T.x; cs.v[sp][4] = T.y; cs.v[sp][5] = T.z;
                cs.v[sp][6] = ln.W * wt; cs.depth[sp] = ln.depth + 1;
                ln.sp = sp + 1;
            }
        }
        if (wr > R(0)) {
            k.sec++;
            const R two = R(2) * cosi;
            ln.d = { d.x + N.x * two, d.y + N.y * two, d.z + N.z * two };
            if (fast_renormalises<R>::value || (c.rules & NT_DEV_RULE_RENORMALIZE)) ln.d = scale(ln.d, Math<R>::rcp(Math<R>::sqrt_(dot(ln.d, ln.d))));
            ln.W = ln.W * wr; ln.depth = ln.depth + 1;
            descend = true;
        } else if (wt > R(0)) {
            ln.d = T; ln.W = ln.W * wt; ln.depth = ln.depth + 1;
            descend = true;
        }
    }
    lane_next_ray<R>(c, ln, q, cs, descend, P, samples, k);
}

// The lane's query has finished: consume its result and arm the next one.
template <typename R>
__device__ __forceinline__ void lane_advance(const Ctx<R, true> &c, Lane<R> &ln, BvhQuery<R> &q, ChildStack<R> &cs,
                                             R *samples, Counters &k) {
    const NtDevScene &s = *c.s;
    const NtSceneView<R> &v = *c.v;
    if (ln.phase == 0) {
        if (q.best.kind < 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) ln.acc[ch] = ln.acc[ch] + ln.W * __ldg(v.globals + 3 + ch);
            lane_next_ray<R>(c, ln, q, cs, false, q.o, samples, k);
            return;
        }
        const R t = q.tb;
        const Hit h = q.best;
        const V3<R> o = q.o, d = ln.d;
        const V3<R> P = { o.x + d.x * t, o.y + d.y * t, o.z + d.z * t };
        V3<R> Ng;
        int mat;
        if (h.kind == 0) {
            R p[4];
            c.ld_sph(h.idx, p);
            const R ir = __ldg(v.sph_invr + h.idx);
            Ng = { (P.x - p[0]) * ir, (P.y - p[1]) * ir, (P.z - p[2]) * ir };
            mat = __ldg(s.sph_mat + h.idx);
        } else if (h.kind == 1) {
            R p[4];
            c.ld_pln(h.idx, p);
            Ng = { p[0], p[1], p[2] };
            mat = __ldg(s.pln_mat + h.idx);
        } else {
            const R *tp = v.tri + (size_t)h.idx * NT_TRI_STRIDE + 9;
            Ng = { __ldg(tp), __ldg(tp + 1), __ldg(tp + 2) };
            mat = __ldg(s.tri_mat + h.idx);
        }
        const R *mp = v.mat + (size_t)mat * NT_MAT_STRIDE;
        R m0[4];
        Ld<R>::g4(mp, m0); // r g b ka
        const bool entering = dot(d, Ng) < R(0);
        ln.mat = mat; ln.entering = entering;
        ln.N = entering ? Ng : V3<R>{ -Ng.x, -Ng.y, -Ng.z };
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) ln.local[ch] = __ldg(v.globals + ch) * (m0[3] * m0[ch]);
        lane_lights_from<R>(c, ln, q, cs, P, 0, samples, k);
        return;
    }
    // occlusion query of light l finished; q.o is the hit point P and q.d the unit vector to the light
    const unsigned l = (unsigned)(ln.phase - 1);
    const V3<R> P = q.o;
    if (!q.found) {
        k.light++;
        const V3<R> L = q.d;
        const R *mp = v.mat + (size_t)ln.mat * NT_MAT_STRIDE;
        R m0[4], m1[4];
        Ld<R>::g4(mp, m0);     // r g b ka
        Ld<R>::g4(mp + 4, m1); // kd ks shininess kr
        const R *lp = v.lights + 6 * l;
        R lc[3] = { __ldg(lp + 3), __ldg(lp + 4), __ldg(lp + 5) };
        if (c.rules & NT_DEV_RULE_ATTENUATE) { // SPEC §8: light colour scaled by 1 / d2 (d2 exactly as the light loop computed it)
            const V3<R> Lv = { __ldg(lp) - P.x, __ldg(lp + 1) - P.y, __ldg(lp + 2) - P.z };
            const R att = Math<R>::rcp(dot(Lv, Lv));
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) lc[ch] = lc[ch] * att;
        }
        const R kdn = m1[0] * ln.ndl;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) ln.local[ch] = ln.local[ch] + lc[ch] * (m0[ch] * kdn);
        const R two = R(2) * ln.ndl;
        const V3<R> Rv = { ln.N.x * two - L.x, ln.N.y * two - L.y, ln.N.z * two - L.z };
        const R rv = -dot(Rv, ln.d);
        if (m1[1] > R(0) && rv > R(0)) {
            const R sterm = m1[1] * Math<R>::pow_(rv, m1[2]);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) ln.local[ch] = ln.local[ch] + lc[ch] * sterm;
        }
    }
    lane_lights_from<R>(c, ln, q, cs, P, l + 1, samples, k);
}

// Sample id -> pixel and sub-sample (the inverse is in resolve_kernel).  sid = (tile*rounds + r)*32 + slot,
// slot = pixel-in-tile * lanes + j, sample index s = r*lanes + j.
struct SampleMap {
    unsigned px, vr, y, sidx;
    bool live;
};
__device__ __forceinline__ SampleMap map_sample(const NtRenderArgs &a, unsigned sid) {
    SampleMap m;
    const unsigned rounds = a.spp / a.lanes;
    const unsigned tile = sid / (rounds * 32), rem = sid % (rounds * 32);
    const unsigned r = rem / 32, slot = rem % 32, pw = slot / a.lanes, j = slot % a.lanes;
    m.px = (tile % a.tiles_x) * a.twx + pw % a.twx;
    m.vr = (tile / a.tiles_x) * a.twy + pw / a.twx;
    m.live = m.px < a.width && m.vr < a.vrows;
    m.y = ((m.vr / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + m.vr % a.band_rows;
    m.sidx = r * a.lanes + j;
    return m;
}

// Persistent-warp trace kernel for BVH scenes: writes one radiance triple per sample.
template <typename R>
__global__ void __launch_bounds__(NT_BLOCK_THREADS, NT_MIN_BLOCKS_BVH)
render_bvh_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtRenderArgs a) {
    __shared__ unsigned long long s_cnt[NT_NCOUNTERS];
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, true> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.eps_lo = (R)a.eps_lo; c.max_depth = a.max_depth; c.rules = a.rules;
    stage_scene<R, true>(s, v, c);

    const unsigned lane = threadIdx.x & 31;
    const unsigned n_sids = a.tiles_x * a.tiles_y * (a.spp / a.lanes) * 32;
    unsigned long long *next_sid = a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS;
    R *samples = (R *)a.samples;
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };

    Lane<R> ln;
    BvhQuery<R> q;
    ChildStack<R> cs;
    int2 bstack[NT_BVH_STACK];
    ln.active = false; ln.sp = 0; ln.phase = 0;
    q.done = true; q.cur = NT_REF_EMPTY; q.sp = 0; q.found = false; q.any = false;
    bool exhausted = false;

    for (;;) {
        // ---- refill: idle lanes claim new samples, one atomic per warp ----
        const unsigned idle = __ballot_sync(0xffffffffu, !ln.active);
        if (!exhausted && (__popc(idle) >= NT_REFILL_THRESHOLD || idle == 0xffffffffu)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(next_sid, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(idle) >= n_sids) exhausted = true;
            if (!ln.active) {
                const unsigned long long sid = base + __popc(idle & ((1u << lane) - 1));
                if (sid < n_sids) {
                    const SampleMap m = map_sample(a, (unsigned)sid);
                    if (m.live) {
                        // SPEC §2: regular n x n grid
                        const unsigned si = m.sidx % a.n, sj = m.sidx / a.n;
                        const R ox = ArgsView<R>::samp_off(a, si), oy = ArgsView<R>::samp_off(a, sj); // (i + 0.5) / n (or i / n, SPEC §8), divided on the host
                        const R fx = (R)m.px + ox, fy = (R)m.y + oy;
                        const V3<R> D = { (ArgsView<R>::cam(a, 3) + ArgsView<R>::cam(a, 6) * fx) + ArgsView<R>::cam(a, 9) * fy,
                                          (ArgsView<R>::cam(a, 4) + ArgsView<R>::cam(a, 7) * fx) + ArgsView<R>::cam(a, 10) * fy,
                                          (ArgsView<R>::cam(a, 5) + ArgsView<R>::cam(a, 8) * fx) + ArgsView<R>::cam(a, 11) * fy };
                        const V3<R> eye = { ArgsView<R>::cam(a, 0), ArgsView<R>::cam(a, 1), ArgsView<R>::cam(a, 2) };
                        ln.d = scale(D, Math<R>::rcp(Math<R>::sqrt_(dot(D, D))));
                        ln.sid = (unsigned)sid; ln.W = R(1); ln.depth = 1; ln.sp = 0; ln.phase = 0; ln.active = true;
                        ln.acc[0] = ln.acc[1] = ln.acc[2] = R(0);
                        k.prim++;
                        primary_query_start<R>(c, q, eye, ln.d, k);
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, ln.active) == 0) {
            if (exhausted) break;
            continue;
        }
        // ---- (A) traversal rounds: descend until (nearly) every traversing lane holds a leaf, then
        //      test the leaves together — the exact primitive tests are the expensive part and must
        //      not run with a handful of lanes ----
        for (;;) {
            for (;;) {
                const bool inner = ln.active && !q.done && q.cur >= 0;
                const unsigned im = __ballot_sync(0xffffffffu, inner);
                if (im == 0) break;
                if (inner) query_inner_step<R>(c, q, bstack, k);
                if (__popc(im) < NT_DESCEND_MIN &&
                    __ballot_sync(0xffffffffu, ln.active && !q.done && ref_is_leaf(q.cur)) != 0) break;
            }
            const bool leaf = ln.active && !q.done && ref_is_leaf(q.cur);
            if (leaf) query_leaf_step<R>(c, q, q.o, q.d, bstack, k);
            if (ln.active) query_second_pass<R>(c, q, q.o, q.d); // drifted directions: the sphere set after the triangle set
            const unsigned parked = __ballot_sync(0xffffffffu, !ln.active || q.done);
            if (parked == 0xffffffffu || __popc(parked) >= NT_ADVANCE_THRESHOLD) break;
        }
        // ---- (B) advance the lanes whose query is complete ----
        if (ln.active && q.done) lane_advance<R>(c, ln, q, cs, samples, k);
    }
    flush_counters(k, a.counters, s_cnt);
}

// Resolve: SPEC §5 — the samples of a pixel are added in sample order, scaled, quantised to RGBA8.
template <typename R>
__global__ void __launch_bounds__(256)
resolve_kernel(const __grid_constant__ NtRenderArgs a) {
    const unsigned px = blockIdx.x * blockDim.x + threadIdx.x, vr = blockIdx.y;
    if (px >= a.width || vr >= a.vrows) return;
    const R *samples = (const R *)a.samples;
    const unsigned rounds = a.spp / a.lanes;
    const unsigned tile = (vr / a.twy) * a.tiles_x + px / a.twx;
    const unsigned pw = (vr % a.twy) * a.twx + px % a.twx;
    R sum[3] = { R(0), R(0), R(0) };
    for (unsigned r = 0; r < rounds; ++r)
        for (unsigned j = 0; j < a.lanes; ++j) {
            const R *sp = samples + 3 * ((size_t)(tile * rounds + r) * 32 + pw * a.lanes + j);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) sum[ch] = sum[ch] + sp[ch];
        }
    const R inv_spp = Math<R>::rcp((R)a.spp);
    unsigned rgba = 0xff000000u;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const R cv = sum[ch] * inv_spp;
        unsigned qv = cv <= R(0) ? 0u : cv >= R(1) ? 255u : (unsigned)(int)(cv * R(255) + R(0.5));
        if ((a.rules & NT_DEV_RULE_TRUNCATE) && cv > R(0) && cv < R(1)) qv = (unsigned)(int)(cv * R(255)); // SPEC §8
        rgba |= qv << (8 * ch);
    }
    const unsigned y = ((vr / a.band_rows) * a.shard_count + a.shard_index) * a.band_rows + vr % a.band_rows;
    const size_t row = a.layout == 1 ? vr : y;
    *(unsigned *)(a.out + row * a.stride + 4 * (size_t)px) = rgba;
}

// Unit-level entry for BVH scenes: nearest hit of arbitrary rays (nt_trace_rays).
template <typename R>
__global__ void __launch_bounds__(NT_BLOCK_THREADS)
trace_bvh_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtTraceArgs a) {
    const NtSceneView<R> &v = *(const NtSceneView<R> *)(sizeof(R) == 8 ? (const void *)&s.v64 : (const void *)&s.v32);
    Ctx<R, true> c;
    c.s = &s; c.v = &v; c.eps = (R)a.eps; c.eps_lo = (R)a.eps_lo; c.max_depth = 1; c.rules = 0;
    stage_scene<R, true>(s, v, c);
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const V3<R> o = { (R)a.origins[3 * i], (R)a.origins[3 * i + 1], (R)a.origins[3 * i + 2] };
    const V3<R> d = { (R)a.dirs[3 * i], (R)a.dirs[3 * i + 1], (R)a.dirs[3 * i + 2] };
    Counters k = { 0, 0, 0, 0, 0, 0, 0, 0 };
    BvhQuery<R> q;
    int2 bstack[NT_BVH_STACK];
    query_start<R>(c, q, o, d, Math<R>::inf(), false, k);
    for (;;) {
        while (!q.done) {
            if (q.cur >= 0) query_inner_step<R>(c, q, bstack, k);
            else query_leaf_step<R>(c, q, q.o, q.d, bstack, k);
        }
        if (q.pass == 1 && !(q.any && q.found)) query_arm<R>(c, q, q.o, q.d, 2); // one ray per thread, no warp to share a sweep with
        if (q.done) break;
    }
    a.t_out[i] = q.best.kind >= 0 ? (double)q.tb : -1.0;
    a.prim_out[i] = q.best.kind >= 0 ? q.best.gid : -1;
}

} // namespace nt
