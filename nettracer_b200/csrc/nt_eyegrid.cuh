// nt_eyegrid.cuh - the shadow-grid construction (nt_shadowgrid.h) for PRIMARY rays, built on the device per render call.
// Every primary ray starts at the eye, so the spheres a primary ray can hit are those whose projection - as seen from the
// eye - contains the ray's direction: grid `nl` of NtDevScene::sgrid lists them per cell, a primary nearest-hit query tests
// its cell's spheres with the exact rule (keeping the nearest, ties to the smallest global id as everywhere) and walks only
// the TRIANGLE set of the tree with that bound (nt_bvh_trace.cuh primary_query_start).  Same result as the walk of the
// whole tree: the lists are conservative by the argument of nt_shadowgrid.h, with the eye in the light's place and rays
// instead of segments (every ball lies in front of the eye's plane).
//
// Eight small launches on the render's stream (~30 us), nothing on the host but the basis:
//   eg_rect    per sphere: its rectangle [tan(th -+ al)] in u and v (binary64), the bounds of all rectangles (atomics on
//              order-preserving integer images of the doubles), the validity conditions of nt_shadowgrid.h
//   eg_setup   one thread: K (halved until a cell is wide enough), origin and scale -> the grid's header
//   eg_count   per sphere: +1 in every cell of its widened rectangle
//   eg_scan1/2 exclusive scan of the K0 x K0 counters (cells beyond K x K stay 0)
//   eg_fill    per sphere: its index into every cell (the scanned offsets become the cells' ends, which are the next
//              cells' starts: offsets are stored shifted by one so that the look-up reads off[c], off[c + 1] as for lights)
//   eg_finish  one thread: the grid is valid unless a condition failed or the lists outgrew their room
// BVH renders of one scene are serialised across streams (nt_api.cu: they share the sample buffer and the wavefront
// workspace), so one grid per scene is enough.
#pragma once
#include "nt_device.h"

namespace nt {

struct NtEyeBasis { double L[3], axis[3], U[3], V[3], scale; };

__device__ __forceinline__ unsigned long long eg_ord(double x) { // order-preserving map double -> u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : b | 0x8000000000000000ull;
}
__device__ __forceinline__ double eg_unord(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? k & 0x7fffffffffffffffull : ~k;
    return __longlong_as_double((long long)b);
}
// acc: [0] umin [1] umax [2] vmin [3] vmax (ordered), [4] invalid flag, [5] total items
static __global__ void eg_rect_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtEyeBasis b) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.ns) return;
    const double *q = s.v64.sph + 4 * (size_t)i;
    // the device projects with the binary32 roundings of the eye and the basis: use exactly those
    const double cv[3] = { q[0] - (double)(float)b.L[0], q[1] - (double)(float)b.L[1], q[2] - (double)(float)b.L[2] };
    const double ax[3] = { (double)(float)b.axis[0], (double)(float)b.axis[1], (double)(float)b.axis[2] };
    const double U[3] = { (double)(float)b.U[0], (double)(float)b.U[1], (double)(float)b.U[2] };
    const double V[3] = { (double)(float)b.V[0], (double)(float)b.V[1], (double)(float)b.V[2] };
    const double rho = sqrt(q[3]) * (1.0 + 1e-6) + 1e-6 * b.scale;
    const double w = cv[0] * ax[0] + cv[1] * ax[1] + cv[2] * ax[2];
    const double x = cv[0] * U[0] + cv[1] * U[1] + cv[2] * U[2], y = cv[0] * V[0] + cv[1] * V[1] + cv[2] * V[2];
    bool ok = w > 1.5 * rho;
    double box[4] = { 0, 0, 0, 0 };
    if (ok) {
        const double hu = hypot(x, w), hv = hypot(y, w);
        const double au = asin(fmin(1.0, rho / hu)), av = asin(fmin(1.0, rho / hv));
        const double tu = atan2(x, w), tv = atan2(y, w);
        ok = fabs(tu) + au <= 1.45 && fabs(tv) + av <= 1.45;
        if (ok) {
            // 1e-12 relative: the device's tan / atan2 / asin are good to a few ulp, the rectangles are widened by 1e-4 later
            box[0] = tan(tu - au); box[1] = tan(tu + au); box[2] = tan(tv - av); box[3] = tan(tv + av);
        }
    }
    if (!ok) { atomicOr(s.eg_acc + 4, 1ull); return; }
    double *dst = s.eg_boxes + 4 * (size_t)i;
    dst[0] = box[0]; dst[1] = box[1]; dst[2] = box[2]; dst[3] = box[3];
    atomicMin(s.eg_acc + 0, eg_ord(box[0])); atomicMax(s.eg_acc + 1, eg_ord(box[1]));
    atomicMin(s.eg_acc + 2, eg_ord(box[2])); atomicMax(s.eg_acc + 3, eg_ord(box[3]));
}

static __global__ void eg_init_kernel(unsigned long long *acc) { acc[threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 2) ? ~0ull : 0ull; }

static __global__ void eg_setup_kernel(const __grid_constant__ NtDevScene s, const __grid_constant__ NtEyeBasis b) {
    NtShadowGrid *g = const_cast<NtShadowGrid *>(s.sgrid) + s.nl;
    const double umin = eg_unord(s.eg_acc[0]), umax = eg_unord(s.eg_acc[1]), vmin = eg_unord(s.eg_acc[2]), vmax = eg_unord(s.eg_acc[3]);
    const double du = fmax(umax - umin, 1e-9), dv = fmax(vmax - vmin, 1e-9);
    unsigned K = s.eg_k0;
    while (K > 1 && fmin(du, dv) / (double)K < 2e-3) K /= 2;
    for (int k = 0; k < 3; ++k) { g->L[k] = (float)b.L[k]; g->axis[k] = (float)b.axis[k]; g->U[k] = (float)b.U[k]; g->V[k] = (float)b.V[k]; }
    g->u0 = (float)umin; g->v0 = (float)vmin;
    g->su = (float)((double)K / du); g->sv = (float)((double)K / dv);
    g->K = K; g->base = 0; g->valid = 0; g->pad = 0;
}

// The cells of a sphere's widened rectangle, as nt_shadowgrid.cpp computes them.
__device__ __forceinline__ void eg_cells(const NtDevScene &s, const NtShadowGrid *g, unsigned i, int &u0, int &u1, int &v0, int &v1) {
    const double *bx = s.eg_boxes + 4 * (size_t)i;
    const double eu = 1e-4 * (1.0 + fmax(fabs(bx[0]), fabs(bx[1]))), ev = 1e-4 * (1.0 + fmax(fabs(bx[2]), fabs(bx[3])));
    const int K = (int)g->K;
    u0 = max((int)floor((bx[0] - eu - (double)g->u0) * (double)g->su), 0); u1 = min((int)floor((bx[1] + eu - (double)g->u0) * (double)g->su), K - 1);
    v0 = max((int)floor((bx[2] - ev - (double)g->v0) * (double)g->sv), 0); v1 = min((int)floor((bx[3] + ev - (double)g->v0) * (double)g->sv), K - 1);
}
template <bool FILL>
static __global__ void eg_count_fill_kernel(const __grid_constant__ NtDevScene s) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.ns || s.eg_acc[4]) return;
    const NtShadowGrid *g = s.sgrid + s.nl;
    uint32_t *off = s.eg_off + 1; // shifted by one: see the header
    uint32_t *items = s.eg_items;
    int u0, u1, v0, v1;
    eg_cells(s, g, i, u0, u1, v0, v1);
    const unsigned K = g->K;
    for (int v = v0; v <= v1; ++v)
        for (int u = u0; u <= u1; ++u) {
            const uint32_t pos = atomicAdd(off + (size_t)v * K + u, 1u);
            if (FILL && pos < s.eg_items_cap) items[pos] = i;
        }
}
// Exclusive scan of n = blocks x 1024 counters in place (off + 1 ...), two launches; the total goes to acc[5].
static __device__ __forceinline__ unsigned eg_block_scan(unsigned v, unsigned *s_warp, unsigned &total) {
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
static __global__ void __launch_bounds__(1024) eg_scan1_kernel(uint32_t *cnt, uint32_t *sums) {
    __shared__ unsigned s_warp[32];
    const unsigned i = blockIdx.x * 1024 + threadIdx.x, c = cnt[i];
    unsigned total;
    const unsigned inc = eg_block_scan(c, s_warp, total);
    cnt[i] = inc - c;
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(1024) eg_scan2_kernel(uint32_t *cnt, const uint32_t *sums, unsigned long long *acc) {
    __shared__ unsigned s_warp[32];
    unsigned total;
    eg_block_scan(threadIdx.x < blockIdx.x ? sums[threadIdx.x] : 0u, s_warp, total);
    cnt[blockIdx.x * 1024 + threadIdx.x] += total;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) acc[5] = (unsigned long long)total + sums[blockIdx.x];
}
static __global__ void eg_finish_kernel(const __grid_constant__ NtDevScene s) {
    NtShadowGrid *g = const_cast<NtShadowGrid *>(s.sgrid) + s.nl;
    g->valid = (s.eg_acc[4] == 0 && s.eg_acc[5] <= (unsigned long long)s.eg_items_cap) ? 1u : 0u;
}

// Host side of a render call: the projection's basis for this eye (as nt_shadowgrid.cpp builds a light's), then the launches.
// `eye`: NtRenderArgs::cam[0..2].  Returns a cudaError_t.
inline int launch_eye_grid(const NtDevScene &s, const double *eye, cudaStream_t st, unsigned *n_launches) {
    NtShadowGrid *g = const_cast<NtShadowGrid *>(s.sgrid) + s.nl;
    NtEyeBasis b;
    double ax[3], an = 0;
    for (int k = 0; k < 3; ++k) { b.L[k] = eye[k]; ax[k] = 0.5 * ((double)s.sph_lo[k] + (double)s.sph_hi[k]) - eye[k]; an += ax[k] * ax[k]; }
    an = sqrt(an);
    double scale = (double)s.max_abs;
    for (int k = 0; k < 3; ++k) scale = scale > fabs(eye[k]) ? scale : fabs(eye[k]);
    b.scale = scale;
    if (!(an > 1e-9 * (1.0 + scale)) || !(an < 1e300)) { // the eye in the middle of the cloud: no grid for this call
        return (int)cudaMemsetAsync(&g->valid, 0, sizeof(uint32_t), st);
    }
    for (int k = 0; k < 3; ++k) ax[k] /= an;
    const int least = fabs(ax[0]) <= fabs(ax[1]) ? (fabs(ax[0]) <= fabs(ax[2]) ? 0 : 2) : (fabs(ax[1]) <= fabs(ax[2]) ? 1 : 2);
    double e[3] = { 0, 0, 0 };
    e[least] = 1.0;
    double U[3] = { ax[1] * e[2] - ax[2] * e[1], ax[2] * e[0] - ax[0] * e[2], ax[0] * e[1] - ax[1] * e[0] };
    const double un = sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
    for (int k = 0; k < 3; ++k) U[k] /= un;
    const double V[3] = { ax[1] * U[2] - ax[2] * U[1], ax[2] * U[0] - ax[0] * U[2], ax[0] * U[1] - ax[1] * U[0] };
    for (int k = 0; k < 3; ++k) { b.axis[k] = ax[k]; b.U[k] = U[k]; b.V[k] = V[k]; }
    const size_t cells = (size_t)s.eg_k0 * s.eg_k0;
    uint32_t *off = s.eg_off;
    cudaMemsetAsync(off, 0, sizeof(uint32_t) * (cells + 2), st);
    eg_init_kernel<<<1, 8, 0, st>>>(s.eg_acc);
    const unsigned T = 256, nb = (s.ns + T - 1) / T, sb = (unsigned)(cells / 1024);
    eg_rect_kernel<<<nb, T, 0, st>>>(s, b);
    eg_setup_kernel<<<1, 1, 0, st>>>(s, b);
    eg_count_fill_kernel<false><<<nb, T, 0, st>>>(s);
    uint32_t *sums = off + cells + 2; // the scan's block totals live behind the offsets
    eg_scan1_kernel<<<sb, 1024, 0, st>>>(off + 1, sums);
    eg_scan2_kernel<<<sb, 1024, 0, st>>>(off + 1, sums, s.eg_acc);
    eg_count_fill_kernel<true><<<nb, T, 0, st>>>(s);
    eg_finish_kernel<<<1, 1, 0, st>>>(s);
    if (n_launches) *n_launches += 8;
    return (int)cudaGetLastError();
}

} // namespace nt
