// nt_bvh_gpu.cu — on-GPU BVH build: an LBVH straight into the kernels' 4-wide node
// format.  Per primitive kind (spheres, triangles — every leaf stays single-kind, as in the host builder):
//   1. float boxes rounded outward + scene bounds (warp-reduced atomics on order-preserving ints)
//   2. 63-bit Morton keys of the box centres, radix-sorted (CUB)
//   3. leaves = runs of `leaf_max` consecutive primitives in Morton order (the BVH-ordered arrays are that order)
//   4. binary radix tree over the leaf keys (Karras 2012), boxes fitted bottom-up (one atomic per node)
//   5. collapse: every inner node at even depth becomes a 128-byte NtBvhNode4 whose slots are its
//      grandchildren (or children that are leaves); indices by prefix sum (CUB)
// The two trees hang off node 0.  Build time for 1M triangles + 10k spheres is a few ms against ~0.5 s for
// the host's binned-SAH build; the tree is of lower quality (no SAH), so traversal is slower — the host build
// stays the default and NT_BVH_BUILD=gpu selects this one (DESIGN.md §4.4 has both numbers).
// Boxes only ever cull, so any valid tree gives bit-identical images (the parity tests run with both).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <math_constants.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nt_bvh.h"
#include "nt_device.h"

namespace {

struct FBox { float lo[3], hi[3]; };

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ __forceinline__ float ord2f(int i) {
    int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// 1. boxes + bounds.  kind 0: spheres [n][4], kind 1: triangles [n][9]
__global__ void prim_boxes_kernel(const double *prims, uint32_t n, int kind, FBox *boxes, int *bounds /*lo[3] hi[3] as ordered ints*/) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    if (i < n) {
        if (kind == 0) {
            const double *s = prims + 4 * (size_t)i;
            for (int a = 0; a < 3; ++a) { lo[a] = __double2float_rd(s[a] - s[3]); hi[a] = __double2float_ru(s[a] + s[3]); }
        } else {
            const double *t = prims + 9 * (size_t)i;
            for (int a = 0; a < 3; ++a) {
                lo[a] = __double2float_rd(fmin(t[a], fmin(t[3 + a], t[6 + a])));
                hi[a] = __double2float_ru(fmax(t[a], fmax(t[3 + a], t[6 + a])));
            }
        }
        FBox b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = lo[a]; b.hi[a] = hi[a]; }
        boxes[i] = b;
    }
    for (int a = 0; a < 3; ++a) {
        int l = f2ord(lo[a]), h = f2ord(hi[a]);
        l = __reduce_min_sync(0xffffffffu, l);
        h = __reduce_max_sync(0xffffffffu, h);
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[a], l); atomicMax(&bounds[3 + a], h); }
    }
}

#ifndef NT_MORTON_CUBE_DEFAULT
#define NT_MORTON_CUBE_DEFAULT 1 // configs[3] strict: GPU-built tree 61.8 -> 58.0 ms (host SAH tree 54.9), profiles/r04_wf_sort.txt
#endif
__device__ __forceinline__ uint64_t spread21(uint64_t x) { // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// 2. Morton keys of the box centres inside the set's bounds
// `cube`: one scale for the three axes (the largest extent).  A flat set - a terrain, a slab of spheres - then keeps the top
// bits of its short axis constant and its coarse order is the 2D order of the surface, instead of being cut at every level
// by an axis along which neighbours do not differ (per-axis scaling gave that axis a third of the key).
__global__ void morton_kernel(const FBox *boxes, uint32_t n, const int *bounds, uint64_t *keys, uint32_t *vals, int cube) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
    float ext = 0.0f;
    for (int a = 0; a < 3; ++a) ext = fmaxf(ext, ord2f(bounds[3 + a]) - ord2f(bounds[a]));
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bounds[a]), hi = cube ? lo + ext : ord2f(bounds[3 + a]);
        const float c = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
        float u = hi > lo ? (c - lo) / (hi - lo) : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        const uint64_t q = (uint64_t)fminf(u * 2097152.0f, 2097151.0f);
        k |= spread21(q) << a;
    }
    keys[i] = k;
    vals[i] = i;
}

// 3. leaves: runs of leaf_max sorted primitives
__global__ void cluster_kernel(const FBox *boxes, const uint32_t *sorted_vals, const uint64_t *sorted_keys, uint32_t n, int leaf_max,
                               uint32_t m, FBox *cbox, uint64_t *ckey) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t first = j * (uint32_t)leaf_max, last = min(n, first + (uint32_t)leaf_max);
    FBox b = boxes[sorted_vals[first]];
    for (uint32_t p = first + 1; p < last; ++p) {
        const FBox o = boxes[sorted_vals[p]];
        for (int a = 0; a < 3; ++a) { b.lo[a] = fminf(b.lo[a], o.lo[a]); b.hi[a] = fmaxf(b.hi[a], o.hi[a]); }
    }
    cbox[j] = b;
    ckey[j] = sorted_keys[first];
}

__device__ __forceinline__ int delta(const uint64_t *keys, int m, int i, int j) {
    if (j < 0 || j >= m) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}

// 4a. Karras 2012: inner node i of the binary radix tree over m sorted keys.  child >= 0: inner, < 0: leaf ~child
__global__ void karras_kernel(const uint64_t *keys, int m, int2 *children, int *parent_inner, int *parent_leaf, int *range_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1) return;
    const int d = delta(keys, m, i, i + 1) - delta(keys, m, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, m, i, i - d);
    int lmax = 2;
    while (delta(keys, m, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, m, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, m, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, m, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int left = min(i, j) == gamma ? ~gamma : gamma;
    const int right = max(i, j) == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    range_lo[i] = min(i, j);
    if (left < 0) parent_leaf[~left] = i; else parent_inner[left] = i;
    if (right < 0) parent_leaf[~right] = i; else parent_inner[right] = i;
    if (i == 0) parent_inner[0] = -1;
}

// ---- 4a'. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) instead of the radix tree ----
// The clusters start as the primitives in Morton order.  Every round each cluster looks at its NT_PLOC_RADIUS neighbours
// on either side and picks the one whose union with it has the smallest surface area; mutually-nearest pairs merge into
// a new inner node (which takes the place of the left partner), the array is compacted, and the next round starts - about
// log_1.6(n) rounds.  An agglomerative build guided by surface area: close to the binned-SAH tree in traversal cost
// (DESIGN.md section 4.4), at a few milliseconds for a million triangles.  Inner node ids are handed out in DEscending order by
// prefix sums, so that the last merge - the root - is node 0 (what emit4_kernel expects) and the tree is deterministic.
#ifndef NT_PLOC_RADIUS
#define NT_PLOC_RADIUS 8 // configs[3]: 8 / 16 / 32 / 64 trace alike (73.2 / 73.2 / 73.4 / 74.7 ms with the old collapse); the search costs r
#endif
__device__ __forceinline__ float union_area(const FBox &a, const FBox &b) {
    const float dx = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]), dy = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]),
                dz = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
    return dx * dy + dy * dz + dz * dx;
}
__global__ void ploc_nearest_kernel(const FBox *cbox, int c, int *nn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const FBox me = cbox[i];
    float best = CUDART_INF_F;
    int bj = i == 0 ? 1 : i - 1;
    const int j0 = max(0, i - NT_PLOC_RADIUS), j1 = min(c - 1, i + NT_PLOC_RADIUS);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const float a = union_area(me, cbox[j]);
        // Ties (a regular mesh is full of them) are broken by a total order on PAIRS, so that the globally best pair is nearest
        // for both of its members and every round merges at least one pair: first the aligned pairs {2k, 2k+1} - a run of
        // equal areas then merges completely in one round, as a balanced tree, instead of one pair per round from its left
        // end -, then the smaller (min index, max index).
        if (a < best) { best = a; bj = j; }
        else if (a == best) {
            const bool al = (i ^ j) == 1, bal = (i ^ bj) == 1;
            if ((al && !bal) || (al == bal && (min(i, j) < min(i, bj) || (min(i, j) == min(i, bj) && max(i, j) < max(i, bj))))) bj = j;
        }
    }
    nn[i] = bj;
}
// flags: low 32 bits = the cluster survives the round (alone or as the merged pair), high 32 bits = it merges (left partner)
__global__ void ploc_flags_kernel(const int *nn, int c, unsigned long long *flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const int j = nn[i];
    const bool mutual = nn[j] == i;
    flags[i] = (mutual && i > j ? 0ull : 1ull) | (mutual && i < j ? 1ull << 32 : 0ull);
}
__global__ void ploc_apply_kernel(const int *nn, int c, const unsigned long long *flags, const unsigned long long *scan, int next_id,
                                  const FBox *cbox_in, const int *cnode_in, const int *ccount_in, FBox *cbox_out, int *cnode_out,
                                  int *ccount_out, int2 *children, int *parent_inner, int *parent_leaf, FBox *nbox, int *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const unsigned long long f = flags[i];
    if (!(f & 1ull)) return; // the right partner of a merging pair: absorbed
    const int pos = (int)(scan[i] & 0xffffffffull);
    if (f >> 32) {
        const int j = nn[i], id = next_id - (int)(scan[i] >> 32);
        const int a = cnode_in[i], b = cnode_in[j];
        FBox u;
        for (int k = 0; k < 3; ++k) { u.lo[k] = fminf(cbox_in[i].lo[k], cbox_in[j].lo[k]); u.hi[k] = fmaxf(cbox_in[i].hi[k], cbox_in[j].hi[k]); }
        const int cnt = ccount_in[i] + ccount_in[j];
        children[id] = make_int2(a, b);
        nbox[id] = u;
        count[id] = cnt;
        if (a < 0) parent_leaf[~a] = id; else parent_inner[a] = id;
        if (b < 0) parent_leaf[~b] = id; else parent_inner[b] = id;
        cbox_out[pos] = u; cnode_out[pos] = id; ccount_out[pos] = cnt;
    } else {
        cbox_out[pos] = cbox_in[i]; cnode_out[pos] = cnode_in[i]; ccount_out[pos] = ccount_in[i];
    }
}
// the clusters the host joined (run_ploc): cluster j's root learns its parent in the top tree
__global__ void ploc_top_parents_kernel(const int *cnode, const int *cparent, int c, int *parent_inner, int *parent_leaf) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c) return;
    const int node = cnode[j];
    if (node < 0) parent_leaf[~node] = cparent[j]; else parent_inner[node] = cparent[j];
}
__global__ void ploc_init_kernel(int m, int *cnode, int *ccount) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) { cnode[i] = ~i; ccount[i] = 1; }
}
// Position of every tree leaf (= primitive) in the depth-first order of the tree, and of every inner node's first leaf:
// the sum, over the ancestors below which the walk comes up from the RIGHT child, of the left sibling's leaf count.
__device__ __forceinline__ int dfs_start(int child, int node, const int2 *children, const int *count, const int *parent_inner) {
    int pos = 0;
    while (node >= 0) {
        const int2 ch = children[node];
        if (ch.y == child) pos += ch.x < 0 ? 1 : count[ch.x];
        child = node;
        node = parent_inner[node];
    }
    return pos;
}
__global__ void ploc_order_kernel(int m, const int2 *children, const int *count, const int *parent_inner, const int *parent_leaf,
                                  const uint32_t *vals_sorted, int *leaf_pos, int *range_lo, uint32_t *order_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) {
        const int pos = dfs_start(~t, parent_leaf[t], children, count, parent_inner);
        leaf_pos[t] = pos;
        order_out[pos] = vals_sorted[t];
    }
    if (t < m - 1) range_lo[t] = dfs_start(t, parent_inner[t], children, count, parent_inner);
}

// 4b. boxes bottom-up: the second thread to arrive at a node fits it and climbs on
__global__ void fit_kernel(int m, const int2 *children, const int *parent_inner, const int *parent_leaf, const FBox *cbox,
                           FBox *nbox, int *visits, int *count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    int node = parent_leaf[j];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&visits[node], 1) == 0) return;
        const int2 ch = children[node];
        const volatile FBox *a = ch.x < 0 ? &cbox[~ch.x] : &nbox[ch.x], *b = ch.y < 0 ? &cbox[~ch.y] : &nbox[ch.y];
        FBox r;
        for (int k = 0; k < 3; ++k) { r.lo[k] = fminf(a->lo[k], b->lo[k]); r.hi[k] = fmaxf(a->hi[k], b->hi[k]); }
        nbox[node] = r;
        const volatile int *vc = count;
        count[node] = (ch.x < 0 ? 1 : vc[ch.x]) + (ch.y < 0 ? 1 : vc[ch.y]);
        node = parent_inner[node];
    }
}

// 5a. depth of every inner node; even depth -> becomes a 4-wide node
__global__ void depth_kernel(int m, const int *parent_inner, const int *count, int leaf_max, uint32_t *flag, int *max_depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1) return;
    if (count[i] <= leaf_max) { flag[i] = 0; return; } // a subtree this small is one leaf (a contiguous Morton run)
    int depth = 0;
    for (int p = parent_inner[i]; p >= 0; p = parent_inner[p]) ++depth;
    flag[i] = (depth & 1) == 0 ? 1u : 0u;
    atomicMax(max_depth, depth);
}

__device__ __forceinline__ int leaf_ref(uint32_t first, uint32_t count, int kind) {
    return -2 - (int)(first | ((count - 1) << 26) | ((uint32_t)kind << 28));
}

// 5b. emit the 4-wide nodes of one set at nodes[offset + index4[i]].  A child is a leaf when it is a single
// primitive or an inner node holding <= leaf_max primitives (its range in Morton order is contiguous).
__global__ void emit4_kernel(int m, int leaf_max, int kind, const int2 *children, const FBox *cbox, const FBox *nbox,
                             const int *count, const int *range_lo, const int *leaf_pos, const uint32_t *flag, const uint32_t *index4,
                             int offset, NtBvhNode4 *nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1 || !flag[i]) return;
    NtBvhNode4 node;
    for (int k = 0; k < 4; ++k) {
        for (int a = 0; a < 3; ++a) { node.lo[a][k] = CUDART_INF_F; node.hi[a][k] = -CUDART_INF_F; }
        node.ref[k] = -1;
        node.pad[k] = 0;
    }
    int ns = 0;
    auto is_leaf = [&](int c) { return c < 0 || count[c] <= leaf_max; };
    auto put = [&](int c) {
        const FBox b = c < 0 ? cbox[~c] : nbox[c];
        for (int a = 0; a < 3; ++a) { node.lo[a][ns] = b.lo[a]; node.hi[a][ns] = b.hi[a]; }
        node.ref[ns] = c < 0 ? leaf_ref((uint32_t)(leaf_pos ? leaf_pos[~c] : ~c), 1u, kind) // PLOC: depth-first position; radix tree: Morton order
                     : count[c] <= leaf_max ? leaf_ref((uint32_t)range_lo[c], (uint32_t)count[c], kind)
                                            : offset + (int)index4[c];
        ++ns;
    };
    const int2 ch = children[i];
    const int cs[2] = { ch.x, ch.y };
    for (int q = 0; q < 2; ++q) {
        if (is_leaf(cs[q])) put(cs[q]);
        else { const int2 g = children[cs[q]]; put(g.x); put(g.y); }
    }
    nodes[offset + (int)index4[i]] = node;
}

// 5'. Collapse of a PLOC tree, as the host builder does it (nt_bvh.cpp emit4): a 4-wide node starts with the two children
// of its binary node and keeps opening the inner child with the LARGEST box until it has four slots.  (The radix-tree
// path's fixed "grandchildren" rule ignores the boxes; on configs[3] the greedy rule is what makes a PLOC tree trace like the SAH
// tree.)  Level-synchronous: queue entry q of a level is the binary node that becomes 4-wide node `base + q`; its inner
// kids are appended to the next level's queue, whose positions are their node ids.  Refs are local to the set.
__global__ void collapse_level_kernel(const int *queue, int n_queue, int base, int next_base, int *next_queue, int *next_count,
                                      int leaf_max, int kind, const int2 *children, const FBox *cbox, const FBox *nbox, const int *count,
                                      const int *range_lo, const int *leaf_pos, NtBvhNode4 *out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queue) return;
    int kids[4];
    int nk = 2;
    { const int2 ch = children[queue[q]]; kids[0] = ch.x; kids[1] = ch.y; }
    auto inner = [&](int c) { return c >= 0 && count[c] > leaf_max; };
    auto area = [&](int c) { const FBox &b = nbox[c]; const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2]; return dx * dy + dy * dz + dz * dx; };
    while (nk < 4) {
        int best = -1;
        float ba = -1.0f;
        for (int k = 0; k < nk; ++k)
            if (inner(kids[k]) && area(kids[k]) > ba) { ba = area(kids[k]); best = k; }
        if (best < 0) break;
        const int2 ch = children[kids[best]];
        for (int k = best; k + 1 < nk; ++k) kids[k] = kids[k + 1]; // erase, then append both (the host's order)
        kids[nk - 1] = ch.x; kids[nk] = ch.y;
        ++nk;
    }
    NtBvhNode4 node;
    for (int k = 0; k < 4; ++k) {
        for (int a = 0; a < 3; ++a) { node.lo[a][k] = CUDART_INF_F; node.hi[a][k] = -CUDART_INF_F; }
        node.ref[k] = -1;
        node.pad[k] = 0;
    }
    for (int k = 0; k < nk; ++k) {
        const int c = kids[k];
        const FBox b = c < 0 ? cbox[~c] : nbox[c];
        for (int a = 0; a < 3; ++a) { node.lo[a][k] = b.lo[a]; node.hi[a][k] = b.hi[a]; }
        if (c < 0) node.ref[k] = leaf_ref((uint32_t)leaf_pos[~c], 1u, kind);
        else if (count[c] <= leaf_max) node.ref[k] = leaf_ref((uint32_t)range_lo[c], (uint32_t)count[c], kind);
        else {
            const int slot = atomicAdd(next_count, 1);
            next_queue[slot] = c;
            node.ref[k] = next_base + slot;
        }
    }
    out[base + q] = node;
}
__global__ void offset_refs_kernel(const NtBvhNode4 *in, int n, int offset, NtBvhNode4 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NtBvhNode4 node = in[i];
    for (int k = 0; k < 4; ++k) if (node.ref[k] >= 0) node.ref[k] += offset;
    out[offset + i] = node;
}

struct SetResult { int ref; FBox box; int nodes4; int max_depth; };

#define CUCHK(x) do { const int e_ = (int)(x); if (e_ != 0) return e_; } while (0)

template <typename T> struct DevBuf {
    T *p = nullptr;
    int alloc(size_t n) { return (int)cudaMalloc((void **)&p, sizeof(T) * (n ? n : 1)); }
    ~DevBuf() { if (p) cudaFree(p); }
};

// Builds one set.  Pass 1 (nodes == nullptr): everything up to the node count.  The caller then allocates the
// node array and calls emit.
struct SetBuild {
    uint32_t n = 0, m = 0;
    int kind = 0, leaf_max = 4;
    DevBuf<FBox> boxes, cbox, nbox;
    DevBuf<int> bounds, parent_inner, parent_leaf, visits, max_depth, count, range_lo;
    DevBuf<uint64_t> keys, keys_sorted, ckey;
    DevBuf<uint32_t> vals, vals_sorted, flag, index4;
    DevBuf<int2> children;
    DevBuf<unsigned char> temp;
    DevBuf<int> leaf_pos;
    DevBuf<NtBvhNode4> tmp4; // PLOC: the set's 4-wide nodes with set-local refs (emit() shifts them)
    DevBuf<uint32_t> order; // PLOC: primitive ids in the depth-first order of the tree (the BVH order)
    bool ploc = true;
    int ploc_rounds = 0;
    int h_bounds[6];
    int nodes4 = 0, depth = 0;
    const uint32_t *order_ptr() const { return ploc && m >= 2 ? order.p : vals_sorted.p; }

    // PLOC rounds (see ploc_nearest_kernel); fills children / parents / nbox / count, root = inner node 0
    int run_ploc(cudaStream_t st) {
        const int T = 256;
        DevBuf<FBox> cb[2];
        DevBuf<int> cn[2], cc[2], nn;
        DevBuf<unsigned long long> flags, scan;
        DevBuf<unsigned char> stemp;
        for (int k = 0; k < 2; ++k) { CUCHK(cb[k].alloc(m)); CUCHK(cn[k].alloc(m)); CUCHK(cc[k].alloc(m)); }
        CUCHK(nn.alloc(m)); CUCHK(flags.alloc(m + 1)); CUCHK(scan.alloc(m + 1));
        size_t sbytes = 0;
        CUCHK(cub::DeviceScan::ExclusiveSum(nullptr, sbytes, flags.p, scan.p, (int)m + 1, st));
        CUCHK(stemp.alloc(sbytes));
        CUCHK(cudaMemcpyAsync(cb[0].p, cbox.p, sizeof(FBox) * m, cudaMemcpyDeviceToDevice, st));
        ploc_init_kernel<<<(m + T - 1) / T, T, 0, st>>>((int)m, cn[0].p, cc[0].p);
        int c = (int)m, next_id = (int)m - 2, cur = 0;
        unsigned long long *h_tot = nullptr;
        CUCHK(cudaMallocHost((void **)&h_tot, 8));
        int rc = 0;
        // The agglomerative tree is good where clusters are small and poor at the top, where a few merges decide how the
        // whole scene is cut (the 8 neighbours in Morton order are all a round ever sees): +25 % box tests per ray against the
        // binned-SAH tree.  So the rounds stop at NT_PLOC_TOP clusters and the HOST joins those with its binned-SAH
        // builder (nt_bvh_build_top: a few thousand boxes, < 1 ms) - the top twelve levels by the surface-area heuristic,
        // everything below by PLOC.  NT_PLOC_TOP=1 builds the whole tree by PLOC (A/B).
        int top = 1024; // configs[3]: 256 -> 68.5 ms, 1024 -> 66.1, 4096 -> 68.1, 16384 -> 70.7 (pure PLOC 73.3, host SAH 58.2)
        if (const char *e = getenv("NT_PLOC_TOP")) top = atoi(e) > 0 ? atoi(e) : 1;
        if ((int)m < 8 * top) top = 1;
        while (c > top && rc == 0) {
            ploc_nearest_kernel<<<(c + T - 1) / T, T, 0, st>>>(cb[cur].p, c, nn.p);
            ploc_flags_kernel<<<(c + T - 1) / T, T, 0, st>>>(nn.p, c, flags.p);
            rc = (int)cudaMemsetAsync(flags.p + c, 0, 8, st); // the scan's last element = the totals
            if (!rc) rc = (int)cub::DeviceScan::ExclusiveSum(stemp.p, sbytes, flags.p, scan.p, c + 1, st);
            if (rc) break;
            ploc_apply_kernel<<<(c + T - 1) / T, T, 0, st>>>(nn.p, c, flags.p, scan.p, next_id, cb[cur].p, cn[cur].p, cc[cur].p, cb[cur ^ 1].p,
                                                            cn[cur ^ 1].p, cc[cur ^ 1].p, children.p, parent_inner.p, parent_leaf.p, nbox.p, count.p);
            rc = (int)cudaMemcpyAsync(h_tot, scan.p + c, 8, cudaMemcpyDeviceToHost, st);
            if (!rc) rc = (int)cudaStreamSynchronize(st);
            if (rc) break;
            const int kept = (int)(*h_tot & 0xffffffffull), merged = (int)(*h_tot >> 32);
            if (merged == 0 || kept != c - merged) { rc = (int)cudaErrorUnknown; break; } // cannot happen: the closest pair is always mutual
            next_id -= merged;
            c = kept;
            cur ^= 1;
            ++ploc_rounds;
        }
        cudaFreeHost(h_tot);
        if (rc) return rc;
        if (c > 1) { // join the remaining clusters on the host: inner nodes 0 .. c - 2 (exactly the ids the rounds have left)
            std::vector<FBox> hb((size_t)c);
            std::vector<int> hn((size_t)c), hc((size_t)c), top_children;
            CUCHK(cudaMemcpyAsync(hb.data(), cb[cur].p, sizeof(FBox) * (size_t)c, cudaMemcpyDeviceToHost, st));
            CUCHK(cudaMemcpyAsync(hn.data(), cn[cur].p, sizeof(int) * (size_t)c, cudaMemcpyDeviceToHost, st));
            CUCHK(cudaMemcpyAsync(hc.data(), cc[cur].p, sizeof(int) * (size_t)c, cudaMemcpyDeviceToHost, st));
            CUCHK(cudaStreamSynchronize(st));
            nt_bvh_build_top(&hb[0].lo[0], c, top_children);
            const int nt_ = c - 1;
            if ((int)top_children.size() != 2 * nt_) return (int)cudaErrorUnknown;
            std::vector<int2> t_children((size_t)nt_);
            std::vector<FBox> t_box((size_t)nt_);
            std::vector<int> t_count((size_t)nt_), t_parent((size_t)nt_, -1), c_parent((size_t)c, -1);
            for (int t = nt_ - 1; t >= 0; --t) { // parents come before children in nt_bvh_build_top's numbering
                FBox u;
                int cnt = 0, ref[2];
                for (int k = 0; k < 2; ++k) {
                    const int ch = top_children[2 * (size_t)t + k];
                    const FBox &b = ch >= 0 ? t_box[(size_t)ch] : hb[(size_t)~ch];
                    cnt += ch >= 0 ? t_count[(size_t)ch] : hc[(size_t)~ch];
                    ref[k] = ch >= 0 ? ch : hn[(size_t)~ch];
                    if (ch >= 0) t_parent[(size_t)ch] = t; else c_parent[(size_t)~ch] = t;
                    if (k == 0) u = b;
                    else for (int a = 0; a < 3; ++a) { u.lo[a] = fminf(u.lo[a], b.lo[a]); u.hi[a] = fmaxf(u.hi[a], b.hi[a]); }
                }
                t_children[(size_t)t] = make_int2(ref[0], ref[1]);
                t_box[(size_t)t] = u;
                t_count[(size_t)t] = cnt;
            }
            CUCHK(cudaMemcpyAsync(children.p, t_children.data(), sizeof(int2) * (size_t)nt_, cudaMemcpyHostToDevice, st));
            CUCHK(cudaMemcpyAsync(nbox.p, t_box.data(), sizeof(FBox) * (size_t)nt_, cudaMemcpyHostToDevice, st));
            CUCHK(cudaMemcpyAsync(count.p, t_count.data(), sizeof(int) * (size_t)nt_, cudaMemcpyHostToDevice, st));
            CUCHK(cudaMemcpyAsync(parent_inner.p, t_parent.data(), sizeof(int) * (size_t)nt_, cudaMemcpyHostToDevice, st));
            CUCHK(cudaMemcpyAsync(nn.p, c_parent.data(), sizeof(int) * (size_t)c, cudaMemcpyHostToDevice, st)); // nn: scratch
            ploc_top_parents_kernel<<<(c + T - 1) / T, T, 0, st>>>(cn[cur].p, nn.p, c, parent_inner.p, parent_leaf.p);
            CUCHK(cudaStreamSynchronize(st)); // the host vectors above are the copies' sources
        }
        const int root_parent = -1;
        CUCHK(cudaMemcpyAsync(parent_inner.p, &root_parent, 4, cudaMemcpyHostToDevice, st));
        CUCHK(leaf_pos.alloc(m)); CUCHK(order.alloc(m));
        ploc_order_kernel<<<(m + T - 1) / T, T, 0, st>>>((int)m, children.p, count.p, parent_inner.p, parent_leaf.p, vals_sorted.p, leaf_pos.p,
                                                        range_lo.p, order.p);
        return (int)cudaGetLastError();
    }

    int run(const double *d_prims, uint32_t n_, int kind_, int leaf_max_, cudaStream_t st) {
        n = n_; kind = kind_; leaf_max = leaf_max_;
        if (const char *e = getenv("NT_BVH_GPU_ALGO")) ploc = strcmp(e, "lbvh") != 0; // "lbvh": the Morton radix tree of round 1
        int morton_cube = NT_MORTON_CUBE_DEFAULT;
        if (const char *e = getenv("NT_MORTON_CUBE")) morton_cube = e[0] != '0';
        if (n == 0) return 0;
        m = n; // one radix-tree leaf per primitive; subtrees of <= leaf_max primitives become the BVH leaves
        const int T = 256;
        CUCHK(boxes.alloc(n)); CUCHK(bounds.alloc(6)); CUCHK(keys.alloc(n)); CUCHK(keys_sorted.alloc(n));
        CUCHK(vals.alloc(n)); CUCHK(vals_sorted.alloc(n)); CUCHK(cbox.alloc(m)); CUCHK(ckey.alloc(m));
        const int init[6] = { INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN };
        CUCHK(cudaMemcpyAsync(bounds.p, init, sizeof init, cudaMemcpyHostToDevice, st));
        prim_boxes_kernel<<<(n + T - 1) / T, T, 0, st>>>(d_prims, n, kind, boxes.p, bounds.p);
        morton_kernel<<<(n + T - 1) / T, T, 0, st>>>(boxes.p, n, bounds.p, keys.p, vals.p, morton_cube);
        size_t tb = 0;
        CUCHK(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)n, 0, 63, st));
        CUCHK(temp.alloc(tb));
        CUCHK(cub::DeviceRadixSort::SortPairs(temp.p, tb, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)n, 0, 63, st));
        cluster_kernel<<<(m + T - 1) / T, T, 0, st>>>(boxes.p, vals_sorted.p, keys_sorted.p, n, 1, m, cbox.p, ckey.p);
        CUCHK(cudaMemcpyAsync(h_bounds, bounds.p, sizeof h_bounds, cudaMemcpyDeviceToHost, st));
        if (m >= 2) {
            const uint32_t ni = m - 1;
            CUCHK(children.alloc(ni)); CUCHK(parent_inner.alloc(ni)); CUCHK(parent_leaf.alloc(m)); CUCHK(visits.alloc(ni));
            CUCHK(nbox.alloc(ni)); CUCHK(flag.alloc(ni)); CUCHK(index4.alloc(ni)); CUCHK(max_depth.alloc(1));
            CUCHK(count.alloc(ni)); CUCHK(range_lo.alloc(ni));
            CUCHK(cudaMemsetAsync(visits.p, 0, sizeof(int) * ni, st));
            CUCHK(cudaMemsetAsync(max_depth.p, 0, sizeof(int), st));
            if (ploc) {
                const int prc = run_ploc(st);
                if (prc == (int)cudaErrorUnknown) ploc = false; // a round without a mutual pair (NaN boxes?): the radix tree always works
                else if (prc) return prc;
            }
            if (!ploc) {
                karras_kernel<<<(ni + T - 1) / T, T, 0, st>>>(ckey.p, (int)m, children.p, parent_inner.p, parent_leaf.p, range_lo.p);
                fit_kernel<<<(m + T - 1) / T, T, 0, st>>>((int)m, children.p, parent_inner.p, parent_leaf.p, cbox.p, nbox.p, visits.p, count.p);
            }
            if (ploc) { // greedy collapse, level by level (collapse_level_kernel)
                DevBuf<int> queue[2], qcount;
                CUCHK(queue[0].alloc(ni)); CUCHK(queue[1].alloc(ni)); CUCHK(qcount.alloc(1)); CUCHK(tmp4.alloc(ni));
                const int root = 0;
                CUCHK(cudaMemcpyAsync(queue[0].p, &root, 4, cudaMemcpyHostToDevice, st));
                int nq = 1, base = 0, cur = 0, levels = 0;
                while (nq > 0) {
                    CUCHK(cudaMemsetAsync(qcount.p, 0, 4, st));
                    collapse_level_kernel<<<(nq + T - 1) / T, T, 0, st>>>(queue[cur].p, nq, base, base + nq, queue[cur ^ 1].p, qcount.p, leaf_max, kind,
                                                                         children.p, cbox.p, nbox.p, count.p, range_lo.p, leaf_pos.p, tmp4.p);
                    int next = 0;
                    CUCHK(cudaMemcpyAsync(&next, qcount.p, 4, cudaMemcpyDeviceToHost, st));
                    CUCHK(cudaStreamSynchronize(st));
                    base += nq; nq = next; cur ^= 1; ++levels;
                }
                nodes4 = base;
                depth = 2 * levels; // the caller turns binary depth into 4-wide depth by halving
                CUCHK(cudaGetLastError());
                return 0;
            }
            depth_kernel<<<(ni + T - 1) / T, T, 0, st>>>((int)m, parent_inner.p, count.p, leaf_max, flag.p, max_depth.p);
            size_t sb = 0;
            DevBuf<unsigned char> stemp;
            CUCHK(cub::DeviceScan::ExclusiveSum(nullptr, sb, flag.p, index4.p, (int)ni, st));
            CUCHK(stemp.alloc(sb));
            CUCHK(cub::DeviceScan::ExclusiveSum(stemp.p, sb, flag.p, index4.p, (int)ni, st));
            uint32_t last_idx = 0, last_flag = 0;
            CUCHK(cudaMemcpyAsync(&last_idx, index4.p + (ni - 1), 4, cudaMemcpyDeviceToHost, st));
            CUCHK(cudaMemcpyAsync(&last_flag, flag.p + (ni - 1), 4, cudaMemcpyDeviceToHost, st));
            CUCHK(cudaMemcpyAsync(&depth, max_depth.p, 4, cudaMemcpyDeviceToHost, st));
            CUCHK(cudaStreamSynchronize(st));
            nodes4 = (int)(last_idx + last_flag);
        } else {
            CUCHK(cudaStreamSynchronize(st));
            nodes4 = 0;
        }
        CUCHK(cudaGetLastError());
        return 0;
    }

    int emit(int offset, NtBvhNode4 *nodes, cudaStream_t st, int *set_ref, FBox *set_box) {
        for (int a = 0; a < 3; ++a) { set_box->lo[a] = INFINITY; set_box->hi[a] = -INFINITY; }
        *set_ref = -1;
        if (n == 0) return 0;
        for (int a = 0; a < 3; ++a) { set_box->lo[a] = ord2f(h_bounds[a]); set_box->hi[a] = ord2f(h_bounds[3 + a]); }
        if (n <= (uint32_t)leaf_max) { // the whole set is one leaf
            *set_ref = -2 - (int)(0u | ((n - 1) << 26) | ((uint32_t)kind << 28));
            return 0;
        }
        const int T = 256;
        if (ploc) {
            offset_refs_kernel<<<(nodes4 + T - 1) / T, T, 0, st>>>(tmp4.p, nodes4, offset, nodes);
            *set_ref = offset;
            CUCHK(cudaGetLastError());
            return 0;
        }
        emit4_kernel<<<(m - 1 + T - 1) / T, T, 0, st>>>((int)m, leaf_max, kind, children.p, cbox.p, nbox.p, count.p, range_lo.p, ploc ? leaf_pos.p : nullptr, flag.p, index4.p, offset, nodes);
        *set_ref = offset; // inner node 0 is the root: depth 0, index 0 of the set
        CUCHK(cudaGetLastError());
        return 0;
    }
};

} // namespace

// Returns 0 or a cudaError_t.  d_spheres / d_triangles: the raw primitives, original order, on the device.
// *d_nodes_out is a cudaMalloc'd array the caller owns.
int nt_bvh_build_gpu(const double *d_spheres, uint32_t ns, const double *d_triangles, uint32_t nt, int leaf_max, void *stream,
                     NtBvhNode4 **d_nodes_out, uint32_t *n_nodes_out, std::vector<int> &sph_order, std::vector<int> &tri_order,
                     float blo[3], float bhi[3], float *max_abs, int *depth4) {
    cudaStream_t st = (cudaStream_t)stream;
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > NT_LEAF_MAX) leaf_max = NT_LEAF_MAX;
    SetBuild sb, tb;
    int rc;
    const bool timing = getenv("NT_BVH_TIMING") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    int leaf_sph = 1; // one sphere per leaf, as the host builder (nt_bvh.cpp): a leaf of scattered small spheres is mostly empty
    if (const char *e = getenv("NT_BVH_LEAF_SPH")) leaf_sph = atoi(e) < 1 ? 1 : atoi(e) > NT_LEAF_MAX ? NT_LEAF_MAX : atoi(e);
    if ((rc = sb.run(d_spheres, ns, 0, leaf_sph, st)) != 0) return rc;
    if ((rc = tb.run(d_triangles, nt, 1, leaf_max, st)) != 0) return rc;
    const uint32_t total = 1 + (uint32_t)sb.nodes4 + (uint32_t)tb.nodes4 + 2; // + the two per-set roots (nt_bvh.h)
    NtBvhNode4 *nodes = nullptr;
    CUCHK(cudaMalloc((void **)&nodes, sizeof(NtBvhNode4) * total));
    int sref, tref;
    FBox sbox, tbox;
    if ((rc = sb.emit(1, nodes, st, &sref, &sbox)) != 0) { cudaFree(nodes); return rc; }
    if ((rc = tb.emit(1 + sb.nodes4, nodes, st, &tref, &tbox)) != 0) { cudaFree(nodes); return rc; }
    NtBvhNode4 root;
    for (int k = 0; k < 4; ++k) {
        const FBox *b = k == 0 ? &sbox : k == 1 ? &tbox : nullptr;
        for (int a = 0; a < 3; ++a) { root.lo[a][k] = b ? b->lo[a] : INFINITY; root.hi[a][k] = b ? b->hi[a] : -INFINITY; }
        root.ref[k] = k == 0 ? sref : k == 1 ? tref : -1;
        root.pad[k] = 0;
    }
    cudaError_t e = cudaMemcpyAsync(nodes, &root, sizeof root, cudaMemcpyHostToDevice, st);
    NtBvhNode4 set_roots[2] = { root, root }; // [total - 2] triangles only (slot 1), [total - 1] spheres only (slot 0)
    for (int set = 0; set < 2; ++set) {
        const int drop = set == 0 ? 0 : 1;
        for (int a = 0; a < 3; ++a) { set_roots[set].lo[a][drop] = INFINITY; set_roots[set].hi[a][drop] = -INFINITY; }
        set_roots[set].ref[drop] = -1;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(nodes + total - 2, set_roots, sizeof set_roots, cudaMemcpyHostToDevice, st);
    sph_order.resize(ns); tri_order.resize(nt);
    if (e == cudaSuccess && ns) e = cudaMemcpyAsync(sph_order.data(), sb.order_ptr(), sizeof(int) * ns, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && nt) e = cudaMemcpyAsync(tri_order.data(), tb.order_ptr(), sizeof(int) * nt, cudaMemcpyDeviceToHost, st);
    if (timing) cudaEventRecord(e1, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(nodes); return (int)e; }
    if (timing) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[nt_bvh_build_gpu] %u spheres + %u triangles -> %u 4-wide nodes: %.3f ms on the stream (boxes, Morton, sort, "
                        "%s, collapse, order download; includes the temporary cudaMallocs; PLOC rounds %d + %d)\n", ns, nt, total, ms,
                tb.ploc ? "PLOC" : "radix tree + fit", sb.ploc_rounds, tb.ploc_rounds);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    float mx = 0;
    for (int a = 0; a < 3; ++a) {
        blo[a] = fminf(sbox.lo[a], tbox.lo[a]); bhi[a] = fmaxf(sbox.hi[a], tbox.hi[a]);
        mx = fmaxf(mx, fmaxf(fabsf(blo[a]), fabsf(bhi[a])));
    }
    *max_abs = mx;
    *depth4 = 2 + (sb.depth > tb.depth ? sb.depth : tb.depth) / 2 + 1;
    *d_nodes_out = nodes;
    *n_nodes_out = total;
    return 0;
}
