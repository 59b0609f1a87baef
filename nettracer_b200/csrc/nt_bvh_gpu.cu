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
__global__ void morton_kernel(const FBox *boxes, uint32_t n, const int *bounds, uint64_t *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = 0;
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
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
                             const int *count, const int *range_lo, const uint32_t *flag, const uint32_t *index4, int offset,
                             NtBvhNode4 *nodes) {
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
        node.ref[ns] = c < 0 ? leaf_ref((uint32_t)~c, 1u, kind)
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
    int h_bounds[6];
    int nodes4 = 0, depth = 0;

    int run(const double *d_prims, uint32_t n_, int kind_, int leaf_max_, cudaStream_t st) {
        n = n_; kind = kind_; leaf_max = leaf_max_;
        if (n == 0) return 0;
        m = n; // one radix-tree leaf per primitive; subtrees of <= leaf_max primitives become the BVH leaves
        const int T = 256;
        CUCHK(boxes.alloc(n)); CUCHK(bounds.alloc(6)); CUCHK(keys.alloc(n)); CUCHK(keys_sorted.alloc(n));
        CUCHK(vals.alloc(n)); CUCHK(vals_sorted.alloc(n)); CUCHK(cbox.alloc(m)); CUCHK(ckey.alloc(m));
        const int init[6] = { INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN };
        CUCHK(cudaMemcpyAsync(bounds.p, init, sizeof init, cudaMemcpyHostToDevice, st));
        prim_boxes_kernel<<<(n + T - 1) / T, T, 0, st>>>(d_prims, n, kind, boxes.p, bounds.p);
        morton_kernel<<<(n + T - 1) / T, T, 0, st>>>(boxes.p, n, bounds.p, keys.p, vals.p);
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
            karras_kernel<<<(ni + T - 1) / T, T, 0, st>>>(ckey.p, (int)m, children.p, parent_inner.p, parent_leaf.p, range_lo.p);
            fit_kernel<<<(m + T - 1) / T, T, 0, st>>>((int)m, children.p, parent_inner.p, parent_leaf.p, cbox.p, nbox.p, visits.p, count.p);
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
        emit4_kernel<<<(m - 1 + T - 1) / T, T, 0, st>>>((int)m, leaf_max, kind, children.p, cbox.p, nbox.p, count.p, range_lo.p, flag.p, index4.p, offset, nodes);
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
    if ((rc = sb.run(d_spheres, ns, 0, leaf_max, st)) != 0) return rc;
    if ((rc = tb.run(d_triangles, nt, 1, leaf_max, st)) != 0) return rc;
    const uint32_t total = 1 + (uint32_t)sb.nodes4 + (uint32_t)tb.nodes4;
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
    sph_order.resize(ns); tri_order.resize(nt);
    if (e == cudaSuccess && ns) e = cudaMemcpyAsync(sph_order.data(), sb.vals_sorted.p, sizeof(int) * ns, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && nt) e = cudaMemcpyAsync(tri_order.data(), tb.vals_sorted.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, st);
    if (timing) cudaEventRecord(e1, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(nodes); return (int)e; }
    if (timing) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[nt_bvh_build_gpu] %u spheres + %u triangles -> %u 4-wide nodes: %.3f ms on the stream (boxes, Morton, sort, "
                        "radix tree, fit, collapse, order download; includes the temporary cudaMallocs)\n", ns, nt, total, ms);
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
