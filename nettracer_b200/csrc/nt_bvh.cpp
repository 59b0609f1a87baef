// nt_bvh.cpp — host-side BVH2 builder (binned SAH, 16 bins) for the bounded primitives of a scene.
// Two trees (spheres, triangles) hang off one root so every leaf holds a single primitive kind and
// indexes a contiguous run of the BVH-ordered device array of that kind.
// Boxes are float, rounded outward: they may only ever cull (SPEC-PROVISIONAL §3 "conservative
// culling only").  The on-GPU builder is nt_bvh_gpu.cu.
#include "nt_bvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; }
    }
    void grow(const Box &b) {
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

inline float f_down(double x) {
    float f = (float)x;
    return (double)f > x ? std::nextafterf(f, -std::numeric_limits<float>::infinity()) : f;
}
inline float f_up(double x) {
    float f = (float)x;
    return (double)f < x ? std::nextafterf(f, std::numeric_limits<float>::infinity()) : f;
}

struct Ref { Box box; int c; int n; };

// Subtrees are built by separate host threads (round 1 built the million-triangle tree of configs[3] on one core in
// 425 ms - most of nt_scene_create).  Nodes come from a preallocated array through an atomic counter, so their ids
// depend on the timing, but the TREE does not (every split is a deterministic function of its primitive range), and the
// 4-wide tree the device traverses is laid out afterwards by a sequential depth-first walk (collapse4): same
// device tree, same traversal, whatever the thread count.
struct Builder {
    std::vector<NtBvhNode> &nodes; // preallocated: one inner node per primitive is an upper bound
    std::atomic<int> &next_node;
    const std::vector<Box> &boxes;
    std::vector<float> cen[3];
    std::vector<int> &order;
    int type_flag, leaf_max, grain, hw;

    Builder(std::vector<NtBvhNode> &n, std::atomic<int> &next, const std::vector<Box> &b, std::vector<int> &o, int tf, int lm)
        : nodes(n), next_node(next), boxes(b), order(o), type_flag(tf), leaf_max(lm) {
        hw = (int)std::max(1u, std::thread::hardware_concurrency());
        grain = std::max<int>(4096, (int)(b.size() / (4 * (size_t)hw))); // ranges larger than this get their own thread
        for (int a = 0; a < 3; ++a) {
            cen[a].resize(b.size());
            for (size_t i = 0; i < b.size(); ++i) cen[a][i] = 0.5f * (b[i].lo[a] + b[i].hi[a]);
        }
    }

    // The few nodes at the top of a large tree hold most primitives each: their passes over [b, e) are cut into chunks
    // worked by separate threads (partial results merged in chunk order, min / max / counts only: same result).
    bool big(int b, int e) const { return e - b > 8 * grain && hw > 1; }
    template <typename F> void chunks(int b, int e, F fn) const {
        const int n = e - b, nth = big(b, e) ? std::min(std::min(16, hw), n / (2 * grain)) : 1;
        if (nth <= 1) { fn(0, b, e); return; }
        std::vector<std::thread> th;
        for (int t = 1; t < nth; ++t) th.emplace_back([=, &fn] { fn(t, b + (int)((long long)n * t / nth), b + (int)((long long)n * (t + 1) / nth)); });
        fn(0, b, b + n / nth);
        for (auto &t : th) t.join();
    }

    Ref build(int b, int e, int depth) {
        Box bounds, cb;
        bounds.reset(); cb.reset();
        if (!big(b, e)) {
            for (int i = b; i < e; ++i) {
                const int p = order[i];
                bounds.grow(boxes[p]);
                for (int a = 0; a < 3; ++a) { cb.lo[a] = std::min(cb.lo[a], cen[a][p]); cb.hi[a] = std::max(cb.hi[a], cen[a][p]); }
            }
        } else {
            Box pb[16], pc[16];
            for (int t = 0; t < 16; ++t) { pb[t].reset(); pc[t].reset(); }
            chunks(b, e, [&](int t, int cb_, int ce_) {
                Box bb, cc;
                bb.reset(); cc.reset();
                for (int i = cb_; i < ce_; ++i) {
                    const int p = order[i];
                    bb.grow(boxes[p]);
                    for (int a = 0; a < 3; ++a) { cc.lo[a] = std::min(cc.lo[a], cen[a][p]); cc.hi[a] = std::max(cc.hi[a], cen[a][p]); }
                }
                pb[t] = bb; pc[t] = cc;
            });
            for (int t = 0; t < 16; ++t) { bounds.grow(pb[t]); cb.grow(pc[t]); }
        }
        const int count = e - b;
        if (count <= leaf_max) return Ref{ bounds, b, count | type_flag };

        int mid = -1;
        if (depth < 40) {
            constexpr int NB = 16;
            float best_cost = std::numeric_limits<float>::infinity();
            int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; ++a) {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0)) continue;
                Box bb[NB]; int bc[NB];
                for (int i = 0; i < NB; ++i) { bb[i].reset(); bc[i] = 0; }
                const float sc = NB / ext;
                if (!big(b, e)) {
                    for (int i = b; i < e; ++i) {
                        const int p = order[i];
                        int bi = (int)((cen[a][p] - cb.lo[a]) * sc);
                        bi = bi < 0 ? 0 : bi >= NB ? NB - 1 : bi;
                        bb[bi].grow(boxes[p]); bc[bi]++;
                    }
                } else {
                    struct Part { Box bb[NB]; int bc[NB]; };
                    std::vector<Part> parts(16);
                    for (Part &pt : parts) for (int i = 0; i < NB; ++i) { pt.bb[i].reset(); pt.bc[i] = 0; }
                    chunks(b, e, [&](int t, int cb_, int ce_) {
                        Part &pt = parts[(size_t)t];
                        for (int i = cb_; i < ce_; ++i) {
                            const int p = order[i];
                            int bi = (int)((cen[a][p] - cb.lo[a]) * sc);
                            bi = bi < 0 ? 0 : bi >= NB ? NB - 1 : bi;
                            pt.bb[bi].grow(boxes[p]); pt.bc[bi]++;
                        }
                    });
                    for (const Part &pt : parts) for (int i = 0; i < NB; ++i) { bb[i].grow(pt.bb[i]); bc[i] += pt.bc[i]; }
                }
                float ra[NB]; int rc[NB];
                Box acc; acc.reset(); int cnt = 0;
                for (int i = NB - 1; i > 0; --i) { acc.grow(bb[i]); cnt += bc[i]; ra[i] = acc.half_area(); rc[i] = cnt; }
                acc.reset(); cnt = 0;
                for (int i = 0; i < NB - 1; ++i) {
                    acc.grow(bb[i]); cnt += bc[i];
                    if (cnt == 0 || rc[i + 1] == 0) continue;
                    const float cost = acc.half_area() * cnt + ra[i + 1] * rc[i + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = i; }
                }
            }
            if (best_axis >= 0) {
                const int a = best_axis;
                const float sc = 16 / (cb.hi[a] - cb.lo[a]), lo = cb.lo[a];
                auto it = std::partition(order.begin() + b, order.begin() + e, [&](int p) {
                    int bi = (int)((cen[a][p] - lo) * sc);
                    bi = bi < 0 ? 0 : bi >= 16 ? 15 : bi;
                    return bi <= best_bin;
                });
                mid = (int)(it - order.begin());
                if (mid == b || mid == e) mid = -1;
            }
        }
        if (mid < 0) { // degenerate or too deep: median split on the widest centroid axis
            int a = 0;
            if (cb.hi[1] - cb.lo[1] > cb.hi[a] - cb.lo[a]) a = 1;
            if (cb.hi[2] - cb.lo[2] > cb.hi[a] - cb.lo[a]) a = 2;
            mid = b + count / 2;
            std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e,
                             [&](int x, int y) { return cen[a][x] < cen[a][y]; });
        }
        const int id = next_node.fetch_add(1);
        Ref l, r;
        if (mid - b > grain && e - mid > grain) {
            std::thread left([&] { l = build(b, mid, depth + 1); });
            r = build(mid, e, depth + 1);
            left.join();
        } else {
            l = build(b, mid, depth + 1);
            r = build(mid, e, depth + 1);
        }
        nt_bvh_set_children(nodes[id], l.box.lo, l.box.hi, l.c, l.n, r.box.lo, r.box.hi, r.c, r.n);
        return Ref{ bounds, id, 0 };
    }
};

Ref build_set(std::vector<NtBvhNode> &nodes, std::atomic<int> &next_node, const std::vector<Box> &boxes, std::vector<int> &order,
              int type_flag, int leaf_max) {
    order.resize(boxes.size());
    for (size_t i = 0; i < boxes.size(); ++i) order[i] = (int)i;
    if (boxes.empty()) {
        Ref r; r.box.reset(); r.c = 0; r.n = -1;
        return r;
    }
    Builder bld(nodes, next_node, boxes, order, type_flag, leaf_max);
    return bld.build(0, (int)boxes.size(), 0);
}

} // namespace

// Binned-SAH binary tree over n boxes, one box per leaf: the top of the GPU builder's tree (nt_bvh_gpu.cu run_ploc), whose
// clusters it joins.  children[2 t], children[2 t + 1] for inner node t of n - 1 (0 = the root, parents before children):
// a child >= 0 is an inner node, a child < 0 is box ~child.  Deterministic; < 1 ms for a few thousand boxes.
void nt_bvh_build_top(const float *boxes6, int n, std::vector<int> &children) {
    children.clear();
    if (n < 2) return;
    std::vector<Box> bx((size_t)n);
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { bx[(size_t)i].lo[a] = boxes6[6 * (size_t)i + a]; bx[(size_t)i].hi[a] = boxes6[6 * (size_t)i + 3 + a]; }
    std::vector<NtBvhNode> nodes((size_t)n + 2, NtBvhNode{});
    std::atomic<int> next_node{ 0 };
    std::vector<int> order;
    const Ref root = build_set(nodes, next_node, bx, order, 0, 1);
    // renumber breadth-first from the root (ids from the builder depend on its threads' timing; the tree does not)
    std::vector<int> queue{ root.c }, new_id((size_t)next_node.load(), -1);
    new_id[(size_t)root.c] = 0;
    for (size_t h = 0; h < queue.size(); ++h) {
        const NtBvhNode &nd = nodes[(size_t)queue[h]];
        const int cs[2] = { nd.c0, nd.c1 }, ns[2] = { nd.n0, nd.n1 };
        for (int k = 0; k < 2; ++k)
            if (ns[k] == 0) { new_id[(size_t)cs[k]] = (int)queue.size(); queue.push_back(cs[k]); }
    }
    children.assign(2 * queue.size(), 0);
    for (size_t h = 0; h < queue.size(); ++h) {
        const NtBvhNode &nd = nodes[(size_t)queue[h]];
        const int cs[2] = { nd.c0, nd.c1 }, ns[2] = { nd.n0, nd.n1 };
        for (int k = 0; k < 2; ++k)
            children[2 * h + k] = ns[k] == 0 ? new_id[(size_t)cs[k]] : ~order[(size_t)((-2 - cs[k]) & 0x3ffffff)]; // leaf ref: first | (count-1) << 26 | kind << 28
    }
}

static int make_ref(int c, int n) {
    if (n == 0) return c;          // inner node
    if (n < 0) return -1;          // empty
    return -2 - (c | (((n & 0xff) - 1) << 26) | (((n >> 8) & 1) << 28));
}

void nt_bvh_set_children(NtBvhNode &n, const float *lo0, const float *hi0, int c0, int n0,
                         const float *lo1, const float *hi1, int c1, int n1) {
    for (int a = 0; a < 3; ++a) { n.lo0[a] = lo0[a]; n.hi0[a] = hi0[a]; n.lo1[a] = lo1[a]; n.hi1[a] = hi1[a]; }
    n.c0 = make_ref(c0, n0); n.c1 = make_ref(c1, n1); n.n0 = n0; n.n1 = n1;
}

// Two more roots behind the tree: copies of node 0 that keep only the slots of one set (the others become empty slots) -
// [n - 2] triangles only, [n - 1] spheres only.  A query whose direction has drifted from unit length (SPEC section 4 does
// not re-normalise) must grow the SPHERE boxes by sqrt(|d|^2 - 1) x reach (nt_bvh_trace.cuh query_arm); triangle tests do
// not depend on |d|, so such a query walks the triangle set with plain boxes and only the sphere set with grown ones.
void nt_bvh_append_set_roots(std::vector<NtBvhNode4> &n4, const int root_kinds[4]) {
    const NtBvhNode4 root = n4[0];
    for (int set = 1; set >= 0; --set) { // triangles first
        NtBvhNode4 r = root;
        for (int k = 0; k < 4; ++k)
            if (root_kinds[k] != set) {
                for (int a = 0; a < 3; ++a) { r.lo[a][k] = std::numeric_limits<float>::infinity(); r.hi[a][k] = -std::numeric_limits<float>::infinity(); }
                r.ref[k] = -1;
            }
        n4.push_back(r);
    }
}

// ---- BVH2 -> BVH4 collapse ----
namespace {
struct Cand { Box box; int ref; int kind = -1; }; // kind: 0 sphere set, 1 triangle set (tracked for the root's slots only)

int emit4(const std::vector<NtBvhNode> &n2, std::vector<NtBvhNode4> &n4, int node2, int depth, int &max_depth, int *root_kinds = nullptr) {
    max_depth = std::max(max_depth, depth);
    std::vector<Cand> kids;
    auto children = [&](int id, Cand out[2]) {
        const NtBvhNode &n = n2[id];
        for (int a = 0; a < 3; ++a) { out[0].box.lo[a] = n.lo0[a]; out[0].box.hi[a] = n.hi0[a]; out[1].box.lo[a] = n.lo1[a]; out[1].box.hi[a] = n.hi1[a]; }
        out[0].ref = n.c0; out[1].ref = n.c1;
    };
    Cand two[2];
    children(node2, two);
    if (root_kinds) { two[0].kind = 0; two[1].kind = 1; } // binary node 0 joins the sphere tree (c0) and the triangle tree (c1)
    for (const Cand &c : two) if (c.ref != -1) kids.push_back(c);
    while (kids.size() < 4) { // open the inner child with the largest box
        int best = -1;
        float area = -1;
        for (size_t i = 0; i < kids.size(); ++i)
            if (kids[i].ref >= 0 && kids[i].box.half_area() > area) { area = kids[i].box.half_area(); best = (int)i; }
        if (best < 0) break;
        const int kind = kids[best].kind;
        children(kids[best].ref, two);
        two[0].kind = two[1].kind = kind;
        kids.erase(kids.begin() + best);
        for (const Cand &c : two) if (c.ref != -1) kids.push_back(c);
    }
    const int id = (int)n4.size();
    n4.emplace_back();
    NtBvhNode4 node;
    for (int k = 0; k < 4; ++k) {
        for (int a = 0; a < 3; ++a) {
            node.lo[a][k] = k < (int)kids.size() ? kids[k].box.lo[a] : std::numeric_limits<float>::infinity();
            node.hi[a][k] = k < (int)kids.size() ? kids[k].box.hi[a] : -std::numeric_limits<float>::infinity();
        }
        node.ref[k] = -1;
        node.pad[k] = 0;
    }
    for (size_t k = 0; k < kids.size(); ++k) node.ref[k] = kids[k].ref >= 0 ? emit4(n2, n4, kids[k].ref, depth + 1, max_depth) : kids[k].ref;
    if (root_kinds) for (int k = 0; k < 4; ++k) root_kinds[k] = k < (int)kids.size() ? kids[k].kind : -1;
    n4[id] = node;
    return id;
}

void collapse4(NtBvhBuild &out) {
    out.nodes4.clear();
    out.nodes4.reserve(out.nodes.size() / 2 + 4);
    out.depth4 = 0;
    int root_kinds[4];
    emit4(out.nodes, out.nodes4, 0, 1, out.depth4, root_kinds);
    nt_bvh_append_set_roots(out.nodes4, root_kinds);
}
} // namespace

void nt_bvh_build(const double *spheres, uint32_t ns, const double *triangles, uint32_t nt,
                  int leaf_max, NtBvhBuild &out) {
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > NT_LEAF_MAX) leaf_max = NT_LEAF_MAX;
    std::vector<Box> sb(ns), tb(nt);
    float max_abs = 0;
    for (uint32_t i = 0; i < ns; ++i) {
        const double *s = spheres + 4 * (size_t)i;
        for (int a = 0; a < 3; ++a) { sb[i].lo[a] = f_down(s[a] - s[3]); sb[i].hi[a] = f_up(s[a] + s[3]); }
    }
    {
        const unsigned nth = nt > 65536 ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
        auto work = [&](uint32_t i0, uint32_t i1) {
            for (uint32_t i = i0; i < i1; ++i) {
                const double *t = triangles + 9 * (size_t)i;
                for (int a = 0; a < 3; ++a) {
                    tb[i].lo[a] = f_down(std::min(t[a], std::min(t[3 + a], t[6 + a])));
                    tb[i].hi[a] = f_up(std::max(t[a], std::max(t[3 + a], t[6 + a])));
                }
            }
        };
        std::vector<std::thread> th;
        for (unsigned k = 1; k < nth; ++k) th.emplace_back(work, (uint32_t)((uint64_t)nt * k / nth), (uint32_t)((uint64_t)nt * (k + 1) / nth));
        work(0, nt / nth);
        for (auto &t : th) t.join();
    }
    for (const auto *v : { &sb, &tb })
        for (const Box &b : *v)
            for (int a = 0; a < 3; ++a) max_abs = std::max(max_abs, std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])));
    const auto T0 = std::chrono::steady_clock::now();
    out.nodes.assign((size_t)ns + nt + 2, NtBvhNode{}); // nodes[0]: the root that joins the two trees
    std::atomic<int> next_node{ 1 };
    Ref rs, rt;
    {
        // spheres: their own leaf size (NT_BVH_LEAF_SPH; see nt_api.cu) - a leaf of several small scattered spheres is a box
        // that is mostly empty, and every ray through it pays the exact sphere tests
        int leaf_sph = 1;
        if (const char *e = getenv("NT_BVH_LEAF_SPH")) leaf_sph = std::min(std::max(atoi(e), 1), NT_LEAF_MAX);
        std::thread spheres_thread([&] { rs = build_set(out.nodes, next_node, sb, out.sph_order, 0, leaf_sph); });
        rt = build_set(out.nodes, next_node, tb, out.tri_order, 0x100, leaf_max);
        spheres_thread.join();
    }
    out.nodes.resize((size_t)next_node.load());
    nt_bvh_set_children(out.nodes[0], rs.box.lo, rs.box.hi, rs.c, rs.n, rt.box.lo, rt.box.hi, rt.c, rt.n);
    out.max_abs = max_abs;
    const auto T1 = std::chrono::steady_clock::now();
    collapse4(out);
    if (getenv("NT_BVH_TIMES")) fprintf(stderr, "bvh: build %.1f ms collapse %.1f ms\n", std::chrono::duration<double, std::milli>(T1 - T0).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T1).count());
    Box all = rs.box;
    all.grow(rt.box);
    for (int a = 0; a < 3; ++a) { out.blo[a] = all.lo[a]; out.bhi[a] = all.hi[a]; }
}
