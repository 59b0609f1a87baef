// nt_bvh.h — host BVH builder interface (see nt_bvh.cpp).
#pragma once
#include <cstdint>
#include <vector>

#include "nt_device.h"

struct NtBvhBuild {
    std::vector<NtBvhNode> nodes;  // binary tree (intermediate), nodes[0] is the root
    std::vector<NtBvhNode4> nodes4; // collapsed 4-wide tree the device traverses, nodes4[0] is the root; the last two entries are
                                    // the per-set roots (triangles only, spheres only: nt_bvh_append_set_roots)
    std::vector<int> sph_order;    // BVH-ordered position -> original sphere index
    std::vector<int> tri_order;    // BVH-ordered position -> original triangle index
    int depth4 = 0;                 // depth of the 4-wide tree (bounds the traversal stack)
    float max_abs = 0;
    float blo[3] = { 0, 0, 0 }, bhi[3] = { 0, 0, 0 }; // union of all primitive boxes
};

void nt_bvh_set_children(NtBvhNode &n, const float *lo0, const float *hi0, int c0, int n0,
                         const float *lo1, const float *hi1, int c1, int n1);
void nt_bvh_build_top(const float *boxes6, int n, std::vector<int> &children);
void nt_bvh_append_set_roots(std::vector<NtBvhNode4> &n4, const int root_kinds[4]);
void nt_bvh_build(const double *spheres, uint32_t ns, const double *triangles, uint32_t nt,
                  int leaf_max, NtBvhBuild &out);

// On-GPU LBVH build straight into the 4-wide node format (nt_bvh_gpu.cu).  Returns 0 or a cudaError_t.
int nt_bvh_build_gpu(const double *d_spheres, uint32_t ns, const double *d_triangles, uint32_t nt, int leaf_max, void *stream,
                     NtBvhNode4 **d_nodes_out, uint32_t *n_nodes_out, std::vector<int> &sph_order, std::vector<int> &tri_order,
                     float blo[3], float bhi[3], float *max_abs, int *depth4);
