// nt_cull.h — conservative culling tables for FLAT scenes (<= 64 bounded primitives, everything staged in
// shared memory).  SPEC-PROVISIONAL §3: "any acceleration structure must return exactly this (conservative
// culling only)".  Three tables, all built once per scene on the host in binary64 with generous margins:
//   bsph  bounding sphere of every bounded primitive (bit j: spheres 0..ns-1, then triangles)
//   lbuf  light buffer (Haines & Greenberg 1986): for every point light a direction cube of K x K cells per
//         face; a cell holds the bit mask of the primitives that a segment from the light in one of the
//         cell's directions can touch.  A shadow query looks its cell up and tests only those primitives.
//   nbr   per sphere i: the primitives whose bounding ball touches ball i.  A secondary ray that starts on
//         sphere i and hits sphere i again (a chord: refraction, internal reflection) can only be stopped
//         earlier by one of those.
// The primary-ray mask (primitives whose bounding sphere meets the frustum of one warp tile) is computed
// inside the render kernel from bsph.
#pragma once
#include <cstdint>
#include <vector>

#define NT_LBUF_K 32       // cells per cube-face edge: 6*32*32 masks of 8 bytes = 48 KB per light
#define NT_LBUF_SUB 4      // host-side refinement: a cell is tested as SUB x SUB sub-cells (tighter masks)
#define NT_CULL_MAX_LIGHTS 16

struct NtCullTables {
    uint32_t k = NT_LBUF_K;
    std::vector<double> bsph;               // [ns+nt][4] cx cy cz r
    std::vector<unsigned long long> lbuf;   // [nl][6][k][k]
    std::vector<unsigned long long> nbr;    // [ns]
};

// Returns false when the scene is not eligible (no bounded primitive, more than 64, too many lights).
bool nt_cull_build(const double *spheres, uint32_t ns, const double *triangles, uint32_t nt, const double *lights,
                   uint32_t nl, NtCullTables &out);
