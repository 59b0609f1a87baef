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
//   prect per render call (it depends on the camera): for every bounded primitive the rectangle of pixels outside
//         which no primary ray can touch its bounding sphere (nt_cull_primary_rects).  A warp tile keeps the
//         primitives whose rectangle it overlaps.
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

// Primary rays D(x, y) = p00 + x dx + y dy from `eye` (cam = eye p00 dx dy, SPEC-PROVISIONAL §2; pixel px holds
// the sample positions x in (px, px + 1)).  rects[j] = { x0, x1, y0, y1 }: every pixel whose samples can touch
// the ball j (radius dilated to r * 1.001 + 1e-3 |c - eye| + margin) lies in [x0, x1] x [y0, y1]; an empty
// rectangle has x0 > x1.  Exact conic bounds in binary64, widened by two pixels; the whole image whenever the
// ball's outline on the image plane is not an ellipse (eye inside the ball, ball crossing the eye's plane).
void nt_cull_primary_rects(const double *bsph, uint32_t nb, const double cam[12], uint32_t width, uint32_t height,
                           double margin, uint16_t *rects);

// Bit l of the result: no plane can stop a strict-mode shadow query (ray epsilon >= eps_min) that starts on a
// bounded primitive and aims at light l.  Per plane: the light is clearly off the plane (s_L = n.L - d), and no point
// of the primitives' exact bounding box is farther than `allowed` on the other side, allowed = 1e-3 eps_min s_L /
// dist_max (dist_max = largest distance from the light to the box).  A point P on the light's side gives a crossing
// parameter t < 0 or t >= dist (1 + 1e-6); a point within `allowed` beyond the plane (a sphere resting on the floor:
// its lowest point, give or take rounding) gives t = dist (-s_P) / (s_L - s_P) <= 1e-3 eps_min: a miss by t > eps.
// All margins are many orders above the rounding errors of the exact rule.  planes [np][4] = nx ny nz d with
// n.x = d on the plane (SPEC-PROVISIONAL §1).  At most 32 lights.
uint32_t nt_cull_plane_free_lights(const double *spheres, uint32_t ns, const double *triangles, uint32_t nt, const double *planes,
                                   uint32_t np, const double *lights, uint32_t nl, double eps_min);

// "Light rooms" (flat scenes, strict and fast mode): per point light the axis-aligned box between the nearest
// axis-aligned planes (normal exactly +-e_k) on either side of the light.  rooms[l] = { lo_x, hi_x, lo_y, hi_y, lo_z,
// hi_z, cap_per_eps, cap_max }.  A shadow query whose origin P satisfies lo_k <= P_k <= hi_k on all three axes and
// whose light distance satisfies dist <= min(eps * cap_per_eps, cap_max) cannot be stopped by ANY axis-aligned plane,
// so the kernel skips the axis lists (general planes are still tested).  Proof, on the doubles the exact rule itself
// works with (SPEC-PROVISIONAL section 3 in its axis form t = fl(fl(p - P_k) / L_k), L_k = fl(fl(l_k - P_k) * fl(1 / dist))),
// for a plane below the light (p < l_k, s = l_k - p; mirrored above), u = 2^-53:
//   P_k > p:  num < 0, so t > 0 needs L_k < 0, i.e. P_k > l_k; then num / L_k = dist (P_k - p) / (P_k - l_k) (1 + 5u') >=
//             dist (1 + s / dist)(1 - 6u) and s / dist >= s / cap_max = 1e-12: t >= dist, not an occluder;
//   P_k == p: num = 0, t = 0: a miss;
//   p - delta <= P_k < p (lo_k = p - delta): 0 < num <= delta, t > 0 needs L_k > 0, L_k >= (s / dist)(1 - 3u), hence
//             t <= delta dist / s (1 + 6u) <= delta eps cap_per_eps / s_min (1 + 6u) = eps / 2 (1 + 6u) < eps: a miss.
// delta = 2^-33 x (largest |plane position| or |light coordinate|): five orders above the rounding of a hit point
// that lies on a wall, so the test passes for (nearly) every query inside the room; s_min = the light's smallest
// distance to an axis-aligned plane; a light ON such a plane gets an empty room (lo = +inf).  The fast mode uses the
// float-rounded box with delta32 = 2^-16 x scale and no distance cap (it has no bit-level contract, SPEC section 7).
void nt_cull_light_rooms(const double *planes, uint32_t np, const double *lights, uint32_t nl, double *rooms, float *rooms32);
