// nt_shadowgrid.h - shadow grids of BVH scenes (host side): per point light, which spheres can lie between the light and
// a point seen from it in a given direction.  Conservative by construction, so that a shadow query may test the listed
// spheres with SPEC section 3's exact rule INSTEAD of walking the sphere set of the tree (nt_bvh_trace.cuh
// shadow_query_start): a sphere that is not listed in a point's cell cannot touch the segment from that point to the light.
//
// Geometry.  Light at L, axis a = unit vector from L to the centre of the spheres' bounds, (U, V) an orthonormal basis of
// the plane perpendicular to a.  A point X in front of the light (w = (X - L).a > 0) projects to (u, v) = ((X - L).U / w,
// (X - L).V / w); every point of the segment from L to X projects to the same (u, v).  The projection of a ball of centre c
// and radius rho with w_c > rho is contained in [tan(th_u - al_u), tan(th_u + al_u)] x [the same in v], th_u = atan2(x_c, w_c),
// al_u = asin(rho / hypot(x_c, w_c)) - exactly the range of X.U / X.a over the ball, because that ratio depends only on the
// ball's shadow in the (U, a) plane, a disc of radius rho around (x_c, w_c).  A ball is listed in every cell its rectangle
// touches after widening it by 1e-4 (1 + |u|) (the device computes (u, v) in binary32: error ~1e-6 (1 + |u|)), with rho = r (1 + 1e-6) + 1e-6 x the
// scene's largest coordinate.  A light gets no grid (valid = 0: its queries walk the whole tree) when a ball reaches
// behind or close to the light's plane (w_c <= 1.5 rho) or projects at more than ~83 degrees from the axis.
#pragma once
#include <cstdint>
#include <vector>
#include "nt_device.h"

// `sph`: [ns][4] cx cy cz r^2 in DEVICE order (the order of NtSceneView::sph); `lights`: [nl][6].  Fills one NtShadowGrid per
// light, and the concatenated offsets / items of all valid lights.  Returns the number of valid grids.
int nt_shadow_grids_build(const double *sph, uint32_t ns, const double *lights, uint32_t nl, double max_abs,
                          std::vector<NtShadowGrid> &grids, std::vector<uint32_t> &off, std::vector<uint32_t> &items);
// Cells per axis of a grid over ns spheres before the per-grid coarsening: a power of two in 64 .. 1024, ~4 sqrt(ns).
uint32_t nt_shadow_grid_k0(uint32_t ns);
