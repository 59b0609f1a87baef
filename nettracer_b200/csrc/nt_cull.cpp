// nt_cull.cpp — host builder of the flat-scene culling tables (see nt_cull.h).  Everything here only has to
// be CONSERVATIVE: a primitive missing from a mask must be impossible to hit by the rays that look the mask
// up, for the binary64 strict mode and for the binary32 fast mode (whose hit points are off by ~1e-6 of the
// scene extent).  Margins are therefore many orders of magnitude above the rounding errors involved:
//   radii       x (1 + 1e-3)  +  1e-5 * scene extent
//   angles      + 2e-3 rad    (the device picks the cell with approximate binary32 arithmetic, error ~1e-6)
#include "nt_cull.h"

#include <algorithm>
#include <cmath>

namespace {

struct CellCone {
    double d[3];   // centre direction of the cell on the canonical face (u, v, 1), normalised
    double theta;  // largest angle between the centre and a corner direction (+ margin)
};

void build_cones(int n, std::vector<CellCone> &out) {
    out.resize((size_t)n * n);
    for (int iv = 0; iv < n; ++iv)
        for (int iu = 0; iu < n; ++iu) {
            const double u0 = -1.0 + 2.0 * iu / n, u1 = -1.0 + 2.0 * (iu + 1) / n;
            const double v0 = -1.0 + 2.0 * iv / n, v1 = -1.0 + 2.0 * (iv + 1) / n;
            CellCone &c = out[(size_t)iv * n + iu];
            const double uc = 0.5 * (u0 + u1), vc = 0.5 * (v0 + v1), inv = 1.0 / std::sqrt(uc * uc + vc * vc + 1.0);
            c.d[0] = uc * inv; c.d[1] = vc * inv; c.d[2] = inv;
            double th = 0;
            const double cu[2] = { u0, u1 }, cv[2] = { v0, v1 };
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const double ic = 1.0 / std::sqrt(cu[a] * cu[a] + cv[b] * cv[b] + 1.0);
                    const double dt = (c.d[0] * cu[a] + c.d[1] * cv[b] + c.d[2]) * ic;
                    th = std::max(th, std::acos(std::min(1.0, std::max(-1.0, dt))));
                }
            c.theta = th + 1e-9;
        }
}

// canonical face direction (a, b, 1) -> world direction of cube face f = 2*axis + (negative ? 1 : 0);
// (a, b) are the two other components in increasing axis order, exactly as the device computes them
inline void face_dir(int f, const double c[3], double o[3]) {
    const int axis = f >> 1, a0 = axis == 0 ? 1 : 0, a1 = axis == 2 ? 1 : 2;
    o[axis] = (f & 1) ? -c[2] : c[2];
    o[a0] = c[0];
    o[a1] = c[1];
}

inline double angle_between(const double a[3], const double b[3]) {
    const double dt = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    return std::acos(std::min(1.0, std::max(-1.0, dt)));
}

} // namespace

bool nt_cull_build(const double *spheres, uint32_t ns, const double *triangles, uint32_t nt, const double *lights,
                   uint32_t nl, NtCullTables &out) {
    const uint32_t nb = ns + nt;
    if (nb == 0 || nb > 64 || nl > NT_CULL_MAX_LIGHTS) return false;
    const int K = NT_LBUF_K, S = NT_LBUF_SUB;
    out.k = (uint32_t)K;
    out.bsph.assign(4 * (size_t)nb, 0.0);
    double extent = 0;
    for (uint32_t i = 0; i < ns; ++i) {
        const double *s = spheres + 4 * (size_t)i;
        double *b = out.bsph.data() + 4 * (size_t)i;
        b[0] = s[0]; b[1] = s[1]; b[2] = s[2]; b[3] = s[3];
    }
    for (uint32_t i = 0; i < nt; ++i) {
        const double *t = triangles + 9 * (size_t)i;
        double *b = out.bsph.data() + 4 * (size_t)(ns + i);
        for (int a = 0; a < 3; ++a) {
            const double lo = std::min(t[a], std::min(t[3 + a], t[6 + a])), hi = std::max(t[a], std::max(t[3 + a], t[6 + a]));
            b[a] = 0.5 * (lo + hi);
        }
        double r2 = 0;
        for (int v = 0; v < 3; ++v) {
            double d2 = 0;
            for (int a = 0; a < 3; ++a) d2 += (t[3 * v + a] - b[a]) * (t[3 * v + a] - b[a]);
            r2 = std::max(r2, d2);
        }
        b[3] = std::sqrt(r2) * (1.0 + 1e-12);
    }
    for (uint32_t j = 0; j < nb; ++j) {
        const double *b = out.bsph.data() + 4 * (size_t)j;
        for (int a = 0; a < 3; ++a) extent = std::max(extent, std::fabs(b[a]) + b[3]);
    }
    for (uint32_t l = 0; l < nl; ++l)
        for (int a = 0; a < 3; ++a) extent = std::max(extent, std::fabs(lights[6 * (size_t)l + a]));
    const double abs_margin = 1e-5 * extent;

    // neighbour masks: balls that touch ball i (own bit excluded: the kernel tests the own sphere first)
    out.nbr.assign(ns, 0ull);
    for (uint32_t i = 0; i < ns; ++i) {
        const double *bi = out.bsph.data() + 4 * (size_t)i;
        for (uint32_t j = 0; j < nb; ++j) {
            if (j == i) continue;
            const double *bj = out.bsph.data() + 4 * (size_t)j;
            const double dx = bi[0] - bj[0], dy = bi[1] - bj[1], dz = bi[2] - bj[2];
            const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (!(dist > (bi[3] + bj[3]) * (1.0 + 1e-3) + 2.0 * abs_margin)) out.nbr[i] |= 1ull << j; // NaN -> kept
        }
    }

    // light buffers
    out.lbuf.assign((size_t)nl * 6 * K * K, 0ull);
    if (nl == 0) return true;
    std::vector<CellCone> coarse, fine;
    build_cones(K, coarse);
    build_cones(K * S, fine);
    for (uint32_t l = 0; l < nl; ++l) {
        const double *lp = lights + 6 * (size_t)l;
        unsigned long long *cells = out.lbuf.data() + (size_t)l * 6 * K * K;
        for (uint32_t j = 0; j < nb; ++j) {
            const double *b = out.bsph.data() + 4 * (size_t)j;
            const unsigned long long bit = 1ull << j;
            const double v[3] = { b[0] - lp[0], b[1] - lp[1], b[2] - lp[2] };
            const double dist = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            const double rj = b[3] * (1.0 + 1e-3) + abs_margin;
            if (!(dist > rj)) { // the light sits inside the (dilated) ball, or NaN: every direction can touch it
                for (size_t c = 0; c < (size_t)6 * K * K; ++c) cells[c] |= bit;
                continue;
            }
            const double alpha = std::asin(std::min(1.0, rj / dist)) + 2e-3;
            const double axis[3] = { v[0] / dist, v[1] / dist, v[2] / dist };
            for (int f = 0; f < 6; ++f)
                for (int iv = 0; iv < K; ++iv)
                    for (int iu = 0; iu < K; ++iu) {
                        const CellCone &cc = coarse[(size_t)iv * K + iu];
                        double wd[3];
                        face_dir(f, cc.d, wd);
                        const double ang = angle_between(axis, wd);
                        if (ang > alpha + cc.theta) continue;
                        bool in = ang + cc.theta <= alpha;
                        for (int sv = 0; sv < S && !in; ++sv)
                            for (int su = 0; su < S && !in; ++su) {
                                const CellCone &fc = fine[(size_t)(iv * S + sv) * (K * S) + (iu * S + su)];
                                face_dir(f, fc.d, wd);
                                in = angle_between(axis, wd) <= alpha + fc.theta;
                            }
                        if (in) cells[((size_t)f * K + iv) * K + iu] |= bit;
                    }
        }
    }
    return true;
}
