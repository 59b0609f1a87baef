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

// The lines through the eye that touch the ball (centre C, radius rho) satisfy (V.D)^2 >= k |D|^2 with V = C - eye,
// k = |V|^2 - rho^2 > 0.  With D = p00 + x dx + y dy this is F(x, y) = a x^2 + 2 b x y + c y^2 + 2 d x + 2 e y + f >= 0,
// M = g g^T - k G (g = A^T V, G = A^T A, A = [dx dy p00]).  When the quadratic part is negative definite the set
// is the inside of an ellipse; for a fixed x, F has a real root in y iff (b x + e)^2 - c (a x^2 + 2 d x + f) >= 0,
// i.e. -D2 x^2 + 2 q x + r >= 0 (D2 = a c - b^2 > 0, q = b e - c d, r = e^2 - c f): x between (q -+ sqrt(q^2 + D2 r)) / D2.
// Lines, not half-lines: a ball behind the eye keeps its rectangle (a superset is all that is needed).
void nt_cull_primary_rects(const double *bsph, uint32_t nb, const double cam[12], uint32_t width, uint32_t height,
                           double margin, uint16_t *rects) {
    const double *E = cam, *P0 = cam + 3, *DX = cam + 6, *DY = cam + 9;
    auto dot3 = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    const double gxx = dot3(DX, DX), gxy = dot3(DX, DY), gx0 = dot3(DX, P0), gyy = dot3(DY, DY), gy0 = dot3(DY, P0), g00 = dot3(P0, P0);
    const uint16_t xmax = (uint16_t)std::min<uint32_t>(width - 1, 65535u), ymax = (uint16_t)std::min<uint32_t>(height - 1, 65535u);
    for (uint32_t j = 0; j < nb; ++j) {
        uint16_t *o = rects + 4 * (size_t)j;
        o[0] = 0; o[1] = xmax; o[2] = 0; o[3] = ymax; // the whole image unless proven smaller
        const double *bs = bsph + 4 * (size_t)j;
        const double V[3] = { bs[0] - E[0], bs[1] - E[1], bs[2] - E[2] };
        const double v2 = dot3(V, V), rho = bs[3] * 1.001 + 1e-3 * std::sqrt(v2) + margin;
        const double k = v2 - rho * rho;
        if (!(k > 0) || !std::isfinite(k)) continue;
        const double vx = dot3(V, DX), vy = dot3(V, DY), v0 = dot3(V, P0);
        const double a = vx * vx - k * gxx, b = vx * vy - k * gxy, c = vy * vy - k * gyy;
        const double d = vx * v0 - k * gx0, e = vy * v0 - k * gy0, f = v0 * v0 - k * g00;
        const double D2 = a * c - b * b;
        if (!(a < 0 && c < 0 && D2 > 1e-9 * (a * c))) continue; // not (safely) an ellipse
        const double qx = b * e - c * d, rx = e * e - c * f, qy = b * d - a * e, ry = d * d - a * f;
        const double sx = std::sqrt(std::max(0.0, qx * qx + D2 * rx)), sy = std::sqrt(std::max(0.0, qy * qy + D2 * ry));
        const double x_lo = (qx - sx) / D2, x_hi = (qx + sx) / D2, y_lo = (qy - sy) / D2, y_hi = (qy + sy) / D2;
        if (!std::isfinite(x_lo + x_hi + y_lo + y_hi)) continue;
        // pixel px holds x in (px, px + 1); two pixels of slack on every side
        const double px_lo = std::floor(x_lo) - 2, px_hi = std::floor(x_hi) + 2, py_lo = std::floor(y_lo) - 2, py_hi = std::floor(y_hi) + 2;
        if (px_hi < 0 || py_hi < 0 || px_lo > xmax || py_lo > ymax) { o[0] = 1; o[1] = 0; o[2] = 1; o[3] = 0; continue; } // off the image
        o[0] = (uint16_t)std::max(0.0, px_lo); o[1] = (uint16_t)std::min((double)xmax, px_hi);
        o[2] = (uint16_t)std::max(0.0, py_lo); o[3] = (uint16_t)std::min((double)ymax, py_hi);
    }
}

uint32_t nt_cull_plane_free_lights(const double *spheres, uint32_t ns, const double *triangles, uint32_t nt, const double *planes,
                                   uint32_t np, const double *lights, uint32_t nl, double eps_min) {
    if (ns + nt == 0 || nl == 0 || nl > 32) return 0;
    // exact bounding box of the bounded primitives (c -+ r is one rounding; triangle vertices are exact)
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY }, ext = 0;
    for (uint32_t j = 0; j < ns; ++j)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], spheres[4 * (size_t)j + a] - spheres[4 * (size_t)j + 3]);
            hi[a] = std::max(hi[a], spheres[4 * (size_t)j + a] + spheres[4 * (size_t)j + 3]);
        }
    for (size_t i = 0; i < 9 * (size_t)nt; ++i) {
        lo[i % 3] = std::min(lo[i % 3], triangles[i]);
        hi[i % 3] = std::max(hi[i % 3], triangles[i]);
    }
    for (int a = 0; a < 3; ++a) ext = std::max(ext, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
    uint32_t out = 0;
    for (uint32_t l = 0; l < nl; ++l) {
        const double *L = lights + 6 * (size_t)l;
        double dmax2 = 0, lext = ext;
        for (int a = 0; a < 3; ++a) {
            const double m = std::max(std::fabs(L[a] - lo[a]), std::fabs(L[a] - hi[a]));
            dmax2 += m * m;
            lext = std::max(lext, std::fabs(L[a]));
        }
        const double dist_max = std::sqrt(dmax2);
        bool free_ = std::isfinite(dist_max) && dist_max > 0;
        for (uint32_t i = 0; i < np && free_; ++i) {
            const double *p = planes + 4 * (size_t)i;
            const double len = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            double sl = (p[0] * L[0] + p[1] * L[1] + p[2] * L[2]) - p[3];
            double smin = -p[3], smax = -p[3]; // range of n.x - d over the box
            for (int a = 0; a < 3; ++a) {
                smin += std::min(p[a] * lo[a], p[a] * hi[a]);
                smax += std::max(p[a] * lo[a], p[a] * hi[a]);
            }
            if (sl < 0) { sl = -sl; const double t = -smin; smin = -smax; smax = t; } // light below: mirror
            const double round = 1e-12 * len * (lext + std::fabs(p[3]) / (len > 0 ? len : 1.0)); // >> rounding of hit points and of n.P - d
            const double allowed = 1e-3 * eps_min * sl / dist_max - 2.0 * round;
            // light clearly off the plane; no point of the box farther than `allowed` on the other side
            free_ = len > 0 && sl > 1e-6 * (std::fabs(smax) + std::fabs(smin) + len * lext) &&
                    (smin >= 2.0 * round || (allowed > 2.0 * round && smin >= -allowed));
        }
        if (free_) out |= 1u << l;
    }
    return out;
}

void nt_cull_light_rooms(const double *planes, uint32_t np, const double *lights, uint32_t nl, double *rooms, float *rooms32) {
    const double inf = HUGE_VAL;
    for (uint32_t l = 0; l < nl; ++l) {
        const double *L = lights + 6 * (size_t)l;
        double lo[3] = { -inf, -inf, -inf }, hi[3] = { inf, inf, inf }, scale = 0, smin = inf;
        bool ok = true;
        for (int a = 0; a < 3; ++a) scale = std::max(scale, std::fabs(L[a]));
        for (uint32_t i = 0; i < np; ++i) {
            const double *p = planes + 4 * (size_t)i;
            for (int k = 0; k < 3; ++k) {
                if (!(std::fabs(p[k]) == 1.0 && p[(k + 1) % 3] == 0.0 && p[(k + 2) % 3] == 0.0)) continue;
                const double pk = p[k] * p[3]; // exact: the plane is x_k = pk
                scale = std::max(scale, std::fabs(pk));
                if (!std::isfinite(pk) || pk == L[k]) { ok = false; continue; }
                smin = std::min(smin, std::fabs(L[k] - pk));
                if (pk < L[k]) lo[k] = std::max(lo[k], pk); else hi[k] = std::min(hi[k], pk);
            }
        }
        double *r = rooms + 8 * (size_t)l;
        float *rf = rooms32 + 8 * (size_t)l;
        if (!(scale > 0) || !std::isfinite(scale)) scale = 1.0;
        const double delta = std::ldexp(scale, -33), delta32 = std::ldexp(scale, -16);
        ok = ok && std::isfinite(L[0]) && std::isfinite(L[1]) && std::isfinite(L[2]);
        if (ok && smin < 1e3 * delta) ok = false; // light (nearly) on a plane: no room
        for (int k = 0; k < 3; ++k) {
            r[2 * k] = ok ? lo[k] - delta : inf;
            r[2 * k + 1] = ok ? hi[k] + delta : -inf;
            // float box: rounded INWARD from the dilated double box (the dilation stays >= delta32 / 2)
            rf[2 * k] = ok ? std::nextafterf((float)(lo[k] - delta32), HUGE_VALF) : HUGE_VALF;
            rf[2 * k + 1] = ok ? std::nextafterf((float)(hi[k] + delta32), -HUGE_VALF) : -HUGE_VALF;
        }
        const bool any = ok && std::isfinite(smin); // no axis-aligned plane at all: nothing to prove, no cap
        r[6] = !ok ? 0.0 : any ? 0.5 * smin / delta : inf;
        r[7] = !ok ? 0.0 : any ? smin * 1e12 : inf;
        rf[6] = ok ? HUGE_VALF : 0.0f;
        rf[7] = ok ? HUGE_VALF : 0.0f;
    }
}
