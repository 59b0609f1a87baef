// nt_shadowgrid.cpp - see nt_shadowgrid.h
#include "nt_shadowgrid.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {
struct Rect { int u0, u1, v0, v1; };
}

uint32_t nt_shadow_grid_k0(uint32_t ns) {
    double mul = 4.0; // cells per axis ~ mul x sqrt(number of spheres); NT_GRID_KMUL: A/B
    if (const char *e = getenv("NT_GRID_KMUL")) { const double v = atof(e); if (v >= 0.5 && v <= 64.0) mul = v; }
    uint32_t k = 64;
    while (k < 1024 && (double)k < mul * std::sqrt((double)ns)) k *= 2;
    return k;
}

int nt_shadow_grids_build(const double *sph, uint32_t ns, const double *lights, uint32_t nl, double max_abs,
                          std::vector<NtShadowGrid> &grids, std::vector<uint32_t> &off, std::vector<uint32_t> &items) {
    grids.assign(nl, NtShadowGrid{});
    off.clear(); items.clear();
    if (ns == 0) return 0;
    double lo[3] = { HUGE_VAL, HUGE_VAL, HUGE_VAL }, hi[3] = { -HUGE_VAL, -HUGE_VAL, -HUGE_VAL };
    for (uint32_t i = 0; i < ns; ++i)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], sph[4 * (size_t)i + a]); hi[a] = std::max(hi[a], sph[4 * (size_t)i + a]); }
    const uint32_t K0 = nt_shadow_grid_k0(ns);
    int n_valid = 0;
    std::vector<double> box(4 * (size_t)ns);
    std::vector<Rect> rect(ns);
    for (uint32_t l = 0; l < nl; ++l) {
        NtShadowGrid &g = grids[l];
        const double *Lp = lights + 6 * (size_t)l;
        double ax[3] = { 0.5 * (lo[0] + hi[0]) - Lp[0], 0.5 * (lo[1] + hi[1]) - Lp[1], 0.5 * (lo[2] + hi[2]) - Lp[2] };
        const double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        if (!(an > 1e-9 * (1.0 + max_abs)) || !std::isfinite(an)) continue;
        for (double &x : ax) x /= an;
        // the device works with the binary32 roundings of the basis: build the lists with exactly those vectors (they need
        // not be orthonormal to the last bit - any three independent vectors define a projection, and the containment
        // argument of the header only needs U, V, a fixed)
        int least = std::fabs(ax[0]) <= std::fabs(ax[1]) ? (std::fabs(ax[0]) <= std::fabs(ax[2]) ? 0 : 2) : (std::fabs(ax[1]) <= std::fabs(ax[2]) ? 1 : 2);
        double e[3] = { 0, 0, 0 };
        e[least] = 1.0;
        double U[3] = { ax[1] * e[2] - ax[2] * e[1], ax[2] * e[0] - ax[0] * e[2], ax[0] * e[1] - ax[1] * e[0] };
        const double un = std::sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
        for (double &x : U) x /= un;
        double V[3] = { ax[1] * U[2] - ax[2] * U[1], ax[2] * U[0] - ax[0] * U[2], ax[0] * U[1] - ax[1] * U[0] };
        float fa[3], fU[3], fV[3], fL[3];
        for (int k = 0; k < 3; ++k) { fa[k] = (float)ax[k]; fU[k] = (float)U[k]; fV[k] = (float)V[k]; fL[k] = (float)Lp[k]; }
        bool ok = true;
        double umin = HUGE_VAL, umax = -HUGE_VAL, vmin = HUGE_VAL, vmax = -HUGE_VAL;
        // the light position is rounded to binary32 on the device as well: that moves every projection by up to
        // ulp(|L|) / w - covered by the dilation of rho below (1e-6 x the largest coordinate >= 8 ulp of it)
        const double scale = std::max(max_abs, std::max(std::fabs(Lp[0]), std::max(std::fabs(Lp[1]), std::fabs(Lp[2]))));
        for (uint32_t i = 0; i < ns && ok; ++i) {
            const double *s = sph + 4 * (size_t)i;
            const double cv[3] = { s[0] - (double)fL[0], s[1] - (double)fL[1], s[2] - (double)fL[2] };
            const double rho = std::sqrt(s[3]) * (1.0 + 1e-6) + 1e-6 * scale;
            const double w = cv[0] * fa[0] + cv[1] * fa[1] + cv[2] * fa[2];
            const double x = cv[0] * fU[0] + cv[1] * fU[1] + cv[2] * fU[2], y = cv[0] * fV[0] + cv[1] * fV[1] + cv[2] * fV[2];
            if (!(w > 1.5 * rho)) { ok = false; break; }
            const double hu = std::hypot(x, w), hv = std::hypot(y, w);
            const double au = std::asin(std::min(1.0, rho / hu)), av = std::asin(std::min(1.0, rho / hv));
            const double tu = std::atan2(x, w), tv = std::atan2(y, w);
            if (std::fabs(tu) + au > 1.45 || std::fabs(tv) + av > 1.45) { ok = false; break; }
            double *b = &box[4 * (size_t)i];
            b[0] = std::tan(tu - au); b[1] = std::tan(tu + au); b[2] = std::tan(tv - av); b[3] = std::tan(tv + av);
            umin = std::min(umin, b[0]); umax = std::max(umax, b[1]); vmin = std::min(vmin, b[2]); vmax = std::max(vmax, b[3]);
        }
        if (!ok) continue;
        const double du = std::max(umax - umin, 1e-9), dv = std::max(vmax - vmin, 1e-9);
        // a cell must stay far wider than the binary32 error of a projected point (~1e-6 (1 + |u|), |u| <= tan 1.45 = 8.2):
        // a distant cluster of spheres gets a coarser grid
        uint32_t K = K0;
        while (K > 1 && std::min(du, dv) / (double)K < 2e-3) K /= 2;
        g.u0 = (float)umin; g.v0 = (float)vmin;
        g.su = (float)((double)K / du); g.sv = (float)((double)K / dv);
        g.K = K; g.base = (uint32_t)off.size();
        for (int k = 0; k < 3; ++k) { g.L[k] = fL[k]; g.axis[k] = fa[k]; g.U[k] = fU[k]; g.V[k] = fV[k]; }
        // cells with the device's own binary32 origin and scale
        auto cell = [&](double t, float t0, float sc) { return (int)std::floor((t - (double)t0) * (double)sc); };
        std::vector<uint32_t> count((size_t)K * K + 1, 0u);
        for (uint32_t i = 0; i < ns; ++i) {
            const double *b = &box[4 * (size_t)i];
            // slack for the device's binary32 projection of a point (error ~1e-6 (1 + |u|)): 1e-4 (1 + |u|), a twentieth of a
            // cell at most (a whole cell on every side listed 25 cells for a 3 x 3 footprint instead of 16)
            const double eu = 1e-4 * (1.0 + std::max(std::fabs(b[0]), std::fabs(b[1]))), ev = 1e-4 * (1.0 + std::max(std::fabs(b[2]), std::fabs(b[3])));
            Rect r = { cell(b[0] - eu, g.u0, g.su), cell(b[1] + eu, g.u0, g.su), cell(b[2] - ev, g.v0, g.sv), cell(b[3] + ev, g.v0, g.sv) };
            r.u0 = std::max(r.u0, 0); r.v0 = std::max(r.v0, 0); r.u1 = std::min(r.u1, (int)K - 1); r.v1 = std::min(r.v1, (int)K - 1);
            rect[i] = r;
            for (int v = r.v0; v <= r.v1; ++v)
                for (int u = r.u0; u <= r.u1; ++u) ++count[(size_t)v * K + u];
        }
        const uint32_t first = (uint32_t)items.size();
        {   // items are addressed with 32 bits: a light whose lists would not fit gets no grid
            unsigned long long total = first;
            for (size_t c = 0; c < (size_t)K * K; ++c) total += count[c];
            if (total >= (1ull << 31)) continue;
        }
        uint32_t run = first;
        for (size_t c = 0; c < (size_t)K * K; ++c) { const uint32_t n = count[c]; count[c] = run; run += n; }
        count[(size_t)K * K] = run;
        off.insert(off.end(), count.begin(), count.end());
        items.resize(run);
        std::vector<uint32_t> cur(count.begin(), count.end() - 1);
        for (uint32_t i = 0; i < ns; ++i) { // ascending sphere index inside every cell
            const Rect &r = rect[i];
            for (int v = r.v0; v <= r.v1; ++v)
                for (int u = r.u0; u <= r.u1; ++u) items[cur[(size_t)v * K + u]++] = i;
        }
        g.valid = 1;
        ++n_valid;
    }
    return n_valid;
}
