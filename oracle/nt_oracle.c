/*
 * nt_oracle.c — TEST INFRASTRUCTURE ONLY.  See nt_oracle.h for the "parity unpinned" statement.
 *
 * Every function cites the SPEC-PROVISIONAL.md section it restates (there is no reference
 * file:line to cite: /root/reference/README:1-3 is the whole reference).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -pthread (see Makefile).  -ffp-contract=off is
 * REQUIRED: the spec forbids fused multiply-add; x86-64 SSE2 gives plain IEEE binary64.
 */
#include "nt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct { double x, y, z; } v3;

/* SPEC §0: dot(a,b) = (a0*b0 + a1*b1) + a2*b2 */
static inline double dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* SPEC §0: cross */
static inline v3 cross3(v3 a, v3 b) {
    v3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
static inline v3 sub3(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline v3 scale3(v3 a, double s) { v3 r = { a.x * s, a.y * s, a.z * s }; return r; }
static inline v3 ld3(const double *p) { v3 r = { p[0], p[1], p[2] }; return r; }

typedef struct { double lo[3], hi[3]; int left, right, start, count; } onode;

typedef struct {
    const nt_scene_desc *d;
    double eps;
    uint32_t max_depth;
    /* derived (SPEC §1) */
    double *sph_r2, *sph_inv_r;
    v3 *tri_e1, *tri_e2, *tri_ng;
    double *mat_inv_ior;
    /* optional BVH over spheres (ids 0..ns-1) and triangles (ids ns+np..) */
    int use_bvh;
    onode *nodes;
    int n_nodes;
    double bvh_ext; /* largest |coordinate| of a bounded primitive */
    uint32_t rules; /* NT_RULE_* bits of the render call (SPEC section 8) */
    int *bvh_prims; /* global primitive ids */
} octx;

typedef struct {
    uint64_t prim, sec, shadow, sph, pln, tri, box, light;
} ocount;

/* ---------------- SPEC §3 intersections ---------------- */

static inline int hit_sphere(const octx *c, uint32_t i, v3 o, v3 dir, double *t_out) {
    const double *s = c->d->spheres + 4 * (size_t)i;
    v3 oc = sub3(o, ld3(s));
    double b = dot3(oc, dir);
    double cc = dot3(oc, oc) - c->sph_r2[i];
    double disc = b * b - cc;
    if (disc < 0) return 0;
    double sq = sqrt(disc);
    double t = -b - sq;
    if (!(t > c->eps)) t = -b + sq;
    if (!(t > c->eps)) return 0;
    *t_out = t;
    return 1;
}

static inline int hit_plane(const octx *c, uint32_t i, v3 o, v3 dir, double *t_out) {
    const double *p = c->d->planes + 4 * (size_t)i;
    v3 n = ld3(p);
    double dn = dot3(n, dir);
    if (dn == 0) return 0;
    double t = (p[3] - dot3(n, o)) / dn;
    if (!(t > c->eps)) return 0;
    *t_out = t;
    return 1;
}

static inline int hit_triangle(const octx *c, uint32_t i, v3 o, v3 dir, double *t_out) {
    const double *tr = c->d->triangles + 9 * (size_t)i;
    v3 e1 = c->tri_e1[i], e2 = c->tri_e2[i];
    v3 p = cross3(dir, e2);
    double det = dot3(e1, p);
    if (det == 0) return 0;
    double inv = 1 / det;
    v3 tv = sub3(o, ld3(tr));
    double u = dot3(tv, p) * inv;
    if (u < 0 || u > 1) return 0;
    v3 q = cross3(tv, e1);
    double v = dot3(dir, q) * inv;
    if (v < 0 || u + v > 1) return 0;
    double t = dot3(e2, q) * inv;
    if (!(t > c->eps)) return 0;
    *t_out = t;
    return 1;
}

static inline int hit_prim(const octx *c, uint32_t gid, v3 o, v3 dir, double *t, ocount *k) {
    const nt_scene_desc *d = c->d;
    if (gid < d->n_spheres) { k->sph++; return hit_sphere(c, gid, o, dir, t); }
    gid -= d->n_spheres;
    if (gid < d->n_planes) { k->pln++; return hit_plane(c, gid, o, dir, t); }
    gid -= d->n_planes;
    k->tri++;
    return hit_triangle(c, gid, o, dir, t);
}

/* Oracle-side BVH: conservative culling only.  Boxes are inflated at build time; a box is
 * skipped only when the slab interval, widened by a relative slack, is empty or entirely beyond
 * the current bound. */
/* `grow`: SPEC section 3's sphere test takes `dir` as a unit vector but section 4 does not re-normalise reflected and
 * refracted directions, so |dir|^2 = L2 drifts from 1 along a mirror chain (1 + 1e-4 at depth 5 is common, each bounce
 * amplifies it).  The rule then accepts "hits" of rays that pass OUTSIDE the sphere: the point o + dir t of a root t lies
 * at distance sqrt(r^2 + (L2 - 1) t^2) <= r + sqrt(L2 - 1) t from the centre.  A box may only cull what the rule cannot
 * hit, so every box is grown by that bound with t <= the largest distance from the origin to any scene point
 * (ray_grow()); for unit directions it is ~1e-8 of the scene size.  (Round 1 missed this: the frame check of bench.py
 * found the oracle's BVH - and the GPU's - losing such hits at depth 4-5.) */
static inline double ray_grow(const octx *c, v3 o, v3 dir) {
    double l2 = dot3(dir, dir);
    if (!(l2 > 1)) return l2 == l2 ? 0 : INFINITY;
    double far_ = 1.7320508075688774 * (fmax(fabs(o.x), fmax(fabs(o.y), fabs(o.z))) + c->bvh_ext);
    return sqrt(l2 - 1) * far_ * (1 + 1e-9);
}
static inline int box_maybe(const onode *n, v3 o, v3 inv, double tmax, double grow) {
    double t0 = -INFINITY, t1 = INFINITY;
    const double oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
    for (int a = 0; a < 3; ++a) {
        double ta = ((n->lo[a] - grow) - oo[a]) * ii[a], tb = ((n->hi[a] + grow) - oo[a]) * ii[a];
        if (ta != ta || tb != tb) continue; /* 0*inf: origin on the slab face, axis undecided */
        if (ta > tb) { double s = ta; ta = tb; tb = s; }
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
    }
    double slack = 1e-9 * (fabs(t0) + fabs(t1)) + 1e-12;
    if (t0 == -INFINITY || t1 == INFINITY) slack = 0;
    if (t0 - slack > t1 + slack) return 0;
    if (t1 + slack < 0) return 0;
    if (t0 - slack > tmax) return 0;
    return 1;
}

/* SPEC §3 nearest hit: smallest t, ties -> smallest global id. */
static int nearest_hit(const octx *c, v3 o, v3 dir, double *t_out, ocount *k) {
    const nt_scene_desc *d = c->d;
    int best = -1;
    double tb = INFINITY, t;
    if (!c->use_bvh) {
        uint32_t n = d->n_spheres + d->n_planes + d->n_triangles;
        for (uint32_t g = 0; g < n; ++g)
            if (hit_prim(c, g, o, dir, &t, k) && t < tb) { tb = t; best = (int)g; }
    } else {
        for (uint32_t i = 0; i < d->n_planes; ++i) {
            k->pln++;
            if (hit_plane(c, i, o, dir, &t) && t < tb) { tb = t; best = (int)(d->n_spheres + i); }
        }
        if (c->n_nodes > 0) {
            v3 inv = { 1 / dir.x, 1 / dir.y, 1 / dir.z };
            const double grow = ray_grow(c, o, dir);
            int stack[128], sp = 0;
            stack[sp++] = 0;
            while (sp) {
                const onode *n = c->nodes + stack[--sp];
                k->box++;
                if (!box_maybe(n, o, inv, tb, grow)) continue;
                if (n->count) {
                    for (int j = 0; j < n->count; ++j) {
                        int g = c->bvh_prims[n->start + j];
                        if (hit_prim(c, (uint32_t)g, o, dir, &t, k) &&
                            (t < tb || (t == tb && g < best))) { tb = t; best = g; }
                    }
                } else { stack[sp++] = n->left; stack[sp++] = n->right; }
            }
        }
    }
    *t_out = tb;
    return best;
}

/* SPEC §3 occlusion: any primitive with a hit (t > eps) and t < dist. */
static int occluded(const octx *c, v3 o, v3 dir, double dist, ocount *k) {
    const nt_scene_desc *d = c->d;
    double t;
    if (!c->use_bvh) {
        uint32_t n = d->n_spheres + d->n_planes + d->n_triangles;
        for (uint32_t g = 0; g < n; ++g)
            if (hit_prim(c, g, o, dir, &t, k) && t < dist) return 1;
        return 0;
    }
    for (uint32_t i = 0; i < d->n_planes; ++i) {
        k->pln++;
        if (hit_plane(c, i, o, dir, &t) && t < dist) return 1;
    }
    if (c->n_nodes > 0) {
        v3 inv = { 1 / dir.x, 1 / dir.y, 1 / dir.z };
        const double grow = ray_grow(c, o, dir);
        int stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const onode *n = c->nodes + stack[--sp];
            k->box++;
            if (!box_maybe(n, o, inv, dist, grow)) continue;
            if (n->count) {
                for (int j = 0; j < n->count; ++j)
                    if (hit_prim(c, (uint32_t)c->bvh_prims[n->start + j], o, dir, &t, k) && t < dist)
                        return 1;
            } else { stack[sp++] = n->left; stack[sp++] = n->right; }
        }
    }
    return 0;
}

/* ---------------- SPEC §4 shading, ray tree in depth-first pre-order ---------------- */

static void trace(const octx *c, v3 o, v3 dir, double W, uint32_t depth, double acc[3], ocount *k) {
    const nt_scene_desc *d = c->d;
    double t;
    int g = nearest_hit(c, o, dir, &t, k);
    if (g < 0) {
        for (int ch = 0; ch < 3; ++ch) acc[ch] = acc[ch] + W * d->background[ch];
        return;
    }
    v3 P = { o.x + dir.x * t, o.y + dir.y * t, o.z + dir.z * t };
    v3 Ng;
    int mat;
    if ((uint32_t)g < d->n_spheres) {
        Ng = scale3(sub3(P, ld3(d->spheres + 4 * (size_t)g)), c->sph_inv_r[g]);
        mat = d->sphere_mat[g];
    } else if ((uint32_t)g < d->n_spheres + d->n_planes) {
        uint32_t i = (uint32_t)g - d->n_spheres;
        Ng = ld3(d->planes + 4 * (size_t)i);
        mat = d->plane_mat[i];
    } else {
        uint32_t i = (uint32_t)g - d->n_spheres - d->n_planes;
        Ng = c->tri_ng[i];
        mat = d->triangle_mat[i];
    }
    const double *m = d->materials + 10 * (size_t)mat;
    const double col[3] = { m[0], m[1], m[2] };
    const double ka = m[3], kd = m[4], ks = m[5], shin = m[6], kr = m[7], kt = m[8], ior = m[9];
    double cosd = dot3(dir, Ng);
    int entering = cosd < 0;
    v3 N = Ng;
    if (!entering) { N.x = -Ng.x; N.y = -Ng.y; N.z = -Ng.z; }

    double local[3];
    for (int ch = 0; ch < 3; ++ch) local[ch] = d->ambient[ch] * (ka * col[ch]);
    for (uint32_t l = 0; l < d->n_lights; ++l) {
        const double *lp = d->lights + 6 * (size_t)l;
        v3 Lv = sub3(ld3(lp), P);
        double d2 = dot3(Lv, Lv);
        double dist = sqrt(d2);
        v3 L = scale3(Lv, 1 / dist);
        double ndl = dot3(N, L);
        if (!(ndl > 0)) continue;
        k->shadow++;
        if (occluded(c, P, L, dist, k)) continue;
        k->light++;
        double kdn = kd * ndl;
        double lc[3] = { lp[3], lp[4], lp[5] };
        if (c->rules & NT_RULE_ATTENUATE_INV_SQUARE) { /* SPEC section 8 */
            double att = 1 / d2;
            for (int ch = 0; ch < 3; ++ch) lc[ch] = lc[ch] * att;
        }
        for (int ch = 0; ch < 3; ++ch) local[ch] = local[ch] + lc[ch] * (col[ch] * kdn);
        double two = 2 * ndl;
        v3 R = { N.x * two - L.x, N.y * two - L.y, N.z * two - L.z };
        double rv = -dot3(R, dir);
        if (ks > 0 && rv > 0) {
            double s = ks * pow(rv, shin);
            for (int ch = 0; ch < 3; ++ch) local[ch] = local[ch] + lc[ch] * s;
        }
    }
    for (int ch = 0; ch < 3; ++ch) acc[ch] = acc[ch] + W * local[ch];

    if (!(depth < c->max_depth)) return;
    double cosi = -dot3(dir, N);
    double wr = kr, wt = 0;
    v3 T = { 0, 0, 0 };
    if (kt > 0) {
        double eta = entering ? c->mat_inv_ior[mat] : ior;
        double kk = 1 - (eta * eta) * (1 - cosi * cosi);
        if (kk < 0) wr = kr + kt;
        else {
            wt = kt;
            double s = eta * cosi - sqrt(kk);
            T.x = dir.x * eta + N.x * s; T.y = dir.y * eta + N.y * s; T.z = dir.z * eta + N.z * s;
        }
    }
    if (wr > 0) {
        double two = 2 * cosi;
        v3 Rd = { dir.x + N.x * two, dir.y + N.y * two, dir.z + N.z * two };
        if (c->rules & NT_RULE_RENORMALIZE) Rd = scale3(Rd, 1 / sqrt(dot3(Rd, Rd))); /* SPEC section 8 */
        k->sec++;
        trace(c, P, Rd, W * wr, depth + 1, acc, k);
    }
    if (wt > 0) {
        if (c->rules & NT_RULE_RENORMALIZE) T = scale3(T, 1 / sqrt(dot3(T, T)));
        k->sec++;
        trace(c, P, T, W * wt, depth + 1, acc, k);
    }
}

/* ---------------- oracle BVH build (median split, double boxes) ---------------- */

typedef struct { double lo[3], hi[3], cen[3]; int gid; } obox;

static int g_axis;
static int cmp_cen(const void *a, const void *b) {
    double x = ((const obox *)a)->cen[g_axis], y = ((const obox *)b)->cen[g_axis];
    return (x > y) - (x < y);
}

static int build_rec(octx *c, obox *b, int start, int count, double margin) {
    int id = c->n_nodes++;
    onode *n = c->nodes + id;
    for (int a = 0; a < 3; ++a) { n->lo[a] = INFINITY; n->hi[a] = -INFINITY; }
    double clo[3] = { INFINITY, INFINITY, INFINITY }, chi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int i = start; i < start + count; ++i)
        for (int a = 0; a < 3; ++a) {
            if (b[i].lo[a] < n->lo[a]) n->lo[a] = b[i].lo[a];
            if (b[i].hi[a] > n->hi[a]) n->hi[a] = b[i].hi[a];
            if (b[i].cen[a] < clo[a]) clo[a] = b[i].cen[a];
            if (b[i].cen[a] > chi[a]) chi[a] = b[i].cen[a];
        }
    for (int a = 0; a < 3; ++a) { n->lo[a] -= margin; n->hi[a] += margin; }
    n->left = n->right = -1; n->start = start; n->count = 0;
    if (count <= 4) {
        n->count = count;
        for (int i = 0; i < count; ++i) c->bvh_prims[start + i] = b[start + i].gid;
        return id;
    }
    int ax = 0;
    if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
    if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
    g_axis = ax;
    qsort(b + start, (size_t)count, sizeof(obox), cmp_cen);
    int half = count / 2;
    int l = build_rec(c, b, start, half, margin);
    int r = build_rec(c, b, start + half, count - half, margin);
    c->nodes[id].left = l; c->nodes[id].right = r;
    return id;
}

static int build_bvh(octx *c) {
    const nt_scene_desc *d = c->d;
    int n = (int)(d->n_spheres + d->n_triangles);
    c->n_nodes = 0;
    if (n == 0) return 0;
    obox *b = (obox *)malloc(sizeof(obox) * (size_t)n);
    c->nodes = (onode *)malloc(sizeof(onode) * (size_t)(2 * n));
    c->bvh_prims = (int *)malloc(sizeof(int) * (size_t)n);
    if (!b || !c->nodes || !c->bvh_prims) { free(b); return -1; }
    double ext = 0;
    int k = 0;
    for (uint32_t i = 0; i < d->n_spheres; ++i, ++k) {
        const double *s = d->spheres + 4 * (size_t)i;
        for (int a = 0; a < 3; ++a) { b[k].lo[a] = s[a] - s[3]; b[k].hi[a] = s[a] + s[3]; b[k].cen[a] = s[a]; }
        b[k].gid = (int)i;
    }
    for (uint32_t i = 0; i < d->n_triangles; ++i, ++k) {
        const double *t = d->triangles + 9 * (size_t)i;
        for (int a = 0; a < 3; ++a) {
            double lo = fmin(t[a], fmin(t[3 + a], t[6 + a])), hi = fmax(t[a], fmax(t[3 + a], t[6 + a]));
            b[k].lo[a] = lo; b[k].hi[a] = hi; b[k].cen[a] = 0.5 * (lo + hi);
        }
        b[k].gid = (int)(d->n_spheres + d->n_planes + i);
    }
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { ext = fmax(ext, fabs(b[i].lo[a])); ext = fmax(ext, fabs(b[i].hi[a])); }
    c->bvh_ext = ext;
    build_rec(c, b, 0, n, 1e-9 * ext + 1e-300);
    free(b);
    return 0;
}

/* ---------------- context ---------------- */

static int validate(const nt_scene_desc *d) {
    if (!d || d->struct_size != sizeof(nt_scene_desc)) return NT_ERR_INVALID;
    if (d->n_materials == 0) return NT_ERR_INVALID;
    for (uint32_t i = 0; i < d->n_spheres; ++i)
        if (d->sphere_mat[i] < 0 || (uint32_t)d->sphere_mat[i] >= d->n_materials) return NT_ERR_INVALID;
    for (uint32_t i = 0; i < d->n_planes; ++i)
        if (d->plane_mat[i] < 0 || (uint32_t)d->plane_mat[i] >= d->n_materials) return NT_ERR_INVALID;
    for (uint32_t i = 0; i < d->n_triangles; ++i)
        if (d->triangle_mat[i] < 0 || (uint32_t)d->triangle_mat[i] >= d->n_materials) return NT_ERR_INVALID;
    return NT_OK;
}

static void ctx_free(octx *c) {
    free(c->sph_r2); free(c->sph_inv_r); free(c->tri_e1); free(c->tri_e2); free(c->tri_ng);
    free(c->mat_inv_ior); free(c->nodes); free(c->bvh_prims);
}

/* SPEC §1 derived quantities. */
static int ctx_init(octx *c, const nt_scene_desc *d, double eps, uint32_t max_depth, int accel) {
    memset(c, 0, sizeof *c);
    int rc = validate(d);
    if (rc) return rc;
    c->d = d;
    c->eps = eps > 0 ? eps : 1e-6;
    c->max_depth = max_depth;
    c->sph_r2 = (double *)malloc(sizeof(double) * (d->n_spheres + 1));
    c->sph_inv_r = (double *)malloc(sizeof(double) * (d->n_spheres + 1));
    c->tri_e1 = (v3 *)malloc(sizeof(v3) * (d->n_triangles + 1));
    c->tri_e2 = (v3 *)malloc(sizeof(v3) * (d->n_triangles + 1));
    c->tri_ng = (v3 *)malloc(sizeof(v3) * (d->n_triangles + 1));
    c->mat_inv_ior = (double *)malloc(sizeof(double) * (d->n_materials + 1));
    if (!c->sph_r2 || !c->sph_inv_r || !c->tri_e1 || !c->tri_e2 || !c->tri_ng || !c->mat_inv_ior) {
        ctx_free(c);
        return NT_ERR_NOMEM;
    }
    for (uint32_t i = 0; i < d->n_spheres; ++i) {
        double r = d->spheres[4 * (size_t)i + 3];
        c->sph_r2[i] = r * r;
        c->sph_inv_r[i] = 1 / r;
    }
    for (uint32_t i = 0; i < d->n_triangles; ++i) {
        const double *t = d->triangles + 9 * (size_t)i;
        v3 e1 = sub3(ld3(t + 3), ld3(t)), e2 = sub3(ld3(t + 6), ld3(t));
        v3 cr = cross3(e1, e2);
        c->tri_e1[i] = e1; c->tri_e2[i] = e2;
        c->tri_ng[i] = scale3(cr, 1 / sqrt(dot3(cr, cr)));
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) c->mat_inv_ior[i] = 1 / d->materials[10 * (size_t)i + 9];
    c->use_bvh = accel != 0;
    if (c->use_bvh && build_bvh(c)) { ctx_free(c); return NT_ERR_NOMEM; }
    return NT_OK;
}

static int isqrt_exact(uint32_t s) {
    for (uint32_t n = 1; n <= 8; ++n) if (n * n == s) return (int)n;
    return 0;
}

static uint32_t shard_rows(uint32_t h, uint32_t band, uint32_t idx, uint32_t cnt) {
    if (band == 0 || cnt == 0 || idx >= cnt) return 0;
    uint32_t nb = (h + band - 1) / band, rows = 0;
    for (uint32_t b = idx; b < nb; b += cnt) {
        uint32_t y0 = b * band, y1 = y0 + band > h ? h : y0 + band;
        rows += y1 - y0;
    }
    return rows;
}

int nto_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const octx *c;
    const nt_render_params *p;
    uint8_t *rgba_out;
    size_t stride;
    double *radiance_out;
    uint16_t *cost_out; /* optional: rays and ray-tree nodes per sample, [h][w][spp][2] (analysis only) */
    uint32_t vrows, row_step, band, scount;
    int n;
    atomic_long next; /* next virtual row to take (dynamic schedule, one row at a time) */
    pthread_mutex_t mu;
    ocount tot;
} ojob;

/* SPEC §2 (sampling) and §5 (pixel) for one image row. */
static void render_row(const ojob *jb, uint32_t vr, ocount *k) {
    const nt_render_params *p = jb->p;
    const nt_camera *cam = &p->camera;
    const int n = jb->n;
    const double inv_spp = 1.0 / (double)p->spp;
    uint32_t kb = vr / jb->band;
    uint32_t y = (kb * jb->scount + p->shard_index) * jb->band + vr % jb->band;
    for (uint32_t x = 0; x < p->width; ++x) {
        double sum[3] = { 0, 0, 0 };
        for (uint32_t s = 0; s < p->spp; ++s) {
            uint32_t i = s % (uint32_t)n, j = s / (uint32_t)n;
            const double half = (p->flags & NT_RULE_SAMPLE_CORNER) ? 0.0 : 0.5; /* SPEC section 8 */
            double ox = ((double)i + half) / (double)n, oy = ((double)j + half) / (double)n;
            double fx = (double)x + ox, fy = (double)y + oy;
            v3 D = { (cam->p00[0] + cam->dx[0] * fx) + cam->dy[0] * fy,
                     (cam->p00[1] + cam->dx[1] * fx) + cam->dy[1] * fy,
                     (cam->p00[2] + cam->dx[2] * fx) + cam->dy[2] * fy };
            v3 dir = scale3(D, 1 / sqrt(dot3(D, D)));
            double acc[3] = { 0, 0, 0 };
            const unsigned long long r0 = k->prim + k->sec + k->shadow, n0 = k->prim + k->sec;
            k->prim++;
            trace(jb->c, ld3(cam->eye), dir, 1.0, 1, acc, k);
            if (jb->cost_out) {
                uint16_t *co = jb->cost_out + (((size_t)y * p->width + x) * p->spp + s) * 2;
                const unsigned long long r = k->prim + k->sec + k->shadow - r0, nn = k->prim + k->sec - n0;
                co[0] = (uint16_t)(r > 65535 ? 65535 : r); co[1] = (uint16_t)(nn > 65535 ? 65535 : nn);
            }
            for (int ch = 0; ch < 3; ++ch) sum[ch] = sum[ch] + acc[ch];
        }
        uint8_t *px = jb->rgba_out
            ? jb->rgba_out + (size_t)(p->layout == NT_LAYOUT_COMPACT ? vr : y) * jb->stride + 4 * (size_t)x
            : NULL;
        for (int ch = 0; ch < 3; ++ch) {
            double cv = sum[ch] * inv_spp;
            if (jb->radiance_out) jb->radiance_out[((size_t)y * p->width + x) * 3 + ch] = cv;
            if (px) px[ch] = cv <= 0 ? 0 : cv >= 1 ? 255
                              : (p->flags & NT_RULE_QUANTIZE_TRUNCATE) ? (uint8_t)(int)(cv * 255) : (uint8_t)(int)(cv * 255 + 0.5);
        }
        if (px) px[3] = 255;
    }
}

static void *worker(void *arg) {
    ojob *jb = (ojob *)arg;
    ocount k = { 0 };
    for (;;) {
        long vr = atomic_fetch_add(&jb->next, (long)jb->row_step);
        if (vr >= (long)jb->vrows) break;
        render_row(jb, (uint32_t)vr, &k);
    }
    pthread_mutex_lock(&jb->mu);
    jb->tot.prim += k.prim; jb->tot.sec += k.sec; jb->tot.shadow += k.shadow; jb->tot.sph += k.sph;
    jb->tot.pln += k.pln; jb->tot.tri += k.tri; jb->tot.box += k.box; jb->tot.light += k.light;
    pthread_mutex_unlock(&jb->mu);
    return NULL;
}

static int render_impl(const nt_scene_desc *desc, const nt_render_params *p, uint8_t *rgba_out,
                       size_t stride, double *radiance_out, uint16_t *cost_out, nt_render_stats *stats, int accel,
                       int n_threads, uint32_t row_step) {
    if (!p || p->struct_size != sizeof(nt_render_params)) return NT_ERR_INVALID;
    int n = isqrt_exact(p->spp);
    if (!n || p->width == 0 || p->height == 0 || p->max_depth < 1 || p->max_depth > NT_MAX_DEPTH)
        return NT_ERR_INVALID;
    uint32_t scount = p->shard_count ? p->shard_count : 1;
    uint32_t band = p->band_rows ? p->band_rows : 1;
    if (p->shard_index >= scount) return NT_ERR_INVALID;
    octx c;
    int rc = ctx_init(&c, desc, p->ray_epsilon, p->max_depth, accel);
    if (rc) return rc;
    c.rules = p->flags & NT_RULE_MASK;
    ojob jb;
    memset(&jb, 0, sizeof jb);
    jb.c = &c; jb.p = p; jb.rgba_out = rgba_out; jb.stride = stride; jb.radiance_out = radiance_out;
    jb.cost_out = cost_out;
    jb.vrows = shard_rows(p->height, band, p->shard_index, scount);
    jb.row_step = row_step ? row_step : 1;
    jb.band = band; jb.scount = scount; jb.n = n;
    atomic_init(&jb.next, 0);
    pthread_mutex_init(&jb.mu, NULL);
    int nt = n_threads > 0 ? n_threads : nto_max_threads();
    if (nt > 1024) nt = 1024;
    if (nt <= 1) worker(&jb);
    else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
        int started = 0;
        for (int i = 0; i < nt && th; ++i)
            if (pthread_create(&th[started], NULL, worker, &jb) == 0) ++started;
        if (!started) worker(&jb);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&jb.mu);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays_primary = jb.tot.prim; stats->rays_secondary = jb.tot.sec;
        stats->rays_shadow = jb.tot.shadow; stats->sphere_tests = jb.tot.sph;
        stats->plane_tests = jb.tot.pln; stats->triangle_tests = jb.tot.tri;
        stats->box_tests = jb.tot.box; stats->light_evals = jb.tot.light;
    }
    ctx_free(&c);
    return NT_OK;
}

int nto_render_radiance(const nt_scene_desc *desc, const nt_render_params *p, uint8_t *rgba_out,
                        size_t stride, double *radiance_out, nt_render_stats *stats, int accel,
                        int n_threads, uint32_t row_step) {
    return render_impl(desc, p, rgba_out, stride, radiance_out, NULL, stats, accel, n_threads, row_step);
}

int nto_sample_costs(const nt_scene_desc *desc, const nt_render_params *p, uint16_t *cost_out,
                     nt_render_stats *stats, int accel, int n_threads) {
    if (!cost_out) return NT_ERR_INVALID;
    return render_impl(desc, p, NULL, 0, NULL, cost_out, stats, accel, n_threads, 1);
}

int nto_render(const nt_scene_desc *desc, const nt_render_params *p, uint8_t *rgba_out, size_t stride,
               nt_render_stats *stats, int accel, int n_threads, uint32_t row_step) {
    return nto_render_radiance(desc, p, rgba_out, stride, NULL, stats, accel, n_threads, row_step);
}

int nto_trace_rays(const nt_scene_desc *desc, uint32_t n, const double *origins, const double *dirs,
                   double ray_epsilon, int accel, double *t_out, int32_t *prim_out) {
    octx c;
    int rc = ctx_init(&c, desc, ray_epsilon, 1, accel);
    if (rc) return rc;
    ocount k = { 0 };
    for (uint32_t i = 0; i < n; ++i) {
        double t;
        int g = nearest_hit(&c, ld3(origins + 3 * (size_t)i), ld3(dirs + 3 * (size_t)i), &t, &k);
        t_out[i] = g < 0 ? -1.0 : t;
        prim_out[i] = g;
    }
    ctx_free(&c);
    return NT_OK;
}
