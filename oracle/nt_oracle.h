/*
 * nt_oracle.h — TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * CPU restatement, in plain C, of SPEC-PROVISIONAL.md (binary64, no FMA, fixed operation order).
 *
 * PARITY UNPINNED: /root/reference holds one file (README:1-3, a URL); there is no NetTracer
 * source, test, golden vector or image to pin this against, and no JVM to run one (SURVEY.md §0,
 * §8(c)).  This oracle therefore restates this repository's own provisional spec, not NetTracer.
 * It supports self-consistency claims only (CUDA path == this file on the same inputs).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so.  Struct types are shared with the public header so both sides see the very
 * same bytes.
 */
#ifndef NT_ORACLE_H
#define NT_ORACLE_H

#include "../include/nettracer_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* accel: 0 = brute force over every primitive (the definition), 1 = oracle's own median-split
 * BVH over spheres+triangles (conservative culling; must equal accel 0 — tests check it).
 * n_threads: OpenMP threads (<=0 = all).  Rows [y0,y1) of the owned virtual rows are rendered
 * (y1 = 0 means all): lets bench.py time a bounded sample.  row_step renders every row_step-th
 * owned row only (others untouched), 0/1 = all. */
int nto_render(const nt_scene_desc *desc, const nt_render_params *params, uint8_t *rgba_out,
               size_t row_stride_bytes, nt_render_stats *stats, int accel, int n_threads,
               uint32_t row_step);

/* Same, but also returns the un-quantised per-pixel radiance (double[h][w][3], may be NULL). */
int nto_render_radiance(const nt_scene_desc *desc, const nt_render_params *params,
                        uint8_t *rgba_out, size_t row_stride_bytes, double *radiance_out,
                        nt_render_stats *stats, int accel, int n_threads, uint32_t row_step);

/* Analysis aid (scripts/sim_tile_schedule.py): per sample the rays cast and the ray-tree nodes visited,
 * cost_out[h][w][spp][2] (uint16, saturating).  No image is written. */
int nto_sample_costs(const nt_scene_desc *desc, const nt_render_params *params, uint16_t *cost_out,
                     nt_render_stats *stats, int accel, int n_threads);

/* Nearest hit of n rays, SPEC-PROVISIONAL §3. */
int nto_trace_rays(const nt_scene_desc *desc, uint32_t n, const double *origins,
                   const double *dirs, double ray_epsilon, int accel, double *t_out,
                   int32_t *prim_out);

int nto_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
