// nt_device.h — structs shared by the host API (nt_api.cu) and the kernels (nt_trace.cuh).
// Data layout in HBM is documented in DESIGN.md §3.
#pragma once
#include <cstddef>
#include <cstdint>

#include <vector_types.h>

#define NT_TRI_STRIDE 12   // v0[3] e1[3] e2[3] ng[3]  (R units; 16-byte aligned rows)
#define NT_MAT_STRIDE 12   // r g b ka kd ks shininess kr kt ior inv_ior pad
#define NT_COUNTER_SLOTS 32
#define NT_NCOUNTERS 11    // primary secondary shadow sphere plane triangle box light + sphere / plane / triangle tests started (instrumented kernel)
#define NT_COUNTER_EXTRA 3 // + next work item, blocks done, sync time-outs
#define NT_BLOCK_THREADS 256
#ifndef NT_MIN_BLOCKS_F64
#define NT_MIN_BLOCKS_F64 4 // resident blocks/SM the flat render kernel is compiled for (64 registers); measured
#endif                      // on configs[2] strict: 2 -> 1.55 ms, 3 -> 1.35 ms, 4 -> 1.30 ms, 5 -> 1.40 ms
#ifndef NT_MIN_BLOCKS_F32
#define NT_MIN_BLOCKS_F32 4 // fast mode with culling: 3 -> 0.68 ms, 4 -> 0.61 ms (before culling: 3 -> 0.72, 4 -> 0.74)
#endif
#ifndef NT_MIN_BLOCKS_SMALL
#define NT_MIN_BLOCKS_SMALL 3      // strict flat kernel, small launches (shards of a multi-GPU frame): 80 registers
#endif
#ifndef NT_SMALL_TILES_PER_WARP
#define NT_SMALL_TILES_PER_WARP 8  // "small" = fewer warp tiles than this per resident warp (1/8 of configs[2]: 6.8; 1/4: 13.7)
#endif
#define NT_BVH_STACK 96     // entries; a 4-wide node pushes up to 3: scene creation checks 3*depth + 4 <= this
#ifndef NT_MIN_BLOCKS_BVH
#define NT_MIN_BLOCKS_BVH 4 // measured best on configs[3] (2: 93.7 ms, 3: 78.2 ms, 4: 75.2 ms) despite spills
#endif
#ifndef NT_ADVANCE_THRESHOLD
#define NT_ADVANCE_THRESHOLD 32 // lanes parked with a finished query before the warp shades (8: 123 ms, 16: 106, 32: 94)
#endif
#ifndef NT_DESCEND_MIN
#define NT_DESCEND_MIN 12     // fewer lanes than this still descending -> go test the waiting leaves (configs[3] f64,
#endif                        // state machine: 2 -> 64.3 ms, 4 -> 63.1, 8 -> 62.9; wavefront: 4 -> 62.5, 8 -> 60.6; last session,
                              // wavefront with sorted shadow tasks: 8 -> 56.0, 12 -> 55.4, 16 -> 55.5, profiles/r04_wf_sort.txt)
#ifndef NT_REFILL_THRESHOLD
#define NT_REFILL_THRESHOLD 32 // idle lanes before the warp claims new samples (8: 105 ms, 24: 78.2, 32: 75.8)
#endif
#define NT_LEAF_MAX 4       // (count-1) is stored in 2 bits of a leaf ref
#define NT_MAX_DEPTH_DEV 16 // == NT_MAX_DEPTH of the public header
// Rule switches, == NT_RULE_* of the public header (nt_api.cu checks): SPEC-PROVISIONAL section 8
#define NT_DEV_RULE_TRUNCATE 2u
#define NT_DEV_RULE_ATTENUATE 4u
#define NT_DEV_RULE_RENORMALIZE 16u

// 64-byte BVH2 node: both child boxes (float, rounded outward) + child refs.
//   q0 = lo0.x lo0.y lo0.z hi0.x | q1 = hi0.y hi0.z lo1.x lo1.y | q2 = lo1.z hi1.x hi1.y hi1.z
//   q3 = c0 c1 n0 n1 (ints).  c = child ref: >= 0 inner node index; -1 empty; <= -2 leaf, with
//   -2 - c = first | (count-1) << 26 | kind << 28  (first = index into the BVH-ordered sphere (kind 0)
//   or triangle (kind 1) array, count 1..4).  n0/n1 keep the builder's count|kind<<8 (diagnostics).
struct NtBvhNode {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int c0, c1, n0, n1;
};
static_assert(sizeof(NtBvhNode) == 64, "node must be 64 bytes");

// 128-byte BVH4 node (one cache line), the layout the kernels traverse.  The host builder collapses its
// binary tree into it (largest-area inner child is replaced by its two children until 4 slots are used).
// Boxes are SoA so that the near/far plane of each axis is ONE 128-bit load chosen by the ray's sign:
//   float4 #0..2 = lo.x lo.y lo.z of the 4 children, #3..5 = hi.x hi.y hi.z, int4 #6 = child refs
//   (same ref encoding as above; empty slots: ref -1 and an inverted box), #7 = padding.
struct NtBvhNode4 {
    float lo[3][4];
    float hi[3][4];
    int ref[4];
    int pad[4];
};
static_assert(sizeof(NtBvhNode4) == 128, "node must be 128 bytes");

template <typename R>
struct NtSceneView {
    const R *sph;      // [ns][4]  cx cy cz r2
    const R *sph_invr; // [ns]
    const R *pln;      // [np][4]  nx ny nz d
    const R *tri;      // [nt][NT_TRI_STRIDE]
    const R *mat;      // [nm][NT_MAT_STRIDE]
    const R *lights;   // [nl][6]
    const R *globals;  // ambient[3] background[3] pad[2]
};

// BVH scenes: per point light a grid over the gnomonic projection of the sphere set as seen from the light (a shadow-map
// frustum), each cell listing the spheres whose projection touches it (nt_shadowgrid.h).  A shadow query tests its cell's few
// spheres with the exact rule and walks only the TRIANGLE set of the tree - the sphere set of configs[3] (10 000 small
// spheres scattered in a slab: mostly empty boxes) was 35 % of all box tests.
struct NtShadowGrid {
    float L[3], axis[3], U[3], V[3]; // light position; the projection's axis and its two plane directions (orthonormal)
    float u0, v0, su, sv;            // cell = floor((u - u0) * su), u = (P - L).U / (P - L).axis
    uint32_t K, base, valid, pad;    // K x K cells; this light's K*K + 1 offsets start at sg_off[base]
};

struct NtDevScene {
    uint32_t ns, np, nt, nm, nl;
    uint32_t use_bvh, n_nodes;
    float max_abs; // largest |coordinate| of any bounded primitive (box-test margin)
    float blo[3], bhi[3]; // bounds of all bounded primitives (float, rounded outward)
    const int *sph_mat, *sph_gid, *pln_mat, *tri_mat, *tri_gid;
    const unsigned *pln_code; // 2 bits per plane, 16 planes per word: 0..2 normal == +-e_k, 3 general (BVH scenes)
    // flat scenes: planes with a normal of exactly +-e_k, grouped by axis k: nax[k] (position p = +-d, index bits)
    // pairs in axl64 / axl32; pgen[ngen] = indices of all other planes
    uint32_t nax[3], ngen;
    const double *axl64;
    const float *axl32;
    const int *pgen;
    // flat scenes, slab == 1: no axis has more than two axis-aligned planes; axs64 / axs32 = [3][2] (position, index bits)
    // pairs, missing entries padded with (NaN, -1): the straight-line nearest-plane routine (nt_trace.cuh planes_nearest_slab)
    uint32_t slab;
    // flat scenes, rooms == 1: room64 / room32 = [nl][8] light rooms (nt_cull.h nt_cull_light_rooms)
    uint32_t rooms;
    uint32_t room_off[2], axs_off[2]; // byte offsets of the staged rooms / slab entries in dynamic shared memory, [0] binary32, [1] binary64
    const double *axs64, *room64;
    const float *axs32, *room32;
    const NtBvhNode4 *nodes;
    // The same construction for PRIMARY rays: grid nl of sgrid is built on the device at the start of every render call from
    // that call's eye (nt_eyegrid.cuh) - a primary ray tests the spheres its cell lists and walks only the triangle set.
    uint32_t eg_on, eg_k0, eg_items_cap;
    uint32_t *eg_off, *eg_items; // the eye grid's own K0*K0 + 2 offsets (+ 1024 block totals of the scan) and item room (device-only memory)
    float sph_lo[3], sph_hi[3];  // bounds of the sphere CENTRES (the projection's axis points from the eye to their middle)
    double *eg_boxes;            // [ns][4] scratch: projected rectangle of every sphere for the current eye
    unsigned long long *eg_acc;  // [8] scratch: bounds (ordered doubles) umin umax vmin vmax, invalid flag
    uint32_t sg_on;              // BVH scenes: shadow grids exist (nt_shadowgrid.h); a light's own `valid` says whether it has one
    const NtShadowGrid *sgrid;   // [nl]
    const uint32_t *sg_off, *sg_items; // per cell: first item; items = sphere indices (device order), ascending per cell
    // flat scenes: conservative culling tables (nt_cull.h); cull == 0 -> every query tests every primitive
    uint32_t cull, lbuf_k;
    float cull_far; // shadow queries whose light is farther than this (max norm) look no light buffer up: see lbuf_mask
    uint32_t lfree; // bit l: no plane can lie between a point of a bounded primitive and light l (nt_cull_plane_free_lights)
    unsigned long long sph_bits, all_bits; // masks of the sphere bits / of all bounded primitives (flat scenes)
    const unsigned long long *lbuf; // [nl][6][lbuf_k][lbuf_k] light buffers
    const unsigned long long *nbr;  // [ns] balls touching ball i
    NtSceneView<double> v64;
    NtSceneView<float> v32;
};

struct NtRenderArgs {
    uint32_t width, height, spp, n, max_depth;
    uint32_t shard_index, shard_count, band_rows, layout, vrows;
    uint32_t lanes;    // lanes per pixel (power of two dividing spp, <= 32)
    uint32_t twx, twy; // warp tile in pixels, twx*twy*lanes == 32
    uint32_t tiles_x, tiles_y; // warp tiles over the virtual image (owned rows only)
    uint32_t n_tiles, tile_rot; // tiles_x * tiles_y; the tile sequence starts at tile tile_rot and wraps (nt_trace.cuh tile_origin)
    double eps;
    double eps_lo;     // fl(eps * (1 - 2^-50)), 0 when eps < 1e-290: quotients provably <= eps skip the division (plane_below_eps)
    double cam[12];    // eye p00 dx dy
    double samp_off[8]; // SPEC §2 sample offsets (i + 0.5) / n, i < n
    double inv_spp;     // 1 / spp
    float camf[12], samp_off_f[8], inv_spp_f; // the same values rounded to binary32 on the host: the fast kernels paid a
                                              // conversion (F2F + flush test, 3 instructions on the XU / FP64 pipes) per use
    alignas(8) uint16_t prect[64][4]; // flat scenes with culling: pixel rectangle x0 x1 y0 y1 of every bounded primitive (nt_cull.h)
    // division-free tile arithmetic (nt_trace.cuh tile_origin / row_to_y): twx, twy, lanes are powers of two
    uint32_t log2_twx, log2_twy, log2_lanes;
    uint32_t band_magic; // ceil(2^32 / band_rows) (band_rows > 1)
    uint32_t n_mul;      // ceil(65536 / n)
    float inv_tiles_x;
    uint8_t *out;
    size_t stride;
    unsigned long long *counters; // [NT_COUNTER_SLOTS][NT_NCOUNTERS] + NT_COUNTER_EXTRA: next warp tile / next sample id,
                                  // blocks that finished (frame-done flag), sync time-outs
    void *samples;                // BVH scenes: per-sample radiance, R[3] each (see nt_bvh_trace.cuh)
    unsigned *n_launches;         // host counter: kernels launched for this frame (may be NULL)
    void *wf;                     // BVH scenes: wavefront workspace (nt_wavefront.cuh); NULL = per-lane state machine
    size_t wf_bytes;
    // frame synchronisation of the multi-GPU exchange (include/nettracer_b200.h nt_frame_sync; nt_sync.cuh): all three are
    // optional (NULL).  Flat scenes: done inside the render kernel; BVH scenes: by sync_kernel before / after the pipeline.
    unsigned *sync_post_ptr;        // system-scope store of sync_post_val when the frame's first kernel starts
    const unsigned *sync_wait_ptr;  // no pixel is stored before (int)(*sync_wait_ptr - sync_wait_val) >= 0
    unsigned *sync_done_ptr;        // release-store of sync_done_val after the frame's last pixel store
    unsigned sync_post_val, sync_wait_val, sync_done_val;
    uint32_t count_executed;        // flat scenes: launch the instrumented kernel (CountersX)
    uint32_t rules;                 // NT_DEV_RULE_* bits (SPEC-PROVISIONAL section 8); 0 = the default rules
};

// per-precision view of the camera / sample arguments (constant bank)
#ifdef __CUDACC__
#define NT_HD __host__ __device__
#else
#define NT_HD
#endif
template <typename R> struct ArgsView;
template <> struct ArgsView<double> {
    static NT_HD inline double cam(const NtRenderArgs &a, int i) { return a.cam[i]; }
    static NT_HD inline double samp_off(const NtRenderArgs &a, unsigned i) { return a.samp_off[i]; }
    static NT_HD inline double inv_spp(const NtRenderArgs &a) { return a.inv_spp; }
};
template <> struct ArgsView<float> {
    static NT_HD inline float cam(const NtRenderArgs &a, int i) { return a.camf[i]; }
    static NT_HD inline float samp_off(const NtRenderArgs &a, unsigned i) { return a.samp_off_f[i]; }
    static NT_HD inline float inv_spp(const NtRenderArgs &a) { return a.inv_spp_f; }
};

struct NtTraceArgs {
    uint32_t n;
    double eps, eps_lo;
    const double *origins, *dirs;
    double *t_out;
    int *prim_out;
};

// Launchers, one translation unit per precision (strict: -fmad=false; fast: FMA + fast math).
int nt_launch_render_f64(const NtDevScene &s, const NtRenderArgs &a, void *stream);
int nt_launch_render_f32(const NtDevScene &s, const NtRenderArgs &a, void *stream);
int nt_launch_trace_f64(const NtDevScene &s, const NtTraceArgs &a, void *stream);
int nt_launch_trace_f32(const NtDevScene &s, const NtTraceArgs &a, void *stream);
size_t nt_flat_smem_bytes(const NtDevScene &s, int precision);
size_t nt_sample_buffer_bytes(const NtDevScene &s, const NtRenderArgs &a, int precision); // 0 for flat scenes
// Wavefront workspace that holds the whole frame in one chunk (0: the scene / parameters do not use the wavefront path)
size_t nt_wavefront_bytes(const NtDevScene &s, const NtRenderArgs &a, int precision);
size_t nt_wavefront_min_bytes(const NtRenderArgs &a, int precision); // workspace for a chunk of 32 samples
