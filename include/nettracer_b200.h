/*
 * nettracer_b200.h — C ABI of the B200 intersect-and-shade path.
 *
 * Drop-in boundary.  The reference interface each entry point would
 * replace CANNOT be cited: /root/reference holds a single README (README:1-3, a URL), so
 * there is no Java render API to bind against.  The surface below is the one SURVEY.md
 * §8(b) proposes — host (Java via Panama FFM / JNI, or any FFI) keeps scene parsing,
 * camera set-up and image output and hands flat arrays across; see INTEGRATION.md.
 *
 * Rules: plain C types only; errors are negative ints (never exceptions); the text of
 * the last error on the calling thread is nt_last_error(); there is no CPU fallback —
 * every compute entry point fails with NT_ERR_NO_DEVICE when no sm_100 GPU is usable.
 * The shading rules are SPEC-PROVISIONAL.md (this repository's own; parity with
 * NetTracer is unpinned).
 */
#ifndef NETTRACER_B200_H
#define NETTRACER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NT_ABI_VERSION 2

enum nt_status {
    NT_OK = 0,
    NT_ERR_INVALID = -1,   /* bad argument / malformed scene */
    NT_ERR_NO_DEVICE = -2, /* no usable sm_100 device */
    NT_ERR_CUDA = -3,      /* CUDA runtime error, text in nt_last_error() */
    NT_ERR_NOMEM = -4,
    NT_ERR_TIMEOUT = -5,   /* a frame-synchronisation flag was not reached in time (a peer is gone) */
    NT_ERR_SYSTEM = -6     /* shm_open / mmap / thread creation failed, errno text in nt_last_error() */
};

enum nt_precision {
    NT_F64_STRICT = 0, /* binary64, no FMA contraction, SPEC-PROVISIONAL §1-6 */
    NT_F32_FAST = 1    /* binary32 fast mode, SPEC-PROVISIONAL §7 */
};

enum nt_layout {
    NT_LAYOUT_FULL = 0,   /* pixel (x,y) at out + y*stride + 4*x; only owned rows are written */
    NT_LAYOUT_COMPACT = 1 /* owned row bands packed back to back, stride bytes per row */
};

/* Camera resolved by the host (SPEC-PROVISIONAL §2). */
typedef struct nt_camera {
    double eye[3];
    double p00[3];
    double dx[3];
    double dy[3];
} nt_camera;

/* Flat row-major arrays owned by the caller; nt_scene_create copies what it needs. */
typedef struct nt_scene_desc {
    uint32_t struct_size; /* sizeof(nt_scene_desc), for ABI evolution */
    uint32_t n_spheres, n_planes, n_triangles, n_materials, n_lights;
    const double *spheres;        /* [n_spheres][4]   cx cy cz r */
    const int32_t *sphere_mat;    /* [n_spheres] */
    const double *planes;         /* [n_planes][4]    nx ny nz d */
    const int32_t *plane_mat;     /* [n_planes] */
    const double *triangles;      /* [n_triangles][9] v0 v1 v2 */
    const int32_t *triangle_mat;  /* [n_triangles] */
    const double *materials;      /* [n_materials][10] r g b ka kd ks shininess kr kt ior (ior > 0) */
    const double *lights;         /* [n_lights][6]    px py pz r g b */
    double ambient[3];
    double background[3];
} nt_scene_desc;

typedef struct nt_render_params {
    uint32_t struct_size; /* sizeof(nt_render_params) */
    uint32_t width, height;
    uint32_t spp;         /* perfect square, 1..64 */
    uint32_t max_depth;   /* 1..NT_MAX_DEPTH; 1 = primary + shadow rays only */
    uint32_t precision;   /* enum nt_precision */
    double ray_epsilon;   /* <= 0 selects the default 1e-6 */
    nt_camera camera;
    /* Image sharding for multi-GPU: rows are cut into bands of band_rows rows; band b is
     * owned by shard (b mod shard_count).  shard_count = 1 renders everything. */
    uint32_t shard_index, shard_count, band_rows;
    uint32_t layout;      /* enum nt_layout */
    uint32_t flags;       /* NT_RENDER_* bits, 0 for production renders */
} nt_render_params;

/* Measurement aid: flat scenes are rendered by an instrumented twin of the kernel that also counts the primitive
 * tests it really starts (nt_render_stats.*_tests_executed); same image, a little slower. */
#define NT_RENDER_COUNT_EXECUTED 1u
/* Rule switches (SPEC-PROVISIONAL.md section 8): rules the reference would dictate and that are [OPEN] while its sources
 * are missing.  0 = the defaults of sections 2-5; every switch is implemented by the CUDA path AND the oracle. */
#define NT_RULE_QUANTIZE_TRUNCATE 2u   /* section 5: q = (int)(c*255) instead of (int)(c*255 + 0.5) */
#define NT_RULE_ATTENUATE_INV_SQUARE 4u /* section 4: a light's colour is scaled by 1/d2 (d2 = squared distance to the light) */
#define NT_RULE_SAMPLE_CORNER 8u       /* section 2: sample offsets i/n, j/n (cell corner) instead of (i+0.5)/n, (j+0.5)/n */
#define NT_RULE_RENORMALIZE 16u        /* section 4: reflected / refracted directions are re-normalised, d * (1/sqrt(dot(d,d))) */
#define NT_RULE_MASK 30u

#define NT_MAX_DEPTH 16

typedef struct nt_render_stats {
    uint64_t rays_primary;
    uint64_t rays_secondary; /* reflection + transmission */
    uint64_t rays_shadow;
    /* Algorithmic work executed, by kind (the roofline numerators). */
    uint64_t sphere_tests;
    uint64_t plane_tests;
    uint64_t triangle_tests;
    uint64_t box_tests;      /* BVH child boxes tested (0 for flat scenes) */
    uint64_t light_evals;    /* unoccluded light contributions shaded */
    /* Tests the kernel really started (flat scenes, NT_RENDER_COUNT_EXECUTED only, else 0): what is left of the
     * brute-force counts above after the conservative culling. */
    uint64_t sphere_tests_executed, plane_tests_executed, triangle_tests_executed;
    double kernel_ms;        /* device time of the render kernel(s), CUDA events; nt_render only */
    double total_ms;         /* host wall time of the whole nt_render call */
} nt_render_stats;

typedef struct nt_scene nt_scene;

/* ---- library ---- */
int nt_abi_version(void);
const char *nt_last_error(void);
int nt_device_count(int *count);

/* ---- scene ---- */
/* Validates, derives per-primitive constants, builds the BVH when the scene has more bounded
 * primitives than fit the flat kernel, and uploads to `device`. */
int nt_scene_create(const nt_scene_desc *desc, int device, nt_scene **out);
void nt_scene_destroy(nt_scene *scene);
/* info[0] = bit 0 uses_bvh | bit 1 BVH built on the GPU (NT_BVH_BUILD=gpu) | bit 2 flat culling tables in use
 *           | bits 8.. BVH build time in us;
 * info[1] = 4-wide BVH nodes; info[2] = device bytes; info[3] = device | kernels launched by the last render call << 32 */
int nt_scene_info(const nt_scene *scene, uint64_t info[4]);

/* Diagnostic, host only (no GPU needed): the conservative culling tables nt_scene_create builds for a flat
 * scene (<= 64 bounded primitives; nettracer_b200/csrc/nt_cull.h).  Bit j of a mask = sphere j, then triangle
 * j - n_spheres.  *k_out = cells per cube-face edge; lbuf_out[n_lights][6][k][k] light buffers (may be NULL);
 * nbr_out[n_spheres] balls touching each sphere; bsph_out[n_spheres + n_triangles][4] bounding spheres.
 * NT_ERR_INVALID when the scene is not eligible.  Tests use it to prove conservativeness on the CPU. */
int nt_cull_tables(const nt_scene_desc *desc, uint32_t *k_out, uint64_t *lbuf_out, size_t lbuf_capacity,
                   uint64_t *nbr_out, double *bsph_out);

/* Diagnostic, host only: the pixel rectangles { x0, x1, y0, y1 } (inclusive; empty when x0 > x1) that nt_render
 * derives from the camera for the primary rays of a flat scene - no primary ray of a pixel outside rectangle j can
 * touch bounded primitive j.  rects_out[n_spheres + n_triangles][4].  Same eligibility as nt_cull_tables. */
int nt_primary_rects(const nt_scene_desc *desc, const nt_render_params *params, uint16_t *rects_out);

/* Diagnostic, host only: bit l of *mask_out is set when nt_scene_create proved that no plane of a flat scene can lie
 * between light l and any point of a bounded primitive (shadow queries from spheres / triangles towards that light
 * then skip the planes).  Same eligibility as nt_cull_tables. */
int nt_plane_free_lights(const nt_scene_desc *desc, uint32_t *mask_out);

/* Diagnostic, host only: the "light rooms" nt_scene_create derives for a flat scene (nettracer_b200/csrc/nt_cull.h
 * nt_cull_light_rooms).  rooms_out[n_lights][8] = { lo_x, hi_x, lo_y, hi_y, lo_z, hi_z, cap_per_eps, cap_max }: a shadow
 * query towards light l whose origin P has lo_k <= P_k <= hi_k on every axis and whose light distance is <= min(ray_epsilon
 * * cap_per_eps, cap_max) cannot be stopped by a plane with a normal of exactly +-e_k, and skips those planes.  Works for
 * any scene (a light without a room gets an empty box); NT_ERR_INVALID for NULL arguments. */
int nt_light_rooms(const nt_scene_desc *desc, double *rooms_out);

/* Diagnostic, host only: the shadow grid nt_scene_create derives for light `light` of a BVH scene (nettracer_b200/csrc/
 * nt_shadowgrid.h) - a K x K grid over the projection of the sphere set as seen from the light, each cell listing the
 * spheres that can lie between the light and a point that projects into it; shadow queries test those with the exact rule
 * and walk only the triangle set of the tree.  params_out[16] = light position, axis, U, V (3 each, binary32 as the device
 * uses them), u0, v0, su, sv: a point X with w = (X - L).axis > 0 falls into cell (floor(((X - L).U / w - u0) su),
 * floor(((X - L).V / w - v0) sv)); outside [0, K) or w <= 0 = no sphere in between.  *k_out = K, or 0 when the light has
 * no grid (its queries walk the whole tree) - nothing else is written then.  off_out[K K + 1] / items_out[*n_items_out]:
 * cell c lists items_out[off_out[c] - off_out[0] .. off_out[c + 1] - off_out[0]), sphere indices of the DESCRIPTION,
 * ascending.  Buffers may be NULL / too small: the call then only reports K and *n_items_out.  Spheres only (no tree is
 * built); NT_ERR_INVALID for a bad light index. */
int nt_shadow_grid(const nt_scene_desc *desc, uint32_t light, float *params_out, uint32_t *k_out, uint32_t *off_out,
                   size_t off_capacity, uint32_t *items_out, size_t items_capacity, size_t *n_items_out);

/* ---- render ---- */
/* Host buffer (pageable or pinned), blocking.  Renders the shard named in params, copies the
 * result to rgba_out.  A PINNED (page-locked) rgba_out receives the pixels straight from the kernel over PCIe (no device
 * frame, no copy).  stats may be NULL - and should be when only the image is wanted: for a flat scene and a pinned frame
 * the call then needs no events and no copy of the work counters, the kernel posts a completion flag into pinned host
 * memory after its last pixel store and the call returns on it (~20 us of a 0.7 ms frame). */
int nt_render(nt_scene *scene, const nt_render_params *params, uint8_t *rgba_out,
              size_t row_stride_bytes, nt_render_stats *stats);

/* Device buffer on the scene's device, asynchronous on `cuda_stream` (a cudaStream_t, may be
 * NULL for the default stream).  rgba_out_dev may be a peer-mapped pointer of another GPU
 * (NVLink): with NT_LAYOUT_FULL every shard can store straight into one remote framebuffer.
 * Concurrency: calls on one scene may be in flight on DIFFERENT streams.  Flat scenes (<= 64 bounded
 * primitives) then really overlap - each call takes its own block of work counters from a small ring;
 * BVH scenes share per-scene scratch buffers, so the library orders a call after the previous one with
 * an event (no overlap, no corruption).  nt_render (host buffer) is blocking and serialised per scene. */
int nt_render_device(nt_scene *scene, const nt_render_params *params, void *rgba_out_dev,
                     size_t row_stride_bytes, void *cuda_stream);
/* Synchronises `cuda_stream` and returns the counters of the last nt_render_device[_sync] call on
 * the scene.  NT_ERR_TIMEOUT when that frame gave up waiting on a nt_frame_sync flag. */
int nt_render_device_stats(nt_scene *scene, void *cuda_stream, nt_render_stats *stats);

/* ---- frame synchronisation of a sharded render (the exchange step of SURVEY.md section 8(e)) ----
 * 32-bit sequence-numbered flags in memory every participant can reach: device memory of the gathering
 * GPU (opened by the other processes with nt_ipc_open) or pinned host memory.  They order the frame
 * exchange on the device, without a collective: each shard's last kernel release-stores "frame f is
 * written" next to the pixels it stored over NVLink, the gathering GPU spins on those flags
 * (nt_flags_wait_device), and its own next frame acknowledges, when it starts, that it has consumed
 * the earlier ones so that a peer may overwrite the buffer of two frames ago.  Flags are compared by
 * signed difference (they may wrap); a wait gives up after ~2 s (NT_ERR_TIMEOUT from the stats call). */
typedef struct nt_frame_sync {
    uint32_t struct_size;           /* sizeof(nt_frame_sync) */
    uint32_t post_at_start_value;
    uint32_t wait_value;
    uint32_t post_when_done_value;
    uint32_t *post_at_start;        /* may be NULL: stored when this frame's first kernel starts, i.e. after all
                                       earlier work of the stream has completed */
    const uint32_t *wait_before_store; /* may be NULL: no pixel is stored before *flag has reached wait_value */
    uint32_t *post_when_done;       /* may be NULL: release-stored after this frame's last pixel store */
} nt_frame_sync;
/* nt_render_device with the three optional flag operations fused into the frame's kernels. */
int nt_render_device_sync(nt_scene *scene, const nt_render_params *params, void *rgba_out_dev,
                          size_t row_stride_bytes, void *cuda_stream, const nt_frame_sync *sync);
/* Enqueues a wait on `cuda_stream` (of the scene's device) until flags[0..n-1] have all reached `value`:
 * one tiny kernel; work enqueued after it sees everything the posters stored before posting.  A wait
 * that gives up is reported by the scene's next nt_render_device_stats as NT_ERR_TIMEOUT. */
int nt_flags_wait_device(nt_scene *scene, const uint32_t *flags, uint32_t n, uint32_t value, void *cuda_stream);

/* ---- several GPUs behind one call (one host process; a Java host binds exactly this) ----
 * The scene is replicated on every listed device; a frame is cut into interleaved row bands
 * (params->band_rows, 0 = 8; shard_index / shard_count / layout of params are ignored), every GPU
 * renders its bands on its own stream, driven by its own host thread, and stores the RGBA8 words
 * straight into the caller's buffer when that is pinned host memory (cudaHostAlloc / cudaHostRegister:
 * every GPU writes over its own PCIe link, no gather, no copy), otherwise into a pinned staging frame
 * of the library that is copied out band by band.  Blocking.  stats: counters summed over the GPUs,
 * kernel_ms = the slowest GPU's kernel. */
typedef struct nt_multi nt_multi;
int nt_multi_create(const nt_scene_desc *desc, const int *devices, int n_devices, nt_multi **out);
void nt_multi_destroy(nt_multi *multi);
int nt_multi_device_count(const nt_multi *multi);
int nt_multi_render(nt_multi *multi, const nt_render_params *params, uint8_t *rgba_out,
                    size_t row_stride_bytes, nt_render_stats *stats);

/* ---- unit-level entry point (parity tests of the intersectors) ---- */
/* Nearest hit of n rays (host arrays origins[n][3], dirs[n][3], used as given — not normalised)
 * against the scene under SPEC-PROVISIONAL §3.  t_out[i] = hit distance or -1; prim_out[i] =
 * global primitive id or -1.  precision as in nt_render_params; ray_epsilon <= 0 = default. */
int nt_trace_rays(nt_scene *scene, uint32_t n, const double *origins, const double *dirs,
                  uint32_t precision, double ray_epsilon, double *t_out, int32_t *prim_out);

/* Rows a shard owns / bytes its compact buffer needs (pure host arithmetic). */
uint32_t nt_shard_rows(uint32_t height, uint32_t band_rows, uint32_t shard_index,
                       uint32_t shard_count);

/* Rank-0 side of the gather: scatter `shard_count` compact buffers, laid out back to back with
 * `shard_stride_bytes` between them, into a full frame.  Device pointers, async on stream. */
int nt_deinterleave_device(const void *compact_all, size_t shard_stride_bytes, void *full_out,
                           size_t row_stride_bytes, uint32_t width, uint32_t height,
                           uint32_t band_rows, uint32_t shard_count, int device,
                           void *cuda_stream);

/* ---- peer framebuffer (NVLink store path) ---- */
/* A plain cudaMalloc'd buffer on `device` (so that its IPC handle has offset 0), and its release. */
int nt_device_malloc(int device, size_t bytes, void **dev_ptr_out);
int nt_device_free(int device, void *dev_ptr);
/* Export / open a CUDA IPC handle (64 bytes) for a device allocation, so that every rank's
 * render kernel can store its pixels directly into rank 0's framebuffer. */
int nt_ipc_export(const void *dev_ptr, int device, uint8_t handle_out[64]);
int nt_ipc_open(const uint8_t handle[64], int device, void **dev_ptr_out);
int nt_ipc_close(void *dev_ptr, int device);

/* ---- one host frame shared by several processes (one process per GPU: torchrun, MPI, ...) ----
 * A POSIX shared-memory segment holding an RGBA8 frame and one flag line per rank, mapped by every
 * rank and page-locked for its GPU, so that nt_render (host pointer = nt_host_frame_pixels) stores
 * each rank's bands straight into the one frame that rank 0's host reads: the end-to-end path of a
 * sharded render without a gather and without a device-to-host copy.  Protocol per frame `seq`
 * (1, 2, ...): every rank: nt_host_frame_wait_ack(seq - 1); nt_render(...); nt_host_frame_post(rank, seq) - or the
 * asynchronous form, nt_render_device_sync with the rank's flag as post_when_done (see nt_host_frame_flag);
 * rank 0: nt_host_frame_wait_all(seq), reads the pixels, nt_host_frame_ack(seq). */
typedef struct nt_host_frame nt_host_frame;
/* create != 0: make the segment (rank 0); 0: attach to an existing one.  device: the GPU of the calling
 * process that will store into it.  name: "/something", as for shm_open. */
int nt_host_frame_open(const char *name, size_t frame_bytes, uint32_t n_ranks, int create, int device,
                       nt_host_frame **out);
uint8_t *nt_host_frame_pixels(nt_host_frame *frame);
/* The flag word of `rank` ("frames this rank has completely stored"), as a pointer both the host and - the segment being
 * page-locked and mapped - this process's GPU can use: nt_render_device_sync(..., rgba_out_dev = nt_host_frame_pixels(),
 * sync.post_when_done = nt_host_frame_flag(frame, rank)) makes the render kernel itself post the flag after its last
 * pixel store, so that the rank's host never has to wait for its GPU (nt_host_frame_wait_all on rank 0 sees it). */
uint32_t *nt_host_frame_flag(nt_host_frame *frame, uint32_t rank);
int nt_host_frame_post(nt_host_frame *frame, uint32_t rank, uint32_t seq);
int nt_host_frame_wait_all(nt_host_frame *frame, uint32_t seq, uint32_t timeout_ms);
int nt_host_frame_ack(nt_host_frame *frame, uint32_t seq);
int nt_host_frame_wait_ack(nt_host_frame *frame, uint32_t seq, uint32_t timeout_ms);
void nt_host_frame_close(nt_host_frame *frame, int unlink_segment);

/* ---- roofline denominators ---- */
typedef struct nt_peaks {
    double f64_fma_gflops;   /* DFMA chain, 2 flops per instruction */
    double f64_nofma_gflops; /* alternating DMUL / DADD, 1 flop per instruction */
    double f32_fma_gflops;   /* FFMA chain */
    double f32_nofma_gflops;
    double sm_clock_mhz_est; /* clock64 ticks / %globaltimer nanoseconds of one block of the DFMA run */
    int sm_count;
} nt_peaks;
/* Issue-rate micro-benchmark on all SMs of `device`: ~50 ms per figure. */
int nt_measure_peaks(int device, nt_peaks *out);

#ifdef __cplusplus
}
#endif
#endif /* NETTRACER_B200_H */
