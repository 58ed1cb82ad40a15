/*
 * rvpt_abi.h — the drop-in boundary of the B200 path-tracing engine.
 *
 * Plain C ABI (no C++/torch types): the POD structs below are byte-identical to
 * the ones RVPT hands to Vulkan today, and the entry points are what the
 * reference's `RVPT::update()` / `RVPT::draw()` would bind instead of the
 * per-frame buffer uploads + `vkCmdDispatch`.
 *
 * Reference interfaces replaced (paths relative to the RVPT tree):
 *   - per-frame uploads            src/rvpt/rvpt.cpp:118-126
 *   - compute dispatch + submit    src/rvpt/rvpt.cpp:350-354, 1005-1039
 *   - scene buffers / BVH build    src/rvpt/rvpt.cpp:84-91, 824-835
 *   - descriptor set 0, bindings   assets/shaders/compute_pass.comp:28-58
 *
 * Error convention: every int-returning call returns 0 on success and a
 * negative RVPT_B200_E* code on failure; the message is available through
 * rvpt_b200_last_error(). Nothing here ever calls exit()/abort().
 *
 * Threading: one ctx = one caller thread at a time. rvpt_b200_render_frame()
 * is asynchronous on the ctx stream; frames serialise on that stream (the
 * temporal accumulation dependency); readbacks synchronise.
 */
#ifndef RVPT_ABI_H
#define RVPT_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RVPT_API
#else
#define RVPT_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------ */
/* POD structs — kept byte-for-byte (sizes 40 / 80 / 32 / 64 / 48).          */
/* ------------------------------------------------------------------------ */

/* src/rvpt/rvpt.h:77-89 ; UBO binding 0, compute_pass.comp:28-40. */
typedef struct rvpt_render_settings
{
    int32_t max_bounces;              /* default 8 */
    int32_t aa;                       /* samples per pixel per frame, default 1 */
    uint32_t current_frame;           /* owned by the caller (rvpt.cpp:102-111) */
    int32_t camera_mode;              /* 0 pinhole, 1 ortho, 2 spherical */
    int32_t top_left_render_mode;     /* integrator index, default 9 (Kajiya) */
    int32_t top_right_render_mode;
    int32_t bottom_left_render_mode;
    int32_t bottom_right_render_mode;
    float split_ratio[2];             /* default 0.5, 0.5 */
} rvpt_render_settings;

/* src/rvpt/camera.cpp:55-66 ; UBO binding 4, compute_pass.comp:44-49.
 * matrix is column-major (matrix[4*c + r]); params = aspect, fov (radians,
 * used as the vertical field), ortho scale, 0. */
typedef struct rvpt_camera_data
{
    float matrix[16];
    float params[4];
} rvpt_camera_data;

/* src/rvpt/bvh.h:12-18 ; SSBO binding 5 ; structs.glsl:9-14.
 * bounds = minx,maxx,miny,maxy,minz,maxz ; leaf iff primitive_count > 0 ;
 * children of an inner node live at first_child_or_primitive and +1. */
typedef struct rvpt_bvh_node
{
    uint32_t first_child_or_primitive;
    uint32_t primitive_count;
    float bounds[6];
} rvpt_bvh_node;

/* src/rvpt/geometry.h:76-111 ; SSBO binding 6 ; structs.glsl:1-7.
 * The three .w lanes carry the host-side face normal (unused by the shader),
 * material_id[0] is the material index stored as a float. */
typedef struct rvpt_triangle
{
    float vertex0[4];
    float vertex1[4];
    float vertex2[4];
    float material_id[4];
} rvpt_triangle;

/* src/rvpt/material.h:9-26 ; SSBO binding 7 ; structs.glsl:22-33.
 * albedo.w = index of refraction, data[0] = type (0 Lambert, 1 mirror,
 * 2 dielectric). */
typedef struct rvpt_material
{
    float albedo[4];
    float emission[4];
    float data[4];
} rvpt_material;

#define RVPT_MAX_BOUNCE_STATS 64

/* Counters of the last render_frame call (all `aa` passes summed) or, after a batched
 * rvpt_b200_render_frames, of its last launch: `frames` frames summed. */
typedef struct rvpt_b200_stats
{
    uint64_t samples;                              /* pixels x aa x frames rendered by this ctx */
    uint64_t segments;                             /* intersect_scene calls = sum of active[] */
    uint64_t active[RVPT_MAX_BOUNCE_STATS];        /* rays traced at bounce b */
    uint32_t kernel_launches;                      /* kernels launched by the last render_frame(s) call */
    uint32_t traversal_order;                      /* 0 reference child order, 1 front to back (see REFERENCE_ORDER) */
    uint32_t frames;                               /* frames the counters cover (1 unless batched) */
    uint32_t reserved;
} rvpt_b200_stats;

/* ------------------------------------------------------------------------ */
/* Error codes                                                              */
/* ------------------------------------------------------------------------ */
#define RVPT_B200_OK 0
#define RVPT_B200_EINVAL (-1)      /* bad argument */
#define RVPT_B200_ECUDA (-2)       /* CUDA runtime error (see last_error) */
#define RVPT_B200_ENOSCENE (-3)    /* render before upload_scene */
#define RVPT_B200_EUNSUPPORTED (-4) /* beyond a limit of this engine (scene > 4 GiB, BVH deeper than the shader's 64-entry stack) */
#define RVPT_B200_ENOMEM (-5)

/* ------------------------------------------------------------------------ */
/* Context flags (rvpt_b200_create)                                          */
/* ------------------------------------------------------------------------ */
/* Accumulate through an 8-bit UNORM temporal image exactly like the
 * reference (rvpt.cpp:759-766, compute_pass.comp:146-148,165): prev is
 * loaded as k/255, the running mean is clamped and stored as round(x*255).
 * Default (flag clear) keeps the running mean in float32. */
#define RVPT_B200_FLAG_ACCUM_RGBA8 0x1u
/* Cover only floor(W/16) x floor(H/16) workgroups like the reference's
 * integer-division dispatch (rvpt.cpp:1035-1036); other pixels stay 0.
 * Default renders every pixel. */
#define RVPT_B200_FLAG_REFERENCE_DISPATCH 0x2u
/* Ignore the BVH and test every triangle in upload order (legacy
 * intersect_triangles semantics; used to cross-check BVH traversal). */
#define RVPT_B200_FLAG_BRUTE_FORCE 0x4u
/* Launch every wave (primary, bounce 1, bounce 2, ...) as its own kernel
 * instead of the single persistent cooperative kernel per frame. Same results;
 * used for per-wave profiling and as a cross-check of the fused kernel. */
#define RVPT_B200_FLAG_UNFUSED 0x8u
/* Do not build the per-CTA direction-octant copies of the BVH nodes (the slab
 * test then uses the reference's per-axis min/max form for every ray). Same
 * results; for A/B measurements. */
#define RVPT_B200_FLAG_NO_OCTANTS 0x10u
/* rvpt_b200_render_frames launches every frame on its own instead of merging the waves of
 * consecutive frames into batched launches. Same results; for A/B measurements. (0x20 was
 * the experimental barrier-free kernel of ABI version 1.) */
#define RVPT_B200_FLAG_NO_BATCH 0x200u
/* Do not use the previous frame's per-bounce ray counts to forecast tiny waves (k_frame
 * then always queues survivors and pays the barrier + tail wave behind them). Same
 * results; for A/B measurements. */
#define RVPT_B200_FLAG_NO_FORECAST 0x40u
/* Walk every BVH in the reference's child order (first, first+1: intersection.glsl:402-406).
 * By default small scenes WITHOUT coincident faces (two coplanar triangles overlapping with
 * positive area, checked at upload) are walked front to back per ray-direction octant: the
 * nearest accepted hit then does not depend on the visiting order (SURVEY.md §8c "BVH
 * dependence"; the parity tests hold both orders to the oracle bit for bit). Scenes with
 * coincident faces — where the strict `t < closest_t` lets the first triangle visited win —
 * always keep the reference's order. rvpt_b200_stats.traversal_order reports the choice. */
#define RVPT_B200_FLAG_REFERENCE_ORDER 0x80u
/* Keep one path queue per wave instead of eight sub-queues (one per direction octant of the
 * queued ray; 8x the queue memory). Same results; for A/B measurements. */
#define RVPT_B200_FLAG_NO_QUEUE_SORT 0x100u

/* rvpt_b200_upload_scene(nodes == NULL) builds its BVH on the GPU (rvpt_b200_build_bvh_gpu)
 * instead of on the host (rvpt_b200_build_bvh). */
#define RVPT_B200_FLAG_GPU_BVH 0x400u
/* Batched launches (rvpt_b200_render_frames) normally render their primary wave pixel block by
 * pixel block with a per-block list of candidate leaves (kernels.cu, primary_phase_beam); this
 * flag makes every primary ray walk the tree instead. Same results; for A/B measurements and
 * cross-checks. */
#define RVPT_B200_FLAG_NO_LEAF_LISTS 0x800u

typedef struct rvpt_b200_ctx rvpt_b200_ctx;

/* ------------------------------------------------------------------------ */
/* Engine lifetime — replaces RVPT::initialize()/shutdown() for the compute  */
/* path (rvpt.cpp:57-94, 407-442).                                           */
/* ------------------------------------------------------------------------ */
RVPT_API int rvpt_b200_create(rvpt_b200_ctx** out, int device, uint32_t width, uint32_t height,
                              uint32_t flags);
RVPT_API void rvpt_b200_destroy(rvpt_b200_ctx* ctx);
RVPT_API const char* rvpt_b200_last_error(const rvpt_b200_ctx* ctx);

/* Pixel-tile partition for multi-GPU rendering: this ctx renders the 16x16
 * tiles with (tile_id % nranks) == rank, tile_id = ty * tiles_x + tx — the
 * reference's workgroup footprint (compute_pass.comp:27). Must be called
 * before the first frame; default is (0, 1). Global (x, y, W) still seed the
 * RNG (util.glsl:35-36), so the assembled image equals the 1-GPU image. */
RVPT_API int rvpt_b200_set_partition(rvpt_b200_ctx* ctx, int rank, int nranks);

/* Launch the frame kernels on a caller-owned CUDA stream (cudaStream_t) —
 * e.g. torch's current stream — instead of the ctx's own. */
RVPT_API int rvpt_b200_set_stream(rvpt_b200_ctx* ctx, void* cuda_stream);

/* ------------------------------------------------------------------------ */
/* Scene upload — replaces the bvh/triangle/material buffer copies           */
/* (rvpt.cpp:123-126) and the build in initialize() (rvpt.cpp:84-86).        */
/* `triangles` must already be in BVH-permuted order, as the reference       */
/* uploads `sorted_triangles`. If nodes == NULL a BVH is built internally     */
/* (rvpt_b200_build_bvh) and the triangles are permuted by it. Host pointers  */
/* are borrowed for the duration of the call only.                           */
/* ------------------------------------------------------------------------ */
RVPT_API int rvpt_b200_upload_scene(rvpt_b200_ctx* ctx, const rvpt_bvh_node* nodes, size_t n_nodes,
                                    const rvpt_triangle* triangles, size_t n_triangles,
                                    const rvpt_material* materials, size_t n_materials);

/* ------------------------------------------------------------------------ */
/* One frame — replaces the settings/camera uniform copies                   */
/* (rvpt.cpp:118-120) + record_compute_command_buffer + submit               */
/* (rvpt.cpp:350-354, 1005-1039). `camera` is the 80-byte block of           */
/* Camera::get_data(). Stateless w.r.t. the frame counter: uses              */
/* settings->current_frame as given.                                         */
/* ------------------------------------------------------------------------ */
RVPT_API int rvpt_b200_render_frame(rvpt_b200_ctx* ctx, const rvpt_render_settings* settings,
                                    const float camera[20]);
/* A progressive batch: the result of n_frames consecutive render_frame calls with
 * current_frame, current_frame+1, ... (the counter rule of rvpt.cpp:102-111 when nothing
 * changes), bit for bit — but only the images after the LAST frame are observable, which lets
 * the engine merge the frames' waves: with aa == 1 and integrator 9 everywhere, as many frames
 * as the path-queue budget allows (<= 64) share ONE kernel launch (scene staged once, one set
 * of grid barriers, samples folded into the running mean in frame order at the end).
 * No host synchronisation. Stats are those of the last launch (stats.frames frames). */
RVPT_API int rvpt_b200_render_frames(rvpt_b200_ctx* ctx, const rvpt_render_settings* settings,
                                     const float camera[20], uint32_t n_frames);
RVPT_API int rvpt_b200_sync(rvpt_b200_ctx* ctx);

/* ------------------------------------------------------------------------ */
/* Read-back (the reference has none: its images stay on the GPU).           */
/* All images are raster order, row 0 first, W*H pixels. With a partition,   */
/* pixels of tiles this rank does not own are returned as 0.                 */
/* ------------------------------------------------------------------------ */
/* result_image, rgba8 (binding 1): W*H*4 bytes, alpha = 0. */
RVPT_API int rvpt_b200_read_output_rgba8(rvpt_b200_ctx* ctx, uint8_t* dst);
/* temporal accumulation as float32 RGBA (w = 0): W*H*4 floats. In
 * ACCUM_RGBA8 mode this is the temporal image decoded as k/255. */
RVPT_API int rvpt_b200_read_accum_f32(rvpt_b200_ctx* ctx, float* dst);
/* Restore the temporal accumulation (checkpoint/resume). */
RVPT_API int rvpt_b200_write_accum_f32(rvpt_b200_ctx* ctx, const float* src);
/* Zero the temporal accumulation and the output image. */
RVPT_API int rvpt_b200_reset_accum(rvpt_b200_ctx* ctx);
/* Counters of the most recent frame (synchronises). */
RVPT_API int rvpt_b200_get_stats(rvpt_b200_ctx* ctx, rvpt_b200_stats* out);

/* Per-kernel device timing (the reference's only counter is the wall-clock
 * Timer around draw(), rvpt.cpp:348,403). While enabled every kernel launch is
 * bracketed by CUDA events on the ctx stream; get_kernel_times() synchronises,
 * returns the sums since the last call and clears them. Serialises nothing,
 * but the event records cost a little: keep it off for throughput runs. */
typedef struct rvpt_b200_kernel_times
{
    double primary_ms;      /* k_frame (whole frame), or k_primary when UNFUSED */
    double bounce_ms;       /* k_bounce launches (UNFUSED only) */
    uint32_t primary_launches;
    uint32_t bounce_launches;
} rvpt_b200_kernel_times;
RVPT_API int rvpt_b200_set_profiling(rvpt_b200_ctx* ctx, int enabled);
RVPT_API int rvpt_b200_get_kernel_times(rvpt_b200_ctx* ctx, rvpt_b200_kernel_times* out);

/* Per-CTA phase time stamps (%globaltimer, ns) of the most recent frame kernel:
 * out[cta * n_slots + k]; k = 0 kernel entry, 1 scene staged, 2 primary wave
 * done (first warp of the CTA), then for every bounce wave b: 2b+1 = past the
 * grid barrier that precedes it, 2b+2 = wave done. 0 = phase not reached.
 * For locating idle time inside the persistent kernel; off by default. */
RVPT_API int rvpt_b200_set_timeline(rvpt_b200_ctx* ctx, int enabled);
RVPT_API int rvpt_b200_get_timeline(rvpt_b200_ctx* ctx, uint64_t* out, size_t capacity,
                                    uint32_t* n_ctas, uint32_t* n_slots);

/* ------------------------------------------------------------------------ */
/* Device-side tile buffers, for the multi-GPU gather (NCCL runs in the      */
/* caller: torch.distributed). Layout: [n_local_tiles_padded][256] pixels,   */
/* local tile j = global tile j*nranks + rank, pixel order inside a tile:    */
/* idx = warp*32 + lane, warp = (py/4)*2 + (px/8), lane = (py%4)*8 + (px%8). */
/* n_local_tiles_padded = ceil(n_tiles / nranks) on every rank.              */
/* ------------------------------------------------------------------------ */
typedef struct rvpt_b200_tile_info
{
    uint32_t width, height;
    uint32_t tiles_x, tiles_y;
    uint32_t rank, nranks;
    uint32_t n_local_tiles;        /* tiles owned by this rank */
    uint32_t n_local_tiles_padded; /* equal on all ranks */
    void* d_accum_tiles;           /* float4 per pixel (uchar4 in ACCUM_RGBA8 mode) */
    void* d_rgba8_tiles;           /* uchar4 per pixel */
    uint64_t accum_bytes;          /* of the padded buffer */
    uint64_t rgba8_bytes;
} rvpt_b200_tile_info;
RVPT_API int rvpt_b200_get_tile_info(rvpt_b200_ctx* ctx, rvpt_b200_tile_info* out);

/* Redirect the per-tile outputs into caller-owned device memory (e.g. this
 * rank's slot of an all-gather buffer) so the frame kernels write the gather
 * payload in place and no pack pass exists. Pass NULL to keep the current
 * buffer. Sizes as reported by rvpt_b200_get_tile_info(). Takes effect for
 * frames launched after the call and does not synchronise: the caller orders
 * the reuse of a buffer against whatever still reads it (e.g. an in-flight
 * gather) with stream/event dependencies. */
RVPT_API int rvpt_b200_set_external_tiles(rvpt_b200_ctx* ctx, void* d_accum_tiles,
                                          void* d_rgba8_tiles);

/* Scatter gathered tile buffers ([nranks][n_local_tiles_padded][256]) into a
 * raster image on the device (the `k_untile` step on the gathering rank).
 * elem_bytes is 4 (rgba8) or 16 (float4). Runs on the ctx stream. */
RVPT_API int rvpt_b200_untile(rvpt_b200_ctx* ctx, const void* d_gathered, void* d_raster,
                              uint32_t elem_bytes, uint32_t nranks);
/* Same, on an explicit CUDA stream (so the scatter can follow an asynchronous
 * gather without blocking the render stream). */
RVPT_API int rvpt_b200_untile_on(rvpt_b200_ctx* ctx, const void* d_gathered, void* d_raster,
                                 uint32_t elem_bytes, uint32_t nranks, void* cuda_stream);

/* ------------------------------------------------------------------------ */
/* Fused gather over NVLink peer memory. Instead of collecting tiles with a  */
/* collective after the frame, the frame kernels of every rank store their    */
/* finished rgba8 pixels straight into ONE raster image that lives on the    */
/* display rank's GPU (peer-mapped through CUDA IPC): the "gather" is the     */
/* kernels' own 32-byte row stores travelling over NVLink while they compute. */
/* The display rank exports its image, the others attach to it; all a frame  */
/* needs afterwards is stream completion on every rank + a barrier.          */
/* ------------------------------------------------------------------------ */
#define RVPT_B200_IPC_HANDLE_BYTES 64
/* Display rank: returns the CUDA IPC handle of this ctx's raster result image
 * (W*H*4 bytes), creating it if the ctx is partitioned; its own tiles are then
 * written into it as well. */
RVPT_API int rvpt_b200_export_output(rvpt_b200_ctx* ctx, unsigned char handle[64]);
/* Other ranks (separate processes on the same node): map the display rank's
 * image and write this rank's pixels into it. */
RVPT_API int rvpt_b200_attach_output(rvpt_b200_ctx* ctx, const unsigned char handle[64]);

/* Double-buffered result image: the read-back of one step overlaps the frames of the next.
 * read_output_rgba8_async() enqueues the device -> host copy of the image the last frames wrote
 * (dst should be pinned host memory) on an internal copy stream, behind those frames, and
 * returns at once; frames launched afterwards fill the ctx's SECOND raster image, so nothing
 * overwrites the one in flight. wait_output() blocks until the last such copy has landed.
 * Needs a ctx that owns the raster image (one GPU, or the display rank). With the fused
 * multi-GPU gather the display rank exports both images (export_output + export_output2), the
 * other ranks attach both (attach_output + attach_output2) and call flip_output() once per
 * step, in step with the display rank's read_output_rgba8_async(). */
RVPT_API int rvpt_b200_read_output_rgba8_async(rvpt_b200_ctx* ctx, uint8_t* dst);
RVPT_API int rvpt_b200_wait_output(rvpt_b200_ctx* ctx);
RVPT_API int rvpt_b200_flip_output(rvpt_b200_ctx* ctx);
RVPT_API int rvpt_b200_export_output2(rvpt_b200_ctx* ctx, unsigned char handle[64]);
RVPT_API int rvpt_b200_attach_output2(rvpt_b200_ctx* ctx, const unsigned char handle[64]);

/* ------------------------------------------------------------------------ */
/* Host-side utilities (no GPU needed).                                      */
/* ------------------------------------------------------------------------ */
/* Binned-SAH BVH builder producing the reference's node format — the
 * counterpart of BinnedBvhBuilder::build_bvh (bvh_builder.cpp:11-199) with
 * its two defects fixed. nodes_out must hold 2*n-1 nodes,
 * prim_indices_out n indices (the permutation: sorted[i] =
 * triangles[prim_indices_out[i]], bvh.h:70-77). */
RVPT_API int rvpt_b200_build_bvh(const rvpt_triangle* triangles, size_t n_triangles,
                                 rvpt_bvh_node* nodes_out, size_t* n_nodes_out,
                                 uint32_t* prim_indices_out);

/* The same contract, built on the GPU (rvpt_b200/csrc/bvh_gpu.cu): a linear BVH — Morton codes of
 * the centroids, radix sort, Karras hierarchy, bottom-up box fit — one triangle per leaf,
 * 2n - 1 nodes, in milliseconds (the host builder above takes 0.9 s for 560 k triangles; its SAH
 * tree traverses faster). Host pointers in and out; build_ms_out (may be NULL) receives the
 * device time of the build kernels. */
RVPT_API int rvpt_b200_build_bvh_gpu(int device, const rvpt_triangle* triangles, size_t n_triangles,
                                     rvpt_bvh_node* nodes_out, size_t* n_nodes_out,
                                     uint32_t* prim_indices_out, float* build_ms_out);

/* Camera::get_data() (camera.cpp:17-25, 55-66) without glm: translation,
 * rotation in degrees (rotation.x about UP, .y about RIGHT, .z about FORWARD),
 * fov in degrees. Writes the 80-byte block. */
RVPT_API void rvpt_b200_camera_data(const float translation[3], const float rotation_deg[3],
                                    float aspect, float fov_deg, float scale, float out[20]);

/* Host only: 1 if two triangles of the list are coplanar and overlap with positive area. */
RVPT_API int rvpt_b200_has_coincident_faces(const rvpt_triangle* triangles, size_t n_triangles);

/* Test / tooling hook (host only): the eight front-to-back octant node arrays the engine
 * uploads next to a scene (device_scene.h). out = 8*n float4 (near x, far x, near y, far y)
 * followed by 8*n float4 (near z, far z, skip link, first triangle | 0xFFFFFFFF) with
 * n = *n_packed_nodes; pass out = NULL to query n. */
RVPT_API int rvpt_b200_octant_layouts(const rvpt_bvh_node* nodes, size_t n_nodes,
                                      const rvpt_triangle* triangles, size_t n_triangles, float* out,
                                      size_t capacity_floats, size_t* n_packed_nodes);

/* ABI / build self-description. */
RVPT_API uint32_t rvpt_b200_abi_version(void);
RVPT_API const char* rvpt_b200_build_info(void);

/* Evaluates the shared arithmetic header on the device for n inputs
 * (test hook: proves device == host bit-for-bit). op: 0 sincos(x) -> (s,c),
 * 1 rand stream from seed bits -> 2 floats, 2 normalize(x,y,z). */
RVPT_API int rvpt_b200_selftest_math(int device, int op, const float* in, size_t n, float* out);

#ifdef __cplusplus
}
#endif

#endif /* RVPT_ABI_H */
