/*
 * device_scene.h — how the scene, the path state and the frame constants are
 * laid out in HBM / shared memory. Shared by the host-side packer (engine.cu)
 * and the kernels (kernels.cu).
 */
#ifndef RVPT_DEVICE_SCENE_H
#define RVPT_DEVICE_SCENE_H

#include <stdint.h>

#define RVPT_TILE_DIM 16u          /* reference workgroup footprint, compute_pass.comp:27 */
#define RVPT_TILE_PIXELS 256u
#define RVPT_NODE_END 0xFFFFFFFFu  /* traversal finished */
#define RVPT_NODE_INNER 0x80000000u /* DevNode::leaf_first of an inner node: this bit | index of its second child */
#define RVPT_TRI_LAST 0x80000000u  /* meta bit: last triangle of its leaf */
/* Binned path queues (closed scenes): a queue is RVPT_SORT_BINS sub-queues — one per direction
 * octant (3 bits) x origin cell (RVPT_SORT_CELL_BITS per axis of the scene's bounding box) —
 * of bin_cap entries each, plus one overflow / unsorted sub-queue that can hold every path. The
 * rays a warp of the next wave loads together (32 consecutive entries of one sub-queue) then walk
 * the same octant's node array. Measured on the Cornell box (profiles/r02_sort_experiments.md):
 * octants alone (CELL_BITS 0) gain 9-12 %; adding 2x2x2 origin cells — which the lockstep model
 * of tools/bvh_cost.py rated at -17 % per later wave — gains nothing there and doubles the
 * primary wave's push time, so the default is 0. */
#ifndef RVPT_SORT_CELL_BITS
#define RVPT_SORT_CELL_BITS 0u
#endif
#define RVPT_SORT_BINS (8u << (3u * RVPT_SORT_CELL_BITS))       /* 8 */
/* batched launches (render_frames): a queued path carries `slot | frame_in_batch << 26` */
#define RVPT_BATCH_SLOT_BITS 26u
#define RVPT_BATCH_SLOT_MASK ((1u << RVPT_BATCH_SLOT_BITS) - 1u)
#define RVPT_MAX_BATCH 64u          /* frames per launch (6 tag bits) */
/* octant node copies in shared memory: byte distance between the two float4 halves of a record */
#define RVPT_OCT_B_OFFSET (100u * 1024u)
#define RVPT_MAX_GROUPS 27 /* frame groups of a batched primary wave (FrameParams::group_start) */
#define RVPT_LIST_WORDS 36u /* u16 per pixel block in FrameParams::leaf_lists */
#define RVPT_LIST_NONE 0xFFFFu /* no list for this block: its rays walk the tree */
#define RVPT_TIMELINE_SLOTS 16u     /* per-CTA phase stamps of the last frame kernel */

/*
 * BVH node, 32 B (two 128-bit loads). Nodes are re-laid-out in the order the
 * reference's stack walk visits them (intersection.glsl:361-413: push
 * first+1, descend into first), so the first child of an inner node is always
 * node+1 and the walk needs no stack: `skip` is where the reference's
 * `stack[--stack_ptr]` pop would land after this subtree.
 */
struct DevNode
{
    float bmin_x, bmax_x, bmin_y, bmax_y; /* bounds[0..3] */
    float bmin_z, bmax_z;                 /* bounds[4..5] */
    uint32_t skip;                        /* next node when this subtree is done / missed */
    uint32_t leaf_first;                  /* first DevTri of a leaf; inner node: RVPT_NODE_INNER | second child
                                           * (the first child is node + 1) */
};

/*
 * Triangle record, 4 x float4: everything intersect_triangle_fast
 * (intersection.glsl:267-323) recomputes per ray but that only depends on the
 * triangle, evaluated once at upload with the same unfused float32 operations:
 *   a = v0.xyz, inv_det          inv_det = 1/(A00*A11 - A01*A10)
 *   b = n.xyz,  A00              n = cross(e0,e1), A00 = dot(e1,e1)
 *   c = e0.xyz, A01              A01 = A10 = -dot(e0,e1)
 *   d = e1.xyz, A11              A11 = dot(e0,e0)
 * Records are stored leaf by leaf in walk order. A second, 16-byte record per triangle
 * (DevTriMeta) holds what shading needs of the hit triangle — normalize(n) of intersect_scene
 * (intersection.glsl:511), evaluated once at upload with rv_normalize — and meta = material
 * index | RVPT_TRI_LAST on the last triangle of a leaf.
 */
struct DevTri
{
    float v0x, v0y, v0z, inv_det;
    float nx, ny, nz, a00;
    float e0x, e0y, e0z, a01;
    float e1x, e1y, e1z, a11;
};

struct DevTriMeta
{
    float unx, uny, unz; /* rv_normalize(n) */
    uint32_t meta;       /* material index | RVPT_TRI_LAST */
};

/* Material, 3 x float4 (convert_old_material, intersection.glsl:45-57). */
struct DevMaterial
{
    float base_r, base_g, base_b, ior;       /* albedo.xyz, albedo.w */
    float emis_r, emis_g, emis_b;            /* emission.xyz */
    int32_t type;                            /* int(data.x) */
    float lam_r, lam_g, lam_b, pad;          /* (base*INV_PI)*PI, integrators.glsl:622 */
};

/*
 * The whole scene is one contiguous 16-byte-aligned blob so one TMA bulk copy
 * (cp.async.bulk) stages it into shared memory:
 *   [DevNode x n_nodes][DevTri x n_tris][DevTriMeta x n_tris][DevMaterial x n_mats]
 *   [ordered octant node arrays, optional]
 */
struct SceneLayout
{
    uint32_t n_nodes, n_tris, n_mats;
    uint32_t off_tris;  /* byte offsets inside the blob */
    uint32_t off_meta;
    uint32_t off_mats;
    uint32_t bytes;     /* of the part above (what the plain kernels stage), multiple of 16 */
    /* Optional block behind it (0 = absent): the eight direction-octant node arrays in
     * FRONT-TO-BACK order — for octant k every inner node's children are laid out so that the
     * one lying earlier along the ray direction is visited first (its own pre-order numbering
     * and skip links; leaves keep their DevTri ranges). Two float4 arrays of 8 * n_nodes
     * entries: (near x, far x, near y, far y) then (near z, far z, skip, leaf). */
    uint32_t off_oct;
};

/*
 * Path state: 4 separate float4 streams (SoA), 64 B per path, indexed by queue
 * slot so every warp moves 4 x 512 contiguous bytes.
 *   q0 = origin.xyz,     accumulation slot (uint bits)
 *   q1 = direction.xyz,  rng state (uint bits)
 *   q2 = throughput.xyz, unused
 *   q3 = radiance (col).xyz, unused
 */
struct PathQueue
{
    float4* q0;
    float4* q1;
    float4* q2;
    float4* q3;
};

/* Per wave, in shared memory: how the wave's groups of L rays (32, or fewer when a small wave is
 * spread over all warps) map to the sub-queues. pre[k] = first group of sub-queue k,
 * pre[BINS + 1] = number of groups, cnt[k] = rays in sub-queue k (k = BINS: overflow / unsorted). */
struct WaveGroups
{
    uint32_t pre[RVPT_SORT_BINS + 2];
    uint32_t cnt[RVPT_SORT_BINS + 1];
    uint32_t count; /* rays of the wave */
    uint32_t L;
};

/*
 * Device counters. Nothing here depends on host-side launch parity, so a CUDA graph that
 * captured any number of launches replays correctly: the LAST CTA to leave a launch
 * (done_ctas) re-zeroes the wave counters for the next launch and, when the launch completes
 * a frame (or a batch of frames), publishes the per-bounce ray counts into `last` and zeroes
 * `stats`. `last` is what the host reads (rvpt_b200_get_stats) and what the next launch's
 * wave-size forecast looks at.
 */
#define RVPT_CHUNK_SHARDS 16u
#define RVPT_QCOUNT_STRIDE 32u /* words between the sub-queue counters of a wave: one 128-byte line each */
struct WaveCounters
{
    /* primary phase work distribution: one counter per shard, 128 B apart so the
     * shards live in different L2 atomic units; shard k hands out chunks k, k+16, ... */
    uint32_t chunk_ctr[RVPT_CHUNK_SHARDS * 32u];
    uint32_t work_ctr[64]; /* k_bounce (one launch per wave): the claimed eighth of wave b */
    /* k_frame's big bounce waves: sharded like chunk_ctr, wave b uses set b & 1 */
    uint32_t bounce_ctr[2][RVPT_CHUNK_SHARDS * 32u];
    /* survivors pushed by bounce b (read by b+1) per sub-queue: [b][k] for sorted appends and the
     * overflow sub-queue (k = RVPT_SORT_BINS), [b][(1 + k) * RVPT_QCOUNT_STRIDE] for unsorted appends
     * spread over the binned sub-queues (a wave uses one of the two; readers add them). A binned
     * sub-queue's counter may run past bin_cap (the excess went to the overflow sub-queue): readers
     * clamp it */
    uint32_t qcount[64][(RVPT_SORT_BINS + 1) * RVPT_QCOUNT_STRIDE];
};
struct FrameStats
{
    unsigned long long active[64]; /* rays traced at bounce b */
};
struct FrameCounters
{
    WaveCounters wave;   /* all zero between launches */
    FrameStats stats;    /* the frame (all aa passes) or batch in flight; zero between frames */
    FrameStats last;     /* the last completed frame / batch */
    uint32_t last_sets;  /* full-image sample sets `last` covers: aa, or the frames of a batch */
    uint32_t done_ctas;  /* CTAs that have left the current launch */
};

/* Per-frame constants, passed by value as a kernel parameter. */
struct FrameParams
{
    /* image / partition */
    uint32_t W, H;                /* full image */
    uint32_t W_eff, H_eff;        /* rendered extent (REFERENCE_DISPATCH rounds down to 16) */
    uint32_t tiles_x, tiles_y, n_tiles;
    uint32_t rank, nranks;
    uint32_t n_local_tiles;
    uint32_t n_chunks;            /* n_local_tiles * 8 warp chunks */
    /* division by tiles_x / n_chunks as multiply-high (two divisions per 32-pixel chunk were 3 %
     * of the frame kernel's instructions): q = (v * magic) >> 40, exact while v * d < 2^40;
     * 0 = the operands of this launch could exceed that, divide */
    unsigned long long tiles_x_magic, n_chunks_magic;
    uint32_t flags;
    /* compute_pass.comp:50-54 */
    float inv_dim_x, inv_dim_y;
    uint32_t frame;
    float frame_f;                /* float(current_frame) */
    float inv_frame1;             /* 1/float(current_frame+1) */
    float keep;                   /* float(min(current_frame, 1)) */
    /* render settings */
    int32_t max_bounces;
    int32_t aa;
    float aa_f;
    int32_t pass;                 /* sample index inside the frame, 0..aa-1 */
    int32_t camera_mode;
    int32_t modes[4];             /* tl, tr, bl, br */
    int32_t all_kajiya;           /* all four quadrants use integrator 9 */
    float split_x, split_y;
    /* camera (camera.glsl) */
    float cam[16];                /* column-major */
    float aspect, hfov, scale;
    float inv_tan_half_fov;       /* w = 1/tan(0.5*hfov), camera.glsl:44 */
    /* buffers */
    const unsigned char* scene;   /* blob */
    SceneLayout layout;
    PathQueue queue[2];
    float4* accum_f32;            /* tile layout, float mode */
    uchar4* accum_u8;             /* tile layout, ACCUM_RGBA8 mode */
    uchar4* out_tiles;            /* rgba8 result, tile layout (partitioned) */
    uchar4* out_raster;           /* rgba8 result, raster (nranks == 1) */
    float4* carry;                /* per-slot (sum.xyz, rng) between aa passes */
    FrameCounters* ctr;
    uint32_t last_of_pass;        /* this launch is the last one that uses the wave counters (fused: always) */
    uint32_t last_of_frame;       /* ... and the last one of the frame / batch: publish the stats */
    /* batched launch (rvpt_b200_render_frames): frames frame .. frame + n_batch - 1 in one launch.
     * Samples are not folded into the running mean as they finish — frames of one pixel must be
     * folded in order — but parked in samples[frame_in_batch * sample_stride + slot]; the resolve
     * phase at the end of the launch folds them in frame order (compute_pass.comp:161-163). */
    uint32_t n_batch;             /* 0 = classic single-frame launch */
    uint32_t sample_stride;       /* slots per frame in `samples` */
    float4* samples;
    uint32_t tail_threshold;      /* waves this small finish inside their threads */
    uint32_t use_forecast;        /* wave-size forecast from the previous launch is meaningful */
    /* binned path queues (device_scene.h, RVPT_SORT_*): entries per binned sub-queue (0: the
     * queues have only the unsorted sub-queue); the unsorted / overflow sub-queue starts at
     * RVPT_SORT_BINS * bin_cap */
    uint32_t bin_cap;
    float sort_lo[3];             /* scene bounding box (root node) */
    float sort_scale[3];          /* cells per unit: (1 << RVPT_SORT_CELL_BITS) / extent */
    unsigned long long* timeline; /* optional [n_ctas][RVPT_TIMELINE_SLOTS] globaltimer stamps */
    /* integrator_Hart (render mode 10) only: the caller's 64-byte triangle records in upload order */
    const float4* raw_tris;
    uint32_t n_raw_tris;
    /* batched launches of shared-memory scenes with octant arrays, pinhole camera: a first phase
     * lists, per 8x4 pixel block, the leaves the block's beam of primary rays can enter
     * (leaf_lists: RVPT_LIST_WORDS u16 per block: count or RVPT_LIST_NONE, octant, 2 unused, 32
     * node offsets), and the primary wave claims (pixel block, frame group) units (kernels.cu,
     * primary_phase_beam); group g covers the frames group_start[g] .. group_start[g + 1] - 1 of
     * the batch. n_groups == 0: the primary wave walks the tree for every ray. */
    uint32_t n_groups;
    uint8_t group_start[RVPT_MAX_GROUPS + 1];
    unsigned short* leaf_lists;
};

#endif
