/*
 * engine.cu — the C-ABI of include/rvpt_abi.h: context, scene packing, frame
 * orchestration, read-backs. This is the code that sits where RVPT::update()
 * copies its per-frame buffers (src/rvpt/rvpt.cpp:118-126) and RVPT::draw()
 * records + submits the compute dispatch (rvpt.cpp:350-354, 1005-1039).
 *
 * There is no CPU fallback: every entry point that renders needs a CUDA
 * device and fails with RVPT_B200_ECUDA otherwise.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rvpt_abi.h"
#include "../../include/rvpt_math.h"
#include "device_scene.h"
#include "kernels.h"

static_assert(sizeof(rvpt_render_settings) == 40, "RenderSettings must stay 40 bytes");
static_assert(sizeof(rvpt_camera_data) == 80, "camera block must stay 80 bytes");
static_assert(sizeof(rvpt_bvh_node) == 32, "BvhNode must stay 32 bytes");
static_assert(sizeof(rvpt_triangle) == 64, "Triangle must stay 64 bytes");
static_assert(sizeof(rvpt_material) == 48, "Material must stay 48 bytes");
static_assert(sizeof(DevNode) == 32 && sizeof(DevTri) == 64 && sizeof(DevMaterial) == 48,
              "device records are 2/4/3 float4");

#define RVPT_ABI_VERSION 2u
/* scenes whose blob fits this budget are staged into shared memory per CTA */
#define RVPT_SMEM_SCENE_LIMIT (224u * 1024u) /* one 1024-thread CTA per SM owns the SM's shared memory (227 KB opt-in max) */

struct rvpt_b200_ctx
{
    int device = 0;
    uint32_t W = 0, H = 0, flags = 0;
    uint32_t tiles_x = 0, tiles_y = 0, n_tiles = 0;
    uint32_t rank = 0, nranks = 1;
    uint32_t n_local_tiles = 0, n_local_padded = 0;
    int num_sms = 0;
    int l2_persist_max = 0; /* cudaDevAttrMaxPersistingL2CacheSize */
    int l2_window_max = 0;  /* cudaDevAttrMaxAccessPolicyWindowSize */
    int grid_frame = 0, grid_primary = 0, grid_bounce = 0;
    uint32_t bin_cap = 0;            /* entries per binned sub-queue (0: unsorted sub-queue only) */
    float sort_lo[3] = {0, 0, 0}, sort_scale[3] = {0, 0, 0}; /* scene box -> origin cells; scale 0: no sorting */
    uint32_t batch_limit = 0;  /* set when an allocation failed: largest batch worth trying */
    uint32_t batch_cap = 0;    /* frames per launch the queues and the sample buffer hold (>= 1 once allocated) */
    size_t queue_budget = (size_t)64 << 30; /* bytes the path queues may take (of 180 GB) */
    uint32_t tail_rays_per_warp = 16; /* waves up to this many rays per resident warp finish in-thread (measured: 2..8 equal, 16 saves a barrier + wave on sparse poses) */
    uint32_t frame_seq = 0;  /* frames / batches launched so far (the forecast needs one behind it) */

    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;

    /* scene */
    unsigned char* d_scene = nullptr;
    unsigned char* h_scene_pinned = nullptr; /* staging copy of the blob */
    size_t scene_capacity = 0;
    cudaEvent_t scene_copied = nullptr;
    size_t scene_blob_bytes = 0;             /* bytes of the packed blob in h_scene_pinned */
    std::vector<unsigned char> scene_inputs; /* the caller's arrays of the last upload (exact compare) */
    SceneLayout layout{};
    bool scene_smem = false;
    bool scene_oct = false; /* direction-octant node copies fit next to the blob */
    bool scene_nested = false; /* every child box lies inside its parent's, all bounds finite (leaf lists) */
    uint32_t frame_group = 16; /* most frames per (pixel block, frame group) unit of a batched primary wave; 0: no leaf lists */
    bool frame_group_fixed = false; /* RVPT_B200_FRAME_GROUP: exactly that many */
    bool have_scene = false;
    /* integrator_Hart (render mode 10) marches against the caller's vertices, not the packed
     * records: the 64-byte triangles in the order the shader's buffer holds them (the caller's, or
     * BVH-permuted when the BVH was built here), copied to the device the first time a frame asks
     * for that integrator */
    std::vector<rvpt_triangle> raw_sorted; /* only when upload_scene built the BVH itself */
    float4* d_raw_tris = nullptr;
    size_t raw_capacity = 0;   /* triangles d_raw_tris can hold */
    bool raw_valid = false;    /* d_raw_tris holds the current scene */
    bool raw_source_ok = false; /* scene_inputs / raw_sorted are the arrays of the scene on the device (the last upload succeeded) */

    /* frame buffers (tile layout) */
    void* d_accum = nullptr;      /* own allocation */
    void* d_out_tiles = nullptr;  /* own allocation */
    void* accum = nullptr;        /* in use (own or external) */
    void* out_tiles = nullptr;
    /* The raster result image exists twice ("front" / "back", the second one on demand) so that an
     * asynchronous read-back of one can overlap the frames that fill the other
     * (rvpt_b200_read_output_rgba8_async / flip_output). [0] is the image every other entry point
     * means. */
    uchar4* d_out_raster = nullptr;  /* [out_cur == 0]: nranks == 1, or exported by the display rank */
    uchar4* d_out_raster2 = nullptr; /* [out_cur == 1] */
    uchar4* peer_out_raster = nullptr;  /* display rank's images, mapped through CUDA IPC */
    uchar4* peer_out_raster2 = nullptr;
    int out_cur = 0;  /* image the next frames write */
    int out_last = 0; /* image the last frames wrote (what read_output_rgba8 returns) */
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rendered = nullptr, ev_copied = nullptr;
    bool copy_pending = false;
    float4* d_carry = nullptr;      /* allocated on first aa > 1 */
    float4* d_samples = nullptr;    /* batched launches: [batch_cap][slots] parked samples */
    unsigned short* d_leaf_lists = nullptr; /* batched launches: RVPT_LIST_WORDS u16 per 8x4 pixel block, on first use */
    PathQueue queue[2]{};
    FrameCounters* d_ctr = nullptr;
    void* d_scratch = nullptr; /* raster-sized float4 staging for read-backs */
    bool buffers_ready = false;
    bool frame_rendered = false;
    int last_max_bounces = 0;
    int last_aa = 0;
    uint32_t last_launches = 0;
    uint32_t call_launches = 0; /* launches of the render_frame(s) call in progress */
    uint32_t last_frames = 0; /* frames the last launch covered (stats) */

    /* per-CTA phase stamps of the last frame kernel (set_timeline) */
    unsigned long long* d_timeline = nullptr;
    int timeline_ctas = 0;

    /* per-kernel timing */
    bool profiling = false;
    struct Timed
    {
        cudaEvent_t a, b;
        int kind; /* 0 primary, 1 bounce */
    };
    std::vector<Timed> timed;      /* recorded, not yet collected */
    std::vector<Timed> event_pool; /* recycled */

    std::string err;
};

namespace
{

int fail(rvpt_b200_ctx* ctx, int code, const char* fmt, ...)
{
    if (ctx)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

#define CU(call)                                                                              \
    do                                                                                        \
    {                                                                                         \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, RVPT_B200_ECUDA, "%s failed: %s (%s:%d)", #call,                 \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)

size_t accum_elem_bytes(const rvpt_b200_ctx* ctx)
{
    return (ctx->flags & RVPT_B200_FLAG_ACCUM_RGBA8) ? 4u : 16u;
}

void free_queues(rvpt_b200_ctx* ctx)
{
    for (int i = 0; i < 2; ++i)
    {
        cudaFree(ctx->queue[i].q0);
        cudaFree(ctx->queue[i].q1);
        cudaFree(ctx->queue[i].q2);
        cudaFree(ctx->queue[i].q3);
        ctx->queue[i] = PathQueue{};
    }
    cudaFree(ctx->d_samples);
    ctx->d_samples = nullptr;
    ctx->bin_cap = 0;
    ctx->batch_cap = 0;
}

void free_frame_buffers(rvpt_b200_ctx* ctx)
{
    cudaFree(ctx->d_accum);
    cudaFree(ctx->d_out_tiles);
    cudaFree(ctx->d_out_raster);
    cudaFree(ctx->d_out_raster2);
    cudaFree(ctx->d_carry);
    cudaFree(ctx->d_ctr);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_leaf_lists);
    ctx->d_leaf_lists = nullptr;
    free_queues(ctx);
    ctx->d_accum = ctx->d_out_tiles = nullptr;
    ctx->accum = ctx->out_tiles = nullptr;
    ctx->d_out_raster = nullptr;
    ctx->d_out_raster2 = nullptr;
    ctx->d_carry = nullptr;
    ctx->d_ctr = nullptr;
    ctx->d_scratch = nullptr;
    ctx->buffers_ready = false;
}

/* Frames one launch may cover with the path queues inside their memory budget. A queue holds
 * 64 B per path; binned (closed-scene ray sorting, device_scene.h) it is RVPT_SORT_BINS
 * sub-queues of an eighth of all paths each plus the overflow sub-queue that can hold them
 * all: (BINS / 8 + 1) x 64 B per pixel and frame, two queues, + 16 B parked sample —
 * 2.4 GB per 1080p frame of a batch, 9.7 GB per 4K frame (of 180 GB). */
#define RVPT_BIN_SHARE 8u /* a binned sub-queue holds 1/8 of a launch's paths */
size_t queue_bytes_per_entry(const rvpt_b200_ctx* ctx)
{
    const bool sorted = !(ctx->flags & (RVPT_B200_FLAG_NO_QUEUE_SORT | RVPT_B200_FLAG_UNFUSED));
    return 2u * 64u * (sorted ? RVPT_SORT_BINS / RVPT_BIN_SHARE + 1u : 1u) + 16u;
}

uint32_t max_batch(const rvpt_b200_ctx* ctx)
{
    const size_t slots = (size_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    if (slots == 0 || slots > RVPT_BATCH_SLOT_MASK) return 1;
    size_t cap = ctx->queue_budget / (slots * queue_bytes_per_entry(ctx));
    cap = std::min<size_t>(cap, RVPT_MAX_BATCH);
    cap = std::min<size_t>(cap, ((size_t)1 << 28) / slots); /* queue indices are 32-bit, binned queues 9x */
    if (ctx->batch_limit) cap = std::min<size_t>(cap, ctx->batch_limit); /* what the device could give */
    return (uint32_t)std::max<size_t>(cap, 1);
}

/* (Re)allocates the path queues and the parked-sample buffer for launches of up to `frames`
 * frames. */
int ensure_batch_capacity(rvpt_b200_ctx* ctx, uint32_t frames)
{
    if (frames <= ctx->batch_cap) return 0;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    free_queues(ctx);
    const size_t slots = (size_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    const size_t per_queue = std::max<size_t>(slots * frames, 1); /* paths of one launch */
    /* binned while that stays below 16 GiB (single frames of huge images keep one sub-queue) */
    bool sorted = !(ctx->flags & (RVPT_B200_FLAG_NO_QUEUE_SORT | RVPT_B200_FLAG_UNFUSED));
    const size_t bin_cap = ((per_queue + RVPT_BIN_SHARE - 1) / RVPT_BIN_SHARE + 31) & ~(size_t)31;
    size_t entries = per_queue + RVPT_SORT_BINS * bin_cap;
    if (sorted && (entries * 128u > std::max(ctx->queue_budget, (size_t)16 << 30) || entries >= ((size_t)1 << 32)))
        sorted = false;
    if (!sorted) entries = per_queue;
    ctx->bin_cap = sorted ? (uint32_t)bin_cap : 0u;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i)
    {
        if (e == cudaSuccess) e = cudaMalloc(&ctx->queue[i].q0, entries * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&ctx->queue[i].q1, entries * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&ctx->queue[i].q2, entries * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&ctx->queue[i].q3, entries * sizeof(float4));
    }
    if (e == cudaSuccess && frames > 1) e = cudaMalloc(&ctx->d_samples, per_queue * sizeof(float4));
    if (e == cudaErrorMemoryAllocation && frames > 1)
    {
        /* the device cannot give that much right now: halve the batch (render_frames re-plans) */
        cudaGetLastError();
        free_queues(ctx);
        ctx->batch_limit = frames / 2u;
        return ensure_batch_capacity(ctx, frames / 2u) ? RVPT_B200_ENOMEM : RVPT_B200_ENOMEM;
    }
    if (e != cudaSuccess)
        return fail(ctx, e == cudaErrorMemoryAllocation ? RVPT_B200_ENOMEM : RVPT_B200_ECUDA,
                    "path queue allocation failed: %s", cudaGetErrorString(e));
    ctx->batch_cap = frames;
    return 0;
}

int ensure_frame_buffers(rvpt_b200_ctx* ctx)
{
    if (ctx->buffers_ready) return 0;
    CU(cudaSetDevice(ctx->device));
    const size_t slots = (size_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    CU(cudaMalloc(&ctx->d_accum, slots * accum_elem_bytes(ctx)));
    CU(cudaMalloc(&ctx->d_out_tiles, slots * 4));
    CU(cudaMemsetAsync(ctx->d_accum, 0, slots * accum_elem_bytes(ctx), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_out_tiles, 0, slots * 4, ctx->stream));
    if (!ctx->accum) ctx->accum = ctx->d_accum;
    if (!ctx->out_tiles) ctx->out_tiles = ctx->d_out_tiles;
    if (ctx->nranks == 1)
    {
        CU(cudaMalloc(&ctx->d_out_raster, (size_t)ctx->W * ctx->H * 4));
        CU(cudaMemsetAsync(ctx->d_out_raster, 0, (size_t)ctx->W * ctx->H * 4, ctx->stream));
    }
    CU(cudaMalloc(&ctx->d_ctr, sizeof(FrameCounters)));
    CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(FrameCounters), ctx->stream));
    ctx->buffers_ready = true;
    return ensure_batch_capacity(ctx, 1);
}

int ensure_scratch(rvpt_b200_ctx* ctx)
{
    if (ctx->d_scratch) return 0;
    /* raster float4 + tile-layout float4 staging */
    const size_t raster = (size_t)ctx->W * ctx->H * sizeof(float4);
    const size_t tiles = (size_t)ctx->n_local_padded * RVPT_TILE_PIXELS * sizeof(float4);
    CU(cudaMalloc(&ctx->d_scratch, raster + tiles));
    return 0;
}

void recompute_partition(rvpt_b200_ctx* ctx)
{
    ctx->tiles_x = (ctx->W + RVPT_TILE_DIM - 1) / RVPT_TILE_DIM;
    ctx->tiles_y = (ctx->H + RVPT_TILE_DIM - 1) / RVPT_TILE_DIM;
    ctx->n_tiles = ctx->tiles_x * ctx->tiles_y;
    ctx->n_local_padded = (ctx->n_tiles + ctx->nranks - 1) / ctx->nranks;
    /* tiles g = j*nranks + rank < n_tiles */
    ctx->n_local_tiles =
        ctx->rank < ctx->n_tiles ? (ctx->n_tiles - ctx->rank + ctx->nranks - 1) / ctx->nranks : 0;
}

/* ---- per-kernel timing ----------------------------------------------------- */

struct ScopedTimer
{
    rvpt_b200_ctx* ctx;
    rvpt_b200_ctx::Timed t{};
    bool on;
    ScopedTimer(rvpt_b200_ctx* c, int kind) : ctx(c), on(c->profiling)
    {
        if (!on) return;
        if (!ctx->event_pool.empty())
        {
            t = ctx->event_pool.back();
            ctx->event_pool.pop_back();
        }
        else if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess)
        {
            on = false;
            return;
        }
        t.kind = kind;
        cudaEventRecord(t.a, ctx->stream);
    }
    ~ScopedTimer()
    {
        if (!on) return;
        cudaEventRecord(t.b, ctx->stream);
        ctx->timed.push_back(t);
    }
};

/* ---- scene packing -------------------------------------------------------- */

struct PackedScene
{
    bool bounds_ordered = true; /* every node has min <= max on every axis (no NaN) */
    bool coincident_faces = false; /* has_coincident_faces(): the walk order can decide a hit */
    bool boxes_nested = false;     /* boxes_are_nested(): a ray that enters a leaf box enters every ancestor's */
    std::vector<DevNode> nodes;
    std::vector<DevTri> tris;
    std::vector<DevTriMeta> meta;
    std::vector<DevMaterial> mats;
};

DevTri make_dev_tri(const rvpt_triangle& t)
{
    /* the per-triangle part of intersect_triangle_fast, intersection.glsl:289-307 */
    const rv_f3 v0 = rv_make(t.vertex0[0], t.vertex0[1], t.vertex0[2]);
    const rv_f3 v1 = rv_make(t.vertex1[0], t.vertex1[1], t.vertex1[2]);
    const rv_f3 v2 = rv_make(t.vertex2[0], t.vertex2[1], t.vertex2[2]);
    const rv_f3 e0 = rv_sub(v1, v0);
    const rv_f3 e1 = rv_sub(v2, v0);
    const rv_f3 n = rv_cross(e0, e1);
    const float a00 = rv_dot(e1, e1);
    const float a01 = -rv_dot(e0, e1);
    const float a10 = -rv_dot(e0, e1);
    const float a11 = rv_dot(e0, e0);
    const float p0 = a00 * a11;
    const float p1 = a01 * a10;
    const float inv_det = 1.0f / (p0 - p1);
    DevTri d;
    d.v0x = v0.x, d.v0y = v0.y, d.v0z = v0.z, d.inv_det = inv_det;
    d.nx = n.x, d.ny = n.y, d.nz = n.z, d.a00 = a00;
    d.e0x = e0.x, d.e0y = e0.y, d.e0z = e0.z, d.a01 = a01;
    d.e1x = e1.x, d.e1y = e1.y, d.e1z = e1.z, d.a11 = a11;
    return d;
}

/* Re-lay the caller's BVH out in the order intersect_bvh's stack walk visits
 * it (intersection.glsl:361-413) and resolve every pop into a skip link. */
int pack_scene(rvpt_b200_ctx* ctx, const rvpt_bvh_node* nodes, size_t n_nodes,
               const rvpt_triangle* tris, size_t n_tris, const rvpt_material* mats, size_t n_mats,
               bool brute_force, PackedScene& out)
{
    for (size_t i = 0; i < n_tris; ++i)
    {
        const float m = tris[i].material_id[0];
        if (!(m >= 0.0f) || (size_t)(int)m >= n_mats)
            return fail(ctx, RVPT_B200_EINVAL, "triangle %zu: material index %g out of range [0,%zu)",
                        i, (double)m, n_mats);
    }
    out.mats.resize(n_mats);
    for (size_t i = 0; i < n_mats; ++i)
    {
        const rvpt_material& m = mats[i];
        DevMaterial& d = out.mats[i];
        d.base_r = m.albedo[0], d.base_g = m.albedo[1], d.base_b = m.albedo[2], d.ior = m.albedo[3];
        d.emis_r = m.emission[0], d.emis_g = m.emission[1], d.emis_b = m.emission[2];
        d.type = (int)m.data[0];
        /* throughput *= mat_eval_Lambert_cos(base_color*INV_PI) = (base*INV_PI)*PI */
        const float l0 = m.albedo[0] * RV_INV_PI, l1 = m.albedo[1] * RV_INV_PI,
                    l2 = m.albedo[2] * RV_INV_PI;
        d.lam_r = l0 * RV_PI, d.lam_g = l1 * RV_PI, d.lam_b = l2 * RV_PI, d.pad = 0.0f;
    }

    auto emit_leaf = [&](uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k)
        {
            out.tris.push_back(make_dev_tri(tris[first + k]));
            uint32_t m = (uint32_t)(int)tris[first + k].material_id[0];
            if (k + 1 == count) m |= RVPT_TRI_LAST;
            /* intersect_scene's normalize(n) (intersection.glsl:511) depends on the triangle only */
            const DevTri& t = out.tris.back();
            const rv_f3 un = rv_normalize(rv_make(t.nx, t.ny, t.nz));
            out.meta.push_back(DevTriMeta{un.x, un.y, un.z, m});
        }
    };

    if (brute_force || nodes == nullptr)
    {
        /* one leaf with unbounded extent: every triangle in upload order */
        DevNode root;
        const float inf = INFINITY;
        root.bmin_x = -inf, root.bmax_x = inf, root.bmin_y = -inf, root.bmax_y = inf;
        root.bmin_z = -inf, root.bmax_z = inf;
        root.skip = RVPT_NODE_END;
        root.leaf_first = 0;
        out.nodes.push_back(root);
        emit_leaf(0, (uint32_t)n_tris);
        return 0;
    }

    if (n_nodes == 0) return fail(ctx, RVPT_B200_EINVAL, "BVH has no nodes");

    /* Pre-order walk, first child first — exactly the order in which the
     * shader's loop pops nodes. `pending` is the shader's stack_ptr on arrival
     * (sentinel included); an inner node reached with 64 entries would write
     * stack[64], which is undefined behaviour in the reference. */
    struct Visit
    {
        uint32_t src;
        uint32_t pending;
    };
    std::vector<Visit> todo;
    std::vector<uint32_t> inner_of; /* per emitted node: 1 if inner */
    todo.push_back({0u, 1u});
    out.nodes.reserve(n_nodes);
    size_t visited = 0;
    while (!todo.empty())
    {
        const Visit f = todo.back();
        todo.pop_back();
        if (f.src >= n_nodes)
            return fail(ctx, RVPT_B200_EINVAL, "BVH child index %u out of range (%zu nodes)", f.src,
                        n_nodes);
        if (++visited > n_nodes)
            return fail(ctx, RVPT_B200_EINVAL, "BVH is not a tree (cycle or shared child)");
        const rvpt_bvh_node& s = nodes[f.src];
        DevNode d;
        d.bmin_x = s.bounds[0], d.bmax_x = s.bounds[1], d.bmin_y = s.bounds[2];
        d.bmax_y = s.bounds[3], d.bmin_z = s.bounds[4], d.bmax_z = s.bounds[5];
        d.skip = RVPT_NODE_END;
        d.leaf_first = RVPT_NODE_INNER;
        /* the octant copies assume ordered bounds; anything else keeps the min/max walk */
        if (!(d.bmin_x <= d.bmax_x && d.bmin_y <= d.bmax_y && d.bmin_z <= d.bmax_z))
            out.bounds_ordered = false;
        if (s.primitive_count > 0)
        {
            if ((size_t)s.first_child_or_primitive + s.primitive_count > n_tris)
                return fail(ctx, RVPT_B200_EINVAL, "BVH leaf %u: triangles [%u,%u) out of range", f.src,
                            s.first_child_or_primitive,
                            s.first_child_or_primitive + s.primitive_count);
            d.leaf_first = (uint32_t)out.tris.size();
            emit_leaf(s.first_child_or_primitive, s.primitive_count);
            inner_of.push_back(0);
        }
        else
        {
            if (f.pending >= 64)
                return fail(ctx, RVPT_B200_EUNSUPPORTED,
                            "BVH needs more than the reference's 64-entry traversal stack");
            /* second child is popped after the first child's whole subtree */
            todo.push_back({s.first_child_or_primitive + 1, f.pending});
            todo.push_back({s.first_child_or_primitive, f.pending + 1});
            inner_of.push_back(1);
        }
        out.nodes.push_back(d);
    }
    /* subtree sizes in reverse pre-order; the node after a subtree is where the
     * shader's pop lands, i.e. the skip link */
    const uint32_t total = (uint32_t)out.nodes.size();
    std::vector<uint32_t> size(total, 1);
    for (uint32_t k = total; k-- > 0;)
    {
        if (inner_of[k])
        {
            const uint32_t c0 = k + 1;
            const uint32_t c1 = c0 + size[c0];
            size[k] = 1 + size[c0] + size[c1];
            out.nodes[k].leaf_first = RVPT_NODE_INNER | c1;
        }
        const uint32_t next = k + size[k];
        out.nodes[k].skip = next < total ? next : RVPT_NODE_END;
    }
    return 0;
}

/* Two triangles that lie in one plane and overlap with positive area are hit at (nearly) the
 * same t by every ray through the overlap: which of them wins then depends on the order the
 * BVH is walked in (strict `t < closest_t`, intersection.glsl:394: the first one visited), and
 * so does whether the second one's box survives the closest_t clip. Such scenes — an object
 * standing on a floor with its bottom face modelled — keep the reference's child order; scenes
 * without such pairs have an order-independent nearest hit (up to last-bit coincidences that
 * the reference's own result shares), and may be walked front to back. O(n^2) on purpose:
 * only scenes small enough for the shared-memory path ask. */
bool has_coincident_faces(const rvpt_triangle* tris, size_t n)
{
    struct Face
    {
        double v[3][3], nrm[3], d, scale;
        bool degenerate;
    };
    std::vector<Face> f(n);
    for (size_t i = 0; i < n; ++i)
    {
        const float* src[3] = {tris[i].vertex0, tris[i].vertex1, tris[i].vertex2};
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) f[i].v[k][a] = src[k][a];
        double e0[3], e1[3];
        for (int a = 0; a < 3; ++a) e0[a] = f[i].v[1][a] - f[i].v[0][a], e1[a] = f[i].v[2][a] - f[i].v[0][a];
        double c[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
        const double len = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        f[i].degenerate = !(len > 0.0);
        f[i].scale = 0.0;
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) f[i].scale = std::max(f[i].scale, std::fabs(f[i].v[k][a]));
        if (f[i].degenerate) continue;
        for (int a = 0; a < 3; ++a) f[i].nrm[a] = c[a] / len;
        f[i].d = f[i].nrm[0] * f[i].v[0][0] + f[i].nrm[1] * f[i].v[0][1] + f[i].nrm[2] * f[i].v[0][2];
    }
    for (size_t i = 0; i < n; ++i)
    {
        if (f[i].degenerate) continue;
        for (size_t j = i + 1; j < n; ++j)
        {
            if (f[j].degenerate) continue;
            const double eps = 1e-5 * std::max(1.0, std::max(f[i].scale, f[j].scale));
            const double dp = f[i].nrm[0] * f[j].nrm[0] + f[i].nrm[1] * f[j].nrm[1] + f[i].nrm[2] * f[j].nrm[2];
            if (std::fabs(dp) < 1.0 - 1e-9) continue; /* planes not parallel */
            bool same_plane = true;
            for (int k = 0; k < 3 && same_plane; ++k)
            {
                const double dist = f[i].nrm[0] * f[j].v[k][0] + f[i].nrm[1] * f[j].v[k][1] +
                                    f[i].nrm[2] * f[j].v[k][2] - f[i].d;
                same_plane = std::fabs(dist) <= eps;
            }
            if (!same_plane) continue;
            /* 2-D separating-axis test in the plane (dominant axis dropped); touching along an
             * edge or at a vertex (two halves of a quad) is not an overlap */
            int drop = 0;
            for (int a = 1; a < 3; ++a)
                if (std::fabs(f[i].nrm[a]) > std::fabs(f[i].nrm[drop])) drop = a;
            const int ax0 = (drop + 1) % 3, ax1 = (drop + 2) % 3;
            double P[2][3][2];
            for (int k = 0; k < 3; ++k)
            {
                P[0][k][0] = f[i].v[k][ax0], P[0][k][1] = f[i].v[k][ax1];
                P[1][k][0] = f[j].v[k][ax0], P[1][k][1] = f[j].v[k][ax1];
            }
            bool separated = false;
            for (int t = 0; t < 2 && !separated; ++t)
                for (int e = 0; e < 3 && !separated; ++e)
                {
                    const double ex = P[t][(e + 1) % 3][0] - P[t][e][0], ey = P[t][(e + 1) % 3][1] - P[t][e][1];
                    const double nx = -ey, ny = ex;
                    const double nl = std::sqrt(nx * nx + ny * ny);
                    if (!(nl > 0.0)) continue;
                    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
                    for (int q = 0; q < 2; ++q)
                        for (int k = 0; k < 3; ++k)
                        {
                            const double pr = (P[q][k][0] * nx + P[q][k][1] * ny) / nl;
                            lo[q] = std::min(lo[q], pr), hi[q] = std::max(hi[q], pr);
                        }
                    separated = hi[0] <= lo[1] + eps || hi[1] <= lo[0] + eps;
                }
            if (!separated) return true;
        }
    }
    return false;
}

/* The eight front-to-back node arrays (device_scene.h, SceneLayout::off_oct) from the packed
 * reference-order nodes. In pre-order a subtree is a contiguous range whose size does not
 * depend on the order its children are visited in, so every layout reuses the reference
 * layout's subtree sizes for its skip links. Children are ordered along the axis on which
 * their box centres differ most; a tie keeps the reference's order. out = 8*n float4 of the
 * first halves, then 8*n float4 of the second halves. */
void build_octant_layouts(const std::vector<DevNode>& ref, std::vector<float>& out)
{
    const uint32_t n = (uint32_t)ref.size();
    out.assign((size_t)n * 8u * 8u, 0.0f);
    float* A = out.data();
    float* B = out.data() + (size_t)n * 8u * 4u;
    auto size_of = [&](uint32_t k) { return (ref[k].skip == RVPT_NODE_END ? n : ref[k].skip) - k; };
    std::vector<uint32_t> todo;
    for (uint32_t oct = 0; oct < 8; ++oct)
    {
        const float sgn[3] = {(oct & 1u) ? -1.0f : 1.0f, (oct & 2u) ? -1.0f : 1.0f, (oct & 4u) ? -1.0f : 1.0f};
        uint32_t idx = 0;
        todo.clear();
        todo.push_back(0u);
        while (!todo.empty())
        {
            const uint32_t k = todo.back();
            todo.pop_back();
            const DevNode& d = ref[k];
            const uint32_t total = size_of(k);
            const uint32_t skip = idx + total < n ? idx + total : RVPT_NODE_END;
            float* a = A + ((size_t)oct * n + idx) * 4u;
            float* b = B + ((size_t)oct * n + idx) * 4u;
            a[0] = (oct & 1u) ? d.bmax_x : d.bmin_x, a[1] = (oct & 1u) ? d.bmin_x : d.bmax_x;
            a[2] = (oct & 2u) ? d.bmax_y : d.bmin_y, a[3] = (oct & 2u) ? d.bmin_y : d.bmax_y;
            b[0] = (oct & 4u) ? d.bmax_z : d.bmin_z, b[1] = (oct & 4u) ? d.bmin_z : d.bmax_z;
            std::memcpy(&b[2], &skip, 4);
            std::memcpy(&b[3], &d.leaf_first, 4);
            ++idx;
            if (d.leaf_first & RVPT_NODE_INNER)
            {
                const uint32_t c0 = k + 1u, c1 = c0 + size_of(c0);
                const DevNode &p = ref[c0], &q = ref[c1];
                const float diff[3] = {(q.bmin_x + q.bmax_x) - (p.bmin_x + p.bmax_x),
                                       (q.bmin_y + q.bmax_y) - (p.bmin_y + p.bmax_y),
                                       (q.bmin_z + q.bmax_z) - (p.bmin_z + p.bmax_z)};
                int ax = 0;
                for (int t = 1; t < 3; ++t)
                    if (std::fabs(diff[t]) > std::fabs(diff[ax])) ax = t;
                const bool c0_first = diff[ax] * sgn[ax] >= 0.0f; /* c1 lies further along the ray */
                todo.push_back(c0_first ? c1 : c0); /* popped second */
                todo.push_back(c0_first ? c0 : c1);
            }
        }
    }
}

/* A scene too large for shared memory is traversed out of L2 (kSmem = false): pin its blob
 * there with an access-policy window so the path-state streams (hundreds of MB per frame)
 * do not evict it. Best effort: failures only cost performance. */
void apply_scene_l2_policy(rvpt_b200_ctx* ctx)
{
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (!ctx->scene_smem && ctx->l2_persist_max > 0)
    {
        const size_t want = std::min<size_t>(ctx->layout.bytes, (size_t)ctx->l2_persist_max);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        attr.accessPolicyWindow.base_ptr = ctx->d_scene;
        attr.accessPolicyWindow.num_bytes = std::min<size_t>(ctx->layout.bytes, (size_t)ctx->l2_window_max);
        attr.accessPolicyWindow.hitRatio =
            ctx->layout.bytes <= want ? 1.0f : (float)want / (float)ctx->layout.bytes;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    else
    {
        attr.accessPolicyWindow.num_bytes = 0; /* disables the window */
        attr.accessPolicyWindow.hitRatio = 0.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError(); /* best effort */
}

/* Every child box inside its parent's box and every bound finite — what a builder produces, but
 * not something the reference's node format promises. The batched primary wave's leaf lists
 * (kernels.cu, build_leaf_list) skip the inner nodes, which is only the walk's result when a ray
 * that passes a leaf's slab test passes every ancestor's. */
bool boxes_are_nested(const std::vector<DevNode>& nodes)
{
    auto inside = [](const DevNode& c, const DevNode& p) {
        return c.bmin_x >= p.bmin_x && c.bmax_x <= p.bmax_x && c.bmin_y >= p.bmin_y && c.bmax_y <= p.bmax_y &&
               c.bmin_z >= p.bmin_z && c.bmax_z <= p.bmax_z;
    };
    for (size_t i = 0; i < nodes.size(); ++i)
    {
        const DevNode& n = nodes[i];
        const float b[6] = {n.bmin_x, n.bmax_x, n.bmin_y, n.bmax_y, n.bmin_z, n.bmax_z};
        for (float v : b)
            if (!std::isfinite(v)) return false;
        if (!(n.leaf_first & RVPT_NODE_INNER)) continue;
        const size_t c1 = i + 1, c2 = n.leaf_first & ~RVPT_NODE_INNER;
        if (c1 >= nodes.size() || c2 >= nodes.size()) return false;
        if (!inside(nodes[c1], n) || !inside(nodes[c2], n)) return false;
    }
    return true;
}

int upload_packed(rvpt_b200_ctx* ctx, const PackedScene& ps)
{
    SceneLayout L{};
    L.n_nodes = (uint32_t)ps.nodes.size();
    L.n_tris = (uint32_t)ps.tris.size();
    L.n_mats = (uint32_t)ps.mats.size();
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t off = align16(ps.nodes.size() * sizeof(DevNode));
    L.off_tris = (uint32_t)off;
    off = align16(off + ps.tris.size() * sizeof(DevTri));
    L.off_meta = (uint32_t)off;
    off = align16(off + ps.meta.size() * sizeof(DevTriMeta));
    L.off_mats = (uint32_t)off;
    off = align16(off + ps.mats.size() * sizeof(DevMaterial));
    if (off > 0xFFFFFFF0u) return fail(ctx, RVPT_B200_EUNSUPPORTED, "scene larger than 4 GiB");
    L.bytes = (uint32_t)off;
    const bool oct = !(ctx->flags & RVPT_B200_FLAG_NO_OCTANTS) && ps.bounds_ordered &&
                     rvpt::frame_smem_bytes(L.bytes, L.n_nodes, L.n_tris, true) <= RVPT_SMEM_SCENE_LIMIT;
    std::vector<float> oct_block;
    if (oct && !(ctx->flags & RVPT_B200_FLAG_REFERENCE_ORDER) && !ps.coincident_faces)
    {
        build_octant_layouts(ps.nodes, oct_block);
        L.off_oct = L.bytes;
    }
    const size_t blob_bytes = (size_t)L.bytes + oct_block.size() * sizeof(float);

    std::vector<unsigned char> blob(blob_bytes, 0);
    if (!oct_block.empty())
        std::memcpy(blob.data() + L.off_oct, oct_block.data(), oct_block.size() * sizeof(float));
    std::memcpy(blob.data(), ps.nodes.data(), ps.nodes.size() * sizeof(DevNode));
    std::memcpy(blob.data() + L.off_tris, ps.tris.data(), ps.tris.size() * sizeof(DevTri));
    std::memcpy(blob.data() + L.off_meta, ps.meta.data(), ps.meta.size() * sizeof(DevTriMeta));
    std::memcpy(blob.data() + L.off_mats, ps.mats.data(), ps.mats.size() * sizeof(DevMaterial));

    CU(cudaSetDevice(ctx->device));
    /* The reference re-uploads its scene buffers every frame (rvpt.cpp:123-126), so this
     * path is kept cheap: the device blob and the pinned staging copy are reused while they
     * are large enough, the copy is ordered on the ctx stream behind the frames that still
     * read the old blob (no host synchronisation before it), and launch geometry is only
     * re-derived when the blob's shape changes. */
    if (blob_bytes > ctx->scene_capacity)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_scene);
        if (ctx->h_scene_pinned) cudaFreeHost(ctx->h_scene_pinned);
        ctx->d_scene = nullptr;
        ctx->h_scene_pinned = nullptr;
        ctx->scene_capacity = 0;
        const size_t cap = (blob_bytes + 4095) & ~(size_t)4095;
        CU(cudaMalloc(&ctx->d_scene, cap));
        CU(cudaMallocHost(&ctx->h_scene_pinned, cap));
        ctx->scene_capacity = cap;
    }
    else
        CU(cudaEventSynchronize(ctx->scene_copied)); /* the previous upload has left the staging copy */
    std::memcpy(ctx->h_scene_pinned, blob.data(), blob_bytes);
    ctx->scene_blob_bytes = blob_bytes;
    CU(cudaMemcpyAsync(ctx->d_scene, ctx->h_scene_pinned, blob_bytes, cudaMemcpyHostToDevice,
                       ctx->stream));
    CU(cudaEventRecord(ctx->scene_copied, ctx->stream));

    {
        /* scene bounding box (root node) -> origin cells of the binned ray sort */
        const DevNode& r = ps.nodes[0];
        const float lo[3] = {r.bmin_x, r.bmin_y, r.bmin_z}, hi[3] = {r.bmax_x, r.bmax_y, r.bmax_z};
        bool ok = true;
        for (int a = 0; a < 3; ++a) ok = ok && std::isfinite(lo[a]) && std::isfinite(hi[a]) && hi[a] >= lo[a];
        for (int a = 0; a < 3; ++a)
        {
            ctx->sort_lo[a] = ok ? lo[a] : 0.0f;
            const float ext = ok ? std::max(hi[a] - lo[a], 1e-20f) : 1.0f;
            ctx->sort_scale[a] = ok ? (float)(1u << RVPT_SORT_CELL_BITS) / ext : 0.0f;
        }
    }
    const bool same_shape = ctx->have_scene && L.bytes == ctx->layout.bytes &&
                            L.n_nodes == ctx->layout.n_nodes && L.n_tris == ctx->layout.n_tris &&
                            oct == ctx->scene_oct;
    ctx->layout = L;
    ctx->scene_oct = oct;
    ctx->scene_nested = oct && boxes_are_nested(ps.nodes);
    if (!same_shape)
    {
        ctx->scene_smem =
            rvpt::frame_smem_bytes(L.bytes, L.n_nodes, L.n_tris, false) <= RVPT_SMEM_SCENE_LIMIT;
        int occ_f = 0, occ_p = 0, occ_b = 0;
        if (ctx->scene_smem) CU(rvpt::configure_kernels(RVPT_SMEM_SCENE_LIMIT));
        CU(rvpt::occupancy(&occ_f, &occ_p, &occ_b, ctx->scene_smem, ctx->scene_oct, L.bytes, L.n_nodes,
                           L.n_tris));
        if (occ_f < 1 || occ_p < 1 || occ_b < 1)
            return fail(ctx, RVPT_B200_ECUDA, "kernels do not fit on an SM (occupancy %d/%d/%d)",
                        occ_f, occ_p, occ_b);
        /* persistent grids: every CTA is resident (a requirement of the cooperative launch) */
        ctx->grid_frame = ctx->num_sms * occ_f;
        ctx->grid_primary = ctx->num_sms * occ_p;
        ctx->grid_bounce = ctx->num_sms * occ_b;
        /* the per-CTA timeline buffer was sized for the previous grid */
        if (ctx->d_timeline && ctx->grid_frame > ctx->timeline_ctas)
        {
            CU(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_timeline);
            ctx->d_timeline = nullptr;
            ctx->timeline_ctas = 0;
            const size_t bytes = (size_t)ctx->grid_frame * RVPT_TIMELINE_SLOTS * sizeof(unsigned long long);
            CU(cudaMalloc(&ctx->d_timeline, bytes));
            CU(cudaMemset(ctx->d_timeline, 0, bytes));
            ctx->timeline_ctas = ctx->grid_frame;
        }
    }
    ctx->have_scene = true;
    apply_scene_l2_policy(ctx);
    return 0;
}

} /* namespace */

/* ======================================================================== */
/* C ABI                                                                     */
/* ======================================================================== */

extern "C" uint32_t rvpt_b200_abi_version(void) { return RVPT_ABI_VERSION; }

extern "C" const char* rvpt_b200_build_info(void)
{
    return "rvpt_b200 sm_100a wavefront path tracer; arithmetic: unfused fp32 (rvpt_math.h); "
           "built " __DATE__;
}

extern "C" const char* rvpt_b200_last_error(const rvpt_b200_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : "null context";
}

extern "C" int rvpt_b200_create(rvpt_b200_ctx** out, int device, uint32_t width, uint32_t height,
                                uint32_t flags)
{
    if (!out) return RVPT_B200_EINVAL;
    *out = nullptr;
    rvpt_b200_ctx* ctx = new (std::nothrow) rvpt_b200_ctx();
    if (!ctx) return RVPT_B200_ENOMEM;
    *out = ctx; /* returned even on failure so last_error() can be read; destroy it */
    if (width == 0 || height == 0 || width > 65536 || height > 65536)
        return fail(ctx, RVPT_B200_EINVAL, "bad image size %ux%u", width, height);
    if (flags & ~(RVPT_B200_FLAG_ACCUM_RGBA8 | RVPT_B200_FLAG_REFERENCE_DISPATCH |
                  RVPT_B200_FLAG_BRUTE_FORCE | RVPT_B200_FLAG_UNFUSED | RVPT_B200_FLAG_NO_OCTANTS |
                  RVPT_B200_FLAG_NO_BATCH | RVPT_B200_FLAG_NO_FORECAST | RVPT_B200_FLAG_REFERENCE_ORDER |
                  RVPT_B200_FLAG_NO_QUEUE_SORT | RVPT_B200_FLAG_GPU_BVH | RVPT_B200_FLAG_NO_LEAF_LISTS))
        return fail(ctx, RVPT_B200_EINVAL, "unknown flags 0x%x", flags);
    if (const char* e = std::getenv("RVPT_B200_TAIL_RAYS_PER_WARP")) /* developer knob (tuning runs) */
        ctx->tail_rays_per_warp = (uint32_t)std::max(0, std::atoi(e));
    if (const char* e = std::getenv("RVPT_B200_FRAME_GROUP")) /* developer knob (tuning runs); 0 = no leaf lists */
        ctx->frame_group = (uint32_t)std::min(64, std::max(0, std::atoi(e))), ctx->frame_group_fixed = true;
    if (const char* e = std::getenv("RVPT_B200_QUEUE_BUDGET_MIB")) /* path-queue memory budget */
        ctx->queue_budget = (size_t)std::max(1, std::atoi(e)) << 20;
    ctx->device = device;
    ctx->W = width;
    ctx->H = height;
    ctx->flags = flags;
    recompute_partition(ctx);

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ctx, RVPT_B200_ECUDA, "no CUDA device: %s (this engine has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return fail(ctx, RVPT_B200_EINVAL, "device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(ctx, RVPT_B200_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    ctx->num_sms = prop.multiProcessorCount;
    ctx->l2_persist_max = prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = prop.accessPolicyMaxWindowSize;
    CU(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->scene_copied, cudaEventDisableTiming));
    ctx->stream = ctx->own_stream;
    return 0;
}

extern "C" void rvpt_b200_destroy(rvpt_b200_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->own_stream)
    {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->peer_out_raster) cudaIpcCloseMemHandle(ctx->peer_out_raster);
        if (ctx->peer_out_raster2) cudaIpcCloseMemHandle(ctx->peer_out_raster2);
        if (ctx->copy_stream)
        {
            cudaStreamSynchronize(ctx->copy_stream);
            cudaStreamDestroy(ctx->copy_stream);
            cudaEventDestroy(ctx->ev_rendered);
            cudaEventDestroy(ctx->ev_copied);
        }
        free_frame_buffers(ctx);
        cudaFree(ctx->d_timeline);
        cudaFree(ctx->d_scene);
        cudaFree(ctx->d_raw_tris);
        if (ctx->h_scene_pinned) cudaFreeHost(ctx->h_scene_pinned);
        if (ctx->scene_copied) cudaEventDestroy(ctx->scene_copied);
        for (auto& t : ctx->timed) ctx->event_pool.push_back(t);
        for (auto& t : ctx->event_pool)
        {
            cudaEventDestroy(t.a);
            cudaEventDestroy(t.b);
        }
        cudaStreamDestroy(ctx->own_stream);
    }
    delete ctx;
}

extern "C" int rvpt_b200_set_partition(rvpt_b200_ctx* ctx, int rank, int nranks)
{
    if (!ctx) return RVPT_B200_EINVAL;
    if (nranks < 1 || rank < 0 || rank >= nranks)
        return fail(ctx, RVPT_B200_EINVAL, "bad partition rank %d of %d", rank, nranks);
    if (ctx->buffers_ready)
        return fail(ctx, RVPT_B200_EINVAL, "set_partition must precede the first frame");
    ctx->rank = (uint32_t)rank;
    ctx->nranks = (uint32_t)nranks;
    recompute_partition(ctx);
    return 0;
}

extern "C" int rvpt_b200_set_stream(rvpt_b200_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return RVPT_B200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (ctx->have_scene) apply_scene_l2_policy(ctx);
    return 0;
}

extern "C" int rvpt_b200_upload_scene(rvpt_b200_ctx* ctx, const rvpt_bvh_node* nodes,
                                      size_t n_nodes, const rvpt_triangle* triangles,
                                      size_t n_triangles, const rvpt_material* materials,
                                      size_t n_materials)
{
    if (!ctx) return RVPT_B200_EINVAL;
    if (!triangles || n_triangles == 0) return fail(ctx, RVPT_B200_EINVAL, "scene has no triangles");
    if (!materials || n_materials == 0) return fail(ctx, RVPT_B200_EINVAL, "scene has no materials");
    if (n_triangles > 0x7FFFFFFFu) return fail(ctx, RVPT_B200_EUNSUPPORTED, "too many triangles");

    const bool brute = (ctx->flags & RVPT_B200_FLAG_BRUTE_FORCE) != 0;
    /* The reference copies its scene buffers every frame (rvpt.cpp:123-126) although they never
     * change after initialize(): when the caller's arrays equal the previous upload byte for
     * byte, the packed blob still sitting in the pinned staging buffer is copied again (the
     * host -> device transfer stays) and the packing work — BVH re-layout, per-triangle
     * precomputation, coincident-face check, octant layouts — is skipped. */
    {
        const size_t nb = nodes ? n_nodes * sizeof(rvpt_bvh_node) : 0;
        const size_t tb = n_triangles * sizeof(rvpt_triangle), mb = n_materials * sizeof(rvpt_material);
        const size_t hdr = 3 * sizeof(size_t);
        const size_t counts[3] = {nodes ? n_nodes : (size_t)-1, n_triangles, n_materials};
        bool same = ctx->have_scene && ctx->scene_blob_bytes != 0 && ctx->scene_inputs.size() == hdr + nb + tb + mb &&
                    std::memcmp(ctx->scene_inputs.data(), counts, hdr) == 0 &&
                    (nb == 0 || std::memcmp(ctx->scene_inputs.data() + hdr, nodes, nb) == 0) &&
                    std::memcmp(ctx->scene_inputs.data() + hdr + nb, triangles, tb) == 0 &&
                    std::memcmp(ctx->scene_inputs.data() + hdr + nb + tb, materials, mb) == 0;
        if (same)
        {
            CU(cudaSetDevice(ctx->device));
            CU(cudaMemcpyAsync(ctx->d_scene, ctx->h_scene_pinned, ctx->scene_blob_bytes, cudaMemcpyHostToDevice,
                               ctx->stream));
            CU(cudaEventRecord(ctx->scene_copied, ctx->stream));
            return 0;
        }
        ctx->scene_blob_bytes = 0; /* invalid until this upload succeeds */
        ctx->raw_valid = false;
        ctx->raw_source_ok = false;
        ctx->raw_sorted.clear();
        ctx->scene_inputs.resize(hdr + nb + tb + mb);
        std::memcpy(ctx->scene_inputs.data(), counts, hdr);
        if (nb) std::memcpy(ctx->scene_inputs.data() + hdr, nodes, nb);
        std::memcpy(ctx->scene_inputs.data() + hdr + nb, triangles, tb);
        std::memcpy(ctx->scene_inputs.data() + hdr + nb + tb, materials, mb);
    }
    PackedScene ps;
    int rc;
    if (!nodes && !brute)
    {
        /* build internally and permute, like RVPT::initialize() (rvpt.cpp:84-86) */
        std::vector<rvpt_bvh_node> built(2 * n_triangles);
        std::vector<uint32_t> perm(n_triangles);
        size_t n_built = 0;
        if (ctx->flags & RVPT_B200_FLAG_GPU_BVH)
            rc = rvpt_b200_build_bvh_gpu(ctx->device, triangles, n_triangles, built.data(), &n_built, perm.data(), nullptr);
        else
            rc = rvpt_b200_build_bvh(triangles, n_triangles, built.data(), &n_built, perm.data());
        if (rc) return fail(ctx, rc, "internal BVH build failed");
        std::vector<rvpt_triangle> sorted(n_triangles);
        for (size_t i = 0; i < n_triangles; ++i) sorted[i] = triangles[perm[i]];
        rc = pack_scene(ctx, built.data(), n_built, sorted.data(), n_triangles, materials,
                        n_materials, false, ps);
        ctx->raw_sorted.swap(sorted);
    }
    else
        rc = pack_scene(ctx, nodes, n_nodes, triangles, n_triangles, materials, n_materials, brute,
                        ps);
    if (rc) return rc;
    /* front-to-back walking needs an order-independent nearest hit; only small scenes can use it */
    if (!(ctx->flags & (RVPT_B200_FLAG_REFERENCE_ORDER | RVPT_B200_FLAG_NO_OCTANTS | RVPT_B200_FLAG_BRUTE_FORCE)) &&
        n_triangles <= 512) /* more triangles never fit the octant arrays (<= 400 nodes) */
        ps.coincident_faces = has_coincident_faces(triangles, n_triangles);
    else
        ps.coincident_faces = true;
    rc = upload_packed(ctx, ps);
    ctx->raw_source_ok = rc == 0;
    return rc;
}

namespace
{

/* The caller's triangle records on the device, for integrator_Hart. */
int ensure_raw_triangles(rvpt_b200_ctx* ctx)
{
    if (ctx->raw_valid) return 0;
    if (!ctx->raw_source_ok)
        return fail(ctx, RVPT_B200_ENOSCENE, "integrator_Hart needs the triangles of the scene on the device, and the last "
                                             "upload_scene failed: upload the scene again");
    const rvpt_triangle* src = nullptr;
    size_t n = 0;
    if (!ctx->raw_sorted.empty())
        src = ctx->raw_sorted.data(), n = ctx->raw_sorted.size();
    else
    {
        /* scene_inputs = [3 counts][nodes][triangles][materials] of the last upload */
        size_t counts[3];
        std::memcpy(counts, ctx->scene_inputs.data(), sizeof(counts));
        const size_t nb = counts[0] == (size_t)-1 ? 0 : counts[0] * sizeof(rvpt_bvh_node);
        src = reinterpret_cast<const rvpt_triangle*>(ctx->scene_inputs.data() + sizeof(counts) + nb);
        n = counts[1];
    }
    CU(cudaSetDevice(ctx->device));
    if (n > ctx->raw_capacity)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_raw_tris);
        ctx->d_raw_tris = nullptr;
        ctx->raw_capacity = 0;
        CU(cudaMalloc(&ctx->d_raw_tris, n * sizeof(rvpt_triangle)));
        ctx->raw_capacity = n;
    }
    /* pageable source: the copy is staged before the call returns, in stream order on the device */
    CU(cudaMemcpyAsync(ctx->d_raw_tris, src, n * sizeof(rvpt_triangle), cudaMemcpyHostToDevice, ctx->stream));
    ctx->raw_valid = true;
    return 0;
}

/* One frame (n_batch == 0: every aa pass, plus the other integrators' pixels) or one batched
 * launch covering frames current_frame .. current_frame + n_batch - 1 (aa == 1, Kajiya only). */
int render_launches(rvpt_b200_ctx* ctx, const rvpt_render_settings* rs, const float camera[20],
                    uint32_t n_batch)
{
    const int modes[4] = {rs->top_left_render_mode, rs->top_right_render_mode,
                          rs->bottom_left_render_mode, rs->bottom_right_render_mode};
    bool all_kajiya = true;
    for (int m : modes) all_kajiya = all_kajiya && m == 9;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    if ((rc = ensure_batch_capacity(ctx, std::max(n_batch, 1u)))) return rc;
    CU(cudaSetDevice(ctx->device));
    if (rs->aa > 1 && !ctx->d_carry)
        CU(cudaMalloc(&ctx->d_carry,
                      (size_t)ctx->n_local_padded * RVPT_TILE_PIXELS * sizeof(float4)));

    FrameParams p{};
    p.W = ctx->W, p.H = ctx->H;
    p.W_eff = ctx->W, p.H_eff = ctx->H;
    if (ctx->flags & RVPT_B200_FLAG_REFERENCE_DISPATCH)
    {
        p.W_eff = (ctx->W / 16u) * 16u; /* rvpt.cpp:1035-1036 */
        p.H_eff = (ctx->H / 16u) * 16u;
    }
    p.tiles_x = ctx->tiles_x, p.tiles_y = ctx->tiles_y, p.n_tiles = ctx->n_tiles;
    p.rank = ctx->rank, p.nranks = ctx->nranks;
    p.n_local_tiles = ctx->n_local_tiles;
    p.n_chunks = ctx->n_local_tiles * 8u;
    {
        /* multiply-high division (device_scene.h): v * d must stay below 2^40 */
        const unsigned long long lim = 1ull << 40;
        const unsigned long long g_max = (unsigned long long)ctx->n_local_padded * ctx->nranks + ctx->nranks;
        const unsigned long long vc_max = (unsigned long long)p.n_chunks * std::max(n_batch, 1u);
        /* ... and the 64-bit product v * magic must not overflow: quotient below 2^23 */
        p.tiles_x_magic = (ctx->tiles_x && g_max * ctx->tiles_x < lim && g_max / ctx->tiles_x < (1ull << 23))
                              ? lim / ctx->tiles_x + 1ull : 0ull;
        p.n_chunks_magic = (p.n_chunks && vc_max * p.n_chunks < lim && vc_max / p.n_chunks < (1ull << 23))
                               ? lim / p.n_chunks + 1ull : 0ull;
    }
    p.flags = ctx->flags;
    p.inv_dim_x = 1.0f / (float)ctx->W;
    p.inv_dim_y = 1.0f / (float)ctx->H;
    p.frame = rs->current_frame;
    p.frame_f = (float)rs->current_frame;
    p.inv_frame1 = 1.0f / (float)(rs->current_frame + 1u);
    p.keep = (float)(rs->current_frame < 1u ? rs->current_frame : 1u);
    p.max_bounces = rs->max_bounces;
    p.aa = rs->aa;
    p.aa_f = (float)rs->aa;
    p.camera_mode = rs->camera_mode;
    for (int i = 0; i < 4; ++i) p.modes[i] = modes[i];
    p.all_kajiya = all_kajiya ? 1 : 0;
    p.split_x = rs->split_ratio[0], p.split_y = rs->split_ratio[1];
    std::memcpy(p.cam, camera, 16 * sizeof(float));
    p.aspect = camera[16], p.hfov = camera[17], p.scale = camera[18];
    p.inv_tan_half_fov = 1.0f / rv_tan(0.5f * p.hfov); /* camera.glsl:44 */
    p.scene = ctx->d_scene;
    p.layout = ctx->layout;
    p.queue[0] = ctx->queue[0], p.queue[1] = ctx->queue[1];
    p.accum_f32 = (ctx->flags & RVPT_B200_FLAG_ACCUM_RGBA8) ? nullptr : (float4*)ctx->accum;
    p.accum_u8 = (ctx->flags & RVPT_B200_FLAG_ACCUM_RGBA8) ? (uchar4*)ctx->accum : nullptr;
    p.out_tiles = (uchar4*)ctx->out_tiles;
    /* raster target: the display rank's image over NVLink, else our own */
    if (ctx->out_cur == 0)
        p.out_raster = ctx->peer_out_raster ? ctx->peer_out_raster : ctx->d_out_raster;
    else
        p.out_raster = ctx->peer_out_raster2 ? ctx->peer_out_raster2 : ctx->d_out_raster2;
    ctx->out_last = ctx->out_cur;
    p.carry = ctx->d_carry;
    p.ctr = ctx->d_ctr;
    p.timeline = ctx->d_timeline;
    /* binned queues (closed scenes): the origin cells need a finite scene box; without one
     * every ray of an octant lands in that octant's cell 0 */
    p.bin_cap = ctx->bin_cap;
    for (int a = 0; a < 3; ++a) p.sort_lo[a] = ctx->sort_lo[a], p.sort_scale[a] = ctx->sort_scale[a];
    p.n_batch = n_batch;
    /* batched primary wave by (pixel block, frame group) with per-block leaf lists: octant arrays in
     * shared memory, pinhole camera (one origin, directions affine in the pixel), nested boxes */
    p.n_groups = 0;
    p.leaf_lists = nullptr;
    if (n_batch > 0 && ctx->scene_smem && ctx->scene_oct && ctx->scene_nested && rs->camera_mode == 0 &&
        ctx->frame_group && p.n_chunks > 0 && !(ctx->flags & RVPT_B200_FLAG_NO_LEAF_LISTS))
    {
        /* Leaf lists: built once per launch by its first phase, used by (pixel block, frame group)
         * units. Groups are short — the wave ends when its last units end (with 32-frame groups one
         * rank of eight waited 64 us of a 330 us primary wave for them), and all a longer group saves
         * is the per-unit pixel arithmetic. */
        if (!ctx->d_leaf_lists)
            CU(cudaMalloc(&ctx->d_leaf_lists,
                          (size_t)ctx->n_local_padded * 8u * RVPT_LIST_WORDS * sizeof(unsigned short)));
        uint32_t G = ctx->frame_group;
        if (!ctx->frame_group_fixed)
        {
            /* same box, C2 on one GPU: groups of 4 / 8 / 16 / 32 frames 29.2 / 29.9 / 30.3 / 30.4
             * Gsamples/s; one rank of eight (tools/timeline.py --nranks 8): 8-frame groups leave a
             * 28 us tail. At least 12 units per resident warp. */
            const uint64_t want = 12ull * (uint64_t)ctx->grid_frame * (rvpt::threads_per_cta() / 32);
            while (G > 4u && (uint64_t)p.n_chunks * ((n_batch + G - 1u) / G) < want) G >>= 1;
        }
        while ((n_batch + G - 1u) / G > RVPT_MAX_GROUPS) G *= 2u;
        uint32_t n = 0, at = 0;
        for (; at < n_batch; at += G) p.group_start[n++] = (uint8_t)at;
        p.group_start[n] = (uint8_t)n_batch;
        p.n_groups = n;
        p.leaf_lists = std::getenv("RVPT_B200_LIST_INLINE") ? nullptr : ctx->d_leaf_lists; /* developer knob */
    }
    p.samples = ctx->d_samples;
    p.sample_stride = ctx->n_local_padded * RVPT_TILE_PIXELS;

    const bool unfused = (ctx->flags & RVPT_B200_FLAG_UNFUSED) != 0;
    /* a wave with at most this many rays per resident warp runs to completion in its threads */
    p.tail_threshold = unfused ? 0u : (uint32_t)ctx->grid_frame * (rvpt::threads_per_cta() / 32) * ctx->tail_rays_per_warp;
    /* the previous frame's per-bounce counts forecast this frame's small waves (k_frame) */
    p.use_forecast = (!unfused && ctx->frame_seq > 0 && !(ctx->flags & RVPT_B200_FLAG_NO_FORECAST)) ? 1u : 0u;

    uint32_t launches = 0;
    for (int pass = 0; pass < rs->aa && p.n_chunks > 0; ++pass)
    {
        p.pass = pass;
        p.last_of_pass = 1u;
        p.last_of_frame = pass == rs->aa - 1 ? 1u : 0u;
        if (!unfused)
        {
            ScopedTimer tm(ctx, 0);
            CU(rvpt::launch_frame(p, ctx->scene_smem, ctx->scene_oct, ctx->grid_frame, ctx->stream));
            ++launches;
        }
        else
        {
            const uint32_t frame_done = p.last_of_frame;
            {
                ScopedTimer tm(ctx, 0);
                p.last_of_pass = rs->max_bounces < 2 ? 1u : 0u; /* the last wave's launch closes the pass */
                p.last_of_frame = p.last_of_pass ? frame_done : 0u;
                CU(rvpt::launch_primary(p, ctx->scene_smem, ctx->grid_primary, ctx->stream));
            }
            ++launches;
            for (int b = 1; b < rs->max_bounces; ++b)
            {
                ScopedTimer tm(ctx, 1);
                p.last_of_pass = b == rs->max_bounces - 1 ? 1u : 0u;
                p.last_of_frame = p.last_of_pass ? frame_done : 0u;
                CU(rvpt::launch_bounce(p, b, ctx->scene_smem, ctx->grid_bounce, ctx->stream));
                ++launches;
            }
        }
    }
    if (!all_kajiya && p.n_chunks > 0)
    {
        /* pixels of the other integrators (split view / debug views): one in-thread pass */
        bool hart = false;
        for (int m : modes) hart = hart || m < 0 || m > 9;
        if (hart)
        {
            if ((rc = ensure_raw_triangles(ctx))) return rc;
            p.raw_tris = ctx->d_raw_tris;
            p.n_raw_tris = (uint32_t)ctx->layout.n_tris;
        }
        p.pass = 0;
        ScopedTimer tm(ctx, 1);
        CU(rvpt::launch_modes(p, ctx->scene_smem, ctx->grid_primary, ctx->stream));
        ++launches;
    }
    ctx->frame_seq++;
    ctx->frame_rendered = true;
    ctx->last_max_bounces = rs->max_bounces;
    ctx->last_aa = rs->aa;
    ctx->last_launches = ctx->call_launches + launches; /* of the whole render_frame(s) call so far */
    ctx->call_launches = ctx->last_launches;
    ctx->last_frames = std::max(n_batch, 1u);
    return 0;
}

int validate_frame(rvpt_b200_ctx* ctx, const rvpt_render_settings* rs, const float camera[20])
{
    if (!rs || !camera) return fail(ctx, RVPT_B200_EINVAL, "null settings/camera");
    if (!ctx->have_scene) return fail(ctx, RVPT_B200_ENOSCENE, "render_frame before upload_scene");
    if (rs->aa < 1) return fail(ctx, RVPT_B200_EINVAL, "aa = %d (reference divides by it)", rs->aa);
    if (rs->max_bounces < 0 || rs->max_bounces > 64)
        return fail(ctx, RVPT_B200_EINVAL, "max_bounces = %d outside [0,64]", rs->max_bounces);
    const int modes[4] = {rs->top_left_render_mode, rs->top_right_render_mode,
                          rs->bottom_left_render_mode, rs->bottom_right_render_mode};
    /* every index outside 0..9 is eval_integrator's default case, integrator_Hart
     * (compute_pass.comp:96-97): one distance evaluation per triangle and march step */
    (void)modes;
    return 0;
}

} /* namespace */

extern "C" int rvpt_b200_render_frame(rvpt_b200_ctx* ctx, const rvpt_render_settings* rs,
                                      const float camera[20])
{
    if (!ctx) return RVPT_B200_EINVAL;
    const int rc = validate_frame(ctx, rs, camera);
    if (rc) return rc;
    ctx->call_launches = 0;
    return render_launches(ctx, rs, camera, 0);
}

extern "C" int rvpt_b200_render_frames(rvpt_b200_ctx* ctx, const rvpt_render_settings* rs,
                                       const float camera[20], uint32_t n_frames)
{
    if (!ctx) return RVPT_B200_EINVAL;
    int rc = validate_frame(ctx, rs, camera);
    if (rc) return rc;
    rvpt_render_settings s = *rs;
    ctx->call_launches = 0;
    const bool kajiya = s.top_left_render_mode == 9 && s.top_right_render_mode == 9 &&
                        s.bottom_left_render_mode == 9 && s.bottom_right_render_mode == 9;
    /* Batched launches: the frames' waves are merged (kernels.cu, k_frame<.., kBatch>). Needs
     * one sample per pixel and frame (the RNG stream of a pixel runs on from sample to sample
     * inside a frame: aa > 1 is sequential) and the wavefront integrator on every pixel. */
    const uint32_t cap = max_batch(ctx);
    const bool batched = n_frames > 1 && s.aa == 1 && kajiya && cap > 1 &&
                         !(ctx->flags & (RVPT_B200_FLAG_NO_BATCH | RVPT_B200_FLAG_UNFUSED));
    if (!batched)
    {
        for (uint32_t i = 0; i < n_frames; ++i)
        {
            if ((rc = render_launches(ctx, &s, camera, 0))) return rc;
            s.current_frame++;
        }
        return 0;
    }
    /* as few launches as the queue budget allows, of equal size */
    uint32_t n_launches = (n_frames + cap - 1) / cap;
    uint32_t per = (n_frames + n_launches - 1) / n_launches;
    for (uint32_t done = 0; done < n_frames;)
    {
        const uint32_t n = std::min(per, n_frames - done);
        rc = render_launches(ctx, &s, camera, n);
        if (rc == RVPT_B200_ENOMEM && ctx->batch_limit && ctx->batch_limit < per)
        {
            per = ctx->batch_limit; /* the queues did not fit the device: smaller batches */
            continue;
        }
        if (rc) return rc;
        s.current_frame += n;
        done += n;
    }
    return 0;
}

extern "C" int rvpt_b200_sync(rvpt_b200_ctx* ctx)
{
    if (!ctx) return RVPT_B200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int rvpt_b200_read_output_rgba8(rvpt_b200_ctx* ctx, uint8_t* dst)
{
    if (!ctx || !dst) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ctx->W * ctx->H * 4;
    if (ctx->d_out_raster && !ctx->peer_out_raster)
    {
        const uchar4* src = (ctx->out_last == 1 && ctx->d_out_raster2) ? ctx->d_out_raster2 : ctx->d_out_raster;
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    else
    {
        rc = ensure_scratch(ctx);
        if (rc) return rc;
        CU(cudaMemsetAsync(ctx->d_scratch, 0, bytes, ctx->stream));
        CU(rvpt::launch_untile(ctx->out_tiles, ctx->d_scratch, 1, ctx->W, ctx->H, ctx->tiles_x,
                               ctx->n_tiles, ctx->nranks, ctx->rank, 1, ctx->n_local_padded,
                               ctx->stream));
        CU(cudaMemcpyAsync(dst, ctx->d_scratch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int rvpt_b200_read_accum_f32(rvpt_b200_ctx* ctx, float* dst)
{
    if (!ctx || !dst) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    rc = ensure_scratch(ctx);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const size_t raster_bytes = (size_t)ctx->W * ctx->H * sizeof(float4);
    const uint64_t slots = (uint64_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    unsigned char* raster = (unsigned char*)ctx->d_scratch;
    const void* tiles = ctx->accum;
    if (ctx->flags & RVPT_B200_FLAG_ACCUM_RGBA8)
    {
        void* staged = raster + raster_bytes;
        CU(rvpt::launch_u8_to_f32(ctx->accum, staged, slots, ctx->stream));
        tiles = staged;
    }
    CU(cudaMemsetAsync(raster, 0, raster_bytes, ctx->stream));
    CU(rvpt::launch_untile(tiles, raster, 4, ctx->W, ctx->H, ctx->tiles_x, ctx->n_tiles, ctx->nranks,
                           ctx->rank, 1, ctx->n_local_padded, ctx->stream));
    CU(cudaMemcpyAsync(dst, raster, raster_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int rvpt_b200_write_accum_f32(rvpt_b200_ctx* ctx, const float* src)
{
    if (!ctx || !src) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    rc = ensure_scratch(ctx);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const size_t raster_bytes = (size_t)ctx->W * ctx->H * sizeof(float4);
    const uint64_t slots = (uint64_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    unsigned char* raster = (unsigned char*)ctx->d_scratch;
    CU(cudaMemcpyAsync(raster, src, raster_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->flags & RVPT_B200_FLAG_ACCUM_RGBA8)
    {
        void* staged = raster + raster_bytes;
        CU(rvpt::launch_tile(raster, staged, 4, ctx->W, ctx->H, ctx->tiles_x, ctx->n_tiles,
                             ctx->nranks, ctx->rank, ctx->n_local_padded, ctx->stream));
        CU(rvpt::launch_f32_to_u8(staged, ctx->accum, slots, ctx->stream));
    }
    else
        CU(rvpt::launch_tile(raster, ctx->accum, 4, ctx->W, ctx->H, ctx->tiles_x, ctx->n_tiles,
                             ctx->nranks, ctx->rank, ctx->n_local_padded, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); /* src is borrowed for the call only */
    return 0;
}

extern "C" int rvpt_b200_reset_accum(rvpt_b200_ctx* ctx)
{
    if (!ctx) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const size_t slots = (size_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    CU(cudaMemsetAsync(ctx->accum, 0, slots * accum_elem_bytes(ctx), ctx->stream));
    CU(cudaMemsetAsync(ctx->out_tiles, 0, slots * 4, ctx->stream));
    if (ctx->d_out_raster)
        CU(cudaMemsetAsync(ctx->d_out_raster, 0, (size_t)ctx->W * ctx->H * 4, ctx->stream));
    if (ctx->d_out_raster2)
        CU(cudaMemsetAsync(ctx->d_out_raster2, 0, (size_t)ctx->W * ctx->H * 4, ctx->stream));
    return 0;
}

extern "C" int rvpt_b200_get_stats(rvpt_b200_ctx* ctx, rvpt_b200_stats* out)
{
    if (!ctx || !out) return RVPT_B200_EINVAL;
    std::memset(out, 0, sizeof(*out));
    if (!ctx->frame_rendered) return 0;
    CU(cudaSetDevice(ctx->device));
    FrameStats h;
    CU(cudaMemcpyAsync(&h, &ctx->d_ctr->last, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < RVPT_MAX_BOUNCE_STATS; ++b)
    {
        out->active[b] = h.active[b];
        out->segments += h.active[b];
    }
    out->samples = ctx->last_max_bounces > 0 ? h.active[0] : 0;
    out->kernel_launches = ctx->last_launches;
    out->traversal_order = ctx->layout.off_oct != 0u ? 1u : 0u;
    out->frames = ctx->last_frames;
    return 0;
}

extern "C" int rvpt_b200_set_profiling(rvpt_b200_ctx* ctx, int enabled)
{
    if (!ctx) return RVPT_B200_EINVAL;
    ctx->profiling = enabled != 0;
    return 0;
}

extern "C" int rvpt_b200_get_kernel_times(rvpt_b200_ctx* ctx, rvpt_b200_kernel_times* out)
{
    if (!ctx || !out) return RVPT_B200_EINVAL;
    std::memset(out, 0, sizeof(*out));
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    for (auto& t : ctx->timed)
    {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, t.a, t.b));
        if (t.kind == 0)
            out->primary_ms += ms, out->primary_launches++;
        else
            out->bounce_ms += ms, out->bounce_launches++;
        ctx->event_pool.push_back(t);
    }
    ctx->timed.clear();
    return 0;
}

extern "C" int rvpt_b200_set_timeline(rvpt_b200_ctx* ctx, int enabled)
{
    if (!ctx) return RVPT_B200_EINVAL;
    if (!ctx->have_scene) return fail(ctx, RVPT_B200_ENOSCENE, "set_timeline before upload_scene");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_timeline);
    ctx->d_timeline = nullptr;
    ctx->timeline_ctas = 0;
    if (enabled)
    {
        const int ctas = ctx->grid_frame;
        const size_t bytes = (size_t)ctas * RVPT_TIMELINE_SLOTS * sizeof(unsigned long long);
        CU(cudaMalloc(&ctx->d_timeline, bytes));
        CU(cudaMemset(ctx->d_timeline, 0, bytes));
        ctx->timeline_ctas = ctas;
    }
    return 0;
}

extern "C" int rvpt_b200_get_timeline(rvpt_b200_ctx* ctx, uint64_t* out, size_t capacity,
                                      uint32_t* n_ctas, uint32_t* n_slots)
{
    if (!ctx || !n_ctas || !n_slots) return RVPT_B200_EINVAL;
    *n_ctas = (uint32_t)ctx->timeline_ctas;
    *n_slots = RVPT_TIMELINE_SLOTS;
    if (!ctx->d_timeline || !out) return 0;
    const size_t n = (size_t)ctx->timeline_ctas * RVPT_TIMELINE_SLOTS;
    if (capacity < n) return fail(ctx, RVPT_B200_EINVAL, "timeline needs %zu entries", n);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, ctx->d_timeline, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    /* stamps of phases a later frame does not reach must not survive */
    CU(cudaMemset(ctx->d_timeline, 0, n * sizeof(uint64_t)));
    return 0;
}

extern "C" int rvpt_b200_get_tile_info(rvpt_b200_ctx* ctx, rvpt_b200_tile_info* out)
{
    if (!ctx || !out) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    out->width = ctx->W, out->height = ctx->H;
    out->tiles_x = ctx->tiles_x, out->tiles_y = ctx->tiles_y;
    out->rank = ctx->rank, out->nranks = ctx->nranks;
    out->n_local_tiles = ctx->n_local_tiles;
    out->n_local_tiles_padded = ctx->n_local_padded;
    out->d_accum_tiles = ctx->accum;
    out->d_rgba8_tiles = ctx->out_tiles;
    const uint64_t slots = (uint64_t)ctx->n_local_padded * RVPT_TILE_PIXELS;
    out->accum_bytes = slots * accum_elem_bytes(ctx);
    out->rgba8_bytes = slots * 4;
    return 0;
}

extern "C" int rvpt_b200_set_external_tiles(rvpt_b200_ctx* ctx, void* d_accum_tiles,
                                            void* d_rgba8_tiles)
{
    if (!ctx) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    if (d_accum_tiles) ctx->accum = d_accum_tiles;
    if (d_rgba8_tiles) ctx->out_tiles = d_rgba8_tiles;
    return 0;
}

extern "C" int rvpt_b200_untile(rvpt_b200_ctx* ctx, const void* d_gathered, void* d_raster,
                                uint32_t elem_bytes, uint32_t nranks)
{
    if (!ctx) return RVPT_B200_EINVAL;
    return rvpt_b200_untile_on(ctx, d_gathered, d_raster, elem_bytes, nranks, ctx->stream);
}

extern "C" int rvpt_b200_untile_on(rvpt_b200_ctx* ctx, const void* d_gathered, void* d_raster,
                                   uint32_t elem_bytes, uint32_t nranks, void* cuda_stream)
{
    if (!ctx || !d_gathered || !d_raster) return RVPT_B200_EINVAL;
    if (elem_bytes != 4 && elem_bytes != 16)
        return fail(ctx, RVPT_B200_EINVAL, "untile element size %u (want 4 or 16)", elem_bytes);
    if (nranks != ctx->nranks)
        return fail(ctx, RVPT_B200_EINVAL, "untile nranks %u != partition %u", nranks, ctx->nranks);
    CU(cudaSetDevice(ctx->device));
    CU(rvpt::launch_untile(d_gathered, d_raster, elem_bytes / 4, ctx->W, ctx->H, ctx->tiles_x,
                           ctx->n_tiles, ctx->nranks, 0, ctx->nranks, ctx->n_local_padded,
                           (cudaStream_t)cuda_stream));
    return 0;
}

extern "C" int rvpt_b200_export_output(rvpt_b200_ctx* ctx, unsigned char handle[64])
{
    if (!ctx || !handle) return RVPT_B200_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == RVPT_B200_IPC_HANDLE_BYTES, "IPC handle size");
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_out_raster)
    {
        CU(cudaMalloc(&ctx->d_out_raster, (size_t)ctx->W * ctx->H * 4));
        CU(cudaMemsetAsync(ctx->d_out_raster, 0, (size_t)ctx->W * ctx->H * 4, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_out_raster));
    std::memcpy(handle, &h, sizeof(h));
    return 0;
}

extern "C" int rvpt_b200_attach_output(rvpt_b200_ctx* ctx, const unsigned char handle[64])
{
    if (!ctx || !handle) return RVPT_B200_EINVAL;
    if (ctx->peer_out_raster) return fail(ctx, RVPT_B200_EINVAL, "an output image is already attached");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_out_raster = (uchar4*)ptr;
    return 0;
}

namespace
{
int ensure_second_image(rvpt_b200_ctx* ctx)
{
    if (ctx->d_out_raster2) return 0;
    if (!ctx->d_out_raster) return fail(ctx, RVPT_B200_EINVAL, "this ctx owns no raster result image");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(&ctx->d_out_raster2, (size_t)ctx->W * ctx->H * 4));
    CU(cudaMemsetAsync(ctx->d_out_raster2, 0, (size_t)ctx->W * ctx->H * 4, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
} /* namespace */

extern "C" int rvpt_b200_export_output2(rvpt_b200_ctx* ctx, unsigned char handle[64])
{
    if (!ctx || !handle) return RVPT_B200_EINVAL;
    unsigned char first[64];
    int rc = rvpt_b200_export_output(ctx, first); /* makes sure the first image exists */
    if (rc) return rc;
    if ((rc = ensure_second_image(ctx))) return rc;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_out_raster2));
    std::memcpy(handle, &h, sizeof(h));
    return 0;
}

extern "C" int rvpt_b200_attach_output2(rvpt_b200_ctx* ctx, const unsigned char handle[64])
{
    if (!ctx || !handle) return RVPT_B200_EINVAL;
    if (!ctx->peer_out_raster) return fail(ctx, RVPT_B200_EINVAL, "attach_output must precede attach_output2");
    if (ctx->peer_out_raster2) return fail(ctx, RVPT_B200_EINVAL, "a second output image is already attached");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_out_raster2 = (uchar4*)ptr;
    return 0;
}

extern "C" int rvpt_b200_flip_output(rvpt_b200_ctx* ctx)
{
    if (!ctx) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    if (ctx->peer_out_raster)
    {
        if (!ctx->peer_out_raster2) return fail(ctx, RVPT_B200_EINVAL, "flip_output needs attach_output2 on an attached ctx");
    }
    else if ((rc = ensure_second_image(ctx)))
        return rc;
    ctx->out_cur ^= 1;
    return 0;
}

extern "C" int rvpt_b200_read_output_rgba8_async(rvpt_b200_ctx* ctx, uint8_t* dst)
{
    if (!ctx || !dst) return RVPT_B200_EINVAL;
    int rc = ensure_frame_buffers(ctx);
    if (rc) return rc;
    if (!ctx->d_out_raster || ctx->peer_out_raster)
        return fail(ctx, RVPT_B200_EINVAL, "read_output_rgba8_async needs a ctx that owns the raster image "
                                           "(one GPU, or the display rank after export_output)");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream)
    {
        CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_rendered, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
    }
    if ((rc = ensure_second_image(ctx))) return rc;
    /* the frames launched from now on must not start before the previous copy has left the image
     * they are about to overwrite (it is the one that copy read) */
    if (ctx->copy_pending) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
    const uchar4* src = ctx->out_last == 1 ? ctx->d_out_raster2 : ctx->d_out_raster;
    CU(cudaEventRecord(ctx->ev_rendered, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered, 0));
    CU(cudaMemcpyAsync(dst, src, (size_t)ctx->W * ctx->H * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
    ctx->copy_pending = true;
    ctx->out_cur = ctx->out_last ^ 1; /* later frames fill the other image */
    return 0;
}

extern "C" int rvpt_b200_wait_output(rvpt_b200_ctx* ctx)
{
    if (!ctx) return RVPT_B200_EINVAL;
    if (!ctx->copy_pending) return 0;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev_copied));
    return 0;
}

extern "C" int rvpt_b200_has_coincident_faces(const rvpt_triangle* triangles, size_t n_triangles)
{
    if (!triangles) return RVPT_B200_EINVAL;
    return has_coincident_faces(triangles, n_triangles) ? 1 : 0;
}

extern "C" int rvpt_b200_octant_layouts(const rvpt_bvh_node* nodes, size_t n_nodes,
                                        const rvpt_triangle* triangles, size_t n_triangles,
                                        float* out, size_t capacity_floats, size_t* n_packed_nodes)
{
    if (!nodes || !triangles || !n_packed_nodes) return RVPT_B200_EINVAL;
    /* one dummy material: pack_scene only validates indices against the count */
    std::vector<rvpt_triangle> tris(triangles, triangles + n_triangles);
    for (auto& t : tris) t.material_id[0] = 0.0f;
    rvpt_material mat{};
    PackedScene ps;
    const int rc = pack_scene(nullptr, nodes, n_nodes, tris.data(), n_triangles, &mat, 1, false, ps);
    if (rc) return rc;
    *n_packed_nodes = ps.nodes.size();
    if (!out) return 0;
    std::vector<float> block;
    build_octant_layouts(ps.nodes, block);
    if (capacity_floats < block.size()) return RVPT_B200_EINVAL;
    std::memcpy(out, block.data(), block.size() * sizeof(float));
    return 0;
}

extern "C" int rvpt_b200_selftest_math(int device, int op, const float* in, size_t n, float* out)
{
    if (!in || !out || n == 0 || op < 0 || op > 3) return RVPT_B200_EINVAL;
    static const size_t in_w[4] = {1, 1, 3, 4}, out_w[4] = {2, 2, 3, 2};
    if (cudaSetDevice(device) != cudaSuccess) return RVPT_B200_ECUDA;
    float *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc(&d_in, n * in_w[op] * 4) != cudaSuccess) return RVPT_B200_ECUDA;
    if (cudaMalloc(&d_out, n * out_w[op] * 4) != cudaSuccess)
    {
        cudaFree(d_in);
        return RVPT_B200_ECUDA;
    }
    int rc = 0;
    if (cudaMemcpy(d_in, in, n * in_w[op] * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        rvpt::launch_selftest(op, d_in, n, d_out, 0) != cudaSuccess ||
        cudaMemcpy(out, d_out, n * out_w[op] * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = RVPT_B200_ECUDA;
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}
