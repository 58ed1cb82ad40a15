/*
 * kernels.cu — the sm_100a wavefront path-tracing kernels.
 *
 * Replaces the reference's per-pixel megakernel (assets/shaders/compute_pass.comp
 * ::main and everything it includes) with two phases
 *
 *   primary     ray generation + bounce 0, fused (camera.glsl:29-99,
 *               compute_pass.comp:121-167, integrators.glsl:547-677 iteration 0)
 *   bounce b    one wave per later bounce over the compacted path queue
 *               (integrators.glsl:574-671 iteration b)
 *
 * run by  k_frame              all waves of a frame in one persistent cooperative launch
 *                              (grid barriers between waves) — the default;
 *         k_primary / k_bounce one launch per wave (RVPT_B200_FLAG_UNFUSED);
 *                              with n_batch > 0 the launch renders a whole progressive batch
 *                              (rvpt_b200_render_frames): the waves of all its frames are merged
 *                              and a resolve phase folds the parked samples in frame order;
 *         k_modes              the reference's other integrators, one thread per pixel.
 *
 * Paths that terminate do the temporal accumulation in place
 * (compute_pass.comp:146-148,161-166); survivors are compacted with a warp
 * ballot into the next SoA queue. The scene (BVH nodes, precomputed triangle
 * records, materials) is staged per CTA into shared memory with one TMA bulk
 * copy; CTAs are persistent and pull work with an atomic counter.
 *
 * Arithmetic follows include/rvpt_math.h (unfused float32, -fmad=false) so the
 * result is bit-identical to oracle/rvpt_oracle.cpp.
 */
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/rvpt_abi.h"
#include "../../include/rvpt_math.h"
#include "device_scene.h"
#include "kernels.h"

namespace
{

#ifndef RVPT_THREADS
/* One 1024-thread CTA per SM: measured 3-7 % faster than 4 x 256 (the scene is staged once
 * per SM and a grid barrier synchronises 148 CTAs instead of 592). */
#define RVPT_THREADS 1024
#endif
constexpr int kThreads = RVPT_THREADS;
constexpr int kWarpsPerCta = kThreads / 32;
#ifndef RVPT_MIN_CTAS
#define RVPT_MIN_CTAS 1 /* resident CTAs per SM the frame kernel is compiled for (64 registers) */
#endif
#ifndef RVPT_CHUNK_GRAIN
#define RVPT_CHUNK_GRAIN 1
#endif
constexpr uint32_t kChunkGrain = RVPT_CHUNK_GRAIN; /* 32-pixel chunks per claim */
#ifndef RVPT_GLOBAL_CTAS
#define RVPT_GLOBAL_CTAS 1 /* resident CTAs per SM of the global-memory-path instantiations */
#endif

#define RV_INF __int_as_float(0x7f800000)

/* A condition every lane of the warp agrees on, said in a way ptxas can see: a branch on a vote
 * needs no convergence barrier. It matters because the triangle tests of a bounce wave sit ten
 * divergent regions deep (wave loop, claim loop, lanes with a ray, node loop, leaf, plane test,
 * ...) and only the barrier registers B0-B7 survive the slow-path call of an IEEE division
 * untouched: with the claim loop's two exits on plain compares the plane test's own region got B8
 * and ptxas saved and restored it (BMOV, a slow instruction) around EVERY plane test. With the
 * exits on votes the queued bounce waves carry no BMOV: Cornell box +5.3 %, C2 +0.6 %, pinned pose
 * +1.7 % (profiles/r02_experiments.md; the same trick on the primary wave's claim loop or the wave
 * loop of k_frame costs C2 1-2 % — a vote per 32-pixel chunk — and stays out). */
__device__ __forceinline__ bool warp_uniform(bool c) { return __all_sync(0xFFFFFFFFu, c); }

/* ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) ----------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    uint32_t done = 0;
    uint32_t spins = 0;
    while (!done)
    {
        /* a copy that never lands must not hang the GPU: fail the launch instead */
        if (++spins > (1u << 24)) __trap();
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

/* Stage the scene blob into shared memory: one elected thread arms the
 * mbarrier with the byte count and issues the bulk copies (<= 32 KB each);
 * everybody waits on the barrier phase. */
__device__ __forceinline__ void tma_region(void* dst, const unsigned char* src, uint32_t bytes, uint64_t* bar)
{
    for (uint32_t off = 0; off < bytes; off += 32768u)
    {
        uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
        tma_bulk_g2s(static_cast<unsigned char*>(dst) + off, src + off, n, bar);
    }
}

/* oct_a / oct_b: shared-memory destinations of the two halves of the front-to-back octant
 * arrays that follow the blob in global memory (oct_bytes each; 0 = none). */
__device__ __forceinline__ void stage_scene(unsigned char* smem_blob, uint64_t* bar,
                                            const unsigned char* gmem_blob, uint32_t bytes,
                                            void* oct_a = nullptr, void* oct_b = nullptr,
                                            uint32_t oct_bytes = 0)
{
    if (threadIdx.x == 0)
    {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes + 2u * oct_bytes);
        tma_region(smem_blob, gmem_blob, bytes, bar);
        if (oct_bytes)
        {
            tma_region(oct_a, gmem_blob + bytes, oct_bytes, bar);
            tma_region(oct_b, gmem_blob + bytes + oct_bytes, oct_bytes, bar);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);
}

/* ---- scene view ---------------------------------------------------------- */

/* Where the scene lives for this kernel instance. kSmem: 32-bit shared-space
 * addresses, read with explicit ld.shared (the generic-pointer form made ptxas
 * re-derive the shared window base — S2UR/ULEA — inside the node loop);
 * otherwise generic pointers into the L2-resident blob, read through the
 * read-only path. All element sizes are powers of two. */
template <bool kSmem>
struct SceneViewT
{
    typedef typename std::conditional<kSmem, uint32_t, uintptr_t>::type addr_t;
    addr_t nodes; /* 2 float4 per node */
    addr_t tris;  /* 4 float4 per triangle */
    addr_t meta;  /* float4 per triangle: unit normal, material index | RVPT_TRI_LAST */
    addr_t mats;  /* 3 float4 per material */
    /* primary-wave copies relative to the shared camera origin (kRel only):
     * node bounds minus origin, and dot(v0 - origin, n) per triangle — the
     * first operations of intersect_aabb / intersect_triangle_fast, which are
     * identical for every primary ray of a pinhole or spherical camera. */
    addr_t rel_nodes;
    addr_t rel_num;
    /* kOct only: eight copies of the node array, one per ray-direction octant, with
     * each axis' two bounds stored (near, far) for that octant — and eight more relative
     * to the camera origin for the primary wave. Each copy is split into two float4 arrays
     * (16-byte stride: incoherent rays then spread over all 32 banks instead of half of them):
     * (near x, far x, near y, far y) at oct_nodes + copy * oct_stride + node * 16 and
     * (near z, far z, skip, leaf) RVPT_OCT_B_OFFSET bytes further. oct_stride = n_nodes * 16. */
    addr_t oct_nodes;
    addr_t oct_rel_nodes;
    uint32_t oct_stride;
    /* kOct only: 64 bytes per warp in the unused gap between the two halves of the octant arrays
     * (0 when the gap is too small): the leaf list of the pixel block a warp is rendering */
    addr_t beam_scratch;
};

template <bool kSmem, typename A>
__device__ __forceinline__ float4 ld_f4(A base, uint32_t idx)
{
    if constexpr (kSmem)
    {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"((uint32_t)base + idx * 16u));
        return v;
    }
    else
        return __ldg(reinterpret_cast<const float4*>(base) + idx);
}

/* shared-memory float4 at byte address `addr + kOff` (compile-time offset: same address register) */
template <uint32_t kOff>
__device__ __forceinline__ float4 lds_f4_off(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr), "n"(kOff));
    return v;
}

template <bool kSmem, typename A>
__device__ __forceinline__ uint32_t ld_u32(A base, uint32_t idx)
{
    if constexpr (kSmem)
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)base + idx * 4u));
        return v;
    }
    else
        return __ldg(reinterpret_cast<const uint32_t*>(base) + idx);
}

template <bool kSmem>
__device__ __forceinline__ SceneViewT<kSmem> make_view(const unsigned char* base, const SceneLayout& L)
{
    typedef typename SceneViewT<kSmem>::addr_t addr_t;
    addr_t b;
    if constexpr (kSmem)
        b = smem_u32(base);
    else
        b = reinterpret_cast<uintptr_t>(base);
    SceneViewT<kSmem> v;
    v.nodes = b;
    v.tris = b + L.off_tris;
    v.meta = b + L.off_meta;
    v.mats = b + L.off_mats;
    v.rel_nodes = 0;
    v.rel_num = 0;
    v.oct_nodes = 0;
    v.oct_rel_nodes = 0;
    v.oct_stride = 0;
    v.beam_scratch = 0;
    return v;
}

/* ---- nearest hit: stackless walk in the reference's visiting order ------- */

/* Walks that do not visit the BVH in the reference's order (front to back) can only disagree
 * with the reference's result (strict `t < closest_t`: the first triangle visited keeps a tie;
 * a leaf box whose slab entry rounds above closest_t is culled even if its triangle's plane t
 * rounds below it) when a SECOND accepted triangle lies within rounding distance of the nearest
 * one: a shared edge, a corner, crossing surfaces. Such walks therefore run RELAXED, at no extra
 * instruction in the node or plane tests and no extra register: the loop carries the CLIP
 * distance — closest_t moved up by RVPT_TIE_ULPS units in the last place (its bit pattern plus
 * a constant: monotonic, exactly invertible; 3.7e-4 to 7.3e-4 relative) — instead of closest_t.
 * Boxes and plane tests are clipped against it, so every triangle within the band of the nearest
 * one is seen, and an accepted triangle that finds or leaves another one inside the band flags
 * the hit AMBIGUOUS (top bit of the triangle index). trace_nearest() re-traces flagged rays
 * (about one in a million) with the reference's own walk and recovers the exact t of the others
 * by subtracting the constant again. The band is far above the rounding error of t (a few ulp;
 * more at grazing incidence) and far below any geometric separation. Found by the full-size
 * parity runs of round 2: without it, 2 of 1.3e8 samples of the pinned pose and 1 of 3.3e7 of
 * the Cornell box picked another triangle than the reference's walk. */
#define RVPT_TIE_ULPS 6144 /* 3 * 2^11 ulp = 3 * 2^-12 relative at the bottom of a binade */
#define RVPT_TRI_AMBIGUOUS 0x80000000u

__device__ __forceinline__ float clip_of(float t) { return __int_as_float(__float_as_int(t) + RVPT_TIE_ULPS); }
__device__ __forceinline__ float exact_of(float clip) { return __int_as_float(__float_as_int(clip) - RVPT_TIE_ULPS); }

__device__ __forceinline__ bool hit_is_ambiguous(uint32_t best_tri)
{
    return (best_tri ^ RVPT_TRI_AMBIGUOUS) < 0x7FFFFFFFu; /* flag set on a valid index (not 0xFFFFFFFF) */
}

/* Leaf: intersect_triangle_fast (intersection.glsl:267-323) on the precomputed
 * records of one leaf; the early-out after the plane test is value-neutral
 * because the acceptance test is a pure conjunction. kRelaxed: best_t is the clip distance. */
template <bool kSmem, bool kRel, bool kRelaxed>
__device__ __forceinline__ void test_leaf(const SceneViewT<kSmem>& sc, rv_f3 o, rv_f3 d, uint32_t i,
                                          float& best_t, uint32_t& best_tri)
{
    uint32_t m;
    do
    {
        const float4 A = ld_f4<kSmem>(sc.tris, 4 * i + 0);
        const float4 B = ld_f4<kSmem>(sc.tris, 4 * i + 1);
        m = ld_u32<kSmem>(sc.meta, 4u * i + 3u);
        const float num = kRel ? __uint_as_float(ld_u32<kSmem>(sc.rel_num, i))
                               : rv_dot(rv_make(A.x - o.x, A.y - o.y, A.z - o.z),
                                        rv_make(B.x, B.y, B.z));
        const float den = rv_dot(d, rv_make(B.x, B.y, B.z));
        const float t = num / den;
        if (0.0f < t && t < best_t)
        {
            const float4 C = ld_f4<kSmem>(sc.tris, 4 * i + 2);
            const float4 D = ld_f4<kSmem>(sc.tris, 4 * i + 3);
            const float tx = t * d.x, ty = t * d.y, tz = t * d.z;
            const rv_f3 p0 = rv_make((o.x + tx) - A.x, (o.y + ty) - A.y, (o.z + tz) - A.z);
            const float bx = rv_dot(p0, rv_make(C.x, C.y, C.z));
            const float by = rv_dot(p0, rv_make(D.x, D.y, D.z));
            const float m0 = B.w * bx, m1 = C.w * by; /* A00*bx + A10*by */
            const float m2 = C.w * bx, m3 = D.w * by; /* A01*bx + A11*by */
            const float u = A.w * (m0 + m1);
            const float v = A.w * (m2 + m3);
            if (0.0f < u && 0.0f < v && u + v < 1.0f)
            {
                if constexpr (kRelaxed)
                {
                    const float tc = clip_of(t); /* t is finite and positive here */
                    if (tc < best_t)
                    {
                        /* new nearest; the previous one may still lie inside its band */
                        best_tri = (best_t <= clip_of(tc)) ? (i | RVPT_TRI_AMBIGUOUS) : i;
                        best_t = tc;
                    }
                    else
                        best_tri |= RVPT_TRI_AMBIGUOUS; /* inside the band of the nearest (or an exact tie) */
                }
                else
                {
                    best_t = t;
                    best_tri = i;
                }
            }
        }
        ++i;
    } while (!(m & RVPT_TRI_LAST));
}

/* The walk. kSorted: `nodes` is the copy for this ray's direction octant, whose
 * records hold (near, far) per axis, so the slab test needs no per-axis min/max:
 * for a finite non-zero invdir and bmin <= bmax, rounding is monotonic and
 * min((bmin-o)*inv, (bmax-o)*inv) IS the product with the near bound. max/min over
 * the three axes, 0 and closest_t are order-independent for non-NaN values. */
template <bool kSmem, bool kRel, bool kSorted>
__device__ __forceinline__ void walk_nearest(const SceneViewT<kSmem>& sc,
                                             const typename SceneViewT<kSmem>::addr_t nodes, rv_f3 o,
                                             rv_f3 d, float ix, float iy, float iz, float& best_t,
                                             uint32_t& best_tri)
{
    /* the octant arrays may be in front-to-back order: relaxed walk, best_t is the clip
     * distance until trace_nearest() turns it back into the exact t */
#ifdef RVPT_PROBE_NO_RELAXED
    constexpr bool kRelaxed = false; /* timing probe only: front-to-back results are not exact */
#else
    constexpr bool kRelaxed = kSorted;
#endif
    uint32_t node = 0;
    while (node != RVPT_NODE_END)
    {
        float4 n0, n1;
        if constexpr (kSorted)
        {
            const uint32_t a = (uint32_t)nodes + node * 16u;
            n0 = lds_f4_off<0>(a);
            n1 = lds_f4_off<RVPT_OCT_B_OFFSET>(a);
        }
        else
        {
            n0 = ld_f4<kSmem>(nodes, 2 * node);
            n1 = ld_f4<kSmem>(nodes, 2 * node + 1);
        }
        float fx, nx, fy, ny, fz, nz;
        if (kRel)
        {
            fx = n0.y * ix, nx = n0.x * ix;
            fy = n0.w * iy, ny = n0.z * iy;
            fz = n1.y * iz, nz = n1.x * iz;
        }
        else
        {
            fx = (n0.y - o.x) * ix, nx = (n0.x - o.x) * ix;
            fy = (n0.w - o.y) * iy, ny = (n0.z - o.y) * iy;
            fz = (n1.y - o.z) * iz, nz = (n1.x - o.z) * iz;
        }
        float t0, t1;
        if (kSorted)
        {
            t0 = fmaxf(fmaxf(nx, ny), fmaxf(nz, 0.0f));
            t1 = fminf(fminf(fx, fy), fminf(fz, best_t));
        }
        else
        {
            t1 = fminf(fmaxf(fx, nx), fminf(fmaxf(fy, ny), fmaxf(fz, nz)));
            t0 = fmaxf(fminf(fx, nx), fmaxf(fminf(fy, ny), fminf(fz, nz)));
            t0 = fmaxf(t0, 0.0f);
            t1 = fminf(t1, best_t);
        }
        const uint32_t skip = __float_as_uint(n1.z);
        const uint32_t leaf = __float_as_uint(n1.w);
        if (t1 >= t0)
        {
            if (!(leaf & RVPT_NODE_INNER))
            {
                test_leaf<kSmem, kRel, kRelaxed>(sc, o, d, leaf, best_t, best_tri);
                node = skip;
            }
            else
                node = node + 1;
        }
        else
            node = skip;
    }
}

/* The reference's walk for the rare ray whose front-to-back result is ambiguous. Out of line
 * on purpose: one copy per kernel, and its registers are not the hot loops' problem. Shared-
 * memory scenes only (the global path always walks the reference's order). */
__device__ __noinline__ uint2 retrace_reference_order(uint32_t nodes, uint32_t tris, uint32_t meta, rv_f3 o, rv_f3 d)
{
    SceneViewT<true> sc;
    sc.nodes = nodes, sc.tris = tris, sc.meta = meta, sc.mats = 0;
    sc.rel_nodes = sc.rel_num = sc.oct_nodes = sc.oct_rel_nodes = 0;
    sc.oct_stride = 0;
    sc.beam_scratch = 0;
    const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    float t = RV_INF;
    uint32_t tri = 0xFFFFFFFFu;
    walk_nearest<true, false, false>(sc, sc.nodes, o, d, ix, iy, iz, t, tri);
    return make_uint2(__float_as_uint(t), tri);
}

template <bool kSmem, bool kRel, bool kOct>
__device__ __forceinline__ void trace_nearest(const SceneViewT<kSmem>& sc, rv_f3 o, rv_f3 d,
                                              float& best_t, uint32_t& best_tri)
{
    /* intersect_aabb (intersection.glsl:327-357): invdir = 1/direction */
    const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    best_t = RV_INF;
    best_tri = 0xFFFFFFFFu;
    if constexpr (kOct)
    {
        /* The sorted copies are exact only while no product can be NaN on one side
         * of a slab alone, i.e. every invdir component is finite and non-zero; the
         * (practically never taken) other case walks the plain copy with the
         * reference's min/max formulation. */
        const float lo = fminf(fminf(fabsf(ix), fabsf(iy)), fabsf(iz));
        const float hi = fmaxf(fmaxf(fabsf(ix), fabsf(iy)), fabsf(iz));
        if (lo > 0.0f && hi < RV_INF)
        {
            const uint32_t oct = (__float_as_uint(ix) >> 31) | ((__float_as_uint(iy) >> 31) << 1) |
                                 ((__float_as_uint(iz) >> 31) << 2);
            uint32_t base = (uint32_t)(kRel ? sc.oct_rel_nodes : sc.oct_nodes) + oct * sc.oct_stride;
            /* opaque to ptxas, which otherwise re-derives this address (shared window base,
             * constant-bank loads, octant bits) inside the node loop to save a register */
            asm volatile("" : "+r"(base));
#ifdef RVPT_PROBE_NO_RELAXED
            walk_nearest<kSmem, kRel, true>(sc, base, o, d, ix, iy, iz, best_t, best_tri);
#else
            walk_nearest<kSmem, kRel, true>(sc, base, o, d, ix, iy, iz, best_t, best_tri);
            if (best_tri != 0xFFFFFFFFu)
            {
                if (hit_is_ambiguous(best_tri))
                    /* a runner-up within rounding distance of the nearest hit: the reference's own
                     * walk decides (plain node array, reference child order, exact clipping) */
                {
                    const uint2 r = retrace_reference_order(sc.nodes, sc.tris, sc.meta, o, d);
                    best_t = __uint_as_float(r.x), best_tri = r.y;
                }
                else
                    best_t = exact_of(best_t);
            }
#endif
        }
        else
            walk_nearest<kSmem, false, false>(sc, sc.nodes, o, d, ix, iy, iz, best_t, best_tri);
    }
    else
        walk_nearest<kSmem, kRel, false>(sc, kRel ? sc.rel_nodes : sc.nodes, o, d, ix, iy, iz, best_t,
                                         best_tri);
}

/* ---- slot <-> pixel ------------------------------------------------------- */

/* v / d with the precomputed magic = floor(2^40 / d) + 1 (engine.cu guarantees v * d < 2^40) */
__device__ __forceinline__ uint32_t div_magic(uint32_t v, uint32_t d, unsigned long long magic)
{
    return magic ? (uint32_t)(((unsigned long long)v * magic) >> 40) : v / d;
}

__device__ __forceinline__ void slot_to_xy(const FrameParams& p, uint32_t slot, uint32_t& x,
                                           uint32_t& y)
{
    const uint32_t local_tile = slot >> 8;
    const uint32_t g = local_tile * p.nranks + p.rank;
    const uint32_t ty = div_magic(g, p.tiles_x, p.tiles_x_magic);
    const uint32_t tx = g - ty * p.tiles_x;
    const uint32_t w = (slot >> 5) & 7u, lane = slot & 31u;
    x = tx * RVPT_TILE_DIM + ((w & 1u) << 3) + (lane & 7u);
    y = ty * RVPT_TILE_DIM + ((w >> 1) << 2) + (lane >> 3);
}

/* compute_pass.comp:146-148, 162-166: running mean with the previous image and
 * the two rgba8 stores, for the frame's mean sample `sampled`. */
/* rv_unorm8_store on the device: saturate (NaN -> 0) and round-to-nearest-even
 * conversion are single instructions; same codes as the header's formulation. */
__device__ __forceinline__ unsigned char dev_unorm8(float x)
{
    return (unsigned char)__float2uint_rn(__saturatef(x) * 255.0f);
}

/* compute_pass.comp:146-148, 161-163: one step of the running mean */
__device__ __forceinline__ rv_f3 fold_sample(rv_f3 prev, rv_f3 sampled, float frame_f, float inv_frame1,
                                             float keep)
{
    const rv_f3 temporal = rv_make(prev.x * keep, prev.y * keep, prev.z * keep);
    return rv_make((temporal.x * frame_f + sampled.x) * inv_frame1,
                   (temporal.y * frame_f + sampled.y) * inv_frame1,
                   (temporal.z * frame_f + sampled.z) * inv_frame1);
}

__device__ __forceinline__ void store_result(const FrameParams& p, uint32_t slot, uchar4 q, uint32_t raster)
{
    if (p.out_raster)
    {
        if (raster == 0xFFFFFFFFu)
        {
            uint32_t x, y;
            slot_to_xy(p, slot, x, y);
            raster = y * p.W + x;
        }
        p.out_raster[raster] = q;
    }
    else
        p.out_tiles[slot] = q;
}

/* `raster` = y*W + x when the caller already knows the pixel, 0xFFFFFFFF otherwise */
__device__ __forceinline__ void accumulate_pixel(const FrameParams& p, uint32_t slot, rv_f3 sampled,
                                                 uint32_t raster = 0xFFFFFFFFu)
{
    rv_f3 prev;
    const bool u8 = (p.flags & RVPT_B200_FLAG_ACCUM_RGBA8) != 0;
    if (u8)
    {
        const uchar4 k = p.accum_u8[slot];
        prev = rv_make(rv_unorm8_load(k.x), rv_unorm8_load(k.y), rv_unorm8_load(k.z));
    }
    else
    {
        const float4 a = p.accum_f32[slot];
        prev = rv_make(a.x, a.y, a.z);
    }
    const rv_f3 acc = fold_sample(prev, sampled, p.frame_f, p.inv_frame1, p.keep);
    const uchar4 q = make_uchar4(dev_unorm8(acc.x), dev_unorm8(acc.y), dev_unorm8(acc.z), 0);
    if (u8)
        p.accum_u8[slot] = q;
    else
        p.accum_f32[slot] = make_float4(acc.x, acc.y, acc.z, 0.0f);
    store_result(p, slot, q, raster);
}


/* compute_pass.comp:134-144: which integrator a pixel uses (4-way split view) */
__device__ __forceinline__ int integrator_of(const FrameParams& p, uint32_t x, uint32_t y)
{
    int idx = p.modes[0];
    const float sx = (float)x * p.inv_dim_x;
    const float sy = (float)y * p.inv_dim_y;
    if (sy > p.split_y)
        idx = sx < p.split_x ? p.modes[2] : p.modes[3];
    else if (sx > p.split_x)
        idx = p.modes[1];
    return idx;
}

/* ---- sample termination: compute_pass.comp:146-148, 157-166 --------------- */

/* The previous running mean of a pixel is prefetched towards L1 when its path
 * starts, so the dependent load in finish_sample does not pay the full HBM/L2
 * latency at the end of the traversal (costs no registers). */
__device__ __forceinline__ void prefetch_prev(const FrameParams& p, uint32_t slot)
{
    if (p.pass != p.aa - 1) return; /* accum is only read by the last pass of a frame */
    const void* a = (p.flags & RVPT_B200_FLAG_ACCUM_RGBA8)
                        ? static_cast<const void*>(p.accum_u8 + slot)
                        : static_cast<const void*>(p.accum_f32 + slot);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
}

/* `tag` = the accumulation slot; in a batched launch (kBatch) slot | frame_in_batch << 26 */
template <bool kBatch>
__device__ __forceinline__ void finish_sample(const FrameParams& p, uint32_t tag, rv_f3 s,
                                              uint32_t rng, uint32_t raster = 0xFFFFFFFFu)
{
    if constexpr (kBatch)
    {
        /* aa == 1: sampled = (vec3(0) + sample) / 1. Parked until the resolve phase folds the
         * frames of the pixel in order; written once, read once (streaming). */
        const uint32_t slot = tag & RVPT_BATCH_SLOT_MASK, fi = tag >> RVPT_BATCH_SLOT_BITS;
        __stcs(&p.samples[(size_t)fi * p.sample_stride + slot],
               make_float4(0.0f + s.x, 0.0f + s.y, 0.0f + s.z, 0.0f));
        return;
    }
    const uint32_t slot = tag;
    /* sampled = vec3(0); sampled += eval_integrator(...) */
    rv_f3 sum;
    if (p.pass == 0)
        sum = rv_make(0.0f + s.x, 0.0f + s.y, 0.0f + s.z);
    else
    {
        const float4 c = p.carry[slot];
        sum = rv_make(c.x + s.x, c.y + s.y, c.z + s.z);
    }
    if (p.pass != p.aa - 1)
    {
        p.carry[slot] = make_float4(sum.x, sum.y, sum.z, __uint_as_float(rng));
        return;
    }
    /* sampled /= aa ; x / 1.0f == x exactly, so the common aa = 1 case skips three divisions */
    const rv_f3 sampled =
        p.aa == 1 ? sum : rv_make(sum.x / p.aa_f, sum.y / p.aa_f, sum.z / p.aa_f);
    accumulate_pixel(p, slot, sampled, raster);
}

/* ---- one iteration of integrator_Kajiya's loop (integrators.glsl:574-671) -- */

struct PathState
{
    rv_f3 o, d, thr, col;
    uint32_t rng;
};

/* Returns true if the path continues (state updated), false if it ended with
 * `sample`. */
template <bool kSmem>
__device__ __forceinline__ bool kajiya_shade(const SceneViewT<kSmem>& sc, PathState& s, rv_f3& sample, float t,
                                             uint32_t tri);

template <bool kSmem, bool kRel, bool kOct>
__device__ __forceinline__ bool kajiya_step(const SceneViewT<kSmem>& sc, PathState& s, rv_f3& sample)
{
    float t;
    uint32_t tri;
    trace_nearest<kSmem, kRel, kOct>(sc, s.o, s.d, t, tri);
    return kajiya_shade<kSmem>(sc, s, sample, t, tri);
}

/* everything of the iteration behind intersect_scene: (t, tri) = the nearest hit, tri = 0xFFFFFFFF: none */
template <bool kSmem>
__device__ __forceinline__ bool kajiya_shade(const SceneViewT<kSmem>& sc, PathState& s, rv_f3& sample, float t,
                                             uint32_t tri)
{
    if (tri == 0xFFFFFFFFu)
    {
        /* :578-579 background */
        const float k = s.d.y * 0.5f + 0.5f;
        const rv_f3 bg = rv_make(rv_mix(1.0f, 0.2f, k), rv_mix(1.0f, 0.3f, k), rv_mix(1.0f, 0.7f, k));
        sample = rv_add(s.col, rv_mul(s.thr, bg));
        return false;
    }

    const float4 U = ld_f4<kSmem>(sc.meta, tri);
    const uint32_t mi = __float_as_uint(U.w) & ~RVPT_TRI_LAST;
    const float4 M0 = ld_f4<kSmem>(sc.mats, 3 * mi + 0);
    const float4 M1 = ld_f4<kSmem>(sc.mats, 3 * mi + 1);
    const int type = __float_as_int(M1.w);

    /* intersect_scene (intersection.glsl:511-513): normalize(n), evaluated at upload */
    rv_f3 normal = rv_make(U.x, U.y, U.z);
    const rv_f3 pos = rv_add(s.o, rv_scale(t, s.d));

    /* :582 */
    s.col = rv_add(s.col, rv_mul(s.thr, rv_make(M1.x, M1.y, M1.z)));

    const rv_f3 dir_in = rv_normalize(s.d);
    const float cos_view = rv_dot(dir_in, normal);
    float cos_in;
    float eta = M0.w;
    if (cos_view > 0.0f)
    {
        cos_in = cos_view;
        normal = rv_neg(normal);
    }
    else
    {
        cos_in = -cos_view;
        eta = 1.0f / eta;
    }

    if (type == 0)
    {
        /* Lambert :617-623, material.glsl:96-108, samples_mapping.glsl:39-60,112-131 */
        const float4 M2 = ld_f4<kSmem>(sc.mats, 3 * mi + 2);
        const float u = rv_rand(&s.rng);
        const float v = rv_rand(&s.rng);
        const float phi = RV_TWO_PI * u;
        const float cos_theta = (1.0f - v) - v;
        const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        float sn, cs;
        rv_sincos(phi, &sn, &cs);
        s.o = rv_add(pos, rv_scale(RV_EPSILON, normal));
        s.d = rv_add(normal, rv_make(sin_theta * cs, sin_theta * sn, cos_theta));
        s.thr = rv_mul(s.thr, rv_make(M2.x, M2.y, M2.z));
        return true;
    }
    if (type == 1)
    {
        /* mirror :625-631 */
        s.o = rv_add(pos, rv_scale(RV_EPSILON, normal));
        s.d = rv_add(dir_in, rv_scale(cos_in + cos_in, normal));
        s.thr = rv_mul(s.thr, rv_make(M0.x, M0.y, M0.z));
        return true;
    }
    if (type == 2)
    {
        /* dielectric :633-665, material.glsl:207-228 */
        const float cos_out_sqr = 1.0f - (eta * eta) * (1.0f - cos_in * cos_in);
        float cos_out = 0.0f;
        bool refl = (cos_out_sqr <= 0.0f);
        if (!refl)
        {
            cos_out = sqrtf(fmaxf(0.0f, cos_out_sqr));
            const float ec = eta * cos_in;
            const float eo = eta * cos_out;
            const float r_perp = (ec - cos_out) / (ec + cos_out);
            const float r_par = (cos_in - eo) / (cos_in + eo);
            const float f_refl = 0.5f * (r_perp * r_perp + r_par * r_par);
            refl = (rv_rand(&s.rng) < f_refl);
        }
        if (refl)
        {
            s.o = rv_add(pos, rv_scale(RV_EPSILON, normal));
            s.d = rv_add(dir_in, rv_scale(cos_in + cos_in, normal));
        }
        else
        {
            s.o = rv_sub(pos, rv_scale(RV_EPSILON, normal));
            s.d = rv_add(rv_scale(eta, dir_in), rv_scale(eta * cos_in - cos_out, normal));
        }
        s.thr = rv_mul(s.thr, rv_make(M0.x, M0.y, M0.z));
        return true;
    }
    /* default :666-667 */
    sample = rv_make(0.0f, 0.0f, 0.0f);
    return false;
}

/* Warp-aggregated append of the surviving lanes to a queue. Unsorted (open scenes: rays of
 * neighbouring pixels, whatever their direction, walk more alike than same-bin rays from all
 * over the image): ballot + one atomicAdd per warp into the unsorted sub-queue, 4 x 512
 * contiguous bytes per full warp. Sorted (closed scenes): the lanes are grouped by the bin of
 * their new ray — direction octant x origin cell — with match.any, each group appends to that
 * bin's sub-queue with one atomicAdd; what does not fit a bin goes to the overflow sub-queue.
 * Which warp traces a path never changes the path: sorting is scheduling only.
 * qcount = the RVPT_SORT_BINS + 1 counters of this wave. All 32 lanes must call. */
__device__ __forceinline__ void push_survivors(const FrameParams& p, const PathQueue& q,
                                               uint32_t* qcount, bool alive, uint32_t slot,
                                               const PathState& s, bool sort)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, alive);
    if (mask == 0) return;
    const uint32_t ovf_base = RVPT_SORT_BINS * p.bin_cap;
    uint32_t i = 0;
    bool spill = alive; /* lanes that go to the unsorted / overflow sub-queue */
    if (sort)
    {
        if (alive)
        {
            constexpr uint32_t kMax = (1u << RVPT_SORT_CELL_BITS) - 1u;
            /* statistics, not rendering arithmetic: any cell assignment gives the same image */
            const uint32_t cx = min((uint32_t)fmaxf((s.o.x - p.sort_lo[0]) * p.sort_scale[0], 0.0f), kMax);
            const uint32_t cy = min((uint32_t)fmaxf((s.o.y - p.sort_lo[1]) * p.sort_scale[1], 0.0f), kMax);
            const uint32_t cz = min((uint32_t)fmaxf((s.o.z - p.sort_lo[2]) * p.sort_scale[2], 0.0f), kMax);
            const uint32_t oct = (__float_as_uint(s.d.x) >> 31) | ((__float_as_uint(s.d.y) >> 31) << 1) |
                                 ((__float_as_uint(s.d.z) >> 31) << 2);
            const uint32_t bin = (((oct << RVPT_SORT_CELL_BITS | cz) << RVPT_SORT_CELL_BITS | cy) << RVPT_SORT_CELL_BITS) | cx;
            const uint32_t peers = __match_any_sync(mask, bin);
            const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&qcount[bin], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            const uint32_t my = base + __popc(peers & ((1u << lane) - 1u));
            spill = my >= p.bin_cap;
            i = bin * p.bin_cap + my;
        }
    }
    else if (p.bin_cap != 0u)
    {
        /* Unsorted, but not through ONE counter: a 64-frame launch of C2 appends 1.8 M times, and
         * same-address atomics retire at one per ~1.5 SM cycles — 1.4 ms of a 2 ms primary wave, so
         * the wave sat on the edge of that limit: whether it fell over it depended on the build and
         * on where the counter happened to live (C2 rendered at 30.3 or at 25.3 Gsamples/s; ncu put
         * 25 % of all stall samples on the shuffle that waits for the append's base index). The
         * binned sub-queues are there anyway: warp w appends to sub-queue w % 8, through a counter
         * of its own 128-byte line (the sorted path keeps its nine counters in ONE line: the up to
         * eight atomics of a warp's push then travel as one request — spreading them cost the
         * Cornell box 6 %). A wave is appended to either sorted or unsorted, never both. */
        const uint32_t bin = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % RVPT_SORT_BINS;
        const uint32_t first = (uint32_t)(__ffs(mask) - 1);
        uint32_t base = 0;
        if (lane == first) base = atomicAdd(&qcount[(1u + bin) * RVPT_QCOUNT_STRIDE], (uint32_t)__popc(mask));
        base = __shfl_sync(0xFFFFFFFFu, base, first);
        const uint32_t my = base + __popc(mask & ((1u << lane) - 1u));
        spill = alive && my >= p.bin_cap;
        i = bin * p.bin_cap + my;
    }
    const uint32_t smask = __ballot_sync(0xFFFFFFFFu, spill);
    if (smask)
    {
        uint32_t base = 0;
        const uint32_t first = (uint32_t)(__ffs(smask) - 1);
        if (lane == first) base = atomicAdd(&qcount[RVPT_SORT_BINS], (uint32_t)__popc(smask));
        base = __shfl_sync(0xFFFFFFFFu, base, first);
        if (spill) i = ovf_base + base + __popc(smask & ((1u << lane) - 1u));
    }
    if (alive)
    {
        /* path state is written once and read once: streaming stores / loads (evict-first)
         * keep L2 for the accumulation image and, for large scenes, the BVH */
        __stcs(&q.q0[i], make_float4(s.o.x, s.o.y, s.o.z, __uint_as_float(slot)));
        __stcs(&q.q1[i], make_float4(s.d.x, s.d.y, s.d.z, __uint_as_float(s.rng)));
        __stcs(&q.q2[i], make_float4(s.thr.x, s.thr.y, s.thr.z, 0.0f));
        __stcs(&q.q3[i], make_float4(s.col.x, s.col.y, s.col.z, 0.0f));
    }
}

/* Deal a wave: reads its sub-queue counters, L rays per group (32, or fewer when the wave is
 * spread over all warps), groups numbered sub-queue by sub-queue. Block-wide (ends with a
 * barrier); wg lives in shared memory and its previous readers are behind the grid barrier /
 * kernel boundary that precedes every wave. Returns the wave's ray count (block-uniform). */
__device__ __forceinline__ uint32_t prepare_wave(const FrameParams& p, WaveGroups& wg, const uint32_t* qcount,
                                                 uint32_t spread_below)
{
    constexpr uint32_t kQ = RVPT_SORT_BINS + 1u;
    for (uint32_t k = threadIdx.x; k < kQ; k += blockDim.x)
    {
        uint32_t c = *reinterpret_cast<const volatile uint32_t*>(&qcount[k]);
        if (k < RVPT_SORT_BINS) c += *reinterpret_cast<const volatile uint32_t*>(&qcount[(1u + k) * RVPT_QCOUNT_STRIDE]);
        wg.cnt[k] = k < RVPT_SORT_BINS ? min(c, p.bin_cap) : c;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t count = 0;
        for (uint32_t k = 0; k < kQ; ++k) count += wg.cnt[k];
        const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
        /* spread: every warp gets one group even though each sub-queue rounds its last group up */
        const uint32_t share = n_warps > 2u * kQ ? n_warps - kQ : n_warps;
        const uint32_t L = count <= spread_below ? max(1u, min(32u, (count + share - 1u) / share)) : 32u;
        uint32_t g = 0;
        for (uint32_t k = 0; k < kQ; ++k)
        {
            wg.pre[k] = g;
            g += (wg.cnt[k] + L - 1u) / L;
        }
        wg.pre[kQ] = g;
        wg.count = count;
        wg.L = L;
    }
    __syncthreads();
    return wg.count;
}

/* mat4 * vec4(x,y,z,w).xyz with the columns summed left to right. */
__device__ __forceinline__ rv_f3 cam_mul(const float* M, float x, float y, float z, float w)
{
    float r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        const float a = M[0 + i] * x;
        const float b = M[4 + i] * y;
        const float c = M[8 + i] * z;
        const float e = M[12 + i] * w;
        float sacc = a + b;
        sacc = sacc + c;
        sacc = sacc + e;
        r[i] = sacc;
    }
    return rv_make(r[0], r[1], r[2]);
}

/* get_camera_ray (compute_pass.comp:102-118, camera.glsl:29-99) */
__device__ __forceinline__ void camera_ray(const FrameParams& p, float cx, float cy, rv_f3& o,
                                           rv_f3& d)
{
    if (p.camera_mode == 0)
    {
        const float u = p.aspect * ((cx + cx) - 1.0f);
        const float v = (cy + cy) - 1.0f;
        o = rv_make(p.cam[12], p.cam[13], p.cam[14]);
        d = rv_normalize(cam_mul(p.cam, u, v, p.inv_tan_half_fov, 0.0f));
    }
    else if (p.camera_mode == 1)
    {
        const float u = p.aspect * ((cx + cx) - 1.0f);
        const float v = (cy + cy) - 1.0f;
        o = cam_mul(p.cam, p.scale * u, p.scale * v, 0.0f, 1.0f);
        d = rv_make(p.cam[8], p.cam[9], p.cam[10]);
    }
    else
    {
        const float phi = cx * RV_TWO_PI;
        const float theta = cy * RV_PI;
        float sp, cp, st, ct;
        rv_sincos(phi, &sp, &cp);
        rv_sincos(theta, &st, &ct);
        o = rv_make(p.cam[12], p.cam[13], p.cam[14]);
        d = cam_mul(p.cam, st * cp, ct, st * sp, 0.0f);
    }
}

/* ======================================================================== */
/* phases                                                                    */
/* ======================================================================== */

/* The last CTA to leave a launch re-zeroes the wave counters for the next launch and, when the
 * launch completes a frame / batch, publishes its per-bounce ray counts (device_scene.h,
 * FrameCounters). Every thread of every CTA calls this as the last thing it does. */
__device__ __forceinline__ void finish_launch(const FrameParams& p)
{
    __shared__ uint32_t is_last;
    __syncthreads(); /* every warp of the CTA is done with the counters */
    if (threadIdx.x == 0)
    {
        __threadfence();
        is_last = atomicAdd(&p.ctr->done_ctas, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (p.last_of_pass)
    {
        uint32_t* w = reinterpret_cast<uint32_t*>(&p.ctr->wave);
        for (uint32_t i = threadIdx.x; i < sizeof(WaveCounters) / 4; i += blockDim.x) w[i] = 0u;
    }
    if (p.last_of_frame && threadIdx.x < 64u)
    {
        p.ctr->last.active[threadIdx.x] = __ldcg(&p.ctr->stats.active[threadIdx.x]);
        p.ctr->stats.active[threadIdx.x] = 0ull;
    }
    if (threadIdx.x == 0)
    {
        if (p.last_of_frame) p.ctr->last_sets = p.n_batch ? p.n_batch : (uint32_t)p.aa;
        p.ctr->done_ctas = 0u;
    }
}

/* Wave-size forecast from the last completed frame / batch: bit b of the result is set when the
 * rays it traced at bounce b, scaled to the full-image sample sets of THIS launch, fit the
 * in-thread tail (<= tail_threshold). A wave whose successor is forecast
 * that small lets its few survivors run on inside their threads instead of queueing them,
 * which saves the grid barrier and the tail wave behind it. Scheduling only: a wrong
 * forecast (camera or scene just changed) costs lane utilisation for one frame, never a
 * different result. `last` is only written by the last CTA of a launch, so every CTA of this
 * launch reads the same values. */
__device__ __forceinline__ unsigned long long forecast_small_waves(const FrameParams& p, uint32_t* mostly_hits)
{
    *mostly_hits = 0u;
    if (!p.use_forecast) return 0ull;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long* a = p.ctr->last.active;
    const unsigned long long sets_prev = max(1u, __ldcg(&p.ctr->last_sets));
    const unsigned long long sets_now = p.n_batch ? p.n_batch : 1u;
    const unsigned long long lim = (unsigned long long)p.tail_threshold * sets_prev;
    const unsigned long long a_lo = __ldcg(&a[lane]), a_hi = __ldcg(&a[lane + 32u]);
    const uint32_t lo = __ballot_sync(0xFFFFFFFFu, a_lo * sets_now <= lim);
    const uint32_t hi = __ballot_sync(0xFFFFFFFFu, a_hi * sets_now <= lim);
    /* bounce rays (depth >= 1) against those that hit and went on (depth >= 2) */
    unsigned long long rays = (lane >= 1u ? a_lo : 0ull) + a_hi, hits = (lane >= 2u ? a_lo : 0ull) + a_hi;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        rays += __shfl_xor_sync(0xFFFFFFFFu, rays, d);
        hits += __shfl_xor_sync(0xFFFFFFFFu, hits, d);
    }
    *mostly_hits = hits * 2ull > rays ? 1u : 0u;
    return ((unsigned long long)hi << 32) | lo;
}

/* generation + bounce 0: compute_pass.comp:121-158, integrators.glsl:574-671 (i = 0) */
template <bool kSmem, bool kRel, bool kOct, bool kBatch = false>
__device__ __forceinline__ void primary_phase(const FrameParams& p, const SceneViewT<kSmem>& sc, bool sort)
{
    WaveCounters& wc = p.ctr->wave;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long traced = 0;

    /* Work distribution: every warp claims one 32-pixel chunk at a time, so sky
     * and geometry balance at the finest grain. One atomic counter cannot hand
     * out 65 k chunks per frame — the L2 atomic unit retires roughly one
     * same-address atomic per 1.5 SM cycles, as long as the whole frame — so the
     * counter is sharded 16 ways (different cache lines, different L2 slices):
     * shard k hands out claims k, k+16, ... of kChunkGrain chunks; a warp starts on shard (warp % 16)
     * and moves to the next shard when its own runs dry (built-in work
     * stealing). The claim for the next chunk is issued before the current one
     * is traced, so its round trip to L2 is off the critical path. */
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    /* claims are kChunkGrain chunks; a batched launch hands out the chunks of all its frames,
     * frame by frame (virtual chunk = frame_in_batch * n_chunks + chunk) */
    const uint32_t n_vchunks = kBatch ? p.n_chunks * p.n_batch : p.n_chunks;
    const uint32_t n_units = (n_vchunks + kChunkGrain - 1) / kChunkGrain;
    uint32_t shard = gwarp % RVPT_CHUNK_SHARDS;
    uint32_t dry = 0; /* shards found exhausted */
    uint32_t claim = 0;
    if (lane == 0) claim = atomicAdd(&wc.chunk_ctr[shard * 32u], 1u);
    for (;;)
    {
        uint32_t unit = __shfl_sync(0xFFFFFFFFu, claim, 0) * RVPT_CHUNK_SHARDS + shard;
        while (unit >= n_units)
        {
            /* this shard is dry: look at all 16 counters with one parallel load and
             * move to the next shard that still has work (or stop if none has) */
            uint32_t v = 0xFFFFFFFFu;
            if (lane < RVPT_CHUNK_SHARDS)
                v = *reinterpret_cast<volatile uint32_t*>(&wc.chunk_ctr[lane * 32u]);
            const bool has_work = lane < RVPT_CHUNK_SHARDS &&
                                  (uint64_t)v * RVPT_CHUNK_SHARDS + lane < n_units;
            uint32_t live = __ballot_sync(0xFFFFFFFFu, has_work);
            if (live == 0) { dry = RVPT_CHUNK_SHARDS; break; }
            /* first live shard after the current one, cyclically */
            const uint32_t rot = (live >> (shard + 1u)) | (live << (RVPT_CHUNK_SHARDS - 1u - shard));
            shard = (shard + 1u + (uint32_t)(__ffs(rot & 0xFFFFu) - 1)) % RVPT_CHUNK_SHARDS;
            if (lane == 0) claim = atomicAdd(&wc.chunk_ctr[shard * 32u], 1u);
            unit = __shfl_sync(0xFFFFFFFFu, claim, 0) * RVPT_CHUNK_SHARDS + shard;
        }
        if (unit >= n_units) break;
        if (lane == 0) claim = atomicAdd(&wc.chunk_ctr[shard * 32u], 1u);

      for (uint32_t vc = unit * kChunkGrain, vc_end = min(vc + kChunkGrain, n_vchunks); vc < vc_end; ++vc)
      {
        const uint32_t fi = kBatch ? div_magic(vc, p.n_chunks, p.n_chunks_magic) : 0u;
        const uint32_t c = kBatch ? vc - fi * p.n_chunks : vc;
        const uint32_t slot = c * 32u + lane;
        const uint32_t tag = kBatch ? (slot | (fi << RVPT_BATCH_SLOT_BITS)) : slot;
        uint32_t x, y;
        slot_to_xy(p, slot, x, y);
        /* pixels of other integrators (split view) are rendered by k_modes */
        const bool inside = (x < p.W_eff) && (y < p.H_eff) &&
                            ((slot >> 8) * p.nranks + p.rank < p.n_tiles) &&
                            (p.all_kajiya || integrator_of(p, x, y) == 9);
        bool alive = false;
        PathState s;
        if (inside)
        {
            if constexpr (!kBatch) prefetch_prev(p, slot);
            /* util.glsl:35-36; later samples of the frame continue the stream */
            if (kBatch || p.pass == 0)
                s.rng = rv_wang_hash(x + y * p.W) + (p.frame + fi);
            else
                s.rng = __float_as_uint(p.carry[slot].w);

            /* compute_pass.comp:153-154 */
            const float jx = rv_rand(&s.rng);
            const float jy = rv_rand(&s.rng);
            const float cx = ((float)x + jx) * p.inv_dim_x;
            float cy = ((float)y + jy) * p.inv_dim_y;
            cy = 1.0f - cy;
            camera_ray(p, cx, cy, s.o, s.d);
            s.thr = rv_make(1.0f, 1.0f, 1.0f);
            s.col = rv_make(0.0f, 0.0f, 0.0f);

            rv_f3 sample = rv_make(0.0f, 0.0f, 0.0f);
            if (p.max_bounces > 0)
            {
                alive = kajiya_step<kSmem, kRel, kOct>(sc, s, sample);
                if (alive && p.max_bounces == 1)
                {
                    alive = false; /* :674-675 ran out of iterations */
                    sample = rv_make(0.0f, 0.0f, 0.0f);
                }
            }
            if (!alive) finish_sample<kBatch>(p, tag, sample, s.rng, y * p.W + x);
        }
        if (p.max_bounces > 0)
            traced += (unsigned long long)__popc(__ballot_sync(0xFFFFFFFFu, inside));
        push_survivors(p, p.queue[0], wc.qcount[0], alive, tag, s, sort);
      }
    }
    if (lane == 0 && traced) atomicAdd(&p.ctr->stats.active[0], traced);
}

/* Resolve a claim issued earlier on `shard` of the sharded global work counters `ctr` into a unit
 * index; a dry shard is left for the next one that still has work (see primary_phase).
 * 0xFFFFFFFF: no primary chunk is left anywhere. */
__device__ __forceinline__ uint32_t resolve_claim(uint32_t* ctr, uint32_t n_units, uint32_t& shard,
                                                  uint32_t claim)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t unit = __shfl_sync(0xFFFFFFFFu, claim, 0) * RVPT_CHUNK_SHARDS + shard;
    while (unit >= n_units)
    {
        uint32_t v = 0xFFFFFFFFu;
        if (lane < RVPT_CHUNK_SHARDS) v = *reinterpret_cast<volatile uint32_t*>(&ctr[lane * 32u]);
        const bool has_work = lane < RVPT_CHUNK_SHARDS && (uint64_t)v * RVPT_CHUNK_SHARDS + lane < n_units;
        const uint32_t live = __ballot_sync(0xFFFFFFFFu, has_work);
        if (live == 0) return 0xFFFFFFFFu;
        const uint32_t rot = (live >> (shard + 1u)) | (live << (RVPT_CHUNK_SHARDS - 1u - shard));
        shard = (shard + 1u + (uint32_t)(__ffs(rot & 0xFFFFu) - 1)) % RVPT_CHUNK_SHARDS;
        if (lane == 0) claim = atomicAdd(&ctr[shard * 32u], 1u);
        unit = __shfl_sync(0xFFFFFFFFu, claim, 0) * RVPT_CHUNK_SHARDS + shard;
    }
    return unit;
}

/* ---- primary wave of a batched launch, one pixel block for several frames --------------- */

/* The primary rays of one 8x4 pixel block form a thin beam from the camera origin, the same beam
 * in every frame of a batch (only the jitter inside each pixel changes). build_leaf_list()
 * collects the leaves whose boxes ANY ray of that beam can enter, in the order of the beam's
 * octant array, into 32 u16 entries of shared memory per warp — once per pixel block and launch
 * (build_all_lists, the launch's first phase, keeps the lists in global memory); the rays of the
 * block, in every frame of the batch, then test exactly those leaf boxes, with the walk's own
 * slab arithmetic, instead of walking the tree.
 *
 * Why this is the walk's result bit for bit: a leaf box lies inside every ancestor's box (the
 * host checks it, SceneLayout::nested), subtraction and multiplication by one invdir are
 * monotonic under rounding, and max / min are monotonic, so for a given ray and clip distance
 *   t0(ancestor) <= t0(leaf)  and  t1(ancestor) >= t1(leaf):
 * a ray that passes a leaf's slab test passes the test of every ancestor, and a leaf whose
 * ancestor fails fails itself. The walk therefore tests the triangles of exactly the leaves
 * whose OWN box passes, in array order — which is what the list loop does, with the same clip
 * evolution (same leaves, same order), the same relaxed bookkeeping and the same re-trace.
 *
 * Why the list is complete: it is built by interval arithmetic over the beam — directions are
 * affine in the pixel coordinate before normalisation (camera.glsl:41-47) and the slab test's
 * outcome does not depend on the length of the direction, so per axis |d| ranges over the
 * corner values; a box is dropped only if the LARGEST possible exit distance lies below the
 * SMALLEST possible entry distance, after widening every quantity by 1e-5 of its scale (the
 * rays' own rounding errors are below 1e-6 of it). Anything unusual — a direction component
 * that changes sign inside the block, more than 32 leaves, a ray outside its beam's octant —
 * takes the ordinary walk. */
#define RVPT_NO_LIST 0xFFFFFFFFu

__device__ __forceinline__ uint32_t build_leaf_list(const FrameParams& p, const SceneViewT<true>& sc, uint32_t x0,
                                                    uint32_t y0, uint32_t scratch, uint32_t& base_out,
                                                    uint32_t& oct_out)
{
    const uint32_t lane = threadIdx.x & 31u;
    base_out = 0u, oct_out = 0u;
    /* pixel block [x0, x0 + 8] x [y0, y0 + 4] (jitter in [0, 1], rand() may return 1), a 64th of a
     * pixel wider on every side; compute_pass.comp:153-154, camera.glsl:41-43 */
    const float pad = 0.015625f;
    const float cxa = ((float)x0 - pad) * p.inv_dim_x, cxb = ((float)x0 + (8.0f + pad)) * p.inv_dim_x;
    const float cya = 1.0f - ((float)y0 - pad) * p.inv_dim_y, cyb = 1.0f - ((float)y0 + (4.0f + pad)) * p.inv_dim_y;
    const float ua = p.aspect * ((cxa + cxa) - 1.0f), ub = p.aspect * ((cxb + cxb) - 1.0f);
    const float va = (cya + cya) - 1.0f, vb = (cyb + cyb) - 1.0f;
    float dlo[3], dhi[3], L = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        const float a = p.cam[i] * ua, b = p.cam[i] * ub, c = p.cam[4 + i] * va, e = p.cam[4 + i] * vb;
        const float wz = p.cam[8 + i] * p.inv_tan_half_fov;
        dlo[i] = (fminf(a, b) + fminf(c, e)) + wz;
        dhi[i] = (fmaxf(a, b) + fmaxf(c, e)) + wz;
        L += fmaxf(fabsf(dlo[i]), fabsf(dhi[i]));
    }
    if (!(L > 0.0f && L < 1e30f)) return RVPT_NO_LIST;
    const float widen = 1e-5f * L, apart = 1e-4f * L;
    uint32_t oct = 0u;
    float inv_lo[3], inv_hi[3];
    uint32_t flip[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        const float lo = dlo[i] - widen, hi = dhi[i] + widen;
        const bool neg = hi < -apart;
        if (!(neg || lo > apart)) return RVPT_NO_LIST; /* the component changes sign (or nearly) inside the block */
        const float alo = neg ? -hi : lo, ahi = neg ? -lo : hi;
        inv_lo[i] = (1.0f / ahi) * 0.99999f;
        inv_hi[i] = (1.0f / alo) * 1.00001f;
        flip[i] = neg ? 0x80000000u : 0u;
        oct |= neg ? (1u << i) : 0u;
    }
    const uint32_t base = (uint32_t)sc.oct_rel_nodes + oct * sc.oct_stride;
    base_out = base, oct_out = oct;

    /* entry / exit bounds of the beam for one record of that array */
    auto beam_misses = [&](const float4& n0, const float4& n1, float eps) -> bool
    {
        const float nn[3] = {__uint_as_float(__float_as_uint(n0.x) ^ flip[0]) - eps,
                             __uint_as_float(__float_as_uint(n0.z) ^ flip[1]) - eps,
                             __uint_as_float(__float_as_uint(n1.x) ^ flip[2]) - eps};
        const float nf[3] = {__uint_as_float(__float_as_uint(n0.y) ^ flip[0]) + eps,
                             __uint_as_float(__float_as_uint(n0.w) ^ flip[1]) + eps,
                             __uint_as_float(__float_as_uint(n1.y) ^ flip[2]) + eps};
        float t0 = 0.0f, t1 = RV_INF;
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            t0 = fmaxf(t0, fminf(nn[i] * inv_lo[i], nn[i] * inv_hi[i]));
            t1 = fminf(t1, fmaxf(nf[i] * inv_lo[i], nf[i] * inv_hi[i]));
        }
        return t1 < t0;
    };

    /* the root first: most pixel blocks of an open scene see no geometry at all */
    const float4 r0 = lds_f4_off<0>(base), r1 = lds_f4_off<RVPT_OCT_B_OFFSET>(base);
    const float R = fmaxf(fmaxf(fmaxf(fabsf(r0.x), fabsf(r0.y)), fmaxf(fabsf(r0.z), fabsf(r0.w))),
                          fmaxf(fabsf(r1.x), fabsf(r1.y)));
    if (!(R < 1e30f)) return RVPT_NO_LIST;
    const float eps = 1e-5f * R;
    if (beam_misses(r0, r1, eps)) return 0u;

    uint32_t count = 0u;
    const uint32_t n_nodes = p.layout.n_nodes;
    for (uint32_t j0 = 0u; j0 < n_nodes; j0 += 32u)
    {
        const uint32_t j = j0 + lane;
        const uint32_t a = base + min(j, n_nodes - 1u) * 16u;
        const float4 n0 = lds_f4_off<0>(a), n1 = lds_f4_off<RVPT_OCT_B_OFFSET>(a);
        const bool pass = j < n_nodes && !(__float_as_uint(n1.w) & RVPT_NODE_INNER) && !beam_misses(n0, n1, eps);
        const uint32_t hits = __ballot_sync(0xFFFFFFFFu, pass);
        const uint32_t pos = count + (uint32_t)__popc(hits & ((1u << lane) - 1u));
        if (pass && pos < 32u)
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(scratch + 2u * pos), "h"((unsigned short)(j * 16u)) : "memory");
        count += (uint32_t)__popc(hits);
    }
    __syncwarp();
    return count > 32u ? RVPT_NO_LIST : count;
}

/* The rays of a listed pixel block: slab test of every listed leaf box + its triangles, in list
 * (= array) order. Same arithmetic as walk_nearest<true, true, true>; best_t is the clip distance
 * of the relaxed walk. */
__device__ __forceinline__ void trace_listed(const SceneViewT<true>& sc, uint32_t base, uint32_t scratch,
                                             uint32_t n_list, rv_f3 o, rv_f3 d, float ix, float iy, float iz,
                                             float& best_t, uint32_t& best_tri)
{
#ifdef RVPT_PROBE_NO_RELAXED
    constexpr bool kRelaxed = false;
#else
    constexpr bool kRelaxed = true;
#endif
    for (uint32_t k = 0u; k < n_list; ++k)
    {
        unsigned short off;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(off) : "r"(scratch + 2u * k));
        const uint32_t a = base + (uint32_t)off;
        const float4 n0 = lds_f4_off<0>(a), n1 = lds_f4_off<RVPT_OCT_B_OFFSET>(a);
        const float fx = n0.y * ix, nx = n0.x * ix;
        const float fy = n0.w * iy, ny = n0.z * iy;
        const float fz = n1.y * iz, nz = n1.x * iz;
        const float t0 = fmaxf(fmaxf(nx, ny), fmaxf(nz, 0.0f));
        const float t1 = fminf(fminf(fx, fy), fminf(fz, best_t));
        if (t1 >= t0) test_leaf<true, true, kRelaxed>(sc, o, d, __float_as_uint(n1.w), best_t, best_tri);
    }
}

/* First phase of such a launch: every pixel block's leaf list, once, into p.leaf_lists (blocks are
 * dealt statically; a grid barrier follows). */
__device__ __forceinline__ void build_all_lists(const FrameParams& p, const SceneViewT<true>& sc)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t scratch = (uint32_t)sc.beam_scratch + (threadIdx.x >> 5) * 64u;
    for (uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < p.n_chunks; c += n_warps)
    {
        uint32_t x, y;
        slot_to_xy(p, c * 32u + lane, x, y);
        __syncwarp(); /* the previous block's entries have been copied out */
        uint32_t base, oct;
        const uint32_t n_list = build_leaf_list(p, sc, x - (lane & 7u), y - (lane >> 3), scratch, base, oct);
        unsigned short* rec = p.leaf_lists + (size_t)c * RVPT_LIST_WORDS;
        if (n_list != RVPT_NO_LIST && lane < n_list)
        {
            unsigned short off;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(off) : "r"(scratch + 2u * lane));
            rec[4u + lane] = off;
        }
        if (lane == 0)
        {
            rec[0] = n_list == RVPT_NO_LIST ? (unsigned short)RVPT_LIST_NONE : (unsigned short)n_list;
            rec[1] = (unsigned short)oct;
        }
    }
}

/* primary_phase for batched launches of scenes with octant arrays and a pinhole camera
 * (p.n_groups > 0): a claimed unit is (pixel block, group of consecutive frames of the batch).
 * What depends on the pixel block only — pixel coordinates, the RNG seed's hash, the leaf list —
 * is computed once per unit. */
__device__ __forceinline__ void primary_phase_beam(const FrameParams& p, const SceneViewT<true>& sc, bool sort)
{
    WaveCounters& wc = p.ctr->wave;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long traced = 0;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t n_units = p.n_chunks * p.n_groups;
    uint32_t scratch = (uint32_t)sc.beam_scratch + (threadIdx.x >> 5) * 64u;
    /* opaque to ptxas, which otherwise re-derives this address (shared window base, constant-bank
     * loads, scene sizes: 14 uniform-datapath instructions) in every iteration of the list loop */
    asm volatile("" : "+r"(scratch));
    const rv_f3 o = rv_make(p.cam[12], p.cam[13], p.cam[14]);
    uint32_t shard = gwarp % RVPT_CHUNK_SHARDS;
    uint32_t claim = 0;
    if (lane == 0) claim = atomicAdd(&wc.chunk_ctr[shard * 32u], 1u);
    for (;;)
    {
        const uint32_t unit = resolve_claim(wc.chunk_ctr, n_units, shard, claim);
        if (unit == 0xFFFFFFFFu) break;
        if (lane == 0) claim = atomicAdd(&wc.chunk_ctr[shard * 32u], 1u);

        const uint32_t g = div_magic(unit, p.n_chunks, p.n_chunks_magic);
        const uint32_t c = unit - g * p.n_chunks;
        const uint32_t slot = c * 32u + lane;
        uint32_t x, y;
        slot_to_xy(p, slot, x, y);
        const bool inside = (x < p.W_eff) && (y < p.H_eff) && ((slot >> 8) * p.nranks + p.rank < p.n_tiles);
        const uint32_t seed = rv_wang_hash(x + y * p.W) + p.frame; /* util.glsl:35-36, + frame_in_batch below */
        const uint32_t fi_begin = p.group_start[g], fi_end = p.group_start[g + 1u];
        if (p.max_bounces > 0)
            traced += (unsigned long long)__popc(__ballot_sync(0xFFFFFFFFu, inside)) * (fi_end - fi_begin);

        /* the block's leaf list: built by the launch's first phase (copied into this warp's
         * shared-memory slots), or here */
        __syncwarp(); /* nobody still reads the previous unit's list */
        uint32_t n_list, oct, base;
        if (p.leaf_lists)
        {
            const unsigned short* rec = p.leaf_lists + (size_t)c * RVPT_LIST_WORDS;
            /* written by this launch: not the read-only path */
            const uint32_t n_rec = __ldcg(rec);
            oct = __ldcg(rec + 1);
            n_list = n_rec == RVPT_LIST_NONE ? RVPT_NO_LIST : n_rec;
            base = (uint32_t)sc.oct_rel_nodes + oct * sc.oct_stride;
            if (n_list != RVPT_NO_LIST && lane < n_list)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(scratch + 2u * lane), "h"(__ldcg(rec + 4u + lane)) : "memory");
            __syncwarp();
        }
        else
            n_list = build_leaf_list(p, sc, x - (lane & 7u), y - (lane >> 3), scratch, base, oct);

        for (uint32_t fi = fi_begin; fi < fi_end; ++fi)
        {
            const uint32_t tag = slot | (fi << RVPT_BATCH_SLOT_BITS);
            bool alive = false;
            PathState s;
            if (inside)
            {
                s.rng = seed + fi;
                /* compute_pass.comp:153-154 */
                const float jx = rv_rand(&s.rng);
                const float jy = rv_rand(&s.rng);
                const float cx = ((float)x + jx) * p.inv_dim_x;
                float cy = ((float)y + jy) * p.inv_dim_y;
                cy = 1.0f - cy;
                camera_ray(p, cx, cy, s.o, s.d);
                s.thr = rv_make(1.0f, 1.0f, 1.0f);
                s.col = rv_make(0.0f, 0.0f, 0.0f);

                rv_f3 sample = rv_make(0.0f, 0.0f, 0.0f);
                if (p.max_bounces > 0)
                {
                    float t = RV_INF;
                    uint32_t tri = 0xFFFFFFFFu;
                    if (n_list != 0u) /* an empty list: no ray of this block enters the root box */
                    {
                        const float ix = 1.0f / s.d.x, iy = 1.0f / s.d.y, iz = 1.0f / s.d.z;
                        const uint32_t my_oct = (__float_as_uint(ix) >> 31) | ((__float_as_uint(iy) >> 31) << 1) |
                                                ((__float_as_uint(iz) >> 31) << 2);
                        const float lo = fminf(fminf(fabsf(ix), fabsf(iy)), fabsf(iz));
                        const float hi = fmaxf(fmaxf(fabsf(ix), fabsf(iy)), fabsf(iz));
                        if (n_list != RVPT_NO_LIST && my_oct == oct && lo > 0.0f && hi < RV_INF)
                        {
                            trace_listed(sc, base, scratch, n_list, o, s.d, ix, iy, iz, t, tri);
#ifndef RVPT_PROBE_NO_RELAXED
                            if (tri != 0xFFFFFFFFu)
                            {
                                if (hit_is_ambiguous(tri))
                                {
                                    const uint2 r = retrace_reference_order(sc.nodes, sc.tris, sc.meta, o, s.d);
                                    t = __uint_as_float(r.x), tri = r.y;
                                }
                                else
                                    t = exact_of(t);
                            }
#endif
                        }
                        else
                            trace_nearest<true, true, true>(sc, s.o, s.d, t, tri);
                    }
                    alive = kajiya_shade<true>(sc, s, sample, t, tri);
                    if (alive && p.max_bounces == 1)
                    {
                        alive = false; /* :674-675 ran out of iterations */
                        sample = rv_make(0.0f, 0.0f, 0.0f);
                    }
                }
                if (!alive) finish_sample<true>(p, tag, sample, s.rng, y * p.W + x);
            }
            push_survivors(p, p.queue[0], wc.qcount[0], alive, tag, s, sort);
        }
    }
    if (lane == 0 && traced) atomicAdd(&p.ctr->stats.active[0], traced);
}


__device__ __forceinline__ void load_path(const PathQueue& q, uint32_t i, PathState& s,
                                          uint32_t& slot)
{
    const float4 a0 = __ldcs(&q.q0[i]);
    const float4 a1 = __ldcs(&q.q1[i]);
    const float4 a2 = __ldcs(&q.q2[i]);
    const float4 a3 = __ldcs(&q.q3[i]);
    s.o = rv_make(a0.x, a0.y, a0.z);
    slot = __float_as_uint(a0.w);
    s.d = rv_make(a1.x, a1.y, a1.z);
    s.rng = __float_as_uint(a1.w);
    s.thr = rv_make(a2.x, a2.y, a2.z);
    s.col = rv_make(a3.x, a3.y, a3.z);
}

/* Iteration b >= 1 of the bounce loop over queue[(b-1)&1] -> queue[b&1].
 *
 * How the wave's rays are dealt to warps depends on its size (all warp-uniform):
 *   sharded   big waves in k_frame: full 32-ray groups, every one claimed from the
 *             sharded counters (same scheme as the primary wave);
 *   (neither) the one-launch-per-wave kernel k_bounce: 7/8 of each warp's share is
 *             dealt statically, the last eighth claimed from one atomic counter;
 *   spread    waves that cannot fill the machine twice over: every warp takes
 *             the same share, L = ceil(count / n_warps) <= 32 lanes at a time,
 *             so a small incoherent wave costs one short batch per warp instead
 *             of a few warps grinding through 32-wide divergent batches;
 *   kInThread the paths run to their end inside their threads instead of going
 *             back through the queue (the tail wave, and a wave whose successor
 *             is forecast to be a tail); rays of the later bounces are counted
 *             as they are traced.
 */
#define RVPT_WAVE_SPREAD 1u
#define RVPT_WAVE_SHARDED 4u /* big wave, every 32-ray group claimed from the sharded counters */
template <bool kSmem, bool kOct, bool kInThread, bool kBatch = false>
__device__ __forceinline__ void bounce_phase(const FrameParams& p, const SceneViewT<kSmem>& sc, int b,
                                             const WaveGroups& wg, uint32_t mode, bool sort)
{
    WaveCounters& wc = p.ctr->wave;
    const PathQueue qin = p.queue[(b - 1) & 1];
    const PathQueue qout = p.queue[b & 1];
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long* active = p.ctr->stats.active;
    const bool spread = (mode & RVPT_WAVE_SPREAD) != 0;
    constexpr bool in_thread = kInThread; /* compile-time: keeps the queueing call sites lean */
    const bool sharded = (mode & RVPT_WAVE_SHARDED) != 0 && !spread;
    uint32_t* shard_ctr = wc.bounce_ctr[b & 1];
    uint32_t shard = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % RVPT_CHUNK_SHARDS;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    /* spread: L lanes per warp and round; otherwise full 32-ray groups (prepare_wave) */
    const uint32_t L = wg.L;
    const uint32_t groups = wg.pre[RVPT_SORT_BINS + 1u];
    const uint32_t per_warp = groups / n_warps;
    const uint32_t static_rounds = spread ? (groups + n_warps - 1) / n_warps : per_warp - (per_warp >> 3);
    const uint32_t dyn_base = static_rounds * n_warps;

    uint32_t claim = 0;
    if (sharded)
    {
        if (lane == 0) claim = atomicAdd(&shard_ctr[shard * 32u], 1u);
    }
    else if (static_rounds == 0 && lane == 0)
        claim = atomicAdd(&wc.work_ctr[b], 1u);
    for (uint32_t round = 0;; ++round)
    {
        uint32_t g;
        if (sharded)
        {
            /* same scheme as the primary wave: the measured barrier wait of a 7/8-static wave
             * was three times that of the fully dynamic primary wave */
            g = resolve_claim(shard_ctr, groups, shard, claim);
            if (warp_uniform(g == 0xFFFFFFFFu)) break;
            if (lane == 0) claim = atomicAdd(&shard_ctr[shard * 32u], 1u);
        }
        else if (round < static_rounds)
            g = round * n_warps + gwarp;
        else if (spread)
            break;
        else
            g = dyn_base + __shfl_sync(0xFFFFFFFFu, claim, 0);
        if (warp_uniform(g >= groups))
        {
            if (spread) continue; /* other warps of this round still have rays; warp-uniform */
            break;
        }
        if (!sharded && !spread && round + 1 >= static_rounds && lane == 0)
            claim = atomicAdd(&wc.work_ctr[b], 1u);

        /* group g -> (sub-queue, group inside it): the last sub-queue whose first group is <= g
         * (binary search over the prefix table in shared memory) */
        uint32_t k = 0;
#pragma unroll
        for (uint32_t step = RVPT_SORT_BINS; step > 0; step >>= 1)
            if (k + step <= RVPT_SORT_BINS && wg.pre[k + step] <= g) k += step;
        const uint32_t i = (g - wg.pre[k]) * L + lane;
        bool alive = false;
        PathState s;
        uint32_t slot = 0;
        if (lane < L && i < wg.cnt[k])
        {
            load_path(qin, k * p.bin_cap + i, s, slot); /* slot: the path's tag */
            if constexpr (!kBatch) prefetch_prev(p, slot);
            rv_f3 sample;
            for (int k = b;; ++k)
            {
                alive = kajiya_step<kSmem, false, kOct>(sc, s, sample);
                if (alive && k == p.max_bounces - 1)
                {
                    alive = false; /* integrators.glsl:674-675: col is discarded */
                    sample = rv_make(0.0f, 0.0f, 0.0f);
                }
                if (!in_thread || !alive) break;
                if (k + 1 < RVPT_MAX_BOUNCE_STATS) atomicAdd(&active[k + 1], 1ull);
            }
            if (!alive) finish_sample<kBatch>(p, slot, sample, s.rng);
        }
        if (!in_thread) push_survivors(p, qout, wc.qcount[b], alive, slot, s, sort);
    }
}

/* ======================================================================== */
/* scene set-up shared by the frame kernels                                   */
/* ======================================================================== */
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* optional per-CTA phase stamps (rvpt_b200_set_timeline): slot k of this CTA */
__device__ __forceinline__ void stamp(const FrameParams& p, uint32_t k)
{
    if (p.timeline && threadIdx.x == 0 && k < RVPT_TIMELINE_SLOTS)
        p.timeline[blockIdx.x * RVPT_TIMELINE_SLOTS + k] = global_ns();
}

/* Stage the blob with TMA and derive the per-frame copies every ray of the CTA
 * shares:
 *   kRel  origin-relative data for the primary wave (the first subtraction of
 *         intersect_aabb and the numerator of the plane test are the same for all
 *         rays of a pinhole / spherical camera: cam.matrix[3].xyz, camera.glsl:46,94)
 *   kOct  eight direction-octant copies of the nodes with (near, far) bounds, absolute
 *         for bounce rays and origin-relative for primary rays. */
template <bool kSmem, bool kRel, bool kOct>
__device__ __forceinline__ SceneViewT<kSmem> setup_scene(const FrameParams& p, unsigned char* smem,
                                                         uint64_t* bar,
                                                         const uint32_t* ordered_bounce = nullptr)
{
    SceneViewT<kSmem> sc;
    if constexpr (kSmem)
    {
        const uint32_t n_nodes = p.layout.n_nodes;
        /* layout of the dynamic shared memory: blob | rel_num (kRel) | node copies */
        unsigned char* extra = smem + p.layout.bytes;
        float* rel_num = reinterpret_cast<float*>(extra);
        if constexpr (kRel) extra += ((size_t)p.layout.n_tris * 4u + 15u) & ~(size_t)15u;
        float4* oct = reinterpret_cast<float4*>(extra);
        constexpr uint32_t kB = RVPT_OCT_B_OFFSET / 16u; /* second half of every octant record */
        const bool ordered = kOct && p.layout.off_oct != 0u;

        if (ordered)
            stage_scene(smem, bar, p.scene, p.layout.bytes, oct, oct + kB, n_nodes * 8u * 16u);
        else
            stage_scene(smem, bar, p.scene, p.layout.bytes);
        sc = make_view<true>(smem, p.layout);
        const rv_f3 o = rv_make(p.cam[12], p.cam[13], p.cam[14]);
        if constexpr (kRel)
        {
            for (uint32_t i = threadIdx.x; i < p.layout.n_tris; i += blockDim.x)
            {
                const float4 A = ld_f4<true>(sc.tris, 4 * i), B = ld_f4<true>(sc.tris, 4 * i + 1);
                rel_num[i] = rv_dot(rv_make(A.x - o.x, A.y - o.y, A.z - o.z), rv_make(B.x, B.y, B.z));
            }
            sc.rel_num = smem_u32(rel_num);
        }
        if constexpr (kOct)
        {
            static_assert(kWarpsPerCta % 8 == 0, "one warp group per octant");
            float4* oct_rel = oct + 8 * (size_t)n_nodes;
            const uint32_t w = threadIdx.x >> 5, k = w & 7u;
            float4* dst = oct + (size_t)n_nodes * k;
            float4* dst_rel = oct_rel + (size_t)n_nodes * k;
            /* bounce rays walk front to back only when most of them hit something (closed scenes):
             * rays that escape gain nothing from the order and a warp's lanes, spread over the eight
             * arrays, would then walk eight different node sequences for nothing */
            const bool keep_ordered_abs = ordered && ordered_bounce && *ordered_bounce != 0u;
            for (uint32_t i = (w >> 3) * 32u + (threadIdx.x & 31u); i < n_nodes; i += (kWarpsPerCta / 8) * 32u)
            {
                if (ordered)
                {
                    /* primary rays: the staged front-to-back arrays minus the camera origin */
                    if constexpr (kRel)
                    {
                        const float4 a = dst[i], b = dst[i + kB];
                        dst_rel[i] = make_float4(a.x - o.x, a.y - o.x, a.z - o.y, a.w - o.y);
                        dst_rel[i + kB] = make_float4(b.x - o.z, b.y - o.z, b.z, b.w);
                    }
                    if (keep_ordered_abs) continue;
                }
                /* reference-order copy for octant k: (near, far) per axis from the plain node */
                const float4 n0 = ld_f4<true>(sc.nodes, 2 * i), n1 = ld_f4<true>(sc.nodes, 2 * i + 1);
                const float ax = (k & 1u) ? n0.y : n0.x, bx = (k & 1u) ? n0.x : n0.y;
                const float ay = (k & 2u) ? n0.w : n0.z, by = (k & 2u) ? n0.z : n0.w;
                const float az = (k & 4u) ? n1.y : n1.x, bz = (k & 4u) ? n1.x : n1.y;
                dst[i] = make_float4(ax, bx, ay, by);
                dst[i + kB] = make_float4(az, bz, n1.z, n1.w);
                if constexpr (kRel)
                {
                    if (!ordered)
                    {
                        dst_rel[i] = make_float4(ax - o.x, bx - o.x, ay - o.y, by - o.y);
                        dst_rel[i + kB] = make_float4(az - o.z, bz - o.z, n1.z, n1.w);
                    }
                }
            }
            sc.oct_nodes = smem_u32(oct);
            sc.oct_rel_nodes = smem_u32(oct_rel);
            sc.oct_stride = n_nodes * 16u;
            /* the first halves of the 16 copies end at n_nodes * 256 bytes, the second halves start
             * RVPT_OCT_B_OFFSET bytes in: what lies between is allocated and unused */
            sc.beam_scratch = n_nodes * 256u + kWarpsPerCta * 64u <= RVPT_OCT_B_OFFSET ? smem_u32(oct) + n_nodes * 256u : 0u;
        }
        else if constexpr (kRel)
        {
            float4* rel_nodes = reinterpret_cast<float4*>(extra);
            for (uint32_t i = threadIdx.x; i < n_nodes; i += blockDim.x)
            {
                const float4 n0 = ld_f4<true>(sc.nodes, 2 * i), n1 = ld_f4<true>(sc.nodes, 2 * i + 1);
                rel_nodes[2 * i] = make_float4(n0.x - o.x, n0.y - o.x, n0.z - o.y, n0.w - o.y);
                rel_nodes[2 * i + 1] = make_float4(n1.x - o.z, n1.y - o.z, n1.z, n1.w);
            }
            sc.rel_nodes = smem_u32(rel_nodes);
        }
        if constexpr (kRel || kOct) __syncthreads();
    }
    else
        sc = make_view<false>(p.scene, p.layout);
    return sc;
}

/* ======================================================================== */
/* resolve phase of a batched launch                                          */
/* ======================================================================== */
/* Folds the parked samples of frames frame .. frame + n_batch - 1 into the running mean, in
 * frame order, with exactly the per-frame arithmetic of compute_pass.comp:146-148,161-166
 * (in ACCUM_RGBA8 mode the temporal image is re-quantised after every frame, as the reference's
 * UNORM8 image is), and writes the result image once. fc = per-frame (float(frame),
 * 1/float(frame+1), float(min(frame,1))) in shared memory. Pure streaming: chunks are dealt
 * statically. */
__device__ __forceinline__ void resolve_phase(const FrameParams& p, const float* fc)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const bool u8 = (p.flags & RVPT_B200_FLAG_ACCUM_RGBA8) != 0;
    for (uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < p.n_chunks; c += n_warps)
    {
        const uint32_t slot = c * 32u + lane;
        uint32_t x, y;
        slot_to_xy(p, slot, x, y);
        if (!((x < p.W_eff) && (y < p.H_eff) && ((slot >> 8) * p.nranks + p.rank < p.n_tiles))) continue;
        rv_f3 acc;
        if (u8)
        {
            const uchar4 k = p.accum_u8[slot];
            acc = rv_make(rv_unorm8_load(k.x), rv_unorm8_load(k.y), rv_unorm8_load(k.z));
        }
        else
        {
            const float4 a = p.accum_f32[slot];
            acc = rv_make(a.x, a.y, a.z);
        }
        uchar4 q = make_uchar4(0, 0, 0, 0);
        const float4* sp = p.samples + slot;
        uint32_t fi = 0;
        /* four loads in flight per thread */
        for (; fi + 4u <= p.n_batch; fi += 4u)
        {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __ldcs(sp + (size_t)(fi + k) * p.sample_stride);
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const float* f = fc + 3u * (fi + k);
                acc = fold_sample(acc, rv_make(v[k].x, v[k].y, v[k].z), f[0], f[1], f[2]);
                if (u8)
                {
                    q = make_uchar4(dev_unorm8(acc.x), dev_unorm8(acc.y), dev_unorm8(acc.z), 0);
                    acc = rv_make(rv_unorm8_load(q.x), rv_unorm8_load(q.y), rv_unorm8_load(q.z));
                }
            }
        }
        for (; fi < p.n_batch; ++fi)
        {
            const float4 v = __ldcs(sp + (size_t)fi * p.sample_stride);
            const float* f = fc + 3u * fi;
            acc = fold_sample(acc, rv_make(v.x, v.y, v.z), f[0], f[1], f[2]);
            if (u8)
            {
                q = make_uchar4(dev_unorm8(acc.x), dev_unorm8(acc.y), dev_unorm8(acc.z), 0);
                acc = rv_make(rv_unorm8_load(q.x), rv_unorm8_load(q.y), rv_unorm8_load(q.z));
            }
        }
        if (u8)
            p.accum_u8[slot] = q;
        else
        {
            q = make_uchar4(dev_unorm8(acc.x), dev_unorm8(acc.y), dev_unorm8(acc.z), 0);
            p.accum_f32[slot] = make_float4(acc.x, acc.y, acc.z, 0.0f);
        }
        store_result(p, slot, q, y * p.W + x);
    }
}

/* ======================================================================== */
/* k_frame: the whole frame (one aa pass) — or, with kBatch, a whole progressive batch of     */
/* frames — in ONE persistent cooperative launch                               */
/* ======================================================================== */
/* kBatch (rvpt_b200_render_frames, aa == 1, Kajiya everywhere): the frames of a batch depend on
 * each other only through each pixel's running mean, so their waves are merged — the primary
 * wave generates the primary rays of ALL frames back to back, bounce wave b traces the depth-b
 * rays of all frames (a queued path carries its frame in the slot tag), finished samples are
 * parked, and one resolve phase folds them per pixel in frame order. The scene is staged once
 * and the launch pays max_bounces grid barriers + one for the resolve, not that many per frame. */
template <bool kSmem, bool kRel, bool kOct, bool kBatch>
__global__ void __launch_bounds__(kThreads, (kSmem ? RVPT_MIN_CTAS : RVPT_GLOBAL_CTAS)) k_frame(const FrameParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();

    __shared__ unsigned long long small_waves;
    __shared__ uint32_t ordered_bounce; /* bounce rays walk the front-to-back arrays this frame */
    __shared__ WaveGroups wg;
    __shared__ float frame_consts[kBatch ? 3 * RVPT_MAX_BATCH : 1];

    stamp(p, 0);
    if (threadIdx.x < 32)
    {
        uint32_t mostly_hits;
        const unsigned long long m = forecast_small_waves(p, &mostly_hits);
        if (threadIdx.x == 0)
        {
            small_waves = m;
            ordered_bounce = mostly_hits;
        }
    }
    if constexpr (kBatch)
    {
        if (threadIdx.x < p.n_batch)
        {
            /* compute_pass.comp:146-148,161-163; rvpt.cpp:102-111 (current_frame++) */
            const uint32_t f = p.frame + threadIdx.x;
            frame_consts[3 * threadIdx.x + 0] = (float)f;
            frame_consts[3 * threadIdx.x + 1] = 1.0f / (float)(f + 1u);
            frame_consts[3 * threadIdx.x + 2] = (float)(f < 1u ? f : 1u);
        }
    }
    const SceneViewT<kSmem> sc = setup_scene<kSmem, kRel, kOct>(p, smem, &bar, &ordered_bounce);
    if constexpr (!kSmem) __syncthreads(); /* the global path stages nothing: publish the forecast */
    stamp(p, 1);

    /* closed scenes (most bounce rays hit something) queue their survivors by bin — direction
     * octant x origin cell; ordered_bounce is visible to everybody since the barriers of
     * setup_scene */
    const bool sort = p.bin_cap != 0u && ordered_bounce != 0u;
    bool listed = false;
    if constexpr (kSmem && kRel && kOct && kBatch) listed = p.n_groups != 0u && sc.beam_scratch != 0u;
    if constexpr (kSmem && kRel && kOct && kBatch)
    {
        if (listed)
        {
            if (p.leaf_lists)
            {
                build_all_lists(p, sc);
                grid.sync();
            }
            primary_phase_beam(p, sc, sort);
        }
        else
            primary_phase<kSmem, kRel, kOct, kBatch>(p, sc, sort);
    }
    else
        primary_phase<kSmem, kRel, kOct, kBatch>(p, sc, sort);
    stamp(p, 2);

    WaveCounters& wc = p.ctr->wave;
    bool synced = false; /* a grid barrier has been passed since the last sample was parked */
    for (int b = 1; b < p.max_bounces; ++b)
    {
        grid.sync(); /* wave b-1 is complete: its survivor counts are final */
        stamp(p, 2 * b + 1);
        const uint32_t n_warps = gridDim.x * kWarpsPerCta;
        const uint32_t count = prepare_wave(p, wg, wc.qcount[b - 1], 64u * n_warps);
        if (count == 0)
        {
            synced = true;
            break;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && b < RVPT_MAX_BOUNCE_STATS)
            atomicAdd(&p.ctr->stats.active[b], (unsigned long long)count);
        if (count <= p.tail_threshold)
        {
            bounce_phase<kSmem, kOct, true, kBatch>(p, sc, b, wg, RVPT_WAVE_SPREAD, sort);
            stamp(p, 2 * b + 2);
            break;
        }
        /* wave b claims from counter set b & 1; the other set (last used by wave b-1, which is
         * complete) is zeroed now for wave b+1 — the grid barrier orders it */
        if (blockIdx.x == 0)
            for (uint32_t i = threadIdx.x; i < RVPT_CHUNK_SHARDS * 32u; i += blockDim.x)
                wc.bounce_ctr[(b + 1) & 1][i] = 0u;
        const uint32_t deal = count <= 64u * n_warps ? RVPT_WAVE_SPREAD : RVPT_WAVE_SHARDED;
        if (b + 1 < p.max_bounces && ((small_waves >> (b + 1)) & 1ull))
        {
            /* forecast: wave b+1 would be a tail anyway — its rays finish here, in their threads */
            bounce_phase<kSmem, kOct, true, kBatch>(p, sc, b, wg, deal, sort);
            stamp(p, 2 * b + 2);
            break;
        }
        bounce_phase<kSmem, kOct, false, kBatch>(p, sc, b, wg, deal, sort);
        stamp(p, 2 * b + 2);
    }
    if constexpr (kBatch)
    {
        if (!synced) grid.sync(); /* every sample of the batch is parked */
        stamp(p, RVPT_TIMELINE_SLOTS - 2u);
        resolve_phase(p, frame_consts);
        stamp(p, RVPT_TIMELINE_SLOTS - 1u);
    }
    finish_launch(p);
}

/* ======================================================================== */
/* unfused variant: one launch per wave (profiling / cross-check)             */
/* ======================================================================== */
template <bool kSmem>
__global__ void __launch_bounds__(kThreads) k_primary(const FrameParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    SceneViewT<kSmem> sc;
    if constexpr (kSmem)
    {
        stage_scene(smem, &bar, p.scene, p.layout.bytes);
        sc = make_view<true>(smem, p.layout);
    }
    else
        sc = make_view<false>(p.scene, p.layout);
    primary_phase<kSmem, false, false>(p, sc, false); /* one launch per wave: push order, no sort */
    finish_launch(p);
}

template <bool kSmem>
__global__ void __launch_bounds__(kThreads) k_bounce(const FrameParams p, const int b)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;

    WaveCounters& wc = p.ctr->wave;
    __shared__ WaveGroups wg;
    const uint32_t n_warps = gridDim.x * kWarpsPerCta;
    const uint32_t count = prepare_wave(p, wg, wc.qcount[b - 1], 64u * n_warps);
    if (count == 0) /* an empty wave costs nothing but the launch */
    {
        finish_launch(p);
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && b < RVPT_MAX_BOUNCE_STATS)
        atomicAdd(&p.ctr->stats.active[b], (unsigned long long)count);

    SceneViewT<kSmem> sc;
    if constexpr (kSmem)
    {
        stage_scene(smem, &bar, p.scene, p.layout.bytes);
        sc = make_view<true>(smem, p.layout);
    }
    else
        sc = make_view<false>(p.scene, p.layout);
    bounce_phase<kSmem, false, false>(p, sc, b, wg, count <= 64u * n_warps ? RVPT_WAVE_SPREAD : 0u, false);
    finish_launch(p);
}

/* ======================================================================== */
/* k_modes: the reference's other integrators (compute_pass.comp:68-99)      */
/* ======================================================================== */
/* Modes 0-8 — binary, color, depth, normal, Utah, ambient occlusion, Appel,
 * Whitted, Cook (integrators.glsl:24-543) — are the reference's teaching /
 * debug views, not the bounce loop: each pixel runs its integrator to the end
 * inside one thread (all `aa` samples, continuing the pixel's RNG stream), then
 * accumulates like any other pixel. Launched only when a quadrant asks for one
 * of them; Kajiya pixels stay on the wavefront path. */

/* intersect_bvh_any (intersection.glsl:417-463): same walk, returns at the
 * first accepted triangle; closest_t never shrinks. */
template <bool kSmem>
__device__ __forceinline__ bool trace_any(const SceneViewT<kSmem>& sc, rv_f3 o, rv_f3 d)
{
    const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    uint32_t node = 0;
    while (node != RVPT_NODE_END)
    {
        const float4 n0 = ld_f4<kSmem>(sc.nodes, 2 * node);
        const float4 n1 = ld_f4<kSmem>(sc.nodes, 2 * node + 1);
        const float fx = (n0.y - o.x) * ix, nx = (n0.x - o.x) * ix;
        const float fy = (n0.w - o.y) * iy, ny = (n0.z - o.y) * iy;
        const float fz = (n1.y - o.z) * iz, nz = (n1.x - o.z) * iz;
        float t1 = fminf(fmaxf(fx, nx), fminf(fmaxf(fy, ny), fmaxf(fz, nz)));
        float t0 = fmaxf(fminf(fx, nx), fmaxf(fminf(fy, ny), fminf(fz, nz)));
        t0 = fmaxf(t0, 0.0f);
        t1 = fminf(t1, RV_INF);
        const uint32_t skip = __float_as_uint(n1.z);
        const uint32_t leaf = __float_as_uint(n1.w);
        if (t1 >= t0)
        {
            if (!(leaf & RVPT_NODE_INNER))
            {
                uint32_t i = leaf, m;
                do
                {
                    const float4 A = ld_f4<kSmem>(sc.tris, 4 * i + 0);
                    const float4 B = ld_f4<kSmem>(sc.tris, 4 * i + 1);
                    const float4 C = ld_f4<kSmem>(sc.tris, 4 * i + 2);
                    const float4 D = ld_f4<kSmem>(sc.tris, 4 * i + 3);
                    m = ld_u32<kSmem>(sc.meta, 4u * i + 3u);
                    const float num = rv_dot(rv_make(A.x - o.x, A.y - o.y, A.z - o.z),
                                             rv_make(B.x, B.y, B.z));
                    const float den = rv_dot(d, rv_make(B.x, B.y, B.z));
                    const float t = num / den;
                    const float tx = t * d.x, ty = t * d.y, tz = t * d.z;
                    const rv_f3 p0 = rv_make((o.x + tx) - A.x, (o.y + ty) - A.y, (o.z + tz) - A.z);
                    const float bx = rv_dot(p0, rv_make(C.x, C.y, C.z));
                    const float by = rv_dot(p0, rv_make(D.x, D.y, D.z));
                    const float m0 = B.w * bx, m1 = C.w * by;
                    const float m2 = C.w * bx, m3 = D.w * by;
                    const float u = A.w * (m0 + m1);
                    const float v = A.w * (m2 + m3);
                    if (0.0f < t && t < RV_INF && 0.0f < u && 0.0f < v && u + v < 1.0f) return true;
                    ++i;
                } while (!(m & RVPT_TRI_LAST));
                node = skip;
            }
            else
                node = node + 1;
        }
        else
            node = skip;
    }
    return false;
}

struct HitInfo /* Isect + Material_new, intersection.glsl:37-72 */
{
    bool hit;
    float t;
    rv_f3 pos, normal; /* normal normalised, zero on a miss (intersect_scene :511-513) */
    rv_f3 base_color, emissive;
    float ior;
    int type;
};

template <bool kSmem>
__device__ __forceinline__ HitInfo intersect_scene_dev(const SceneViewT<kSmem>& sc, rv_f3 o, rv_f3 d)
{
    HitInfo h;
    uint32_t tri;
    trace_nearest<kSmem, false, false>(sc, o, d, h.t, tri);
    h.hit = tri != 0xFFFFFFFFu;
    h.pos = rv_make(0.0f, 0.0f, 0.0f);
    h.normal = rv_make(0.0f, 0.0f, 0.0f);
    h.base_color = h.emissive = rv_make(0.0f, 0.0f, 0.0f);
    h.ior = 0.0f;
    h.type = -1;
    if (h.hit)
    {
        const float4 U = ld_f4<kSmem>(sc.meta, tri);
        const uint32_t mi = __float_as_uint(U.w) & ~RVPT_TRI_LAST;
        const float4 M0 = ld_f4<kSmem>(sc.mats, 3 * mi + 0);
        const float4 M1 = ld_f4<kSmem>(sc.mats, 3 * mi + 1);
        h.normal = rv_make(U.x, U.y, U.z);
        h.pos = rv_add(o, rv_scale(h.t, d));
        h.base_color = rv_make(M0.x, M0.y, M0.z);
        h.ior = M0.w;
        h.emissive = rv_make(M1.x, M1.y, M1.z);
        h.type = __float_as_int(M1.w);
    }
    return h;
}

__device__ __forceinline__ rv_f3 splat3(float v) { return rv_make(v, v, v); }

__device__ __forceinline__ rv_f3 sky_no_remap(rv_f3 d) /* mix(white, blue, ray.direction.y) */
{
    return rv_make(rv_mix(1.0f, 0.2f, d.y), rv_mix(1.0f, 0.3f, d.y), rv_mix(1.0f, 0.7f, d.y));
}

/* cosine-weighted scatter around n: material.glsl:96-108, samples_mapping.glsl:39-60,112-131 */
__device__ __forceinline__ rv_f3 scatter_lambert(rv_f3 n, uint32_t* rng)
{
    const float u = rv_rand(rng);
    const float v = rv_rand(rng);
    const float phi = RV_TWO_PI * u;
    const float cos_theta = (1.0f - v) - v;
    const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float sn, cs;
    rv_sincos(phi, &sn, &cs);
    return rv_add(n, rv_make(sin_theta * cs, sin_theta * sn, cos_theta));
}

struct SurfaceDev
{
    rv_f3 pos, normal, dir_in;
    float cos_in, eta;
};

/* normal flip / eta selection shared by Whitted, Cook (and Kajiya): integrators.glsl:290-316 */
__device__ __forceinline__ SurfaceDev surface_of(rv_f3 d, const HitInfo& h)
{
    SurfaceDev s;
    s.pos = h.pos;
    s.normal = h.normal;
    s.dir_in = rv_normalize(d);
    const float cos_view = rv_dot(s.dir_in, s.normal);
    s.eta = h.ior;
    if (cos_view > 0.0f)
    {
        s.cos_in = cos_view;
        s.normal = rv_neg(s.normal);
    }
    else
    {
        s.cos_in = -cos_view;
        s.eta = 1.0f / s.eta;
    }
    return s;
}

/* mirror / dielectric branches, integrators.glsl:338-377; false = unknown material */
__device__ __forceinline__ bool specular_bounce(const SurfaceDev& s, const HitInfo& h, uint32_t* rng,
                                                rv_f3* o, rv_f3* d, rv_f3* thr)
{
    if (h.type == 1)
    {
        *o = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
        *d = rv_add(s.dir_in, rv_scale(s.cos_in + s.cos_in, s.normal));
    }
    else if (h.type == 2)
    {
        const float cos_out_sqr = 1.0f - (s.eta * s.eta) * (1.0f - s.cos_in * s.cos_in);
        float cos_out = 0.0f;
        bool refl = (cos_out_sqr <= 0.0f);
        if (!refl)
        {
            cos_out = sqrtf(fmaxf(0.0f, cos_out_sqr));
            const float ec = s.eta * s.cos_in, eo = s.eta * cos_out;
            const float r_perp = (ec - cos_out) / (ec + cos_out);
            const float r_par = (s.cos_in - eo) / (s.cos_in + eo);
            const float f_refl = 0.5f * (r_perp * r_perp + r_par * r_par);
            refl = (rv_rand(rng) < f_refl);
        }
        if (refl)
        {
            *o = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
            *d = rv_add(s.dir_in, rv_scale(s.cos_in + s.cos_in, s.normal));
        }
        else
        {
            *o = rv_sub(s.pos, rv_scale(RV_EPSILON, s.normal));
            *d = rv_add(rv_scale(s.eta, s.dir_in), rv_scale(s.eta * s.cos_in - cos_out, s.normal));
        }
    }
    else
        return false;
    *thr = rv_mul(*thr, h.base_color);
    return true;
}

template <bool kSmem>
__device__ rv_f3 eval_mode(const SceneViewT<kSmem>& sc, int mode, rv_f3 o, rv_f3 d, int max_bounces,
                           uint32_t* rng)
{
    /* normalize(vec3(0.5,1,0.3)) as folded into the shipped SPIR-V (rvpt_math.h) */
    const rv_f3 light_dir = rv_make(RV_LIGHT_DIR_X, RV_LIGHT_DIR_Y, RV_LIGHT_DIR_Z);
    if (mode == 0) /* binary :24-38 */
        return splat3(trace_any<kSmem>(sc, o, d) ? 1.0f : 0.0f);
    if (mode == 7 || mode == 8)
    {
        /* Whitted :254-403, Cook :407-543 */
        rv_f3 col = mode == 7 ? splat3(0.1f) : splat3(0.0f);
        rv_f3 thr = splat3(1.0f);
        for (int i = 0; i < max_bounces; ++i)
        {
            HitInfo h = intersect_scene_dev<kSmem>(sc, o, d);
            if (!h.hit) return rv_add(col, rv_mul(thr, sky_no_remap(d)));
            col = rv_add(col, rv_mul(thr, h.emissive));
            const SurfaceDev s = surface_of(d, h);
            if (h.type == 0)
            {
                const rv_f3 o2 = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
                if (mode == 7)
                {
                    if (trace_any<kSmem>(sc, o2, light_dir)) return col;
                    const float cos_light = fmaxf(0.0f, rv_dot(light_dir, s.normal));
                    const rv_f3 lit = rv_mul(rv_mul(thr, h.base_color), splat3(1.0f));
                    return rv_add(col, rv_scale(cos_light, lit));
                }
                const rv_f3 d2 = scatter_lambert(s.normal, rng);
                const rv_f3 lam = rv_scale(RV_PI, rv_scale(RV_INV_PI, h.base_color));
                thr = rv_mul(thr, lam);
                h = intersect_scene_dev<kSmem>(sc, o2, d2);
                if (!h.hit) return rv_add(col, rv_mul(thr, sky_no_remap(d2)));
                return rv_add(col, rv_mul(thr, h.emissive));
            }
            if (!specular_bounce(s, h, rng, &o, &d, &thr)) return splat3(0.0f);
        }
        return splat3(0.0f);
    }

    const HitInfo h = intersect_scene_dev<kSmem>(sc, o, d);
    if (mode == 1) return h.hit ? h.base_color : splat3(0.0f); /* color :42-59 */
    if (mode == 2)                                              /* depth :63-82 */
    {
        const float len = sqrtf(rv_dot(d, d));
        return splat3(1.0f / (len * h.t));
    }
    if (mode == 3) /* normal :86-102 */
    {
        const float k = 0.5f * (h.hit ? 1.0f : 0.0f);
        return rv_make(0.5f * h.normal.x + k, 0.5f * h.normal.y + k, 0.5f * h.normal.z + k);
    }
    if (mode == 4) /* Utah :106-148 */
    {
        if (!h.hit) return sky_no_remap(d);
        const rv_f3 col = rv_add(splat3(0.1f), h.emissive);
        const rv_f3 n = rv_dot(d, h.normal) < 0.0f ? h.normal : rv_neg(h.normal);
        const float cos_light = fmaxf(0.0f, rv_dot(light_dir, n));
        return rv_add(col, rv_scale(cos_light, rv_mul(h.base_color, splat3(1.0f))));
    }
    if (mode == 5) /* ambient occlusion :152-200, nrays = max_bounces */
    {
        if (!h.hit) return splat3(0.0f);
        const rv_f3 n = rv_dot(d, h.normal) < 0.0f ? h.normal : rv_neg(h.normal);
        float acc = 0.0f;
        for (int i = 0; i < max_bounces; ++i)
        {
            const rv_f3 o2 = rv_add(h.pos, rv_scale(RV_EPSILON, n));
            const rv_f3 d2 = scatter_lambert(n, rng);
            acc += trace_any<kSmem>(sc, o2, d2) ? 1.0f : 0.0f;
        }
        return splat3(1.0f - acc / (float)max_bounces);
    }
    /* mode 6: Appel :204-250 */
    if (!h.hit) return splat3(1.0f);
    const rv_f3 dir_in = rv_normalize(d);
    const rv_f3 n = rv_dot(dir_in, h.normal) > 0.0f ? rv_neg(h.normal) : h.normal;
    if (trace_any<kSmem>(sc, rv_add(h.pos, rv_scale(RV_EPSILON, n)), light_dir)) return splat3(0.0f);
    return splat3(1.0f * fmaxf(0.0f, rv_dot(light_dir, n)));
}

/* Mode 10 and every other index outside 0..9: integrator_Hart (integrators.glsl:681-693), the
 * sphere tracer of distance_functions.glsl:70-116 shown as a heat map of its iteration count.
 * It marches against EVERY triangle of the buffer (no BVH), so it reads the caller's vertices —
 * p.raw_tris, the 64-byte records in upload order — through the read-only path: all lanes of a
 * warp ask for the same address, one broadcast transaction per record. */
__device__ __forceinline__ float sign_glsl(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

/* distance_functions.glsl:36-61 (squared distance; the caller takes the root) */
__device__ __forceinline__ float distance_triangle_sq(rv_f3 p, rv_f3 a, rv_f3 b, rv_f3 c)
{
    const rv_f3 ba = rv_sub(b, a), pa = rv_sub(p, a);
    const rv_f3 cb = rv_sub(c, b), pb = rv_sub(p, b);
    const rv_f3 ac = rv_sub(a, c), pc = rv_sub(p, c);
    const rv_f3 nor = rv_cross(ba, ac);
    const float side = (sign_glsl(rv_dot(rv_cross(ba, nor), pa)) + sign_glsl(rv_dot(rv_cross(cb, nor), pb))) +
                       sign_glsl(rv_dot(rv_cross(ac, nor), pc));
    if (side < 2.0f)
    {
        const rv_f3 q0 = rv_sub(rv_scale(clamp01(rv_dot(ba, pa) / rv_dot(ba, ba)), ba), pa);
        const rv_f3 q1 = rv_sub(rv_scale(clamp01(rv_dot(cb, pb) / rv_dot(cb, cb)), cb), pb);
        const rv_f3 q2 = rv_sub(rv_scale(clamp01(rv_dot(ac, pc) / rv_dot(ac, ac)), ac), pc);
        return fminf(fminf(rv_dot(q0, q0), rv_dot(q1, q1)), rv_dot(q2, q2));
    }
    const float h = rv_dot(nor, pa);
    return (h * h) / rv_dot(nor, nor);
}

__device__ __noinline__ rv_f3 integrator_hart(const FrameParams& p, rv_f3 o, rv_f3 d)
{
    const float4* __restrict__ T = p.raw_tris;
    rv_f3 pos = rv_add(o, rv_scale(0.0f, d)); /* p = origin + mint * direction, mint = 0 */
    int i = 0;
    for (; i < 32; ++i) /* MARCH_ITER, compute_pass.comp:10 */
    {
        float radius = RV_INF;
        for (uint32_t j = 0; j < p.n_raw_tris; ++j)
        {
            const float4 a = __ldg(T + 4 * j), b = __ldg(T + 4 * j + 1), c = __ldg(T + 4 * j + 2);
            const float dist = sqrtf(distance_triangle_sq(pos, rv_make(a.x, a.y, a.z), rv_make(b.x, b.y, b.z),
                                                          rv_make(c.x, c.y, c.z)));
            radius = radius < dist ? radius : dist; /* min_idx: lhs.x < rhs.x ? lhs : rhs */
        }
        radius = fminf(RV_INF, radius);
        if (radius < 0.1f || radius > RV_INF) break; /* MARCH_EPS; maxt = INF */
        pos = rv_add(pos, rv_scale(radius, d));
    }
    const float g = (float)i / 31.0f;
    return rv_make(g, g, g);
}

template <bool kSmem>
__global__ void __launch_bounds__(kThreads) k_modes(const FrameParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    SceneViewT<kSmem> sc;
    if constexpr (kSmem)
    {
        stage_scene(smem, &bar, p.scene, p.layout.bytes);
        sc = make_view<true>(smem, p.layout);
    }
    else
        sc = make_view<false>(p.scene, p.layout);

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * kWarpsPerCta;
    for (uint32_t c = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); c < p.n_chunks; c += n_warps)
    {
        const uint32_t slot = c * 32u + lane;
        uint32_t x, y;
        slot_to_xy(p, slot, x, y);
        if (!((x < p.W_eff) && (y < p.H_eff) && ((slot >> 8) * p.nranks + p.rank < p.n_tiles))) continue;
        const int mode = integrator_of(p, x, y);
        if (mode == 9) continue; /* on the wavefront path */
        uint32_t rng = rv_wang_hash(x + y * p.W) + p.frame;
        rv_f3 sum = rv_make(0.0f, 0.0f, 0.0f);
        for (int i = 0; i < p.aa; ++i)
        {
            const float jx = rv_rand(&rng);
            const float jy = rv_rand(&rng);
            const float cx = ((float)x + jx) * p.inv_dim_x;
            float cy = ((float)y + jy) * p.inv_dim_y;
            cy = 1.0f - cy;
            rv_f3 o, d;
            camera_ray(p, cx, cy, o, d);
            sum = rv_add(sum, (mode >= 0 && mode <= 8) ? eval_mode<kSmem>(sc, mode, o, d, p.max_bounces, &rng)
                                                       : integrator_hart(p, o, d));
        }
        accumulate_pixel(p, slot, rv_make(sum.x / p.aa_f, sum.y / p.aa_f, sum.z / p.aa_f), y * p.W + x);
    }
}

/* ======================================================================== */
/* tile <-> raster                                                           */
/* ======================================================================== */
/* src: [n_src_ranks][n_local_padded][256] elements of `words` 32-bit words;
 * element (r, j, q) is pixel q of global tile j*nranks + first_rank + r. */
__global__ void k_untile(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                         uint32_t words, uint32_t W, uint32_t H, uint32_t tiles_x, uint32_t n_tiles,
                         uint32_t nranks, uint32_t first_rank, uint32_t n_src_ranks,
                         uint32_t n_local_padded)
{
    const uint64_t total = (uint64_t)n_src_ranks * n_local_padded * RVPT_TILE_PIXELS;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (uint64_t)gridDim.x * blockDim.x)
    {
        const uint32_t q = (uint32_t)(e & 255u);
        const uint64_t tj = e >> 8;
        const uint32_t j = (uint32_t)(tj % n_local_padded);
        const uint32_t r = (uint32_t)(tj / n_local_padded);
        const uint32_t g = j * nranks + first_rank + r;
        if (g >= n_tiles) continue;
        const uint32_t ty = g / tiles_x, tx = g - ty * tiles_x;
        const uint32_t w = q >> 5, lane = q & 31u;
        const uint32_t x = tx * RVPT_TILE_DIM + ((w & 1u) << 3) + (lane & 7u);
        const uint32_t y = ty * RVPT_TILE_DIM + ((w >> 1) << 2) + (lane >> 3);
        if (x >= W || y >= H) continue;
        const uint64_t d = ((uint64_t)y * W + x) * words;
        if (words == 4)
            reinterpret_cast<uint4*>(dst)[d >> 2] = reinterpret_cast<const uint4*>(src)[e];
        else
            for (uint32_t k = 0; k < words; ++k) dst[d + k] = src[e * words + k];
    }
}

/* inverse of k_untile for one rank's local tiles (checkpoint restore) */
__global__ void k_tile(const uint32_t* __restrict__ raster, uint32_t* __restrict__ tiles,
                       uint32_t words, uint32_t W, uint32_t H, uint32_t tiles_x, uint32_t n_tiles,
                       uint32_t nranks, uint32_t rank, uint32_t n_local_padded)
{
    const uint64_t total = (uint64_t)n_local_padded * RVPT_TILE_PIXELS;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (uint64_t)gridDim.x * blockDim.x)
    {
        const uint32_t q = (uint32_t)(e & 255u);
        const uint32_t j = (uint32_t)(e >> 8);
        const uint32_t g = j * nranks + rank;
        uint32_t x = 0, y = 0;
        bool ok = g < n_tiles;
        if (ok)
        {
            const uint32_t ty = g / tiles_x, tx = g - ty * tiles_x;
            const uint32_t w = q >> 5, lane = q & 31u;
            x = tx * RVPT_TILE_DIM + ((w & 1u) << 3) + (lane & 7u);
            y = ty * RVPT_TILE_DIM + ((w >> 1) << 2) + (lane >> 3);
            ok = x < W && y < H;
        }
        for (uint32_t k = 0; k < words; ++k)
            tiles[e * words + k] = ok ? raster[((uint64_t)y * W + x) * words + k] : 0u;
    }
}

/* rgba8 temporal image -> float4 (k/255) for read_accum in ACCUM_RGBA8 mode */
__global__ void k_u8_to_f32(const uchar4* __restrict__ src, float4* __restrict__ dst, uint64_t n)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (uint64_t)gridDim.x * blockDim.x)
    {
        const uchar4 k = src[e];
        dst[e] = make_float4(rv_unorm8_load(k.x), rv_unorm8_load(k.y), rv_unorm8_load(k.z), 0.0f);
    }
}
__global__ void k_f32_to_u8(const float4* __restrict__ src, uchar4* __restrict__ dst, uint64_t n)
{
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (uint64_t)gridDim.x * blockDim.x)
    {
        const float4 a = src[e];
        dst[e] = make_uchar4(dev_unorm8(a.x), dev_unorm8(a.y), dev_unorm8(a.z), 0);
    }
}

/* ---- shared-header self test ---------------------------------------------- */
__global__ void k_selftest(int op, const float* __restrict__ in, size_t n, float* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0)
    {
        float s, c;
        rv_sincos(in[i], &s, &c);
        out[2 * i] = s;
        out[2 * i + 1] = c;
    }
    else if (op == 1)
    {
        uint32_t st = __float_as_uint(in[i]);
        out[2 * i] = rv_rand(&st);
        out[2 * i + 1] = rv_rand(&st);
    }
    else if (op == 2)
    {
        const rv_f3 r = rv_normalize(rv_make(in[3 * i], in[3 * i + 1], in[3 * i + 2]));
        out[3 * i] = r.x;
        out[3 * i + 1] = r.y;
        out[3 * i + 2] = r.z;
    }
    else if (op == 3)
    {
        /* contraction probe + tan: in = (a, b, c, x) */
        out[2 * i] = (float)rv_contract_probe(in[4 * i], in[4 * i + 1], in[4 * i + 2]);
        out[2 * i + 1] = 1.0f / rv_tan(0.5f * in[4 * i + 3]);
    }
}

} /* namespace */

/* ======================================================================== */
/* launch wrappers (called from engine.cu)                                   */
/* ======================================================================== */
namespace rvpt
{

static size_t smem_bytes_for(const FrameParams& p, bool smem) { return smem ? p.layout.bytes : 0; }

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t configure_kernels(size_t max_dynamic_smem)
{
    cudaError_t e;
    if ((e = set_smem(k_frame<true, true, true, false>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, false, true, false>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, true, false, false>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, false, false, false>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, true, true, true>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, false, true, true>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, true, false, true>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_frame<true, false, false, true>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_primary<true>, max_dynamic_smem)) != cudaSuccess) return e;
    if ((e = set_smem(k_bounce<true>, max_dynamic_smem)) != cudaSuccess) return e;
    return set_smem(k_modes<true>, max_dynamic_smem);
}

size_t frame_smem_bytes(size_t scene_bytes, uint32_t n_nodes, uint32_t n_tris, bool oct)
{
    /* blob + one float per triangle + the derived node copies: 8 absolute + 8
     * origin-relative octant copies, or one origin-relative copy */
    const size_t rel_num = ((size_t)n_tris * 4u + 15u) & ~(size_t)15u;
    if (!oct) return scene_bytes + rel_num + (size_t)n_nodes * 32u;
    /* 16 copies of the first record halves, and the second halves RVPT_OCT_B_OFFSET further */
    if ((size_t)n_nodes * 16u * 16u > RVPT_OCT_B_OFFSET) return ~(size_t)0; /* does not fit: no octant copies */
    return scene_bytes + rel_num + RVPT_OCT_B_OFFSET + (size_t)n_nodes * 16u * 16u;
}

cudaError_t occupancy(int* frame_ctas_per_sm, int* primary_ctas_per_sm, int* bounce_ctas_per_sm,
                      bool smem, bool oct, size_t scene_bytes, uint32_t n_nodes, uint32_t n_tris)
{
    const size_t dyn = smem ? scene_bytes : 0;
    cudaError_t e;
    int classic = 0, batch = 0;
    if (smem)
    {
        const size_t fdyn = frame_smem_bytes(scene_bytes, n_nodes, n_tris, oct);
        if (oct)
        {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&classic, k_frame<true, true, true, false>, kThreads, fdyn);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&batch, k_frame<true, true, true, true>, kThreads, fdyn);
        }
        else
        {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&classic, k_frame<true, true, false, false>, kThreads, fdyn);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&batch, k_frame<true, true, false, true>, kThreads, fdyn);
        }
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(primary_ctas_per_sm, k_primary<true>,
                                                          kThreads, dyn);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(bounce_ctas_per_sm, k_bounce<true>,
                                                          kThreads, dyn);
    }
    else
    {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&classic, k_frame<false, false, false, false>, kThreads, dyn);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&batch, k_frame<false, false, false, true>, kThreads, dyn);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(primary_ctas_per_sm, k_primary<false>,
                                                          kThreads, dyn);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(bounce_ctas_per_sm, k_bounce<false>,
                                                          kThreads, dyn);
    }
    /* one grid size for both flavours of the frame kernel (tail thresholds, timeline) */
    *frame_ctas_per_sm = classic < batch ? classic : batch;
    return e;
}

template <bool kBatch>
static const void* frame_kernel(const FrameParams& p, bool smem, bool oct)
{
    if (!smem) return (const void*)k_frame<false, false, false, kBatch>;
    /* the ortho camera has a different origin per ray: no shared-origin copies */
    const bool rel = p.camera_mode != 1;
    if (oct)
        return rel ? (const void*)k_frame<true, true, true, kBatch> : (const void*)k_frame<true, false, true, kBatch>;
    return rel ? (const void*)k_frame<true, true, false, kBatch> : (const void*)k_frame<true, false, false, kBatch>;
}

/* p.n_batch > 0 selects the batched flavour (frames p.frame .. p.frame + n_batch - 1) */
cudaError_t launch_frame(const FrameParams& p, bool smem, bool oct, int grid, cudaStream_t st)
{
    void* args[] = {const_cast<FrameParams*>(&p)};
    const size_t dyn = smem ? frame_smem_bytes(p.layout.bytes, p.layout.n_nodes, p.layout.n_tris, oct) : 0;
    const void* k = p.n_batch ? frame_kernel<true>(p, smem, oct) : frame_kernel<false>(p, smem, oct);
    return cudaLaunchCooperativeKernel(k, dim3(grid), dim3(kThreads), args, dyn, st);
}

cudaError_t launch_primary(const FrameParams& p, bool smem, int grid, cudaStream_t st)
{
    if (smem)
        k_primary<true><<<grid, kThreads, smem_bytes_for(p, true), st>>>(p);
    else
        k_primary<false><<<grid, kThreads, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_bounce(const FrameParams& p, int b, bool smem, int grid, cudaStream_t st)
{
    if (smem)
        k_bounce<true><<<grid, kThreads, smem_bytes_for(p, true), st>>>(p, b);
    else
        k_bounce<false><<<grid, kThreads, 0, st>>>(p, b);
    return cudaGetLastError();
}

cudaError_t launch_modes(const FrameParams& p, bool smem, int grid, cudaStream_t st)
{
    if (smem)
        k_modes<true><<<grid, kThreads, smem_bytes_for(p, true), st>>>(p);
    else
        k_modes<false><<<grid, kThreads, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_untile(const void* src, void* dst, uint32_t words, uint32_t W, uint32_t H,
                          uint32_t tiles_x, uint32_t n_tiles, uint32_t nranks, uint32_t first_rank,
                          uint32_t n_src_ranks, uint32_t n_local_padded, cudaStream_t st)
{
    k_untile<<<592, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, words, W, H, tiles_x, n_tiles,
                                  nranks, first_rank, n_src_ranks, n_local_padded);
    return cudaGetLastError();
}

cudaError_t launch_tile(const void* raster, void* tiles, uint32_t words, uint32_t W, uint32_t H,
                        uint32_t tiles_x, uint32_t n_tiles, uint32_t nranks, uint32_t rank,
                        uint32_t n_local_padded, cudaStream_t st)
{
    k_tile<<<592, 256, 0, st>>>((const uint32_t*)raster, (uint32_t*)tiles, words, W, H, tiles_x,
                                n_tiles, nranks, rank, n_local_padded);
    return cudaGetLastError();
}

cudaError_t launch_u8_to_f32(const void* src, void* dst, uint64_t n, cudaStream_t st)
{
    k_u8_to_f32<<<592, 256, 0, st>>>((const uchar4*)src, (float4*)dst, n);
    return cudaGetLastError();
}

cudaError_t launch_f32_to_u8(const void* src, void* dst, uint64_t n, cudaStream_t st)
{
    k_f32_to_u8<<<592, 256, 0, st>>>((const float4*)src, (uchar4*)dst, n);
    return cudaGetLastError();
}

cudaError_t launch_selftest(int op, const float* in, size_t n, float* out, cudaStream_t st)
{
    k_selftest<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op, in, n, out);
    return cudaGetLastError();
}

int threads_per_cta() { return kThreads; }

} /* namespace rvpt */
