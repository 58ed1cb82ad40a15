/*
 * bvh_gpu.cu — BVH construction on the GPU ("host BVH build as a correct, fast
 * builder, CPU then GPU"), emitting the reference's node format (src/rvpt/bvh.h:12-18: children
 * at first and first + 1, leaf iff primitive_count > 0, bounds minx,maxx,miny,maxy,minz,maxz)
 * and the primitive permutation of Bvh::permute_primitives (bvh.h:70-77).
 *
 * A linear BVH: 30-bit Morton codes of the triangle centroids (the reference's builder also bins
 * centroids, bvh_builder.h:20-27), an LSD radix sort (8-bit digits, one warp per tile: match.any
 * ranks the equal digits of 32 keys stably), the Karras 2012 hierarchy — every inner node finds
 * its key range and split from common-prefix lengths, fully parallel — and a bottom-up fit of the
 * boxes (the second thread to reach a node merges its children). One triangle per leaf, 2n - 1
 * nodes, depth <= 30 + log2(n) < the reference's 64-entry traversal stack. The tree is not the
 * reference's (its builder asserts on most inputs) and not the SAH tree of
 * bvh_build.cpp (which traverses ~1.3x faster); the contract is "a valid BVH in the reference's
 * format", built in milliseconds: 560 k triangles in a few ms against 0.9 s on the host.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/rvpt_abi.h"

namespace
{

#define CU_OK(call)                       \
    do                                    \
    {                                     \
        if ((call) != cudaSuccess)        \
        {                                 \
            rc = RVPT_B200_ECUDA;         \
            goto done;                    \
        }                                 \
    } while (0)

/* float atomics through the ordered-int trick (min for lo, max for hi) */
__device__ __forceinline__ void atomic_min_f(float* a, float v)
{
    if (v >= 0.0f)
        atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else
        atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v)
{
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}

struct Tri64
{
    float4 v0, v1, v2, mat;
};

/* Triangle::center() (geometry.h:103-110): (v0 + v1 + v2) * (1/3) */
__device__ __forceinline__ float3 centroid(const Tri64& t)
{
    const float k = 1.0f / 3.0f;
    return make_float3(((t.v0.x + t.v1.x) + t.v2.x) * k, ((t.v0.y + t.v1.y) + t.v2.y) * k,
                       ((t.v0.z + t.v1.z) + t.v2.z) * k);
}

__global__ void k_centroid_bounds(const Tri64* tris, uint32_t n, float* box /* lo xyz, hi xyz */)
{
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float3 c = centroid(tris[i]);
        if (c.x == c.x && c.y == c.y && c.z == c.z) /* NaN centroids (degenerate input) do not shape the grid */
        {
            lo[0] = fminf(lo[0], c.x), lo[1] = fminf(lo[1], c.y), lo[2] = fminf(lo[2], c.z);
            hi[0] = fmaxf(hi[0], c.x), hi[1] = fmaxf(hi[1], c.y), hi[2] = fmaxf(hi[2], c.z);
        }
    }
    for (int a = 0; a < 3; ++a)
    {
        for (int d = 16; d > 0; d >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], d));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], d));
        }
        if ((threadIdx.x & 31u) == 0)
        {
            atomic_min_f(&box[a], lo[a]);
            atomic_max_f(&box[3 + a], hi[a]);
        }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) /* 10 bits -> every third bit */
{
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void k_morton(const Tri64* tris, uint32_t n, const float* box, uint32_t* keys, uint32_t* vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 c = centroid(tris[i]);
    uint32_t q[3];
    const float cc[3] = {c.x, c.y, c.z};
    for (int a = 0; a < 3; ++a)
    {
        const float ext = box[3 + a] - box[a];
        float f = ext > 0.0f ? (cc[a] - box[a]) / ext : 0.0f;
        f = fminf(fmaxf(f, 0.0f), 1.0f); /* NaN -> 0 */
        q[a] = min((uint32_t)(f * 1024.0f), 1023u);
    }
    keys[i] = (spread10(q[0]) << 2) | (spread10(q[1]) << 1) | spread10(q[2]);
    vals[i] = i;
}

/* ---- LSD radix sort, 8-bit digits. Tiles of kTile consecutive elements, one warp per tile. ---- */
constexpr uint32_t kTile = 2048;

__global__ void k_radix_hist(const uint32_t* keys, uint32_t n, uint32_t shift, uint32_t n_tiles, uint32_t* hist /* [256][n_tiles] */)
{
    __shared__ uint32_t cnt[256];
    const uint32_t tile = blockIdx.x;
    for (uint32_t d = threadIdx.x; d < 256; d += blockDim.x) cnt[d] = 0;
    __syncthreads();
    const uint32_t lo = tile * kTile, hi = min(lo + kTile, n);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&cnt[(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < 256; d += blockDim.x) hist[d * n_tiles + tile] = cnt[d];
}

/* exclusive scan of hist in (digit, tile) order, one block */
__global__ void k_radix_scan(uint32_t* hist, uint32_t total)
{
    __shared__ uint32_t part[1024];
    const uint32_t per = (total + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, total);
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t run = 0;
        for (uint32_t t = 0; t < blockDim.x; ++t)
        {
            const uint32_t v = part[t];
            part[t] = run;
            run += v;
        }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i)
    {
        const uint32_t v = hist[i];
        hist[i] = run;
        run += v;
    }
}

/* stable scatter: one warp walks its tile 32 keys at a time */
__global__ void k_radix_scatter(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                                uint32_t n, uint32_t shift, uint32_t n_tiles, const uint32_t* offsets)
{
    __shared__ uint32_t base[256];
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    for (uint32_t d = lane; d < 256; d += 32) base[d] = offsets[d * n_tiles + tile];
    __syncwarp();
    const uint32_t lo = tile * kTile, hi = min(lo + kTile, n);
    for (uint32_t i0 = lo; i0 < hi; i0 += 32)
    {
        const uint32_t i = i0 + lane;
        const bool on = i < hi;
        const uint32_t k = on ? keys_in[i] : 0u, v = on ? vals_in[i] : 0u;
        const uint32_t d = on ? ((k >> shift) & 255u) : 256u + lane; /* idle lanes match nobody */
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (on) pos = base[d] + rank;
        __syncwarp();
        if (on && rank == (uint32_t)__popc(peers) - 1u) base[d] = pos + 1u; /* the last of the group advances the digit */
        __syncwarp();
        if (on)
        {
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
    }
}

/* ---- Karras 2012 ---- */
__device__ __forceinline__ int delta(const uint32_t* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j); /* equal codes: the index breaks the tie */
    return __clz(a ^ b);
}

/* inner node i in [0, n-1): children[2i], children[2i+1] (index < n-1: inner, else leaf + (n-1)); parents */
__global__ void k_hierarchy(const uint32_t* keys, int n, uint32_t* children, uint32_t* parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) > delta(keys, n, i, i - 1) ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1)
    {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t left = lo == gamma ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    const uint32_t right = hi == gamma + 1 ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    children[2 * i] = left, children[2 * i + 1] = right;
    parent[left] = (uint32_t)i << 1;
    parent[right] = ((uint32_t)i << 1) | 1u;
}

/* Bottom-up box fit + emission in the reference's format. Node ids: inner i, leaf n-1+k. The
 * children of inner node i live at output slots 1 + 2i and 2 + 2i, the root (inner 0) at 0. */
__global__ void k_fit_and_emit(const Tri64* tris, const uint32_t* vals, int n, const uint32_t* children,
                               const uint32_t* parent, uint32_t* visits, float* boxes /* [2n-1][6] */,
                               rvpt_bvh_node* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Tri64 t = tris[vals[k]];
    /* Triangle::aabb(): AABB(v0).expand(v1).expand(v2) (geometry.h:97-102) */
    float b[6] = {fminf(fminf(t.v0.x, t.v1.x), t.v2.x), fmaxf(fmaxf(t.v0.x, t.v1.x), t.v2.x),
                  fminf(fminf(t.v0.y, t.v1.y), t.v2.y), fmaxf(fmaxf(t.v0.y, t.v1.y), t.v2.y),
                  fminf(fminf(t.v0.z, t.v1.z), t.v2.z), fmaxf(fmaxf(t.v0.z, t.v1.z), t.v2.z)};
    uint32_t id = (uint32_t)(n - 1 + k);
    {
        const uint32_t slot = n == 1 ? 0u : 1u + parent[id];
        rvpt_bvh_node nd;
        nd.first_child_or_primitive = (uint32_t)k;
        nd.primitive_count = 1;
        for (int a = 0; a < 6; ++a) nd.bounds[a] = b[a];
        out[slot] = nd;
        for (int a = 0; a < 6; ++a) boxes[(size_t)id * 6 + a] = b[a];
    }
    while (id != 0u)
    {
        const uint32_t p = parent[id] >> 1;
        __threadfence();
        if (atomicAdd(&visits[p], 1u) == 0u) return; /* the sibling subtree is not done: its thread finishes p */
        __threadfence();
        const uint32_t sib = children[2 * p] == id ? children[2 * p + 1] : children[2 * p];
        const volatile float* sb = boxes + (size_t)sib * 6;
        b[0] = fminf(b[0], sb[0]), b[1] = fmaxf(b[1], sb[1]), b[2] = fminf(b[2], sb[2]);
        b[3] = fmaxf(b[3], sb[3]), b[4] = fminf(b[4], sb[4]), b[5] = fmaxf(b[5], sb[5]);
        for (int a = 0; a < 6; ++a) boxes[(size_t)p * 6 + a] = b[a];
        rvpt_bvh_node nd;
        nd.first_child_or_primitive = 1u + 2u * p;
        nd.primitive_count = 0;
        for (int a = 0; a < 6; ++a) nd.bounds[a] = b[a];
        out[p == 0u ? 0u : 1u + parent[p]] = nd;
        id = p;
    }
}

} /* namespace */

extern "C" int rvpt_b200_build_bvh_gpu(int device, const rvpt_triangle* triangles, size_t n_triangles,
                                       rvpt_bvh_node* nodes_out, size_t* n_nodes_out, uint32_t* prim_indices_out,
                                       float* build_ms_out)
{
    if (!triangles || !nodes_out || !n_nodes_out || !prim_indices_out || n_triangles == 0 || n_triangles > 0x3FFFFFFFu)
        return RVPT_B200_EINVAL;
    static_assert(sizeof(Tri64) == sizeof(rvpt_triangle), "triangle layout");
    const uint32_t n = (uint32_t)n_triangles;
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    int rc = RVPT_B200_OK;
    Tri64* d_tris = nullptr;
    uint32_t *d_keys[2] = {nullptr, nullptr}, *d_vals[2] = {nullptr, nullptr}, *d_hist = nullptr;
    uint32_t *d_children = nullptr, *d_parent = nullptr, *d_visits = nullptr;
    float *d_box = nullptr, *d_boxes = nullptr;
    rvpt_bvh_node* d_nodes = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const float box_init[6] = {3.4e38f, 3.4e38f, 3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    const size_t n_nodes = 2 * (size_t)n - 1;
    int cur = 0;

    if (cudaSetDevice(device) != cudaSuccess) return RVPT_B200_ECUDA;
    CU_OK(cudaMalloc(&d_tris, (size_t)n * sizeof(Tri64)));
    for (int k = 0; k < 2; ++k)
    {
        CU_OK(cudaMalloc(&d_keys[k], (size_t)n * 4));
        CU_OK(cudaMalloc(&d_vals[k], (size_t)n * 4));
    }
    CU_OK(cudaMalloc(&d_hist, (size_t)256 * n_tiles * 4));
    CU_OK(cudaMalloc(&d_children, (size_t)2 * n * 4));
    CU_OK(cudaMalloc(&d_parent, n_nodes * 4));
    CU_OK(cudaMalloc(&d_visits, (size_t)n * 4));
    CU_OK(cudaMalloc(&d_box, 6 * sizeof(float)));
    CU_OK(cudaMalloc(&d_boxes, n_nodes * 6 * sizeof(float)));
    CU_OK(cudaMalloc(&d_nodes, n_nodes * sizeof(rvpt_bvh_node)));
    CU_OK(cudaMemcpy(d_tris, triangles, (size_t)n * sizeof(Tri64), cudaMemcpyHostToDevice));
    CU_OK(cudaMemcpy(d_box, box_init, sizeof(box_init), cudaMemcpyHostToDevice));
    CU_OK(cudaMemset(d_visits, 0, (size_t)n * 4));
    CU_OK(cudaMemset(d_parent, 0, n_nodes * 4));
    CU_OK(cudaEventCreate(&e0));
    CU_OK(cudaEventCreate(&e1));
    CU_OK(cudaEventRecord(e0, 0));

    k_centroid_bounds<<<296, 256>>>(d_tris, n, d_box);
    k_morton<<<(n + 255) / 256, 256>>>(d_tris, n, d_box, d_keys[0], d_vals[0]);
    for (uint32_t shift = 0; shift < 32; shift += 8)
    {
        k_radix_hist<<<n_tiles, 256>>>(d_keys[cur], n, shift, n_tiles, d_hist);
        k_radix_scan<<<1, 1024>>>(d_hist, 256 * n_tiles);
        k_radix_scatter<<<n_tiles, 32>>>(d_keys[cur], d_vals[cur], d_keys[cur ^ 1], d_vals[cur ^ 1], n, shift, n_tiles, d_hist);
        cur ^= 1;
    }
    if (n > 1) k_hierarchy<<<(n - 1 + 255) / 256, 256>>>(d_keys[cur], (int)n, d_children, d_parent);
    k_fit_and_emit<<<(n + 255) / 256, 256>>>(d_tris, d_vals[cur], (int)n, d_children, d_parent, d_visits, d_boxes, d_nodes);
    CU_OK(cudaEventRecord(e1, 0));
    CU_OK(cudaGetLastError());
    CU_OK(cudaEventSynchronize(e1));
    if (build_ms_out) CU_OK(cudaEventElapsedTime(build_ms_out, e0, e1));
    CU_OK(cudaMemcpy(nodes_out, d_nodes, n_nodes * sizeof(rvpt_bvh_node), cudaMemcpyDeviceToHost));
    CU_OK(cudaMemcpy(prim_indices_out, d_vals[cur], (size_t)n * 4, cudaMemcpyDeviceToHost));
    *n_nodes_out = n_nodes;
done:
    cudaFree(d_tris);
    for (int k = 0; k < 2; ++k)
    {
        cudaFree(d_keys[k]);
        cudaFree(d_vals[k]);
    }
    cudaFree(d_hist);
    cudaFree(d_children);
    cudaFree(d_parent);
    cudaFree(d_visits);
    cudaFree(d_box);
    cudaFree(d_boxes);
    cudaFree(d_nodes);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}
