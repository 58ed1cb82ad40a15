/*
 * bvh_build.cpp — host-side binned-SAH BVH builder emitting the reference's
 * node format (src/rvpt/bvh.h:12-18: children at first and first+1, leaf iff
 * primitive_count > 0, bounds minx,maxx,miny,maxy,minz,maxz).
 *
 * Counterpart of BinnedBvhBuilder (src/rvpt/bvh_builder.cpp:11-199) with the
 * same knobs (16 bins, leaves of 2..8 triangles unless SAH says otherwise,
 * median fallback) but not its tree: the reference partitions with a squared
 * bin index (bvh_builder.cpp:41-47) and mis-parenthesises its median
 * (bvh_builder.cpp:167), so it asserts / recurses forever on the built-in
 * scene. The contract here is "a valid BVH in the reference's
 * format", which is all the shader's traversal relies on.
 */
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/rvpt_abi.h"

namespace
{

constexpr size_t kMinLeaf = 2;  /* nodes below this are never split (bvh_builder.h:47) */
constexpr size_t kMaxLeaf = 8;  /* nodes above this are always split (bvh_builder.h:49) */
constexpr int kBins = 16;       /* bvh_builder.h:51 */
constexpr int kMaxSahDepth = 32; /* beyond this only median splits: bounds the depth under 64 */

struct Box
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    float hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const float* p)
    {
        for (int a = 0; a < 3; ++a)
        {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    void grow(const Box& b)
    {
        for (int a = 0; a < 3; ++a)
        {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    float half_area() const
    {
        const float dx = std::max(hi[0] - lo[0], 0.0f), dy = std::max(hi[1] - lo[1], 0.0f),
                    dz = std::max(hi[2] - lo[2], 0.0f);
        return dx * (dy + dz) + dy * dz;
    }
};

struct Builder
{
    const std::vector<Box>& boxes;
    const std::vector<float>& centers; /* 3 per primitive */
    std::vector<uint32_t>& indices;
    std::vector<rvpt_bvh_node>& nodes;
    float traversal_cost; /* cost of one box test relative to one triangle test */

    static void set_bounds(rvpt_bvh_node& n, const Box& b)
    {
        n.bounds[0] = b.lo[0], n.bounds[1] = b.hi[0];
        n.bounds[2] = b.lo[1], n.bounds[3] = b.hi[1];
        n.bounds[4] = b.lo[2], n.bounds[5] = b.hi[2];
    }

    void build(uint32_t node_index, size_t begin, size_t end, int depth)
    {
        const size_t count = end - begin;
        Box bounds, cbounds;
        for (size_t i = begin; i < end; ++i)
        {
            bounds.grow(boxes[indices[i]]);
            cbounds.grow(&centers[3 * indices[i]]);
        }
        set_bounds(nodes[node_index], bounds);
        nodes[node_index].first_child_or_primitive = (uint32_t)begin;
        nodes[node_index].primitive_count = (uint32_t)count;
        if (count < kMinLeaf) return;

        /* binned SAH over the centroid bounds */
        float best_cost = FLT_MAX;
        int best_axis = -1, best_bin = 0;
        float best_scale = 0.0f;
        if (depth < kMaxSahDepth)
        {
            for (int axis = 0; axis < 3; ++axis)
            {
                const float extent = cbounds.hi[axis] - cbounds.lo[axis];
                if (!(extent > 0.0f)) continue;
                const float scale = (float)kBins / extent;
                Box bin_box[kBins];
                size_t bin_count[kBins] = {};
                for (size_t i = begin; i < end; ++i)
                {
                    const uint32_t p = indices[i];
                    const int b = bin_of(centers[3 * p + axis], cbounds.lo[axis], scale);
                    bin_box[b].grow(boxes[p]);
                    bin_count[b]++;
                }
                float left_cost[kBins];
                Box acc;
                size_t n = 0;
                for (int b = 0; b < kBins; ++b)
                {
                    acc.grow(bin_box[b]);
                    n += bin_count[b];
                    left_cost[b] = n ? acc.half_area() * (float)n : 0.0f;
                }
                Box racc;
                size_t rn = 0;
                for (int b = kBins - 1; b > 0; --b)
                {
                    racc.grow(bin_box[b]);
                    rn += bin_count[b];
                    const size_t ln = count - rn;
                    if (rn == 0 || ln == 0) continue; /* a split must separate something */
                    const float cost = racc.half_area() * (float)rn + left_cost[b - 1];
                    if (cost < best_cost)
                    {
                        best_cost = cost;
                        best_axis = axis;
                        best_bin = b;
                        best_scale = scale;
                    }
                }
            }
        }

        size_t mid;
        /* SAH: a leaf costs count * area triangle tests; a split costs the two
         * children's tests plus `traversal_cost` extra node tests of this box */
        const float leaf_cost = bounds.half_area() * (float)count;
        if (best_axis >= 0) best_cost += traversal_cost * bounds.half_area();
        if (best_axis < 0 || best_cost >= leaf_cost)
        {
            if (count <= kMaxLeaf) return; /* a leaf is fine */
            /* median split along the widest centroid axis */
            int axis = 0;
            for (int a = 1; a < 3; ++a)
                if (cbounds.hi[a] - cbounds.lo[a] > cbounds.hi[axis] - cbounds.lo[axis]) axis = a;
            mid = begin + (count >> 1);
            std::nth_element(indices.begin() + begin, indices.begin() + mid, indices.begin() + end,
                             [&](uint32_t a, uint32_t b) {
                                 const float ca = centers[3 * a + axis], cb = centers[3 * b + axis];
                                 return ca < cb || (ca == cb && a < b);
                             });
        }
        else
        {
            const float lo = cbounds.lo[best_axis];
            auto it = std::stable_partition(
                indices.begin() + begin, indices.begin() + end, [&](uint32_t p) {
                    return bin_of(centers[3 * p + best_axis], lo, best_scale) < best_bin;
                });
            mid = (size_t)(it - indices.begin());
        }

        const uint32_t first_child = (uint32_t)nodes.size();
        nodes.emplace_back();
        nodes.emplace_back();
        nodes[node_index].first_child_or_primitive = first_child;
        nodes[node_index].primitive_count = 0;
        build(first_child, begin, mid, depth + 1);
        build(first_child + 1, mid, end, depth + 1);
    }

    static int bin_of(float c, float lo, float scale)
    {
        const int b = (int)((c - lo) * scale);
        return std::min(kBins - 1, std::max(0, b));
    }
};

} /* namespace */

extern "C" int rvpt_b200_build_bvh(const rvpt_triangle* triangles, size_t n_triangles,
                                   rvpt_bvh_node* nodes_out, size_t* n_nodes_out,
                                   uint32_t* prim_indices_out)
{
    if (!triangles || n_triangles == 0 || !nodes_out || !n_nodes_out || !prim_indices_out)
        return RVPT_B200_EINVAL;
    if (n_triangles > 0x7FFFFFFFu) return RVPT_B200_EUNSUPPORTED;

    /* Triangle::aabb() / center(), geometry.h:98-110 */
    std::vector<Box> boxes(n_triangles);
    std::vector<float> centers(3 * n_triangles);
    for (size_t i = 0; i < n_triangles; ++i)
    {
        const rvpt_triangle& t = triangles[i];
        boxes[i].grow(t.vertex0);
        boxes[i].grow(t.vertex1);
        boxes[i].grow(t.vertex2);
        for (int a = 0; a < 3; ++a)
            centers[3 * i + a] = (t.vertex0[a] + t.vertex1[a] + t.vertex2[a]) * (1.0f / 3.0f);
    }
    std::vector<uint32_t> indices(n_triangles);
    std::iota(indices.begin(), indices.end(), 0u);
    std::vector<rvpt_bvh_node> nodes;
    nodes.reserve(2 * n_triangles);
    nodes.emplace_back();

    /* traversal cost 0 = the reference's criterion (bvh_builder.cpp:154-162): split whenever the
     * children's SAH cost beats the leaf's. Measured on B200: 0.5-1.0 moves frame time by < 5 %
     * either way (fewer node tests, more triangle tests), so the reference's choice stays. */
    Builder b{boxes, centers, indices, nodes, 0.0f};
    b.build(0, 0, n_triangles, 0);

    std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(rvpt_bvh_node));
    std::memcpy(prim_indices_out, indices.data(), n_triangles * sizeof(uint32_t));
    *n_nodes_out = nodes.size();
    return 0;
}
