/*
 * ref_layout.cpp — compiles the reference's own host headers, UNMODIFIED and where they lie
 * (/root/reference/src/rvpt/{geometry,material,bvh,bvh_builder}.h, bvh.cpp, bvh_builder.cpp),
 * against the glm stand-in of oracle/ref_shim and exports what the C ABI must agree with:
 * sizeof / offsetof of Triangle, Material, BvhNode (geometry.h:76-111, material.h:9-26,
 * bvh.h:12-58), the bytes the reference's constructors produce, Bvh::permute_primitives, and
 * the reference's BinnedBvhBuilder (bvh_builder.cpp:11-199; it asserts / crashes on many inputs,
 * SURVEY.md 2.2 — callers run it in a subprocess).
 * Built by oracle/Makefile into oracle/_ref/ (git-ignored); TEST INFRASTRUCTURE.
 */
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "bvh.h"
#include "bvh_builder.h"
#include "geometry.h"
#include "material.h"

#define EXPORT extern "C" __attribute__((visibility("default")))

/* out[] = sizeof(Triangle), offsets of vertex0, vertex1, vertex2, material_id,
 *         sizeof(Material), offsets of albedo, emission, data,
 *         sizeof(BvhNode), offsets of first_child_or_primitive, primitive_count, bounds */
EXPORT int ref_layout(uint32_t* out, int capacity)
{
    const uint32_t v[] = {
        (uint32_t)sizeof(Triangle), (uint32_t)offsetof(Triangle, vertex0), (uint32_t)offsetof(Triangle, vertex1),
        (uint32_t)offsetof(Triangle, vertex2), (uint32_t)offsetof(Triangle, material_id),
        (uint32_t)sizeof(Material), (uint32_t)offsetof(Material, albedo), (uint32_t)offsetof(Material, emission),
        (uint32_t)offsetof(Material, data),
        (uint32_t)sizeof(BvhNode), (uint32_t)offsetof(BvhNode, first_child_or_primitive),
        (uint32_t)offsetof(BvhNode, primitive_count), (uint32_t)offsetof(BvhNode, bounds)};
    const int n = (int)(sizeof(v) / sizeof(v[0]));
    if (capacity < n) return -1;
    std::memcpy(out, v, sizeof(v));
    return n;
}

/* Triangle(v0, v1, v2, material_id) -> 64 bytes */
EXPORT void ref_make_triangle(const float v0[3], const float v1[3], const float v2[3], int material_id, void* out64)
{
    const Triangle t(glm::vec3(v0[0], v0[1], v0[2]), glm::vec3(v1[0], v1[1], v1[2]), glm::vec3(v2[0], v2[1], v2[2]),
                     material_id);
    std::memcpy(out64, &t, sizeof(t));
}

/* Material(albedo, emission, type) -> 48 bytes */
EXPORT void ref_make_material(const float albedo[4], const float emission[4], int type, void* out48)
{
    const Material m(glm::vec4(albedo[0], albedo[1], albedo[2], albedo[3]),
                     glm::vec4(emission[0], emission[1], emission[2], emission[3]), (Material::Type)type);
    std::memcpy(out48, &m, sizeof(m));
}

/* BvhNode::AABBProxy::operator= : (min, max) -> bounds[6] = minx,maxx,miny,maxy,minz,maxz (bvh.h:40-45) */
EXPORT void ref_node_bounds(const float mn[3], const float mx[3], void* out32)
{
    BvhNode n{};
    n.aabb() = AABB(glm::vec3(mn[0], mn[1], mn[2]), glm::vec3(mx[0], mx[1], mx[2]));
    std::memcpy(out32, &n, sizeof(n));
}

/* BinnedBvhBuilder().build_bvh(triangles) + Bvh::permute_primitives (rvpt.cpp:84-86).
 * nodes_out: capacity 2n, sorted_out: n triangles. Returns the node count. May assert()/crash. */
EXPORT int ref_build_bvh(const void* triangles, size_t n, void* nodes_out, void* sorted_out, uint32_t* indices_out)
{
    std::vector<Triangle> tris(n);
    std::memcpy(tris.data(), triangles, n * sizeof(Triangle));
    BinnedBvhBuilder builder;
    const Bvh bvh = builder.build_bvh(tris);
    const std::vector<Triangle> sorted = bvh.permute_primitives(tris);
    std::memcpy(nodes_out, bvh.nodes.data(), bvh.nodes.size() * sizeof(BvhNode));
    std::memcpy(sorted_out, sorted.data(), n * sizeof(Triangle));
    std::memcpy(indices_out, bvh.primitive_indices.data(), n * sizeof(uint32_t));
    return (int)bvh.nodes.size();
}
