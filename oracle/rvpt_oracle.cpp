/*
 * rvpt_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's per-pixel path-tracing shader
 * (assets/shaders/compute_pass.comp and its includes), written line by line
 * from the GLSL. Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this library; nothing under rvpt_b200/ does.
 *
 * PARITY PINNED to the reference's own shipped build of this path: the repo holds no tests,
 * golden images or known-answer vectors, and its Vulkan path cannot run here (no loader / ICD),
 * but it ships the shader COMPILED — assets/shaders/compute_pass.comp.spv. oracle/spirv_vm.cpp
 * executes that binary with the reference's bindings; this oracle equals its output bit for bit
 * on every case of tests/golden/make_spirv_golden.py (all configs' shapes, both image formats,
 * every integrator and camera; live in tests/test_spirv_pin.py next to the reference,
 * through committed digests elsewhere; oracle/spirv_to_cpp.py additionally compiles the same
 * binary for the host, oracle/_ref/libref_shader.so). That run found one deviation in round 2 — glslang had
 * folded normalize(vec3(0.5,1,0.3)) in double precision (RV_LIGHT_DIR_* in rvpt_math.h).
 * What stays unpinned is what SPIR-V itself leaves to the Vulkan driver: summation order of
 * dot / mat*vec, accuracy of sin / cos / tan / normalize, UNORM8 rounding (next paragraph).
 *
 * Driver-defined float behaviour (summation order of dot/cross/normalize/mix/
 * mat*vec, sin/cos/tan) is fixed by include/rvpt_math.h; GLSL min/max on NaN is
 * undefined and resolved as IEEE minNum/maxNum (fminf/fmaxf).
 *
 * Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -std=c++17 -pthread -shared -fPIC
 */
#include <atomic>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/rvpt_abi.h"
#include "../include/rvpt_math.h"

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace
{
const float INF = std::numeric_limits<float>::infinity(); /* compute_pass.comp:12 */

struct Ray /* structs.glsl:16-20 */
{
    rv_f3 origin;
    rv_f3 direction;
};

struct MaterialNew /* intersection.glsl:37-43 */
{
    int type;
    rv_f3 base_color;
    rv_f3 emissive;
    float ior;
};

struct Isect /* intersection.glsl:59-72 */
{
    float t;
    rv_f3 pos;
    rv_f3 normal;
    float u, v;
    MaterialNew mat;
};

struct Scene
{
    const rvpt_bvh_node* nodes;
    size_t n_nodes;
    const rvpt_triangle* tris;
    size_t n_tris;
    const rvpt_material* mats;
    size_t n_mats;
    bool brute_force;
};

struct Frame
{
    const rvpt_render_settings* rs;
    const float* cam; /* 16 matrix (column-major) + 4 params */
    uint32_t W, H;
    float inv_dim_x, inv_dim_y;    /* compute_pass.comp:51 */
    float current_frame;           /* :53 */
    float inv_current_frame;       /* :54 */
};

/* intersection.glsl:45-57 */
MaterialNew convert_old_material(const rvpt_material& m)
{
    MaterialNew r;
    r.type = (int)m.data[0];
    r.base_color = rv_make(m.albedo[0], m.albedo[1], m.albedo[2]);
    r.emissive = rv_make(m.emission[0], m.emission[1], m.emission[2]);
    r.ior = m.albedo[3];
    return r;
}

/* mat4 * vec4(x, y, z, w).xyz, columns summed left to right. */
rv_f3 mat_mul_xyz(const float* M, float x, float y, float z, float w)
{
    float r[3];
    for (int i = 0; i < 3; ++i)
    {
        float a = M[0 + i] * x;
        float b = M[4 + i] * y;
        float c = M[8 + i] * z;
        float d = M[12 + i] * w;
        float s = a + b;
        s = s + c;
        s = s + d;
        r[i] = s;
    }
    return rv_make(r[0], r[1], r[2]);
}

/* camera.glsl:29-51 */
Ray camera_pinhole_ray(const float* cam, float x, float y)
{
    float aspect = cam[16];
    float hfov = cam[17];
    float u = aspect * ((x + x) - 1.0f);
    float v = (y + y) - 1.0f;
    float w = 1.0f / rv_tan(0.5f * hfov);
    Ray r;
    r.origin = rv_make(cam[12], cam[13], cam[14]);
    r.direction = rv_normalize(mat_mul_xyz(cam, u, v, w, 0.0f));
    return r;
}

/* camera.glsl:55-76 */
Ray camera_ortho_ray(const float* cam, float x, float y)
{
    float scale_x = cam[18];
    float scale_y = cam[18];
    float aspect = cam[16];
    float u = aspect * ((x + x) - 1.0f);
    float v = (y + y) - 1.0f;
    Ray r;
    r.origin = mat_mul_xyz(cam, scale_x * u, scale_y * v, 0.0f, 1.0f);
    r.direction = rv_make(cam[8], cam[9], cam[10]);
    return r;
}

/* camera.glsl:80-99 with util.glsl:94-113 (unit_spherical_to_cartesian).xzy */
Ray camera_spherical_ray(const float* cam, float x, float y)
{
    float phi = x * RV_TWO_PI;
    float theta = y * RV_PI;
    float sp, cp, st, ct;
    rv_sincos(phi, &sp, &cp);
    rv_sincos(theta, &st, &ct);
    /* vec3(sin_theta * vec2(cos(phi), sin(phi)), cos(theta)).xzy */
    float lx = st * cp;
    float ly = st * sp;
    float lz = ct;
    Ray r;
    r.origin = rv_make(cam[12], cam[13], cam[14]);
    r.direction = mat_mul_xyz(cam, lx, lz, ly, 0.0f);
    return r;
}

/* compute_pass.comp:102-118 */
Ray get_camera_ray(const float* cam, int camera_idx, float u, float v)
{
    switch (camera_idx)
    {
        case 0: return camera_pinhole_ray(cam, u, v);
        case 1: return camera_ortho_ray(cam, u, v);
        default: return camera_spherical_ray(cam, u, v);
    }
}

/* intersection.glsl:267-323 */
bool intersect_triangle_fast(const Ray& ray, rv_f3 v0, rv_f3 v1, rv_f3 v2, float mint, float maxt,
                             Isect& info)
{
    rv_f3 e0 = rv_sub(v1, v0);
    rv_f3 e1 = rv_sub(v2, v0);
    rv_f3 n = rv_cross(e0, e1);

    float t = rv_dot(rv_sub(v0, ray.origin), n) / rv_dot(ray.direction, n);
    rv_f3 p = rv_add(ray.origin, rv_scale(t, ray.direction));

    rv_f3 p0 = rv_sub(p, v0);

    float bx = rv_dot(p0, e0);
    float by = rv_dot(p0, e1);

    /* mat2 A_adj = mat2(dot(e1,e1), -dot(e0,e1), -dot(e0,e1), dot(e0,e0)) — column major */
    float a00 = rv_dot(e1, e1);
    float a01 = -rv_dot(e0, e1);
    float a10 = -rv_dot(e0, e1);
    float a11 = rv_dot(e0, e0);

    float inv_det = 1.0f / (a00 * a11 - a01 * a10);

    /* uv = inv_det * (A_adj * b) ; A_adj*b = col0*b.x + col1*b.y */
    float mu = a00 * bx + a10 * by;
    float mv = a01 * bx + a11 * by;
    float u = inv_det * mu;
    float v = inv_det * mv;

    bool isect = mint < t && t < maxt && 0.0f < u && 0.0f < v && u + v < 1.0f;

    info.t = isect ? t : INF;
    info.pos = rv_add(ray.origin, rv_scale(info.t, ray.direction));
    info.normal = n;
    info.u = u;
    info.v = v;
    return isect;
}

/* intersection.glsl:327-357 — t0/t1 are declared double there, but every
 * operand is a float so the comparisons are value-identical. */
bool intersect_aabb(const Ray& ray, rv_f3 aabb_min, rv_f3 aabb_max, float mint, float maxt)
{
    rv_f3 invdir =
        rv_make(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);

    rv_f3 f = rv_mul(rv_sub(aabb_max, ray.origin), invdir);
    rv_f3 n = rv_mul(rv_sub(aabb_min, ray.origin), invdir);

    rv_f3 tmax = rv_make(fmaxf(f.x, n.x), fmaxf(f.y, n.y), fmaxf(f.z, n.z));
    rv_f3 tmin = rv_make(fminf(f.x, n.x), fminf(f.y, n.y), fminf(f.z, n.z));

    float t1 = fminf(tmax.x, fminf(tmax.y, tmax.z));
    float t0 = fmaxf(tmin.x, fmaxf(tmin.y, tmin.z));

    t0 = fmaxf(t0, mint);
    t1 = fminf(t1, maxt);

    return t1 >= t0;
}

/* intersection.glsl:361-413 */
bool intersect_bvh(const Scene& sc, const Ray& ray, float mint, float maxt, Isect& info,
                   bool* stack_overflow)
{
    uint32_t stack[64];
    int stack_ptr = 0;

    float closest_t = maxt;
    info.t = INF;
    info.pos = rv_make(0, 0, 0);
    info.normal = rv_make(0, 0, 0);

    stack[stack_ptr++] = ~0u;
    uint32_t stack_top = 0;
    while (stack_top != ~0u)
    {
        const rvpt_bvh_node& node = sc.nodes[stack_top];
        rv_f3 node_min = rv_make(node.bounds[0], node.bounds[2], node.bounds[4]);
        rv_f3 node_max = rv_make(node.bounds[1], node.bounds[3], node.bounds[5]);
        if (!intersect_aabb(ray, node_min, node_max, mint, closest_t))
        {
            stack_top = stack[--stack_ptr];
            continue;
        }

        uint32_t first_child_or_primitive = node.first_child_or_primitive;
        if (node.primitive_count > 0)
        {
            for (uint32_t i = first_child_or_primitive, n = i + node.primitive_count; i < n; ++i)
            {
                const rvpt_triangle& tri = sc.tris[i];
                rv_f3 v0 = rv_make(tri.vertex0[0], tri.vertex0[1], tri.vertex0[2]);
                rv_f3 v1 = rv_make(tri.vertex1[0], tri.vertex1[1], tri.vertex1[2]);
                rv_f3 v2 = rv_make(tri.vertex2[0], tri.vertex2[1], tri.vertex2[2]);
                Isect temp;
                if (intersect_triangle_fast(ray, v0, v1, v2, mint, closest_t, temp))
                {
                    info = temp;
                    info.mat = convert_old_material(sc.mats[(int)tri.material_id[0]]);
                    closest_t = temp.t;
                }
            }
            stack_top = stack[--stack_ptr];
        }
        else
        {
            if (stack_ptr >= 64)
            {
                *stack_overflow = true; /* undefined behaviour in the shader */
                return false;
            }
            stack[stack_ptr++] = first_child_or_primitive + 1;
            stack_top = first_child_or_primitive;
        }
    }
    return closest_t < maxt;
}

/* Nearest hit over the triangle list in upload order, strict t < closest_t so
 * the first triangle wins ties (legacy intersect_triangles semantics,
 * intersection.glsl:708-752; the BVH-independent definition of SURVEY §8c). */
bool intersect_list(const Scene& sc, const Ray& ray, float mint, float maxt, Isect& info)
{
    float closest_t = maxt;
    info.t = INF;
    info.pos = rv_make(0, 0, 0);
    info.normal = rv_make(0, 0, 0);
    for (size_t i = 0; i < sc.n_tris; ++i)
    {
        const rvpt_triangle& tri = sc.tris[i];
        rv_f3 v0 = rv_make(tri.vertex0[0], tri.vertex0[1], tri.vertex0[2]);
        rv_f3 v1 = rv_make(tri.vertex1[0], tri.vertex1[1], tri.vertex1[2]);
        rv_f3 v2 = rv_make(tri.vertex2[0], tri.vertex2[1], tri.vertex2[2]);
        Isect temp;
        if (intersect_triangle_fast(ray, v0, v1, v2, mint, closest_t, temp))
        {
            info = temp;
            info.mat = convert_old_material(sc.mats[(int)tri.material_id[0]]);
            closest_t = temp.t;
        }
    }
    return closest_t < maxt;
}

/* intersection.glsl:417-463 — early out at the first accepted triangle */
bool intersect_bvh_any(const Scene& sc, const Ray& ray, float mint, float maxt, bool* stack_overflow)
{
    uint32_t stack[64];
    int stack_ptr = 0;
    float closest_t = maxt;

    stack[stack_ptr++] = ~0u;
    uint32_t stack_top = 0;
    while (stack_top != ~0u)
    {
        const rvpt_bvh_node& node = sc.nodes[stack_top];
        rv_f3 node_min = rv_make(node.bounds[0], node.bounds[2], node.bounds[4]);
        rv_f3 node_max = rv_make(node.bounds[1], node.bounds[3], node.bounds[5]);
        if (!intersect_aabb(ray, node_min, node_max, mint, closest_t))
        {
            stack_top = stack[--stack_ptr];
            continue;
        }
        uint32_t first_child_or_primitive = node.first_child_or_primitive;
        if (node.primitive_count > 0)
        {
            for (uint32_t i = first_child_or_primitive, n = i + node.primitive_count; i < n; ++i)
            {
                const rvpt_triangle& tri = sc.tris[i];
                Isect temp;
                if (intersect_triangle_fast(ray, rv_make(tri.vertex0[0], tri.vertex0[1], tri.vertex0[2]),
                                            rv_make(tri.vertex1[0], tri.vertex1[1], tri.vertex1[2]),
                                            rv_make(tri.vertex2[0], tri.vertex2[1], tri.vertex2[2]),
                                            mint, closest_t, temp))
                    return true;
            }
            stack_top = stack[--stack_ptr];
        }
        else
        {
            if (stack_ptr >= 64)
            {
                *stack_overflow = true;
                return false;
            }
            stack[stack_ptr++] = first_child_or_primitive + 1;
            stack_top = first_child_or_primitive;
        }
    }
    return false;
}

/* intersection.glsl:467-485 */
bool intersect_scene_any(const Scene& sc, const Ray& ray, float mint, float maxt,
                         bool* stack_overflow)
{
    if (sc.brute_force)
    {
        for (size_t i = 0; i < sc.n_tris; ++i)
        {
            const rvpt_triangle& tri = sc.tris[i];
            Isect temp;
            if (intersect_triangle_fast(ray, rv_make(tri.vertex0[0], tri.vertex0[1], tri.vertex0[2]),
                                        rv_make(tri.vertex1[0], tri.vertex1[1], tri.vertex1[2]),
                                        rv_make(tri.vertex2[0], tri.vertex2[1], tri.vertex2[2]), mint,
                                        maxt, temp))
                return true;
        }
        return false;
    }
    return intersect_bvh_any(sc, ray, mint, maxt, stack_overflow);
}

/* intersection.glsl:489-517 */
bool intersect_scene(const Scene& sc, const Ray& ray, float mint, float maxt, Isect& info,
                     bool* stack_overflow)
{
    (void)maxt;
    float closest_t = INF;
    info.t = closest_t;
    info.pos = rv_make(0, 0, 0);
    info.normal = rv_make(0, 0, 0);
    Isect temp;

    bool hit = sc.brute_force ? intersect_list(sc, ray, mint, closest_t, temp)
                              : intersect_bvh(sc, ray, mint, closest_t, temp, stack_overflow);
    if (hit)
    {
        closest_t = temp.t;
        info = temp;
    }

    info.normal = closest_t < INF ? rv_normalize(info.normal) : rv_make(0, 0, 0);
    info.pos = closest_t < INF ? rv_add(ray.origin, rv_scale(info.t, ray.direction))
                               : rv_make(0, 0, 0);
    return closest_t < INF;
}

/* samples_mapping.glsl:39-60 */
rv_f3 map_uniform_sphere(float u, float v)
{
    float phi = RV_TWO_PI * u;
    float cos_theta = (1.0f - v) - v;
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float s, c;
    rv_sincos(phi, &s, &c);
    return rv_make(sin_theta * c, sin_theta * s, cos_theta);
}

/* samples_mapping.glsl:112-131 */
rv_f3 map_cosine_hemisphere_simple(float u, float v, rv_f3 n)
{
    return rv_add(n, map_uniform_sphere(u, v));
}

/* material.glsl:96-108 — first rand() is u, second is v (SPIR-V order). */
rv_f3 mat_scatter_Lambert_cos(rv_f3 normal, uint32_t* rng)
{
    float u = rv_rand(rng);
    float v = rv_rand(rng);
    return map_cosine_hemisphere_simple(u, v, normal);
}

/* material.glsl:78-92 */
rv_f3 mat_eval_Lambert_cos(rv_f3 diffuse) { return rv_scale(RV_PI, diffuse); }

/* material.glsl:207-228 */
float frensel_reflectance(float cos_in, float cos_out, float eta)
{
    float r_perp = (eta * cos_in - cos_out) / (eta * cos_in + cos_out);
    float r_parallel = (cos_in - eta * cos_out) / (cos_in + eta * cos_out);
    return 0.5f * (r_perp * r_perp + r_parallel * r_parallel);
}

struct Counters
{
    uint64_t active[RVPT_MAX_BOUNCE_STATS];
    bool stack_overflow;
};

/* integrators.glsl:547-677 */
rv_f3 integrator_Kajiya(const Scene& sc, Ray primary_ray, float mint, float maxt, int nbounce,
                        uint32_t* rng, Counters* ctr)
{
    Ray ray = primary_ray;
    Isect info;

    rv_f3 col = rv_make(0, 0, 0);
    rv_f3 throughput = rv_make(1, 1, 1);
    rv_f3 white = rv_make(1, 1, 1);
    rv_f3 blue = rv_make(0.2f, 0.3f, 0.7f);

    for (int i = 0; i < nbounce; ++i)
    {
        if (i < RVPT_MAX_BOUNCE_STATS) ctr->active[i]++;
        if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow))
        {
            float t = ray.direction.y * 0.5f + 0.5f;
            rv_f3 bg = rv_make(rv_mix(white.x, blue.x, t), rv_mix(white.y, blue.y, t),
                               rv_mix(white.z, blue.z, t));
            return rv_add(col, rv_mul(throughput, bg));
        }

        col = rv_add(col, rv_mul(throughput, info.mat.emissive));

        rv_f3 pos = info.pos;
        rv_f3 normal = info.normal;
        rv_f3 dir_in = rv_normalize(ray.direction);
        rv_f3 pos_out;
        rv_f3 dir_out;

        float cos_view = rv_dot(dir_in, normal);
        float cos_in;
        bool flipped_normal = cos_view > 0.0f;
        float eta = info.mat.ior;
        if (flipped_normal)
        {
            cos_in = cos_view;
            normal = rv_neg(normal);
        }
        else
        {
            cos_in = -cos_view;
            eta = 1.0f / eta;
        }

        switch (info.mat.type)
        {
            case 0: /* Lambert */
                pos_out = rv_add(pos, rv_scale(RV_EPSILON, normal));
                dir_out = mat_scatter_Lambert_cos(normal, rng);
                throughput =
                    rv_mul(throughput, mat_eval_Lambert_cos(rv_scale(RV_INV_PI, info.mat.base_color)));
                break;

            case 1: /* perfect mirror */
                pos_out = rv_add(pos, rv_scale(RV_EPSILON, normal));
                dir_out = rv_add(dir_in, rv_scale(cos_in + cos_in, normal));
                throughput = rv_mul(throughput, info.mat.base_color);
                break;

            case 2: /* dielectric */
            {
                float cos_out_sqr = 1.0f - (eta * eta) * (1.0f - cos_in * cos_in);
                float cos_out = 0.0f, f_refl;

                bool refl = (cos_out_sqr <= 0.0f);
                if (!refl)
                {
                    cos_out = sqrtf(fmaxf(0.0f, cos_out_sqr));
                    f_refl = frensel_reflectance(cos_in, cos_out, eta);
                    refl = (rv_rand(rng) < f_refl);
                }

                if (refl)
                {
                    pos_out = rv_add(pos, rv_scale(RV_EPSILON, normal));
                    dir_out = rv_add(dir_in, rv_scale(cos_in + cos_in, normal));
                }
                else
                {
                    pos_out = rv_sub(pos, rv_scale(RV_EPSILON, normal));
                    dir_out = rv_add(rv_scale(eta, dir_in),
                                     rv_scale(eta * cos_in - cos_out, normal));
                }
                throughput = rv_mul(throughput, info.mat.base_color);
                break;
            }
            default: return rv_make(0, 0, 0);
        }

        ray.origin = pos_out;
        ray.direction = dir_out;
    }

    return rv_make(0, 0, 0);
}

rv_f3 splat(float v) { return rv_make(v, v, v); }

/* the directional light of the Utah / Appel / Whitted models: normalize(vec3(0.5,1,0.3)) as the
 * shipped SPIR-V holds it (constant-folded by glslang; see RV_LIGHT_DIR_* in rvpt_math.h) */
rv_f3 light_direction() { return rv_make(RV_LIGHT_DIR_X, RV_LIGHT_DIR_Y, RV_LIGHT_DIR_Z); }

/* integrators.glsl:24-38 */
rv_f3 integrator_binary(const Scene& sc, const Ray& ray, float mint, float maxt, Counters* ctr)
{
    return splat(intersect_scene_any(sc, ray, mint, maxt, &ctr->stack_overflow) ? 1.0f : 0.0f);
}

/* integrators.glsl:42-59 */
rv_f3 integrator_color(const Scene& sc, const Ray& ray, float mint, float maxt, Counters* ctr)
{
    Isect info;
    if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow)) return splat(0.0f);
    return info.mat.base_color;
}

/* integrators.glsl:63-82 */
rv_f3 integrator_depth(const Scene& sc, const Ray& ray, float mint, float maxt, Counters* ctr)
{
    Isect info;
    intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow);
    float len = sqrtf(rv_dot(ray.direction, ray.direction));
    float inv_dist = 1.0f / (len * info.t);
    return splat(inv_dist);
}

/* integrators.glsl:86-102 */
rv_f3 integrator_normal(const Scene& sc, const Ray& ray, float mint, float maxt, Counters* ctr)
{
    Isect info;
    float isect = intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow) ? 1.0f : 0.0f;
    float h = 0.5f * isect;
    return rv_make(0.5f * info.normal.x + h, 0.5f * info.normal.y + h, 0.5f * info.normal.z + h);
}

/* integrators.glsl:106-148 */
rv_f3 integrator_Utah(const Scene& sc, const Ray& ray, float mint, float maxt, Counters* ctr)
{
    rv_f3 light_intensity = splat(1.0f);
    rv_f3 light_dir = light_direction();
    rv_f3 ambient = splat(0.1f);
    Isect info;
    rv_f3 col = ambient;
    if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow))
    {
        float t = ray.direction.y;
        return rv_make(rv_mix(1.0f, 0.2f, t), rv_mix(1.0f, 0.3f, t), rv_mix(1.0f, 0.7f, t));
    }
    col = rv_add(col, info.mat.emissive);
    rv_f3 normal = info.normal;
    normal = rv_dot(ray.direction, normal) < 0.0f ? normal : rv_neg(normal);
    float cos_light = fmaxf(0.0f, rv_dot(light_dir, normal));
    return rv_add(col, rv_scale(cos_light, rv_mul(info.mat.base_color, light_intensity)));
}

/* integrators.glsl:152-200 */
rv_f3 integrator_ao(const Scene& sc, const Ray& ray, float mint, float maxt, int nrays,
                    uint32_t* rng, Counters* ctr)
{
    Isect info;
    bool isect = intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow);
    if (!isect) return splat(0.0f);
    rv_f3 normal = info.normal;
    normal = rv_dot(ray.direction, normal) < 0.0f ? normal : rv_neg(normal);
    float acc = 0.0f;
    for (int i = 0; i < nrays; ++i)
    {
        Ray new_ray;
        new_ray.origin = rv_add(info.pos, rv_scale(RV_EPSILON, normal));
        float u = rv_rand(rng);
        float v = rv_rand(rng);
        new_ray.direction = map_cosine_hemisphere_simple(u, v, normal);
        acc += intersect_scene_any(sc, new_ray, mint, maxt, &ctr->stack_overflow) ? 1.0f : 0.0f;
    }
    return splat(1.0f - acc / (float)nrays);
}

/* integrators.glsl:204-250 */
rv_f3 integrator_Appel(const Scene& sc, const Ray& ray, float mint, float maxt, Counters* ctr)
{
    rv_f3 light_dir = light_direction();
    Isect info;
    if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow)) return splat(1.0f);
    rv_f3 dir_in = rv_normalize(ray.direction);
    float cos_view = rv_dot(dir_in, info.normal);
    rv_f3 normal = cos_view > 0.0f ? rv_neg(info.normal) : info.normal;
    Ray shadow_ray;
    shadow_ray.origin = rv_add(info.pos, rv_scale(RV_EPSILON, normal));
    shadow_ray.direction = light_dir;
    if (intersect_scene_any(sc, shadow_ray, 0.0f, INF, &ctr->stack_overflow)) return splat(0.0f);
    float cos_light = fmaxf(0.0f, rv_dot(light_dir, normal));
    return splat(1.0f * cos_light);
}

/* Shared by Whitted and Cook: integrators.glsl:290-316 / :439-465 (normal flip, eta)
 * and the mirror / dielectric branches :338-377 / :487-527, identical to Kajiya's. */
struct Surface
{
    rv_f3 pos, normal, dir_in;
    float cos_in, eta;
};

Surface surface_of(const Ray& ray, const Isect& info)
{
    Surface s;
    s.pos = info.pos;
    s.normal = info.normal;
    s.dir_in = rv_normalize(ray.direction);
    float cos_view = rv_dot(s.dir_in, s.normal);
    s.eta = info.mat.ior;
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

/* returns false for an unknown material type */
bool specular_bounce(const Surface& s, const MaterialNew& mat, uint32_t* rng, Ray* ray,
                     rv_f3* throughput)
{
    rv_f3 pos_out, dir_out;
    if (mat.type == 1)
    {
        pos_out = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
        dir_out = rv_add(s.dir_in, rv_scale(s.cos_in + s.cos_in, s.normal));
    }
    else if (mat.type == 2)
    {
        float cos_out_sqr = 1.0f - (s.eta * s.eta) * (1.0f - s.cos_in * s.cos_in);
        float cos_out = 0.0f;
        bool refl = (cos_out_sqr <= 0.0f);
        if (!refl)
        {
            cos_out = sqrtf(fmaxf(0.0f, cos_out_sqr));
            float f_refl = frensel_reflectance(s.cos_in, cos_out, s.eta);
            refl = (rv_rand(rng) < f_refl);
        }
        if (refl)
        {
            pos_out = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
            dir_out = rv_add(s.dir_in, rv_scale(s.cos_in + s.cos_in, s.normal));
        }
        else
        {
            pos_out = rv_sub(s.pos, rv_scale(RV_EPSILON, s.normal));
            dir_out = rv_add(rv_scale(s.eta, s.dir_in),
                             rv_scale(s.eta * s.cos_in - cos_out, s.normal));
        }
    }
    else
        return false;
    *throughput = rv_mul(*throughput, mat.base_color);
    ray->origin = pos_out;
    ray->direction = dir_out;
    return true;
}

rv_f3 sky_no_remap(const Ray& ray) /* mix(white, blue, ray.direction.y), :284 / :433 */
{
    float t = ray.direction.y;
    return rv_make(rv_mix(1.0f, 0.2f, t), rv_mix(1.0f, 0.3f, t), rv_mix(1.0f, 0.7f, t));
}

/* integrators.glsl:254-403 */
rv_f3 integrator_Whitted(const Scene& sc, Ray ray, float mint, float maxt, int nbounce,
                         uint32_t* rng, Counters* ctr)
{
    rv_f3 light_intensity = splat(1.0f);
    rv_f3 light_dir = light_direction();
    rv_f3 col = splat(0.1f); /* ambient */
    rv_f3 throughput = splat(1.0f);
    Isect info;
    for (int i = 0; i < nbounce; ++i)
    {
        if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow))
            return rv_add(col, rv_mul(throughput, sky_no_remap(ray)));
        col = rv_add(col, rv_mul(throughput, info.mat.emissive));
        Surface s = surface_of(ray, info);
        if (info.mat.type == 0)
        {
            Ray shadow_ray;
            shadow_ray.origin = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
            shadow_ray.direction = light_dir;
            if (intersect_scene_any(sc, shadow_ray, 0.0f, INF, &ctr->stack_overflow)) return col;
            float cos_light = fmaxf(0.0f, rv_dot(light_dir, s.normal));
            rv_f3 lit = rv_scale(cos_light, rv_mul(rv_mul(throughput, info.mat.base_color), light_intensity));
            return rv_add(col, lit);
        }
        if (!specular_bounce(s, info.mat, rng, &ray, &throughput)) return splat(0.0f);
    }
    return splat(0.0f);
}

/* integrators.glsl:407-543 */
rv_f3 integrator_Cook(const Scene& sc, Ray ray, float mint, float maxt, int nbounce, uint32_t* rng,
                      Counters* ctr)
{
    rv_f3 col = splat(0.0f);
    rv_f3 throughput = splat(1.0f);
    Isect info;
    for (int i = 0; i < nbounce; ++i)
    {
        if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow))
            return rv_add(col, rv_mul(throughput, sky_no_remap(ray)));
        col = rv_add(col, rv_mul(throughput, info.mat.emissive));
        Surface s = surface_of(ray, info);
        if (info.mat.type == 0)
        {
            ray.origin = rv_add(s.pos, rv_scale(RV_EPSILON, s.normal));
            ray.direction = mat_scatter_Lambert_cos(s.normal, rng);
            throughput = rv_mul(throughput,
                                mat_eval_Lambert_cos(rv_scale(RV_INV_PI, info.mat.base_color)));
            if (!intersect_scene(sc, ray, mint, maxt, info, &ctr->stack_overflow))
                return rv_add(col, rv_mul(throughput, sky_no_remap(ray)));
            return rv_add(col, rv_mul(throughput, info.mat.emissive));
        }
        if (!specular_bounce(s, info.mat, rng, &ray, &throughput)) return splat(0.0f);
    }
    return splat(0.0f);
}

/* distance_functions.glsl:36-61: distance from p to the triangle (a, b, c). sign() is
 * GLSL.std.450 FSign (+1 / -1 / 0; NaN -> 0 as in oracle/spirv_vm.cpp), clamp() is FClamp =
 * min(max(x, 0), 1), both arms of the ?: are pure, so which of them the binary evaluates
 * does not matter. */
float sign_of(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
float dot2(rv_f3 v) { return rv_dot(v, v); }

float distance_triangle(rv_f3 p, rv_f3 a, rv_f3 b, rv_f3 c)
{
    const rv_f3 ba = rv_sub(b, a), pa = rv_sub(p, a);
    const rv_f3 cb = rv_sub(c, b), pb = rv_sub(p, b);
    const rv_f3 ac = rv_sub(a, c), pc = rv_sub(p, c);
    const rv_f3 nor = rv_cross(ba, ac);
    const float s0 = sign_of(rv_dot(rv_cross(ba, nor), pa));
    const float s1 = sign_of(rv_dot(rv_cross(cb, nor), pb));
    const float s2 = sign_of(rv_dot(rv_cross(ac, nor), pc));
    float v;
    if ((s0 + s1) + s2 < 2.0f)
    {
        const float e0 = dot2(rv_sub(rv_scale(clamp01(rv_dot(ba, pa) / dot2(ba)), ba), pa));
        const float e1 = dot2(rv_sub(rv_scale(clamp01(rv_dot(cb, pb) / dot2(cb)), cb), pb));
        const float e2 = dot2(rv_sub(rv_scale(clamp01(rv_dot(ac, pc) / dot2(ac)), ac), pc));
        v = fminf(fminf(e0, e1), e2);
    }
    else
    {
        const float h = rv_dot(nor, pa);
        v = (h * h) / dot2(nor);
    }
    return sqrtf(v);
}

/* distance_functions.glsl:70-116, sphere tracing against EVERY triangle in buffer order.
 * MARCH_ITER 32, MARCH_EPS 0.1 (compute_pass.comp:10-11). Returns the iteration count, the
 * only field integrator_Hart looks at. */
int intersect_scene_st_iter(const Scene& sc, const Ray& ray, float mint, float maxt)
{
    float t = mint;
    rv_f3 p = rv_add(ray.origin, rv_scale(t, ray.direction));
    int i;
    for (i = 0; i < 32; ++i)
    {
        float t_radius = INF; /* t_radius_idx.x; the index lane never reaches the output */
        for (size_t j = 0; j < sc.n_tris; ++j)
        {
            const rvpt_triangle& tri = sc.tris[j];
            const float dist = distance_triangle(p, rv_make(tri.vertex0[0], tri.vertex0[1], tri.vertex0[2]),
                                                 rv_make(tri.vertex1[0], tri.vertex1[1], tri.vertex1[2]),
                                                 rv_make(tri.vertex2[0], tri.vertex2[1], tri.vertex2[2]));
            t_radius = t_radius < dist ? t_radius : dist; /* min_idx :63-66: lhs.x < rhs.x ? lhs : rhs */
        }
        const float min_radius = fminf(INF, t_radius); /* min(s_radius_idx.x, t_radius_idx.x), s = INF */
        if (min_radius < 0.1f || min_radius > maxt) return i;
        t = t + min_radius;
        p = rv_add(p, rv_scale(min_radius, ray.direction));
    }
    return i;
}

/* integrators.glsl:681-693: the sphere tracer's iteration count as a grey level, iter / 31
 * (32 / 31 when the march ran out of iterations). */
rv_f3 integrator_Hart(const Scene& sc, const Ray& ray, float mint, float maxt)
{
    return splat((float)intersect_scene_st_iter(sc, ray, mint, maxt) / 31.0f);
}

/* compute_pass.comp:68-99: every index outside 0..9 (negative ones too) is integrator_Hart. */
rv_f3 eval_integrator(const Scene& sc, int idx, const Ray& ray, int max_bounces, uint32_t* rng,
                      Counters* ctr, bool* ok)
{
    switch (idx)
    {
        case 0: return integrator_binary(sc, ray, 0.0f, INF, ctr);
        case 1: return integrator_color(sc, ray, 0.0f, INF, ctr);
        case 2: return integrator_depth(sc, ray, 0.0f, INF, ctr);
        case 3: return integrator_normal(sc, ray, 0.0f, INF, ctr);
        case 4: return integrator_Utah(sc, ray, 0.0f, INF, ctr);
        case 5: return integrator_ao(sc, ray, 0.0f, INF, max_bounces, rng, ctr);
        case 6: return integrator_Appel(sc, ray, 0.0f, INF, ctr);
        case 7: return integrator_Whitted(sc, ray, 0.0f, INF, max_bounces, rng, ctr);
        case 8: return integrator_Cook(sc, ray, 0.0f, INF, max_bounces, rng, ctr);
        case 9: return integrator_Kajiya(sc, ray, 0.0f, INF, max_bounces, rng, ctr);
        default: (void)ok; return integrator_Hart(sc, ray, 0.0f, INF);
    }
}

/* compute_pass.comp:121-167 for one pixel. (The bool result is a leftover of the time when
 * integrator_Hart was not restated: every integrator index is now.) */
bool shade_pixel(const Scene& sc, const Frame& fr, uint32_t x, uint32_t y, rv_f3 prev_in,
                 rv_f3* out, Counters* ctr)
{
    const rvpt_render_settings& rs = *fr.rs;

    /* util.glsl:35-36 */
    uint32_t p_idx = x + y * fr.W;
    uint32_t rng_state = rv_wang_hash(p_idx) + rs.current_frame;

    /* compute_pass.comp:134-144 */
    int integrator_idx = rs.top_left_render_mode;
    float split_x = (float)x * fr.inv_dim_x;
    float split_y = (float)y * fr.inv_dim_y;
    if (split_y > rs.split_ratio[1])
    {
        if (split_x < rs.split_ratio[0])
            integrator_idx = rs.bottom_left_render_mode;
        else
            integrator_idx = rs.bottom_right_render_mode;
    }
    else if (split_x > rs.split_ratio[0])
        integrator_idx = rs.top_right_render_mode;

    /* :146-148 */
    float keep = (float)(rs.current_frame < 1u ? rs.current_frame : 1u);
    rv_f3 temporal = rv_scale(keep, prev_in);

    rv_f3 sampled = rv_make(0, 0, 0);
    for (int i = 0; i < rs.aa; i++)
    {
        float jx = rv_rand(&rng_state);
        float jy = rv_rand(&rng_state);
        float cx = ((float)x + jx) * fr.inv_dim_x;
        float cy = ((float)y + jy) * fr.inv_dim_y;
        cy = 1.0f - cy;

        Ray ray = get_camera_ray(fr.cam, rs.camera_mode, cx, cy);
        bool ok = true;
        sampled = rv_add(sampled, eval_integrator(sc, integrator_idx, ray, rs.max_bounces,
                                                  &rng_state, ctr, &ok));
        if (!ok) return false;
    }

    float aa_f = (float)rs.aa;
    sampled = rv_make(sampled.x / aa_f, sampled.y / aa_f, sampled.z / aa_f);
    /* :162-163 (temporal * current_frame + sampled) * inv_current_frame */
    rv_f3 acc = rv_make((temporal.x * fr.current_frame + sampled.x) * fr.inv_current_frame,
                  (temporal.y * fr.current_frame + sampled.y) * fr.inv_current_frame,
                  (temporal.z * fr.current_frame + sampled.z) * fr.inv_current_frame);
    *out = acc;
    return true;
}

} /* namespace */

/* ------------------------------------------------------------------------ */
/* C entry points (ctypes)                                                   */
/* ------------------------------------------------------------------------ */

/* flags: same bits as RVPT_B200_FLAG_* (accum rgba8, reference dispatch,
 * brute force). Renders rows [y_begin, y_end) of one frame.
 * Float mode: accum (W*H*4 floats) is read as the previous running mean and
 * overwritten. RGBA8 mode: temporal (W*H*4 bytes) likewise. result (W*H*4
 * bytes, may be NULL) receives the UNORM8 store of the new value.
 * active[64] is incremented per traced segment. Returns 0, -4 for an
 * unsupported integrator, -6 for a BVH stack overflow. */
ORACLE_API int rvpt_oracle_render_rows(const rvpt_bvh_node* nodes, size_t n_nodes,
                                       const rvpt_triangle* tris, size_t n_tris,
                                       const rvpt_material* mats, size_t n_mats,
                                       const rvpt_render_settings* settings, const float* camera,
                                       uint32_t width, uint32_t height, uint32_t flags,
                                       uint32_t y_begin, uint32_t y_end, float* accum,
                                       uint8_t* temporal, uint8_t* result, uint64_t* active,
                                       int nthreads)
{
    Scene sc{nodes, n_nodes, tris, n_tris, mats, n_mats,
             (flags & RVPT_B200_FLAG_BRUTE_FORCE) != 0 || nodes == nullptr};
    Frame fr;
    fr.rs = settings;
    fr.cam = camera;
    fr.W = width;
    fr.H = height;
    fr.inv_dim_x = 1.0f / (float)width;
    fr.inv_dim_y = 1.0f / (float)height;
    fr.current_frame = (float)settings->current_frame;
    fr.inv_current_frame = 1.0f / (float)(settings->current_frame + 1u);

    const bool rgba8_mode = (flags & RVPT_B200_FLAG_ACCUM_RGBA8) != 0;
    uint32_t x_end = width, y_cap = height;
    if (flags & RVPT_B200_FLAG_REFERENCE_DISPATCH)
    {
        x_end = (width / 16u) * 16u; /* rvpt.cpp:1035-1036 */
        y_cap = (height / 16u) * 16u;
    }
    if (y_end > y_cap) y_end = y_cap;
    if (y_begin > y_end) y_begin = y_end;

    if (nthreads < 1) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;

    std::atomic<uint32_t> next_band{0};
    const uint32_t band = 4;
    const uint32_t n_bands = (y_end - y_begin + band - 1) / band;
    std::atomic<int> status{0};
    std::vector<Counters> ctrs((size_t)nthreads);
    for (auto& c : ctrs) std::memset(&c, 0, sizeof(c));

    auto worker = [&](int tid) {
        Counters& ctr = ctrs[(size_t)tid];
        for (;;)
        {
            uint32_t b = next_band.fetch_add(1);
            if (b >= n_bands) break;
            uint32_t y0 = y_begin + b * band;
            uint32_t y1 = y0 + band < y_end ? y0 + band : y_end;
            for (uint32_t y = y0; y < y1; ++y)
                for (uint32_t x = 0; x < x_end; ++x)
                {
                    size_t p = (size_t)y * width + x;
                    rv_f3 prev;
                    if (rgba8_mode)
                        prev = rv_make(rv_unorm8_load(temporal[4 * p + 0]),
                                       rv_unorm8_load(temporal[4 * p + 1]),
                                       rv_unorm8_load(temporal[4 * p + 2]));
                    else
                        prev = rv_make(accum[4 * p + 0], accum[4 * p + 1], accum[4 * p + 2]);
                    rv_f3 out;
                    if (!shade_pixel(sc, fr, x, y, prev, &out, &ctr))
                    {
                        status.store(RVPT_B200_EUNSUPPORTED);
                        return;
                    }
                    uint8_t q[4] = {(uint8_t)rv_unorm8_store(out.x), (uint8_t)rv_unorm8_store(out.y),
                                    (uint8_t)rv_unorm8_store(out.z), 0};
                    if (rgba8_mode)
                        std::memcpy(temporal + 4 * p, q, 4);
                    else
                    {
                        accum[4 * p + 0] = out.x;
                        accum[4 * p + 1] = out.y;
                        accum[4 * p + 2] = out.z;
                        accum[4 * p + 3] = 0.0f;
                    }
                    if (result) std::memcpy(result + 4 * p, q, 4);
                }
        }
    };

    if (nthreads == 1)
        worker(0);
    else
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
        for (auto& th : pool) th.join();
    }

    bool overflow = false;
    for (auto& c : ctrs)
    {
        overflow |= c.stack_overflow;
        if (active)
            for (int i = 0; i < RVPT_MAX_BOUNCE_STATS; ++i) active[i] += c.active[i];
    }
    if (status.load() != 0) return status.load();
    return overflow ? -6 : 0;
}

/* ---- function-level probes for unit tests ------------------------------- */

ORACLE_API uint32_t rvpt_oracle_wang_hash(uint32_t seed) { return rv_wang_hash(seed); }

/* Seeds like util.glsl:35-36 and writes n xorshift states + n rand() floats. */
ORACLE_API void rvpt_oracle_rand_stream(uint32_t x, uint32_t y, uint32_t width, uint32_t frame,
                                        uint32_t n, uint32_t* states, float* values)
{
    uint32_t s = rv_wang_hash(x + y * width) + frame;
    for (uint32_t i = 0; i < n; ++i)
    {
        values[i] = rv_rand(&s);
        states[i] = s;
    }
}

ORACLE_API void rvpt_oracle_sincos(const float* x, size_t n, float* s, float* c)
{
    for (size_t i = 0; i < n; ++i) rv_sincos(x[i], &s[i], &c[i]);
}

ORACLE_API void rvpt_oracle_normalize(const float* v, size_t n, float* out)
{
    for (size_t i = 0; i < n; ++i)
    {
        rv_f3 r = rv_normalize(rv_make(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
        out[3 * i] = r.x;
        out[3 * i + 1] = r.y;
        out[3 * i + 2] = r.z;
    }
}

/* out6 = origin.xyz, direction.xyz */
ORACLE_API void rvpt_oracle_camera_ray(const float* camera, int mode, float u, float v, float* out6)
{
    Ray r = get_camera_ray(camera, mode, u, v);
    out6[0] = r.origin.x;
    out6[1] = r.origin.y;
    out6[2] = r.origin.z;
    out6[3] = r.direction.x;
    out6[4] = r.direction.y;
    out6[5] = r.direction.z;
}

/* ray6 = origin, direction; tri9 = v0, v1, v2. out6 = t, n.xyz, u, v. */
ORACLE_API int rvpt_oracle_intersect_triangle(const float* ray6, const float* tri9, float mint,
                                              float maxt, float* out6)
{
    Ray r{rv_make(ray6[0], ray6[1], ray6[2]), rv_make(ray6[3], ray6[4], ray6[5])};
    Isect info;
    bool hit = intersect_triangle_fast(r, rv_make(tri9[0], tri9[1], tri9[2]),
                                       rv_make(tri9[3], tri9[4], tri9[5]),
                                       rv_make(tri9[6], tri9[7], tri9[8]), mint, maxt, info);
    out6[0] = info.t;
    out6[1] = info.normal.x;
    out6[2] = info.normal.y;
    out6[3] = info.normal.z;
    out6[4] = info.u;
    out6[5] = info.v;
    return hit ? 1 : 0;
}

ORACLE_API int rvpt_oracle_intersect_aabb(const float* ray6, const float* min3, const float* max3,
                                          float mint, float maxt)
{
    Ray r{rv_make(ray6[0], ray6[1], ray6[2]), rv_make(ray6[3], ray6[4], ray6[5])};
    return intersect_aabb(r, rv_make(min3[0], min3[1], min3[2]),
                          rv_make(max3[0], max3[1], max3[2]), mint, maxt)
               ? 1
               : 0;
}

/* Nearest hit for a batch of rays. out per ray: t (inf on miss), normal.xyz
 * (normalised), material index as float. Returns the number of hits. */
ORACLE_API int64_t rvpt_oracle_intersect_scene(const rvpt_bvh_node* nodes, size_t n_nodes,
                                               const rvpt_triangle* tris, size_t n_tris,
                                               const rvpt_material* mats, size_t n_mats,
                                               int brute_force, const float* rays6, size_t n_rays,
                                               float* out5)
{
    Scene sc{nodes, n_nodes, tris, n_tris, mats, n_mats, brute_force != 0 || nodes == nullptr};
    int64_t hits = 0;
    for (size_t i = 0; i < n_rays; ++i)
    {
        Ray r{rv_make(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]),
              rv_make(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5])};
        Isect info;
        bool overflow = false;
        bool hit = intersect_scene(sc, r, 0.0f, INF, info, &overflow);
        if (overflow) return -6;
        out5[5 * i + 0] = info.t;
        out5[5 * i + 1] = info.normal.x;
        out5[5 * i + 2] = info.normal.y;
        out5[5 * i + 3] = info.normal.z;
        out5[5 * i + 4] = hit ? (float)info.mat.type : -1.0f;
        hits += hit ? 1 : 0;
    }
    return hits;
}

ORACLE_API float rvpt_oracle_fresnel(float cos_in, float cos_out, float eta)
{
    return frensel_reflectance(cos_in, cos_out, eta);
}

ORACLE_API int rvpt_oracle_contract_probe(void)
{
    volatile float a = RV_PROBE_A, b = RV_PROBE_B, c = RV_PROBE_C;
    return rv_contract_probe(a, b, c);
}

ORACLE_API int rvpt_oracle_hardware_threads(void)
{
    return (int)std::thread::hardware_concurrency();
}
