/*
 * rvpt_math.h — the arithmetic contract shared by the sm_100a kernels and the
 * CPU oracle.
 *
 * GLSL leaves the rounding / summation order of dot, cross, normalize, mix,
 * matrix*vector and sin/cos/tan to the driver (SURVEY.md Appendix A, items
 * marked †). This header fixes ONE order for each of them: left-to-right,
 * every operation separately rounded IEEE binary32 (no FMA contraction),
 * correctly rounded sqrt/div, and one polynomial sincos. Compile device code
 * with `-fmad=false` and host code with `-ffp-contract=off`; rv_contract_probe()
 * lets tests prove both were honoured.
 *
 * Only primitives live here. The algorithms (triangle test, AABB test, BVH
 * walk, integrator, camera) are written twice on purpose: straight from the
 * GLSL in oracle/rvpt_oracle.cpp, and restructured for the GPU in
 * rvpt_b200/csrc/kernels.cu.
 */
#ifndef RVPT_MATH_H
#define RVPT_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#else
#define RV_HD inline
#endif

/* compute_pass.comp:5-12 — un-suffixed GLSL literals are float32. */
#define RV_PI 3.14159274101257324f      /* 0x40490FDB */
#define RV_TWO_PI 6.28318548202514648f  /* 0x40C90FDB */
#define RV_INV_PI 0.318309873342514038f /* 0x3EA2F983 */
#define RV_EPSILON 0.005f               /* 0x3BA3D70A */
/* normalize(vec3(0.5, 1, 0.3)), the directional light of the Utah / Appel / Whitted integrators
 * (integrators.glsl:121,218,274): glslang folded it at compile time in double precision, so the
 * shipped compute_pass.comp.spv holds these three constants (%1585-%1587) — z is one ulp below
 * what a float32 normalize returns (0x3E84B0B1). Found by running the binary (oracle/spirv_vm.cpp). */
#define RV_LIGHT_DIR_X 0.431934207677841187f /* 0x3EDD267B */
#define RV_LIGHT_DIR_Y 0.863868415355682373f /* 0x3F5D267B */
#define RV_LIGHT_DIR_Z 0.259160518646240234f /* 0x3E84B0B0 */

struct rv_f3
{
    float x, y, z;
};

RV_HD rv_f3 rv_make(float x, float y, float z)
{
    rv_f3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
RV_HD rv_f3 rv_add(rv_f3 a, rv_f3 b) { return rv_make(a.x + b.x, a.y + b.y, a.z + b.z); }
RV_HD rv_f3 rv_sub(rv_f3 a, rv_f3 b) { return rv_make(a.x - b.x, a.y - b.y, a.z - b.z); }
RV_HD rv_f3 rv_mul(rv_f3 a, rv_f3 b) { return rv_make(a.x * b.x, a.y * b.y, a.z * b.z); }
RV_HD rv_f3 rv_scale(float s, rv_f3 a) { return rv_make(s * a.x, s * a.y, s * a.z); }
RV_HD rv_f3 rv_neg(rv_f3 a) { return rv_make(-a.x, -a.y, -a.z); }

/* dot: ((x*x' + y*y') + z*z'), three roundings for the products, two for the sums. */
RV_HD float rv_dot(rv_f3 a, rv_f3 b)
{
    float p0 = a.x * b.x;
    float p1 = a.y * b.y;
    float p2 = a.z * b.z;
    float s = p0 + p1;
    return s + p2;
}

/* cross per the GLSL spec: (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y). */
RV_HD rv_f3 rv_cross(rv_f3 a, rv_f3 b)
{
    float x0 = a.y * b.z, x1 = b.y * a.z;
    float y0 = a.z * b.x, y1 = b.z * a.x;
    float z0 = a.x * b.y, z1 = b.x * a.y;
    return rv_make(x0 - x1, y0 - y1, z0 - z1);
}

/* normalize: v * (1 / sqrt(dot(v, v))) — the inversesqrt lowering. */
RV_HD rv_f3 rv_normalize(rv_f3 v)
{
    float inv = 1.0f / sqrtf(rv_dot(v, v));
    return rv_make(v.x * inv, v.y * inv, v.z * inv);
}

/* mix(a, b, t) = a*(1-t) + b*t (GLSL spec formula). */
RV_HD float rv_mix(float a, float b, float t)
{
    float w = 1.0f - t;
    float l = a * w;
    float r = b * t;
    return l + r;
}

/* ---- RNG: util.glsl:25-50 ------------------------------------------------ */
RV_HD uint32_t rv_wang_hash(uint32_t seed)
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}

RV_HD uint32_t rv_xorshift(uint32_t* state)
{
    uint32_t s = *state;
    s ^= (s << 13);
    s ^= (s >> 17);
    s ^= (s << 5);
    *state = s;
    return s;
}

/* float(uint) (round to nearest even) / 2^32 — may return exactly 1.0f. */
RV_HD float rv_rand(uint32_t* state)
{
    uint32_t s = rv_xorshift(state);
#if defined(__CUDA_ARCH__)
    return __uint2float_rn(s) / 4294967296.0f;
#else
    return (float)s / 4294967296.0f;
#endif
}

/* ---- sin / cos ------------------------------------------------------------
 * Quadrant reduction by pi/2 with a three-term Cody-Waite split (q*C1 and
 * q*C2 are exact for |q| < 2^13 because C1/C2 carry 8/11 significant bits),
 * then the classic single-precision minimax polynomials on [-pi/4, pi/4].
 * Max error vs the real functions is < 2 ulp on [-2pi, 2pi] without FMA —
 * far inside Vulkan's 2^-11 absolute bound for sin/cos.
 */
RV_HD void rv_sincos(float x, float* s_out, float* c_out)
{
    if (!(fabsf(x) <= 8192.0f))
    {
        /* out of the supported domain (also NaN / inf) */
        float n = x - x;
        n = n / n;
        *s_out = n;
        *c_out = n;
        return;
    }
    float q = rintf(x * 0.636619747f); /* 2/pi */
    float r = x - q * 1.5703125f;
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.54978995489188216e-8f;
    int qi = (int)q;
    float z = r * r;

    /* sin(r) = r + r*z*(S1 + z*(S2 + z*S3)) */
    float ps = -1.9515295891e-4f * z;
    ps = ps + 8.3321608736e-3f;
    ps = ps * z;
    ps = ps - 1.6666654611e-1f;
    ps = ps * z;
    ps = ps * r;
    float sr = ps + r;

    /* cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + z*C3)) */
    float pc = 2.443315711809948e-5f * z;
    pc = pc - 1.388731625493765e-3f;
    pc = pc * z;
    pc = pc + 4.166664568298827e-2f;
    pc = pc * z;
    pc = pc * z;
    float hz = 0.5f * z;
    pc = pc - hz;
    float cr = pc + 1.0f;

    float s, c;
    if (qi & 1)
    {
        s = cr;
        c = sr;
    }
    else
    {
        s = sr;
        c = cr;
    }
    if (qi & 2) s = -s;
    if ((qi + 1) & 2) c = -c;
    *s_out = s;
    *c_out = c;
}

RV_HD float rv_tan(float x)
{
    float s, c;
    rv_sincos(x, &s, &c);
    return s / c;
}

/* UNORM8 conversions of the rgba8 images (rvpt.cpp:759-766): store =
 * round-to-nearest-even of clamp(x,0,1)*255, load = k/255. NaN stores 0. */
RV_HD uint32_t rv_unorm8_store(float x)
{
    float c = x;
    if (!(c > 0.0f)) c = 0.0f;
    if (c > 1.0f) c = 1.0f;
    return (uint32_t)(int)rintf(c * 255.0f);
}
RV_HD float rv_unorm8_load(uint32_t k) { return (float)k / 255.0f; }

/* Returns 1 if a*b+c was evaluated with two roundings (what this header
 * requires), 0 if the compiler contracted it into an FMA. With the RV_PROBE_*
 * values a*b = 1 - 2^-26 exactly, which rounds to 1.0f, so the unfused sum is
 * 0 while an FMA returns -2^-26. Call it with run-time (volatile) operands. */
RV_HD int rv_contract_probe(float a, float b, float c)
{
    float r = a * b + c;
    return r == 0.0f ? 1 : 0;
}
#define RV_PROBE_A 1.0001220703125f /* 1 + 2^-13 */
#define RV_PROBE_B 0.9998779296875f /* 1 - 2^-13 */
#define RV_PROBE_C (-1.0f)

#endif /* RVPT_MATH_H */
