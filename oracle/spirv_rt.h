/*
 * spirv_rt.h — runtime of the C++ that oracle/spirv_to_cpp.py generates from the reference's
 * shipped compute shader (TEST INFRASTRUCTURE). Same value model and the same driver-defined
 * operations as the interpreter oracle/spirv_vm.cpp: every value a run of 32-bit words, f64 two
 * words, the arithmetic contract of include/rvpt_math.h for what SPIR-V leaves to the driver.
 * Compile the generated file with -ffp-contract=off.
 */
#ifndef RVPT_SPIRV_RT_H
#define RVPT_SPIRV_RT_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <cstring>

#include "../include/rvpt_math.h"

struct Ptr
{
    uint8_t* p;
    uint32_t tag; /* bit 0: buffer memory (explicit layout); bits 8..: matrix stride of the enclosing member */
};

struct RtImage
{
    float* f32;  /* W*H*4 floats, or */
    uint8_t* u8; /* W*H*4 bytes with UNORM8 conversion (the reference's image format) */
    int W, H;
};

struct RtBindings
{
    const uint8_t* buf[8];
    size_t bytes[8];
    RtImage img[8];
};

struct Ctx
{
    const RtBindings* bind;
    uint32_t priv[256]; /* Private / Input / UniformConstant variables of the invocation */
    Ptr g[4096];        /* global variables by id */
};

static inline float asf(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t asu(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline double asd(const uint32_t* p) { double d; std::memcpy(&d, p, 8); return d; }
static inline void putd(uint32_t* p, double d) { std::memcpy(p, &d, 8); }
[[noreturn]] static inline void rt_unreachable() { abort(); }

static inline void rt_image_read(Ctx& cx, uint32_t h, int x, int y, uint32_t* out)
{
    const RtImage& im = cx.bind->img[h & 7u];
    if (x < 0 || y < 0 || x >= im.W || y >= im.H)
    {
        out[0] = out[1] = out[2] = out[3] = 0;
        return;
    }
    const size_t i = ((size_t)y * im.W + x) * 4;
    for (int c = 0; c < 4; ++c) out[c] = asu(im.u8 ? rv_unorm8_load(im.u8[i + c]) : im.f32[i + c]);
}
static inline void rt_image_write(Ctx& cx, uint32_t h, int x, int y, const uint32_t* v)
{
    const RtImage& im = cx.bind->img[h & 7u];
    if (x < 0 || y < 0 || x >= im.W || y >= im.H) return;
    const size_t i = ((size_t)y * im.W + x) * 4;
    for (int c = 0; c < 4; ++c)
    {
        if (im.u8)
            im.u8[i + c] = (uint8_t)rv_unorm8_store(asf(v[c]));
        else
            im.f32[i + c] = asf(v[c]);
    }
}
static inline void rt_image_size(Ctx& cx, uint32_t h, uint32_t* out)
{
    out[0] = (uint32_t)cx.bind->img[h & 7u].W;
    out[1] = (uint32_t)cx.bind->img[h & 7u].H;
}

/* sum of column * component, left to right (OpMatrixTimesVector) */
template <int kCols, int kRows>
static inline void rt_mat_vec(uint32_t* r, const uint32_t* m, const uint32_t* v)
{
    uint32_t t[4];
    for (int i = 0; i < kRows; ++i)
    {
        float acc = asf(m[i]) * asf(v[0]);
        for (int c = 1; c < kCols; ++c)
        {
            const float p = asf(m[c * kRows + i]) * asf(v[c]);
            acc = acc + p;
        }
        t[i] = asu(acc);
    }
    std::memcpy(r, t, 4 * kRows);
}
template <int kN>
static inline void rt_dot(uint32_t* r, const uint32_t* x, const uint32_t* y)
{
    float acc = asf(x[0]) * asf(y[0]);
    for (int k = 1; k < kN; ++k)
    {
        const float p = asf(x[k]) * asf(y[k]);
        acc = acc + p;
    }
    r[0] = asu(acc);
}

/* GLSL.std.450: kInst = instruction number, kN = result components, kXN = components of operand 0 */
template <int kInst, int kN, int kXN>
static inline void rt_ext(uint32_t* r, const uint32_t* x)
{
    if (kInst == 4) { for (int k = 0; k < kN; ++k) r[k] = asu(fabsf(asf(x[k]))); }
    else if (kInst == 6) { for (int k = 0; k < kN; ++k) { const float v = asf(x[k]); r[k] = asu(v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f)); } }
    else if (kInst == 13 || kInst == 14)
    {
        for (int k = 0; k < kN; ++k)
        {
            float s, c;
            rv_sincos(asf(x[k]), &s, &c);
            r[k] = asu(kInst == 13 ? s : c);
        }
    }
    else if (kInst == 15) { for (int k = 0; k < kN; ++k) r[k] = asu(rv_tan(asf(x[k]))); }
    else if (kInst == 31) { for (int k = 0; k < kN; ++k) r[k] = asu(sqrtf(asf(x[k]))); }
    else if (kInst == 66)
    {
        if (kXN == 1) { r[0] = asu(fabsf(asf(x[0]))); return; }
        float p = asf(x[0]) * asf(x[0]);
        for (int k = 1; k < kXN; ++k) { const float q = asf(x[k]) * asf(x[k]); p = p + q; }
        r[0] = asu(sqrtf(p));
    }
    else if (kInst == 69)
    {
        const rv_f3 c = rv_normalize(rv_make(asf(x[0]), asf(x[1]), asf(x[2])));
        r[0] = asu(c.x), r[1] = asu(c.y), r[2] = asu(c.z);
    }
    else
        abort();
}
template <int kInst, int kN, int kXN>
static inline void rt_ext(uint32_t* r, const uint32_t* x, const uint32_t* y)
{
    if (kInst == 37) { for (int k = 0; k < kN; ++k) r[k] = asu(fminf(asf(x[k]), asf(y[k]))); }
    else if (kInst == 40) { for (int k = 0; k < kN; ++k) r[k] = asu(fmaxf(asf(x[k]), asf(y[k]))); }
    else if (kInst == 38) { for (int k = 0; k < kN; ++k) r[k] = x[k] < y[k] ? x[k] : y[k]; }
    else if (kInst == 68)
    {
        const rv_f3 c = rv_cross(rv_make(asf(x[0]), asf(x[1]), asf(x[2])), rv_make(asf(y[0]), asf(y[1]), asf(y[2])));
        uint32_t t[3] = {asu(c.x), asu(c.y), asu(c.z)};
        std::memcpy(r, t, 12);
    }
    else
        abort();
}
template <int kInst, int kN, int kXN>
static inline void rt_ext(uint32_t* r, const uint32_t* x, const uint32_t* y, const uint32_t* z)
{
    if (kInst == 43) { for (int k = 0; k < kN; ++k) r[k] = asu(fminf(fmaxf(asf(x[k]), asf(y[k])), asf(z[k]))); }
    else if (kInst == 46) { for (int k = 0; k < kN; ++k) r[k] = asu(rv_mix(asf(x[k]), asf(y[k]), asf(z[k]))); }
    else
        abort();
}
/* f64 results: only FMin / FMax occur (intersect_aabb) */
template <int kInst, int kN, int kXN>
static inline void rt_ext64(uint32_t* r, const uint32_t* x, const uint32_t* y)
{
    for (int k = 0; k < kN; ++k)
    {
        const double a = asd(x + 2 * k), b = asd(y + 2 * k);
        putd(r + 2 * k, kInst == 37 ? fmin(a, b) : fmax(a, b));
    }
}

#endif
