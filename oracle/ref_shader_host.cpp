/*
 * ref_shader_host.cpp — C API around the C++ that oracle/spirv_to_cpp.py generates from the
 * reference's shipped compute_pass.comp.spv (TEST INFRASTRUCTURE). Linked with the generated file
 * into oracle/_ref/libref_shader.so (`make -C oracle ref_shader`, next to the reference tree only).
 * One invocation per pixel, rows spread over host threads; every invocation only touches its own
 * texel of the two storage images.
 */
#include <thread>
#include <vector>

#include "spirv_rt.h"

extern "C" void ref_shader_invoke(Ctx* cx, uint32_t x, uint32_t y);
extern "C" void ref_shader_bind(Ctx* cx);
extern "C" uint32_t ref_shader_max_id();

#define API extern "C" __attribute__((visibility("default")))

struct rvpt_ref_shader
{
    RtBindings bind{};
};

API rvpt_ref_shader* rvpt_ref_shader_create()
{
    if (ref_shader_max_id() > 4096) return nullptr; /* Ctx::g is sized for the shipped module */
    return new rvpt_ref_shader();
}
API void rvpt_ref_shader_destroy(rvpt_ref_shader* s) { delete s; }
API void rvpt_ref_shader_bind_buffer(rvpt_ref_shader* s, int binding, const void* data, size_t bytes)
{
    if (binding < 0 || binding >= 8) return;
    s->bind.buf[binding] = (const uint8_t*)data;
    s->bind.bytes[binding] = bytes;
}
API void rvpt_ref_shader_bind_image(rvpt_ref_shader* s, int binding, void* data, int W, int H, int unorm8)
{
    if (binding < 0 || binding >= 8) return;
    RtImage& im = s->bind.img[binding];
    im.W = W, im.H = H;
    im.u8 = unorm8 ? (uint8_t*)data : nullptr;
    im.f32 = unorm8 ? nullptr : (float*)data;
}
API int rvpt_ref_shader_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
API int rvpt_ref_shader_dispatch(rvpt_ref_shader* s, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, int nthreads)
{
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    const uint32_t rows = y1 > y0 ? y1 - y0 : 0;
    if ((uint32_t)nthreads > rows) nthreads = rows ? (int)rows : 1;
    auto work = [&](int t) {
        Ctx* cx = new Ctx();
        cx->bind = &s->bind;
        ref_shader_bind(cx);
        for (uint32_t y = y0 + (uint32_t)t; y < y1; y += (uint32_t)nthreads)
            for (uint32_t x = x0; x < x1; ++x) ref_shader_invoke(cx, x, y);
        delete cx;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    return 0;
}
