/*
 * spirv_vm.cpp — a small SPIR-V interpreter for the reference's SHIPPED compute shader.
 *
 * TEST INFRASTRUCTURE (like everything under oracle/): it exists to pin the CPU oracle
 * (oracle/rvpt_oracle.cpp) to an artefact the reference itself holds. The reference ships
 * the hot path compiled: assets/shaders/compute_pass.comp.spv (glslang output, in sync with
 * the GLSL sources). No Vulkan loader / ICD exists in this image, so the binary cannot run on
 * a driver here; this interpreter executes it instead, one invocation (= one pixel) at a time,
 * with the descriptor bindings of the reference (rvpt.cpp:646-655, compute_pass.comp:28-58):
 *
 *   binding 0  RenderSettings UBO      binding 4  Camera UBO
 *   binding 1  result image (rgba8)    binding 5  BvhNode[] SSBO
 *   binding 2  temporal image (rgba8)  binding 6  Triangle[] SSBO
 *   binding 3  random floats (dead)    binding 7  Material[] SSBO
 *
 * Everything the binary fixes is taken from the binary: control flow, the order of
 * operations, the order of the rand() draws, constants, struct layouts (Offset / ArrayStride
 * / MatrixStride decorations), f32 vs f64 arithmetic. What SPIR-V leaves to the driver —
 * the summation order of OpDot / OpMatrixTimesVector, the accuracy of the GLSL.std.450
 * Sin / Cos / Tan / Normalize / Length / Cross / FMix, UNORM8 image conversion — follows the
 * arithmetic contract of include/rvpt_math.h, the same one oracle and kernels use (every f32
 * operation separately rounded, left to right). Nothing under /root/reference is copied: the
 * .spv is read at run time from the path the caller passes (tests/golden/make_spirv_golden.py
 * turns its outputs into committed fixtures for the GPU box, where the reference is absent).
 *
 * Supported: exactly the opcodes the shipped module uses (84 of them) + the 14 GLSL.std.450
 * instructions it calls; anything else aborts the run with an error message.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "../include/rvpt_math.h"

namespace
{

enum Kind { K_VOID, K_BOOL, K_INT, K_FLOAT, K_VEC, K_MAT, K_IMAGE, K_ARRAY, K_RTARRAY, K_STRUCT, K_PTR, K_FUNC };

struct Type
{
    Kind kind = K_VOID;
    uint32_t width = 32;     /* scalar bits */
    uint32_t elem = 0;       /* component / column / element / pointee type id */
    uint32_t count = 0;      /* components, columns, array length */
    uint32_t storage = 0;    /* pointer storage class */
    std::vector<uint32_t> members;
    uint32_t flat = 0;       /* size in 32-bit words of a value of this type inside the VM */
    /* explicit layout (decorations) for types living in buffers */
    std::vector<uint32_t> member_offset, member_matrix_stride;
    uint32_t array_stride = 0;
};

struct Func
{
    uint32_t id = 0, first_word = 0, frame_words = 0;
    std::vector<uint32_t> params;
};

struct Module
{
    std::vector<uint32_t> w;
    uint32_t bound = 0;
    std::vector<Type> types;            /* by id */
    std::vector<uint8_t> is_type, is_global;
    std::vector<uint32_t> off;          /* by id: word offset in the global pool or in the frame */
    std::vector<uint32_t> type_of;      /* by id: result type */
    std::vector<uint32_t> var_store;    /* by id: offset of a variable's backing storage */
    std::vector<uint32_t> label_pos;    /* by id: word index of the instruction after OpLabel */
    std::vector<int32_t> func_index;    /* by id */
    std::vector<Func> funcs;
    std::vector<uint32_t> gpool;        /* constants + global pointers (template, copied per machine) */
    uint32_t priv_words = 0;            /* Private-storage variables */
    struct GVar { uint32_t id, storage, pointee, store_off; int set, binding, builtin; };
    std::vector<GVar> gvars;
    std::vector<int> deco_binding, deco_builtin;
    uint32_t entry = 0;
    uint32_t glsl_ext = 0;
    std::string err;
};

struct Image
{
    float* f32 = nullptr;    /* W*H*4 floats (mode 0) */
    uint8_t* u8 = nullptr;   /* W*H*4 bytes (mode 1: UNORM8 storage like the reference's images) */
    int W = 0, H = 0;
};

struct Bindings
{
    const uint8_t* buf[8] = {};
    size_t bytes[8] = {};
    Image img[8];
};

/* pointer value inside the VM: 3 words = host address (lo, hi), tag (bit 0: external layout,
 * bits 8.. matrix stride of the enclosing member) */
inline void put_ptr(uint32_t* dst, const void* p, uint32_t tag)
{
    const uint64_t a = (uint64_t)(uintptr_t)p;
    dst[0] = (uint32_t)a, dst[1] = (uint32_t)(a >> 32), dst[2] = tag;
}
inline uint8_t* get_ptr(const uint32_t* src) { return (uint8_t*)(uintptr_t)((uint64_t)src[0] | ((uint64_t)src[1] << 32)); }

inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline double asd(const uint32_t* p) { double d; memcpy(&d, p, 8); return d; }
inline void putd(uint32_t* p, double d) { memcpy(p, &d, 8); }

bool fail(Module& m, const char* fmt, uint32_t a = 0, uint32_t b = 0)
{
    char buf[256];
    snprintf(buf, sizeof buf, fmt, a, b);
    if (m.err.empty()) m.err = buf;
    return false;
}

uint32_t flat_size(Module& m, uint32_t tid)
{
    Type& t = m.types[tid];
    if (t.flat) return t.flat;
    switch (t.kind)
    {
        case K_VOID: case K_FUNC: t.flat = 0; break;
        case K_BOOL: case K_INT: case K_IMAGE: t.flat = 1; break;
        case K_FLOAT: t.flat = t.width / 32; break;
        case K_VEC: case K_MAT: case K_ARRAY: t.flat = t.count * flat_size(m, t.elem); break;
        case K_RTARRAY: t.flat = 0; break;
        case K_STRUCT: { uint32_t s = 0; for (uint32_t x : t.members) s += flat_size(m, x); t.flat = s; break; }
        case K_PTR: t.flat = 3; break;
    }
    return t.flat;
}

bool parse(Module& m)
{
    const std::vector<uint32_t>& w = m.w;
    if (w.size() < 5 || w[0] != 0x07230203u) return fail(m, "not a SPIR-V module");
    m.bound = w[3];
    const uint32_t B = m.bound;
    m.types.assign(B, Type());
    m.is_type.assign(B, 0), m.is_global.assign(B, 0);
    m.off.assign(B, 0), m.type_of.assign(B, 0), m.var_store.assign(B, 0), m.label_pos.assign(B, 0);
    m.func_index.assign(B, -1);
    m.deco_binding.assign(B, -1), m.deco_builtin.assign(B, -1);
    std::vector<uint32_t> const_len(B, 0); /* array lengths need constant values */

    /* pass 1: decorations, types, constants, global variables */
    Func* cur = nullptr;
    for (size_t i = 5; i < w.size();)
    {
        const uint32_t wc = w[i] >> 16, op = w[i] & 0xFFFFu;
        if (wc == 0) return fail(m, "zero-length instruction");
        const uint32_t* a = &w[i + 1];
        switch (op)
        {
            case 11: /* OpExtInstImport */
                if (strcmp((const char*)&a[1], "GLSL.std.450") == 0) m.glsl_ext = a[0];
                break;
            case 15: /* OpEntryPoint */ m.entry = a[1]; break;
            case 71: /* OpDecorate */
                if (a[1] == 33) m.deco_binding[a[0]] = (int)a[2];
                if (a[1] == 11) m.deco_builtin[a[0]] = (int)a[2];
                if (a[1] == 6) m.types[a[0]].array_stride = a[2];
                break;
            case 72: /* OpMemberDecorate */
            {
                Type& t = m.types[a[0]];
                if (t.member_offset.size() <= a[1]) t.member_offset.resize(a[1] + 1, 0), t.member_matrix_stride.resize(a[1] + 1, 0);
                if (a[2] == 35) t.member_offset[a[1]] = a[3];
                if (a[2] == 7) t.member_matrix_stride[a[1]] = a[3];
                if (a[2] == 4) return fail(m, "RowMajor matrices are not supported");
                break;
            }
            case 19: m.types[a[0]].kind = K_VOID, m.is_type[a[0]] = 1; break;
            case 20: m.types[a[0]].kind = K_BOOL, m.is_type[a[0]] = 1; break;
            case 21: m.types[a[0]].kind = K_INT, m.types[a[0]].width = a[1], m.is_type[a[0]] = 1; break;
            case 22: m.types[a[0]].kind = K_FLOAT, m.types[a[0]].width = a[1], m.is_type[a[0]] = 1; break;
            case 23: m.types[a[0]].kind = K_VEC, m.types[a[0]].elem = a[1], m.types[a[0]].count = a[2], m.is_type[a[0]] = 1; break;
            case 24: m.types[a[0]].kind = K_MAT, m.types[a[0]].elem = a[1], m.types[a[0]].count = a[2], m.is_type[a[0]] = 1; break;
            case 25: m.types[a[0]].kind = K_IMAGE, m.is_type[a[0]] = 1; break;
            case 28: m.types[a[0]].kind = K_ARRAY, m.types[a[0]].elem = a[1], m.types[a[0]].count = const_len[a[2]], m.is_type[a[0]] = 1; break;
            case 29: m.types[a[0]].kind = K_RTARRAY, m.types[a[0]].elem = a[1], m.is_type[a[0]] = 1; break;
            case 30:
            {
                Type& t = m.types[a[0]];
                t.kind = K_STRUCT, m.is_type[a[0]] = 1;
                t.members.assign(a + 1, a + wc - 1);
                t.member_offset.resize(t.members.size(), 0), t.member_matrix_stride.resize(t.members.size(), 0);
                break;
            }
            case 32: m.types[a[1 - 1]].kind = K_PTR, m.types[a[0]].storage = a[1], m.types[a[0]].elem = a[2], m.is_type[a[0]] = 1; break;
            case 33: m.types[a[0]].kind = K_FUNC, m.is_type[a[0]] = 1; break;
            case 41: case 42: case 43: case 44: /* constants */
            {
                const uint32_t id = a[1], n = flat_size(m, a[0]);
                m.type_of[id] = a[0], m.is_global[id] = 1, m.off[id] = (uint32_t)m.gpool.size();
                if (op == 41) m.gpool.push_back(1);
                else if (op == 42) m.gpool.push_back(0);
                else if (op == 43)
                {
                    for (uint32_t k = 0; k < n; ++k) m.gpool.push_back(a[2 + k]);
                    const_len[id] = a[2];
                }
                else
                {
                    std::vector<uint32_t> v;
                    for (uint32_t k = 2; k < wc - 1; ++k)
                    {
                        const uint32_t c = a[k], cn = flat_size(m, m.type_of[c]);
                        for (uint32_t q = 0; q < cn; ++q) v.push_back(m.gpool[m.off[c] + q]);
                    }
                    if (v.size() != n) return fail(m, "constant composite %u has the wrong size", id);
                    m.gpool.insert(m.gpool.end(), v.begin(), v.end());
                }
                break;
            }
            case 54: /* OpFunction */
                m.funcs.push_back(Func());
                cur = &m.funcs.back();
                cur->id = a[1];
                m.func_index[a[1]] = (int32_t)m.funcs.size() - 1;
                break;
            case 56: cur = nullptr; break;
            case 59: /* OpVariable */
                if (!cur)
                {
                    const uint32_t id = a[1], ptype = a[0], sc = a[2];
                    Module::GVar g{id, sc, m.types[ptype].elem, 0, 0, m.deco_binding[id], m.deco_builtin[id]};
                    m.type_of[id] = ptype, m.is_global[id] = 1, m.off[id] = (uint32_t)m.gpool.size();
                    m.gpool.push_back(0), m.gpool.push_back(0), m.gpool.push_back(0);
                    if (sc == 6 || sc == 1 || sc == 0) /* Private, Input, UniformConstant: VM-side storage */
                    {
                        g.store_off = m.priv_words;
                        m.priv_words += flat_size(m, g.pointee);
                        if (wc > 4) return fail(m, "global initialisers are not supported (%u)", id);
                    }
                    m.gvars.push_back(g);
                }
                break;
            default: break;
        }
        i += wc;
    }

    /* pass 2: frame layout of every function, label positions */
    cur = nullptr;
    uint32_t fi = 0;
    for (size_t i = 5; i < w.size();)
    {
        const uint32_t wc = w[i] >> 16, op = w[i] & 0xFFFFu;
        const uint32_t* a = &w[i + 1];
        if (op == 54)
        {
            cur = &m.funcs[fi++];
            cur->first_word = (uint32_t)(i + wc);
            cur->frame_words = 0;
        }
        else if (op == 56)
            cur = nullptr;
        else if (cur)
        {
            uint32_t rtype = 0, rid = 0;
            switch (op)
            {
                case 55: rtype = a[0], rid = a[1]; cur->params.push_back(rid); break;
                case 248: m.label_pos[a[0]] = (uint32_t)(i + wc); break;
                case 59:
                    rtype = a[0], rid = a[1];
                    m.var_store[rid] = cur->frame_words;
                    cur->frame_words += flat_size(m, m.types[rtype].elem);
                    if (wc > 4) return fail(m, "variable initialisers are not supported (%u)", rid);
                    break;
                case 12: case 57: case 61: case 65: case 68: case 79: case 80: case 81: case 98: case 104:
                case 110: case 111: case 112: case 115: case 124: case 127: case 128: case 129: case 130:
                case 131: case 132: case 133: case 136: case 142: case 145: case 148: case 166: case 167:
                case 168: case 169: case 171: case 172: case 176: case 177: case 184: case 186: case 188:
                case 190: case 194: case 196: case 198: case 245:
                    rtype = a[0], rid = a[1];
                    break;
                case 62: case 99: case 246: case 247: case 249: case 250: case 251: case 253: case 254: case 255:
                    break;
                default: return fail(m, "unsupported opcode %u in function %u", op, cur->id);
            }
            if (rid)
            {
                m.type_of[rid] = rtype;
                m.off[rid] = cur->frame_words;
                cur->frame_words += flat_size(m, rtype);
            }
        }
        i += wc;
    }
    if (m.func_index[m.entry] < 0) return fail(m, "entry point not found");
    return true;
}

struct Machine
{
    Module& m;
    const Bindings& bind;
    std::vector<uint32_t> gpool, priv, arena;
    uint64_t executed = 0;
    std::string err;

    Machine(Module& mod, const Bindings& b) : m(mod), bind(b), gpool(mod.gpool), priv(mod.priv_words + 4, 0), arena(1u << 20, 0)
    {
        for (const Module::GVar& g : m.gvars)
        {
            if (g.storage == 6 || g.storage == 1 || g.storage == 0)
                put_ptr(&gpool[m.off[g.id]], &priv[g.store_off], 0);
            else /* Uniform (UBO / SSBO blocks) */
            {
                const int b2 = g.binding;
                put_ptr(&gpool[m.off[g.id]], (b2 >= 0 && b2 < 8) ? bind.buf[b2] : nullptr, 1);
            }
            if (g.storage == 0) priv[g.store_off] = (uint32_t)g.binding; /* image handle = binding */
        }
    }

    inline uint32_t* V(uint32_t* frame, uint32_t id) { return m.is_global[id] ? &gpool[m.off[id]] : frame + m.off[id]; }

    bool die(const char* fmt, uint32_t a = 0, uint32_t b = 0)
    {
        char buf[256];
        snprintf(buf, sizeof buf, fmt, a, b);
        if (err.empty()) err = buf;
        return false;
    }

    /* external (decorated) memory -> flat VM value */
    void load_ext(uint32_t tid, const uint8_t* p, uint32_t* out, uint32_t mstride)
    {
        const Type& t = m.types[tid];
        switch (t.kind)
        {
            case K_BOOL: case K_INT: memcpy(out, p, 4); break;
            case K_FLOAT: memcpy(out, p, t.width / 8); break;
            case K_VEC: memcpy(out, p, 4u * t.flat); break;
            case K_MAT:
                for (uint32_t c = 0; c < t.count; ++c) memcpy(out + c * m.types[t.elem].flat, p + c * mstride, 4u * m.types[t.elem].flat);
                break;
            case K_ARRAY:
                for (uint32_t k = 0; k < t.count; ++k) load_ext(t.elem, p + k * t.array_stride, out + k * m.types[t.elem].flat, mstride);
                break;
            case K_STRUCT:
            {
                uint32_t o = 0;
                for (size_t k = 0; k < t.members.size(); ++k)
                {
                    load_ext(t.members[k], p + t.member_offset[k], out + o, t.member_matrix_stride[k]);
                    o += m.types[t.members[k]].flat;
                }
                break;
            }
            default: break;
        }
    }

    bool image_read(uint32_t handle, int x, int y, uint32_t* out)
    {
        if (handle >= 8) return die("bad image handle %u", handle);
        const Image& im = bind.img[handle];
        if (x < 0 || y < 0 || x >= im.W || y >= im.H)
        {
            out[0] = out[1] = out[2] = out[3] = 0; /* robust-access style: out of bounds reads 0 */
            return true;
        }
        const size_t i = ((size_t)y * im.W + x) * 4;
        for (int c = 0; c < 4; ++c)
            out[c] = asu(im.u8 ? rv_unorm8_load(im.u8[i + c]) : im.f32[i + c]);
        return true;
    }
    bool image_write(uint32_t handle, int x, int y, const uint32_t* v)
    {
        if (handle >= 8) return die("bad image handle %u", handle);
        const Image& im = bind.img[handle];
        if (x < 0 || y < 0 || x >= im.W || y >= im.H) return true;
        const size_t i = ((size_t)y * im.W + x) * 4;
        for (int c = 0; c < 4; ++c)
        {
            if (im.u8)
                im.u8[i + c] = rv_unorm8_store(asf(v[c]));
            else
                im.f32[i + c] = asf(v[c]);
        }
        return true;
    }

    bool ext_inst(uint32_t inst, uint32_t rtype, uint32_t* r, uint32_t* frame, const uint32_t* ops)
    {
        const Type& rt = m.types[rtype];
        const bool f64 = (rt.kind == K_FLOAT ? rt.width : m.types[rt.elem].width) == 64;
        const uint32_t n = f64 ? rt.flat / 2 : rt.flat;
        const uint32_t* x = V(frame, ops[0]);
        if (f64)
        {
            if (inst != 37 && inst != 40) return die("f64 GLSL.std.450 instruction %u", inst);
            const uint32_t* y = V(frame, ops[1]);
            for (uint32_t k = 0; k < n; ++k)
            {
                const double a = asd(x + 2 * k), b = asd(y + 2 * k);
                /* NaN operands are undefined in GLSL.std.450; resolved as IEEE minNum/maxNum like the f32 case */
                putd(r + 2 * k, inst == 37 ? fmin(a, b) : fmax(a, b));
            }
            return true;
        }
        switch (inst)
        {
            case 4: for (uint32_t k = 0; k < n; ++k) r[k] = asu(fabsf(asf(x[k]))); return true;
            case 6: for (uint32_t k = 0; k < n; ++k) { const float v = asf(x[k]); r[k] = asu(v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f)); } return true;
            case 13: case 14:
                for (uint32_t k = 0; k < n; ++k)
                {
                    float s, c;
                    rv_sincos(asf(x[k]), &s, &c);
                    r[k] = asu(inst == 13 ? s : c);
                }
                return true;
            case 15: for (uint32_t k = 0; k < n; ++k) r[k] = asu(rv_tan(asf(x[k]))); return true;
            case 31: for (uint32_t k = 0; k < n; ++k) r[k] = asu(sqrtf(asf(x[k]))); return true;
            case 37: case 40:
            {
                const uint32_t* y = V(frame, ops[1]);
                for (uint32_t k = 0; k < n; ++k)
                    r[k] = asu(inst == 37 ? fminf(asf(x[k]), asf(y[k])) : fmaxf(asf(x[k]), asf(y[k])));
                return true;
            }
            case 38:
            {
                const uint32_t* y = V(frame, ops[1]);
                for (uint32_t k = 0; k < n; ++k) r[k] = x[k] < y[k] ? x[k] : y[k];
                return true;
            }
            case 43: /* FClamp = min(max(x, lo), hi) */
            {
                const uint32_t *lo = V(frame, ops[1]), *hi = V(frame, ops[2]);
                for (uint32_t k = 0; k < n; ++k) r[k] = asu(fminf(fmaxf(asf(x[k]), asf(lo[k])), asf(hi[k])));
                return true;
            }
            case 46: /* FMix = x*(1-a) + y*a */
            {
                const uint32_t *y = V(frame, ops[1]), *t = V(frame, ops[2]);
                for (uint32_t k = 0; k < n; ++k) r[k] = asu(rv_mix(asf(x[k]), asf(y[k]), asf(t[k])));
                return true;
            }
            case 66: /* Length */
            {
                const uint32_t xn = m.types[m.type_of[ops[0]]].flat;
                if (xn == 1) { r[0] = asu(fabsf(asf(x[0]))); return true; }
                float p = asf(x[0]) * asf(x[0]);
                for (uint32_t k = 1; k < xn; ++k) { const float q = asf(x[k]) * asf(x[k]); p = p + q; }
                r[0] = asu(sqrtf(p));
                return true;
            }
            case 68: /* Cross */
            {
                const uint32_t* y = V(frame, ops[1]);
                const rv_f3 c = rv_cross(rv_make(asf(x[0]), asf(x[1]), asf(x[2])), rv_make(asf(y[0]), asf(y[1]), asf(y[2])));
                r[0] = asu(c.x), r[1] = asu(c.y), r[2] = asu(c.z);
                return true;
            }
            case 69: /* Normalize */
            {
                if (n != 3) return die("normalize of a %u-vector", n);
                const rv_f3 c = rv_normalize(rv_make(asf(x[0]), asf(x[1]), asf(x[2])));
                r[0] = asu(c.x), r[1] = asu(c.y), r[2] = asu(c.z);
                return true;
            }
            default: return die("unsupported GLSL.std.450 instruction %u", inst);
        }
    }

    /* runs function `fidx` with its frame at arena[base...]; params already stored */
    bool run(int fidx, uint32_t base, uint32_t* ret)
    {
        const Func& fn = m.funcs[fidx];
        if ((size_t)base + fn.frame_words + 64 > arena.size()) return die("VM stack overflow");
        uint32_t* frame = &arena[base];
        const std::vector<uint32_t>& w = m.w;
        uint32_t pc = fn.first_word;
        uint32_t cur_label = 0, prev_label = 0;
        for (;;)
        {
            const uint32_t wc = w[pc] >> 16, op = w[pc] & 0xFFFFu;
            const uint32_t* a = &w[pc + 1];
            ++executed;
            uint32_t next = pc + wc;
            switch (op)
            {
                case 55: case 246: case 247: break; /* params are in place; merges are hints */
                case 248: prev_label = cur_label, cur_label = a[0]; break;
                case 249: next = m.label_pos[a[0]], prev_label = cur_label, cur_label = a[0]; break;
                case 250:
                {
                    const uint32_t t = V(frame, a[0])[0] ? a[1] : a[2];
                    next = m.label_pos[t], prev_label = cur_label, cur_label = t;
                    break;
                }
                case 251:
                {
                    const uint32_t sel = V(frame, a[0])[0];
                    uint32_t t = a[1];
                    for (uint32_t k = 2; k + 1 < wc - 0 && k + 1 <= wc - 1; k += 2)
                        if (a[k] == sel) { t = a[k + 1]; break; }
                    next = m.label_pos[t], prev_label = cur_label, cur_label = t;
                    break;
                }
                case 253: return true;
                case 254:
                {
                    const uint32_t n = m.types[m.type_of[a[0]]].flat;
                    memcpy(ret, V(frame, a[0]), 4u * n);
                    return true;
                }
                case 255: return die("OpUnreachable executed");
                case 245: /* OpPhi */
                {
                    uint32_t* r = V(frame, a[1]);
                    const uint32_t n = m.types[a[0]].flat;
                    bool found = false;
                    for (uint32_t k = 2; k + 1 < wc; k += 2)
                        if (a[k + 1] == prev_label)
                        {
                            memcpy(r, V(frame, a[k]), 4u * n);
                            found = true;
                            break;
                        }
                    if (!found) return die("OpPhi %u: no incoming edge from %u", a[1], prev_label);
                    break;
                }
                case 59: /* OpVariable (Function) */
                    put_ptr(V(frame, a[1]), frame + m.var_store[a[1]], 0);
                    break;
                case 61: /* OpLoad */
                {
                    uint32_t* r = V(frame, a[1]);
                    const uint32_t* pv = V(frame, a[2]);
                    const uint8_t* p = get_ptr(pv);
                    if (!p) return die("load through a null pointer (%u)", a[2]);
                    if (pv[2] & 1u)
                        load_ext(a[0], p, r, pv[2] >> 8);
                    else
                        memcpy(r, p, 4u * m.types[a[0]].flat);
                    break;
                }
                case 62: /* OpStore */
                {
                    const uint32_t* pv = V(frame, a[0]);
                    if (pv[2] & 1u) return die("store to a buffer (%u): the shader only writes images", a[0]);
                    memcpy(get_ptr(pv), V(frame, a[1]), 4u * m.types[m.type_of[a[1]]].flat);
                    break;
                }
                case 65: /* OpAccessChain */
                {
                    const uint32_t* bv = V(frame, a[2]);
                    uint8_t* p = get_ptr(bv);
                    const bool ext = (bv[2] & 1u) != 0;
                    uint32_t mstride = bv[2] >> 8;
                    uint32_t tid = m.types[m.type_of[a[2]]].elem;
                    for (uint32_t k = 3; k < wc - 1; ++k)
                    {
                        const uint32_t idx = V(frame, a[k])[0];
                        const Type& t = m.types[tid];
                        switch (t.kind)
                        {
                            case K_STRUCT:
                            {
                                if (idx >= t.members.size()) return die("member index out of range");
                                if (ext)
                                    p += t.member_offset[idx], mstride = t.member_matrix_stride[idx];
                                else
                                {
                                    uint32_t o = 0;
                                    for (uint32_t q = 0; q < idx; ++q) o += m.types[t.members[q]].flat;
                                    p += 4u * o;
                                }
                                tid = t.members[idx];
                                break;
                            }
                            case K_ARRAY: case K_RTARRAY:
                                if (t.kind == K_ARRAY && idx >= t.count) return die("array index %u out of range (%u)", idx, t.count);
                                p += ext ? (size_t)idx * t.array_stride : (size_t)idx * 4u * m.types[t.elem].flat;
                                tid = t.elem;
                                break;
                            case K_MAT:
                                p += ext ? (size_t)idx * mstride : (size_t)idx * 4u * m.types[t.elem].flat;
                                tid = t.elem;
                                break;
                            case K_VEC:
                                if (idx >= t.count) return die("vector index out of range");
                                p += (size_t)idx * (m.types[t.elem].width / 8);
                                tid = t.elem;
                                break;
                            default: return die("access chain into a scalar");
                        }
                    }
                    put_ptr(V(frame, a[1]), p, (ext ? 1u : 0u) | (mstride << 8));
                    break;
                }
                case 68: /* OpArrayLength: structure pointer, member index */
                {
                    const uint32_t* bv = V(frame, a[2]);
                    const uint32_t sid = m.types[m.type_of[a[2]]].elem;
                    const Type& st = m.types[sid];
                    const Type& arr = m.types[st.members[a[3]]];
                    int binding = -1;
                    for (const Module::GVar& g : m.gvars)
                        if (g.id == a[2]) binding = g.binding;
                    if (binding < 0 || binding >= 8) return die("OpArrayLength on an unbound buffer");
                    (void)bv;
                    V(frame, a[1])[0] = (uint32_t)((bind.bytes[binding] - st.member_offset[a[3]]) / arr.array_stride);
                    break;
                }
                case 57: /* OpFunctionCall */
                {
                    const int callee = m.func_index[a[2]];
                    if (callee < 0) return die("call to unknown function %u", a[2]);
                    const Func& cf = m.funcs[callee];
                    const uint32_t cbase = base + fn.frame_words;
                    if ((size_t)cbase + cf.frame_words + 64 > arena.size()) return die("VM stack overflow");
                    if (cf.params.size() != wc - 4) return die("argument count mismatch calling %u", a[2]);
                    for (size_t k = 0; k < cf.params.size(); ++k)
                        memcpy(&arena[cbase + m.off[cf.params[k]]], V(frame, a[3 + k]), 4u * m.types[m.type_of[cf.params[k]]].flat);
                    uint32_t tmp[64];
                    if (m.types[a[0]].flat > 64) return die("return value too large");
                    if (!run(callee, cbase, tmp)) return false;
                    memcpy(V(frame, a[1]), tmp, 4u * m.types[a[0]].flat);
                    break;
                }
                case 12: /* OpExtInst */
                    if (a[2] != m.glsl_ext) return die("unknown extended instruction set");
                    if (!ext_inst(a[3], a[0], V(frame, a[1]), frame, a + 4)) return false;
                    break;
                case 79: /* OpVectorShuffle */
                {
                    uint32_t* r = V(frame, a[1]);
                    const uint32_t *x = V(frame, a[2]), *y = V(frame, a[3]);
                    const uint32_t nx = m.types[m.type_of[a[2]]].count;
                    if (m.types[m.types[a[0]].elem].width != 32) return die("64-bit shuffle");
                    uint32_t tmp[4];
                    for (uint32_t k = 4; k < wc - 1; ++k)
                        tmp[k - 4] = a[k] == 0xFFFFFFFFu ? 0u : (a[k] < nx ? x[a[k]] : y[a[k] - nx]);
                    memcpy(r, tmp, 4u * (wc - 5));
                    break;
                }
                case 80: /* OpCompositeConstruct */
                {
                    uint32_t tmp[64];
                    uint32_t o = 0;
                    for (uint32_t k = 2; k < wc - 1; ++k)
                    {
                        const uint32_t n = m.types[m.type_of[a[k]]].flat;
                        if (o + n > 64) return die("composite too large");
                        memcpy(tmp + o, V(frame, a[k]), 4u * n);
                        o += n;
                    }
                    if (o != m.types[a[0]].flat) return die("composite construct %u: size mismatch", a[1]);
                    memcpy(V(frame, a[1]), tmp, 4u * o);
                    break;
                }
                case 81: /* OpCompositeExtract (literal indices) */
                {
                    uint32_t tid = m.type_of[a[2]];
                    uint32_t o = 0;
                    for (uint32_t k = 3; k < wc - 1; ++k)
                    {
                        const Type& t = m.types[tid];
                        const uint32_t idx = a[k];
                        if (t.kind == K_STRUCT)
                        {
                            for (uint32_t q = 0; q < idx; ++q) o += m.types[t.members[q]].flat;
                            tid = t.members[idx];
                        }
                        else
                        {
                            o += idx * m.types[t.elem].flat;
                            tid = t.elem;
                        }
                    }
                    uint32_t tmp[64];
                    const uint32_t n = m.types[a[0]].flat;
                    memcpy(tmp, V(frame, a[2]) + o, 4u * n);
                    memcpy(V(frame, a[1]), tmp, 4u * n);
                    break;
                }
                case 98: /* OpImageRead: image, coordinate */
                {
                    const uint32_t* c = V(frame, a[3]);
                    if (!image_read(V(frame, a[2])[0], (int32_t)c[0], (int32_t)c[1], V(frame, a[1]))) return false;
                    break;
                }
                case 99: /* OpImageWrite: image, coordinate, texel */
                {
                    const uint32_t* c = V(frame, a[1]);
                    if (!image_write(V(frame, a[0])[0], (int32_t)c[0], (int32_t)c[1], V(frame, a[2]))) return false;
                    break;
                }
                case 104: /* OpImageQuerySize */
                {
                    const uint32_t h = V(frame, a[2])[0];
                    if (h >= 8) return die("bad image handle");
                    uint32_t* r = V(frame, a[1]);
                    r[0] = (uint32_t)bind.img[h].W, r[1] = (uint32_t)bind.img[h].H;
                    break;
                }
                default:
                    if (!alu(op, a, wc, frame)) return false;
                    break;
            }
            pc = next;
        }
    }

    bool alu(uint32_t op, const uint32_t* a, uint32_t wc, uint32_t* frame)
    {
        (void)wc;
        const Type& rt = m.types[a[0]];
        uint32_t* r = V(frame, a[1]);
        const uint32_t* x = V(frame, a[2]);
        const Type& xt = m.types[m.type_of[a[2]]];
        const uint32_t xw = (xt.kind == K_VEC || xt.kind == K_MAT) ? m.types[xt.kind == K_MAT ? m.types[xt.elem].elem : xt.elem].width : xt.width;
        const uint32_t rw = (rt.kind == K_VEC) ? m.types[rt.elem].width : rt.width;
        const uint32_t n = rt.kind == K_VEC ? rt.count : 1;
        const uint32_t* y = nullptr;
        switch (op)
        {
            case 110: for (uint32_t k = 0; k < n; ++k) r[k] = (uint32_t)(int32_t)asf(x[k]); return true;            /* ConvertFToS (rtz) */
            case 111: for (uint32_t k = 0; k < n; ++k) r[k] = asu((float)(int32_t)x[k]); return true;               /* ConvertSToF */
            case 112: for (uint32_t k = 0; k < n; ++k) r[k] = asu((float)x[k]); return true;                        /* ConvertUToF (rn) */
            case 115: /* FConvert */
                for (uint32_t k = 0; k < n; ++k)
                {
                    if (xw == 32 && rw == 64) putd(r + 2 * k, (double)asf(x[k]));
                    else if (xw == 64 && rw == 32) r[k] = asu((float)asd(x + 2 * k));
                    else return die("FConvert %u -> %u bits", xw, rw);
                }
                return true;
            case 124: memcpy(r, x, 4u * rt.flat); return true;                                                       /* Bitcast */
            case 127: if (rw != 32) return die("f64 negate"); for (uint32_t k = 0; k < n; ++k) r[k] = asu(-asf(x[k])); return true;
            case 168: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] ? 0u : 1u; return true;                           /* LogicalNot */
            default: break;
        }
        y = V(frame, a[3]);
        const bool f64 = xw == 64;
        switch (op)
        {
            case 128: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] + y[k]; return true;
            case 130: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] - y[k]; return true;
            case 132: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] * y[k]; return true;
            case 129: case 131: case 133: case 136:
                if (f64) return die("f64 arithmetic (opcode %u)", op);
                for (uint32_t k = 0; k < n; ++k)
                {
                    const float p = asf(x[k]), q = asf(y[k]);
                    /* volatile: every operation separately rounded, no contraction across iterations */
                    volatile float v = op == 129 ? p + q : op == 131 ? p - q : op == 133 ? p * q : p / q;
                    r[k] = asu(v);
                }
                return true;
            case 142: /* VectorTimesScalar */
                for (uint32_t k = 0; k < n; ++k) { volatile float v = asf(x[k]) * asf(y[0]); r[k] = asu(v); }
                return true;
            case 145: /* MatrixTimesVector: sum of column * component, left to right */
            {
                const uint32_t cols = xt.count, rows = m.types[xt.elem].count;
                uint32_t tmp[4];
                for (uint32_t i = 0; i < rows; ++i)
                {
                    volatile float acc = asf(x[i]) * asf(y[0]);
                    for (uint32_t c = 1; c < cols; ++c)
                    {
                        volatile float p = asf(x[c * rows + i]) * asf(y[c]);
                        acc = acc + p;
                    }
                    tmp[i] = asu(acc);
                }
                memcpy(r, tmp, 4u * rows);
                return true;
            }
            case 148: /* Dot: left to right */
            {
                volatile float acc = asf(x[0]) * asf(y[0]);
                for (uint32_t k = 1; k < xt.count; ++k)
                {
                    volatile float p = asf(x[k]) * asf(y[k]);
                    acc = acc + p;
                }
                r[0] = asu(acc);
                return true;
            }
            case 166: for (uint32_t k = 0; k < n; ++k) r[k] = (x[k] || y[k]) ? 1u : 0u; return true;
            case 167: for (uint32_t k = 0; k < n; ++k) r[k] = (x[k] && y[k]) ? 1u : 0u; return true;
            case 169: /* Select: condition, a, b */
            {
                const uint32_t* z = V(frame, a[4]);
                const uint32_t per = rt.flat / n;
                uint32_t tmp[8];
                const bool scalar_cond = m.types[m.type_of[a[2]]].kind == K_BOOL;
                for (uint32_t k = 0; k < n; ++k)
                    for (uint32_t q = 0; q < per; ++q)
                        tmp[k * per + q] = (x[scalar_cond ? 0 : k] ? y : z)[k * per + q];
                memcpy(r, tmp, 4u * rt.flat);
                return true;
            }
            case 171: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] != y[k]; return true;
            case 172: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] > y[k]; return true;
            case 176: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] < y[k]; return true;
            case 177: for (uint32_t k = 0; k < n; ++k) r[k] = (int32_t)x[k] < (int32_t)y[k]; return true;
            case 184: case 186: case 188: case 190:
                for (uint32_t k = 0; k < n; ++k)
                {
                    bool v;
                    if (f64)
                    {
                        const double p = asd(x + 2 * k), q = asd(y + 2 * k);
                        v = op == 184 ? p < q : op == 186 ? p > q : op == 188 ? p <= q : p >= q;
                    }
                    else
                    {
                        const float p = asf(x[k]), q = asf(y[k]);
                        v = op == 184 ? p < q : op == 186 ? p > q : op == 188 ? p <= q : p >= q;
                    }
                    r[k] = v ? 1u : 0u;
                }
                return true;
            case 194: for (uint32_t k = 0; k < n; ++k) r[k] = y[k] < 32 ? x[k] >> y[k] : 0u; return true;
            case 196: for (uint32_t k = 0; k < n; ++k) r[k] = y[k] < 32 ? x[k] << y[k] : 0u; return true;
            case 198: for (uint32_t k = 0; k < n; ++k) r[k] = x[k] ^ y[k]; return true;
            default: return die("unsupported opcode %u", op);
        }
    }

    bool invoke(uint32_t gx, uint32_t gy, uint32_t gz)
    {
        for (const Module::GVar& g : m.gvars)
            if (g.builtin == 28) /* GlobalInvocationId */
                priv[g.store_off] = gx, priv[g.store_off + 1] = gy, priv[g.store_off + 2] = gz;
            else if (g.storage == 1)
                return die("unsupported input builtin %u", (uint32_t)g.builtin);
        uint32_t ret[4];
        return run(m.func_index[m.entry], 0, ret);
    }
};

} /* namespace */

/* ------------------------------------------------------------------------------------------ */
/* C API (ctypes: oracle/spirv_vm.py)                                                            */
/* ------------------------------------------------------------------------------------------ */
struct rvpt_spirv_vm
{
    Module mod;
    Bindings bind;
    std::string err;
    uint64_t executed = 0;
};

extern "C" {

__attribute__((visibility("default"))) rvpt_spirv_vm* rvpt_spirv_vm_create(const uint32_t* words, size_t n_words)
{
    rvpt_spirv_vm* vm = new rvpt_spirv_vm();
    vm->mod.w.assign(words, words + n_words);
    if (!parse(vm->mod)) vm->err = vm->mod.err;
    return vm;
}

__attribute__((visibility("default"))) void rvpt_spirv_vm_destroy(rvpt_spirv_vm* vm) { delete vm; }

__attribute__((visibility("default"))) const char* rvpt_spirv_vm_error(const rvpt_spirv_vm* vm) { return vm->err.c_str(); }

__attribute__((visibility("default"))) void rvpt_spirv_vm_bind_buffer(rvpt_spirv_vm* vm, int binding, const void* data, size_t bytes)
{
    if (binding < 0 || binding >= 8) return;
    vm->bind.buf[binding] = (const uint8_t*)data, vm->bind.bytes[binding] = bytes;
}

/* unorm8 != 0: `data` is W*H*4 bytes with UNORM8 load/store conversion (the reference's image
 * format, rvpt.cpp:759-766, 803-811); else W*H*4 floats stored as they are */
__attribute__((visibility("default"))) void rvpt_spirv_vm_bind_image(rvpt_spirv_vm* vm, int binding, void* data, int W, int H, int unorm8)
{
    if (binding < 0 || binding >= 8) return;
    Image& im = vm->bind.img[binding];
    im.W = W, im.H = H;
    im.u8 = unorm8 ? (uint8_t*)data : nullptr;
    im.f32 = unorm8 ? nullptr : (float*)data;
}

/* Runs the entry point for every invocation (x, y, 0) with x in [x0, x1), y in [y0, y1) — the
 * reference dispatches ceil-free W/16 x H/16 groups of 16x16 (rvpt.cpp:1035-1036), the caller
 * passes that extent. Rows are spread over `nthreads` threads (0 = hardware concurrency); every
 * invocation only touches its own texel of each image. Returns 0, or -1 with an error message. */
__attribute__((visibility("default"))) int rvpt_spirv_vm_dispatch(rvpt_spirv_vm* vm, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, int nthreads)
{
    if (!vm->err.empty()) return -1;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    const uint32_t rows = y1 > y0 ? y1 - y0 : 0;
    if ((uint32_t)nthreads > rows) nthreads = rows ? (int)rows : 1;
    std::vector<std::string> errs(nthreads);
    std::vector<uint64_t> counts(nthreads, 0);
    auto work = [&](int t) {
        Machine mach(vm->mod, vm->bind);
        for (uint32_t y = y0 + (uint32_t)t; y < y1; y += (uint32_t)nthreads)
            for (uint32_t x = x0; x < x1; ++x)
                if (!mach.invoke(x, y, 0))
                {
                    errs[t] = mach.err;
                    return;
                }
        counts[t] = mach.executed;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    for (int t = 0; t < nthreads; ++t)
    {
        vm->executed += counts[t];
        if (!errs[t].empty())
        {
            vm->err = errs[t];
            return -1;
        }
    }
    return 0;
}

__attribute__((visibility("default"))) uint64_t rvpt_spirv_vm_executed(const rvpt_spirv_vm* vm) { return vm->executed; }

/* Reflection for layout tests: byte offset of member `member` of the struct type named `name`
 * (OpName), its array stride when it is the element of a runtime array; -1 if unknown. */
__attribute__((visibility("default"))) int rvpt_spirv_vm_member_offset(const rvpt_spirv_vm* vm, uint32_t type_id, uint32_t member)
{
    if (type_id >= vm->mod.bound) return -1;
    const Type& t = vm->mod.types[type_id];
    if (t.kind != K_STRUCT || member >= t.member_offset.size()) return -1;
    return (int)t.member_offset[member];
}

} /* extern "C" */
