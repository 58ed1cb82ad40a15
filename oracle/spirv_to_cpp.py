"""SPIR-V -> C++ translator for the reference's SHIPPED compute shader (TEST INFRASTRUCTURE).

    python oracle/spirv_to_cpp.py [compute_pass.comp.spv] [out.cpp]

Where `oracle/spirv_vm.cpp` interprets `assets/shaders/compute_pass.comp.spv`, this turns the same
binary into straight-line C++ — one C++ function per SPIR-V function, one word array per result
id, `goto` per branch — which `oracle/Makefile` (`make ref_shader`) compiles into
`oracle/_ref/libref_shader.so`: the reference's own implementation of the hot path, built for the
host CPU from the artefact the reference ships (what lavapipe would do with LLVM). Generated source
and library live only under `oracle/_ref/` (git-ignored: they are a translation of the reference's
binary, not this repo's code); the library travels to the GPU box.

The value model is the interpreter's (every value a run of 32-bit words; f64 two words; a pointer
is {host address, tag}; buffer loads follow the Offset / ArrayStride / MatrixStride decorations)
and the driver-defined operations (dot / matrix product summation order, sin / cos / tan /
normalize / length / cross / mix, UNORM8 image conversion) come from the same runtime header,
`oracle/spirv_rt.h`, i.e. from `include/rvpt_math.h`. tests/test_spirv_pin.py holds the
translation to the interpreter and to the oracle bit for bit.
"""
from __future__ import annotations

import struct
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEFAULT_SPV = Path("/root/reference/assets/shaders/compute_pass.comp.spv")

VOID, BOOL, INT, FLOAT, VEC, MAT, IMAGE, ARRAY, RTARRAY, STRUCT, PTR, FUNC = range(12)


class Type:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.width = kw.get("width", 32)
        self.elem = kw.get("elem", 0)
        self.count = kw.get("count", 0)
        self.storage = kw.get("storage", 0)
        self.members = kw.get("members", [])
        self.member_offset = {}
        self.member_mstride = {}
        self.array_stride = 0


class Module:
    def __init__(self, path: Path):
        data = path.read_bytes()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        assert w[0] == 0x07230203
        self.bound = w[3]
        self.ins = []
        i = 5
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            self.ins.append((op, list(w[i + 1:i + wc])))
            i += wc
        self.types: dict[int, Type] = {}
        self.type_of: dict[int, int] = {}
        self.const_words: dict[int, list[int]] = {}
        self.binding, self.builtin = {}, {}
        self.glsl = 0
        self.entry = 0
        self.gvars = []  # (id, ptr type, storage)
        pending_deco = []
        for op, a in self.ins:
            if op == 11 and b"GLSL.std.450" in b"".join(struct.pack("<I", x) for x in a[1:]):
                self.glsl = a[0]
            elif op == 15:
                self.entry = a[1]
            elif op == 71:
                pending_deco.append(("d", a))
            elif op == 72:
                pending_deco.append(("m", a))
        cur_fn = None
        for op, a in self.ins:
            if op == 19:
                self.types[a[0]] = Type(VOID)
            elif op == 20:
                self.types[a[0]] = Type(BOOL)
            elif op == 21:
                self.types[a[0]] = Type(INT, width=a[1])
            elif op == 22:
                self.types[a[0]] = Type(FLOAT, width=a[1])
            elif op == 23:
                self.types[a[0]] = Type(VEC, elem=a[1], count=a[2])
            elif op == 24:
                self.types[a[0]] = Type(MAT, elem=a[1], count=a[2])
            elif op == 25:
                self.types[a[0]] = Type(IMAGE)
            elif op == 28:
                self.types[a[0]] = Type(ARRAY, elem=a[1], count=self.const_words[a[2]][0])
            elif op == 29:
                self.types[a[0]] = Type(RTARRAY, elem=a[1])
            elif op == 30:
                self.types[a[0]] = Type(STRUCT, members=a[1:])
            elif op == 32:
                self.types[a[0]] = Type(PTR, storage=a[1], elem=a[2])
            elif op == 33:
                self.types[a[0]] = Type(FUNC)
            elif op in (41, 42):
                self.type_of[a[1]] = a[0]
                self.const_words[a[1]] = [1 if op == 41 else 0]
            elif op == 43:
                self.type_of[a[1]] = a[0]
                self.const_words[a[1]] = list(a[2:])
            elif op == 44:
                self.type_of[a[1]] = a[0]
                ws = []
                for c in a[2:]:
                    ws += self.const_words[c]
                self.const_words[a[1]] = ws
            elif op == 54:
                cur_fn = a[1]
            elif op == 56:
                cur_fn = None
            elif op == 59 and cur_fn is None:
                self.type_of[a[1]] = a[0]
                self.gvars.append((a[1], a[0], a[2]))
        for kind, a in pending_deco:
            if kind == "d":
                if a[1] == 33:
                    self.binding[a[0]] = a[2]
                elif a[1] == 11:
                    self.builtin[a[0]] = a[2]
                elif a[1] == 6 and a[0] in self.types:
                    self.types[a[0]].array_stride = a[2]
            else:
                t = self.types.get(a[0])
                if t is None:
                    continue
                if a[2] == 35:
                    t.member_offset[a[1]] = a[3]
                elif a[2] == 7:
                    t.member_mstride[a[1]] = a[3]

    def flat(self, tid: int) -> int:
        t = self.types[tid]
        if t.kind in (VOID, FUNC, RTARRAY):
            return 0
        if t.kind in (BOOL, INT, IMAGE):
            return 1
        if t.kind == FLOAT:
            return t.width // 32
        if t.kind in (VEC, MAT, ARRAY):
            return t.count * self.flat(t.elem)
        if t.kind == STRUCT:
            return sum(self.flat(m) for m in t.members)
        return 0  # pointers are C++ Ptr objects, not words

    def scalar_width(self, tid: int) -> int:
        t = self.types[tid]
        while t.kind in (VEC, MAT, ARRAY):
            t = self.types[t.elem]
        return t.width

    def is_ptr(self, tid: int) -> bool:
        return self.types[tid].kind == PTR


ARITH = {129: "+", 131: "-", 133: "*", 136: "/"}
ICMP = {171: ("!=", "u"), 172: (">", "u"), 176: ("<", "u"), 177: ("<", "s")}
FCMP = {184: "<", 186: ">", 188: "<=", 190: ">="}


class Gen:
    def __init__(self, m: Module):
        self.m = m
        self.out: list[str] = []
        self.ext_loaders: dict[int, str] = {}
        self.gvar_ids = {g[0] for g in m.gvars}

    def emit(self, s: str = ""):
        self.out.append(s)

    def val(self, i: int) -> str:
        """expression of type `const uint32_t*` (or Ptr for pointer-typed ids)"""
        if i in self.m.const_words:
            return f"k{i}"
        if i in self.gvar_ids:
            return f"cx.g[{i}]"
        return f"v{i}"

    # ---- buffer loaders following the layout decorations ----
    def ext_loader(self, tid: int) -> str:
        if tid in self.ext_loaders:
            return self.ext_loaders[tid]
        m, t = self.m, self.m.types[tid]
        name = f"ldx{tid}"
        self.ext_loaders[tid] = name
        body = []
        if t.kind in (BOOL, INT, FLOAT, VEC):
            body.append(f"std::memcpy(o, p, {4 * m.flat(tid)});")
        elif t.kind == MAT:
            cf = m.flat(t.elem)
            for c in range(t.count):
                body.append(f"std::memcpy(o + {c * cf}, p + {c} * ms, {4 * cf});")
        elif t.kind == ARRAY:
            ef = m.flat(t.elem)
            sub = self.ext_loader(t.elem)
            body.append(f"for (uint32_t k = 0; k < {t.count}; ++k) {sub}(p + k * {t.array_stride}u, ms, o + k * {ef});")
        elif t.kind == STRUCT:
            off = 0
            for k, mt in enumerate(t.members):
                sub = self.ext_loader(mt)
                body.append(f"{sub}(p + {t.member_offset.get(k, 0)}, {t.member_mstride.get(k, 0)}u, o + {off});")
                off += m.flat(mt)
        else:
            raise NotImplementedError(f"buffer load of type {tid}")
        self.loader_defs.append(f"static inline void {name}(const uint8_t* p, uint32_t ms, uint32_t* o) {{ (void)ms; "
                                + " ".join(body) + " }")
        return name

    def generate(self) -> str:
        m = self.m
        self.loader_defs: list[str] = []
        fn_bodies: list[str] = []
        protos: list[str] = []
        # constants
        consts = [f"static const uint32_t k{i}[{max(len(ws), 1)}] = {{{', '.join(hex(x) for x in ws) or '0'}}};"
                  for i, ws in sorted(m.const_words.items())]
        # private / input / image globals live in the per-invocation context
        priv_off, priv_words = {}, 0
        for gid, ptype, sc in m.gvars:
            if sc in (6, 1, 0):
                priv_off[gid] = priv_words
                priv_words += max(m.flat(m.types[ptype].elem), 1)
        self.priv_off = priv_off
        # functions
        fn = None
        for idx, (op, a) in enumerate(m.ins):
            if op == 54:
                fn = {"id": a[1], "ret": a[0], "params": [], "body": [], "locals": [], "phis": {}, "labels": []}
            elif op == 56:
                fn_bodies.append(self.function(fn))
                protos.append(self.proto(fn) + ";")
                fn = None
            elif fn is not None:
                fn["body"].append((op, a))
        head = [
            "// GENERATED by oracle/spirv_to_cpp.py from the reference's compute_pass.comp.spv — do not commit.",
            '#include "../spirv_rt.h"', "namespace {",
            f"constexpr uint32_t kPrivWords = {priv_words};",
        ] + consts
        init = ["static void bind_globals(Ctx& cx) {"]
        for gid, ptype, sc in m.gvars:
            if sc in (6, 1, 0):
                init.append(f"    cx.g[{gid}] = Ptr{{reinterpret_cast<uint8_t*>(cx.priv + {priv_off[gid]}), 0u}};")
                if sc == 0:
                    init.append(f"    cx.priv[{priv_off[gid]}] = {m.binding.get(gid, 0)}u;")
            else:
                b = m.binding.get(gid, -1)
                init.append(f"    cx.g[{gid}] = Ptr{{const_cast<uint8_t*>(cx.bind->buf[{b}]), 1u}};")
        init.append("}")
        builtin = ["static void set_invocation(Ctx& cx, uint32_t x, uint32_t y, uint32_t z) {"]
        for gid, ptype, sc in m.gvars:
            if m.builtin.get(gid) == 28:
                builtin.append(f"    cx.priv[{priv_off[gid]}] = x; cx.priv[{priv_off[gid] + 1}] = y; cx.priv[{priv_off[gid] + 2}] = z;")
        builtin.append("}")
        tail = ["} // namespace", "",
                f'extern "C" __attribute__((visibility("default"))) void ref_shader_invoke(Ctx* cx, uint32_t x, uint32_t y)',
                "{", "    set_invocation(*cx, x, y, 0);", f"    f{m.entry}(*cx);", "}",
                'extern "C" __attribute__((visibility("default"))) void ref_shader_bind(Ctx* cx) { bind_globals(*cx); }',
                f'extern "C" __attribute__((visibility("default"))) uint32_t ref_shader_max_id() {{ return {m.bound}; }}']
        return "\n".join(head + protos + self.loader_defs + init + builtin + fn_bodies + tail) + "\n"

    def proto(self, fn) -> str:
        m = self.m
        ps = ["Ctx& cx"]
        for pid, ptid in fn["param_list"]:
            ps.append(f"Ptr v{pid}" if m.is_ptr(ptid) else f"const uint32_t* v{pid}")
        if m.flat(fn["ret"]):
            ps.append("uint32_t* ret")
        return f"static void f{fn['id']}({', '.join(ps)})"

    def function(self, fn) -> str:
        m = self.m
        lines: list[str] = []
        decl: list[str] = []
        fn["param_list"] = []
        # phi sources: target label -> [(phi id, {pred label: value id})]
        phis: dict[int, list] = {}
        cur = None
        for op, a in fn["body"]:
            if op == 248:
                cur = a[0]
            elif op == 245:
                phis.setdefault(cur, []).append((a[1], a[0], {a[k + 1]: a[k] for k in range(2, len(a), 2)}))

        def jump(src_label, target):
            s = ""
            for pid, ptid, srcs in phis.get(target, []):
                n = m.flat(ptid)
                s += f"std::memcpy(v{pid}, {self.val(srcs[src_label])}, {4 * n}); "
            return s + f"goto L{target};"

        cur = None
        for op, a in fn["body"]:
            if op == 55:
                fn["param_list"].append((a[1], a[0]))
                m.type_of[a[1]] = a[0]
                continue
            if op == 248:
                cur = a[0]
                lines.append(f"L{a[0]}:;")
                continue
            if op in (246, 247):
                continue
            rid = rtype = None
            if op in (12, 57, 59, 61, 65, 68, 79, 80, 81, 98, 104, 110, 111, 112, 115, 124, 127, 128, 129, 130, 131, 132, 133,
                      136, 142, 145, 148, 166, 167, 168, 169, 171, 172, 176, 177, 184, 186, 188, 190, 194, 196, 198, 245):
                rtype, rid = a[0], a[1]
                m.type_of[rid] = rtype
                if m.is_ptr(rtype):
                    decl.append(f"Ptr v{rid};")
                else:
                    decl.append(f"uint32_t v{rid}[{max(m.flat(rtype), 1)}];")
            n = m.flat(rtype) if rtype is not None and not m.is_ptr(rtype) else 0
            V = self.val
            if op == 245:
                pass  # assigned on the incoming edges
            elif op == 249:
                lines.append(jump(cur, a[0]))
            elif op == 250:
                lines.append(f"if ({V(a[0])}[0]) {{ {jump(cur, a[1])} }} else {{ {jump(cur, a[2])} }}")
            elif op == 251:
                s = f"switch ({V(a[0])}[0]) {{ "
                for k in range(2, len(a), 2):
                    s += f"case {a[k]}u: {{ {jump(cur, a[k + 1])} }} "
                s += f"default: {{ {jump(cur, a[1])} }} }}"
                lines.append(s)
            elif op == 253:
                lines.append("return;")
            elif op == 254:
                lines.append(f"std::memcpy(ret, {V(a[0])}, {4 * m.flat(m.type_of[a[0]])}); return;")
            elif op == 255:
                lines.append("rt_unreachable();")
            elif op == 59:
                pointee = m.types[rtype].elem
                decl.append(f"uint32_t s{rid}[{max(m.flat(pointee), 1)}];")
                lines.append(f"v{rid} = Ptr{{reinterpret_cast<uint8_t*>(s{rid}), 0u}};")
            elif op == 61:
                ptid = m.type_of[a[2]]
                src = V(a[2])
                if m.types[ptid].storage == 2:
                    lines.append(f"{self.ext_loader(rtype)}({src}.p, {src}.tag >> 8, v{rid});")
                else:
                    lines.append(f"std::memcpy(v{rid}, {src}.p, {4 * n});")
            elif op == 62:
                dst = V(a[0])
                lines.append(f"std::memcpy({dst}.p, {V(a[1])}, {4 * m.flat(m.type_of[a[1]])});")
            elif op == 65:
                base = V(a[2])
                ptid = m.type_of[a[2]]
                ext = m.types[ptid].storage == 2
                tid = m.types[ptid].elem
                const_off, dyn, mstride = 0, [], None
                for ix in a[3:]:
                    t = m.types[tid]
                    cidx = m.const_words[ix][0] if ix in m.const_words else None
                    if t.kind == STRUCT:
                        assert cidx is not None
                        if ext:
                            const_off += t.member_offset.get(cidx, 0)
                            mstride = t.member_mstride.get(cidx, 0)
                        else:
                            const_off += 4 * sum(m.flat(x) for x in t.members[:cidx])
                        tid = t.members[cidx]
                    else:
                        if t.kind in (ARRAY, RTARRAY):
                            stride = t.array_stride if ext else 4 * m.flat(t.elem)
                        elif t.kind == MAT:
                            stride = None if ext else 4 * m.flat(t.elem)
                        else:  # vector component
                            stride = m.types[t.elem].width // 8
                        if cidx is not None and stride is not None:
                            const_off += cidx * stride
                        elif stride is not None:
                            dyn.append(f"(size_t){V(ix)}[0] * {stride}u")
                        else:  # column of a buffer matrix: runtime matrix stride carried by the pointer tag
                            ms = f"{mstride}u" if mstride is not None else f"({base}.tag >> 8)"
                            dyn.append(f"(size_t){V(ix)}[0] * {ms}")
                        tid = t.elem
                off = " + ".join([str(const_off)] + dyn)
                tag = f"(1u | ({mstride}u << 8))" if (ext and mstride is not None) else f"{base}.tag"
                lines.append(f"v{rid} = Ptr{{{base}.p + {off}, {tag}}};")
            elif op == 68:
                sid = m.types[m.type_of[a[2]]].elem
                st = m.types[sid]
                arr = m.types[st.members[a[3]]]
                b = m.binding.get(a[2], 0)
                lines.append(f"v{rid}[0] = (uint32_t)((cx.bind->bytes[{b}] - {st.member_offset.get(a[3], 0)}u) / {arr.array_stride}u);")
            elif op == 57:
                callee = a[2]
                args = ["cx"] + [V(x) for x in a[3:]]
                if m.flat(rtype):
                    args.append(f"v{rid}")
                lines.append(f"f{callee}({', '.join(args)});")
            elif op == 12:
                inst = a[3]
                ops = a[4:]
                w = m.scalar_width(rtype)
                cnt = n // (w // 32)
                xn = m.flat(m.type_of[ops[0]])
                args = ", ".join(V(x) for x in ops)
                lines.append(f"rt_ext{'64' if w == 64 else ''}<{inst}, {cnt}, {xn}>(v{rid}, {args});")
            elif op == 79:
                nx = m.types[m.type_of[a[2]]].count
                for k, c in enumerate(a[4:]):
                    src = "0u" if c == 0xFFFFFFFF else (f"{V(a[2])}[{c}]" if c < nx else f"{V(a[3])}[{c - nx}]")
                    lines.append(f"t_[{k}] = {src};")
                lines.append(f"std::memcpy(v{rid}, t_, {4 * (len(a) - 4)});")
            elif op == 80:
                off = 0
                for x in a[2:]:
                    k = m.flat(m.type_of[x])
                    lines.append(f"std::memcpy(t_ + {off}, {V(x)}, {4 * k});")
                    off += k
                lines.append(f"std::memcpy(v{rid}, t_, {4 * off});")
            elif op == 81:
                tid, off = m.type_of[a[2]], 0
                for ix in a[3:]:
                    t = m.types[tid]
                    if t.kind == STRUCT:
                        off += sum(m.flat(x) for x in t.members[:ix])
                        tid = t.members[ix]
                    else:
                        off += ix * m.flat(t.elem)
                        tid = t.elem
                lines.append(f"std::memcpy(t_, {V(a[2])} + {off}, {4 * n}); std::memcpy(v{rid}, t_, {4 * n});")
            elif op == 98:
                lines.append(f"rt_image_read(cx, {V(a[2])}[0], (int32_t){V(a[3])}[0], (int32_t){V(a[3])}[1], v{rid});")
            elif op == 99:
                lines.append(f"rt_image_write(cx, {V(a[0])}[0], (int32_t){V(a[1])}[0], (int32_t){V(a[1])}[1], {V(a[2])});")
            elif op == 104:
                lines.append(f"rt_image_size(cx, {V(a[2])}[0], v{rid});")
            elif op in (110, 111, 112, 127, 168):
                f = {110: "(uint32_t)(int32_t)asf({x})", 111: "asu((float)(int32_t){x})", 112: "asu((float){x})",
                     127: "asu(-asf({x}))", 168: "({x} ? 0u : 1u)"}[op]
                for k in range(n):
                    lines.append(f"v{rid}[{k}] = " + f.format(x=f"{V(a[2])}[{k}]") + ";")
            elif op == 115:
                xw, rw = m.scalar_width(m.type_of[a[2]]), m.scalar_width(rtype)
                cnt = n // (rw // 32)
                for k in range(cnt):
                    if xw == 32 and rw == 64:
                        lines.append(f"putd(v{rid} + {2 * k}, (double)asf({V(a[2])}[{k}]));")
                    else:
                        lines.append(f"v{rid}[{k}] = asu((float)asd({V(a[2])} + {2 * k}));")
            elif op == 124:
                lines.append(f"std::memcpy(v{rid}, {V(a[2])}, {4 * n});")
            elif op in (128, 130, 132, 194, 196, 198):
                sym = {128: "+", 130: "-", 132: "*", 198: "^"}.get(op)
                for k in range(n):
                    x, y = f"{V(a[2])}[{k}]", f"{V(a[3])}[{k}]"
                    if op == 194:
                        lines.append(f"v{rid}[{k}] = {y} < 32u ? {x} >> {y} : 0u;")
                    elif op == 196:
                        lines.append(f"v{rid}[{k}] = {y} < 32u ? {x} << {y} : 0u;")
                    else:
                        lines.append(f"v{rid}[{k}] = {x} {sym} {y};")
            elif op in ARITH:
                assert m.scalar_width(rtype) == 32
                for k in range(n):
                    lines.append(f"v{rid}[{k}] = asu(asf({V(a[2])}[{k}]) {ARITH[op]} asf({V(a[3])}[{k}]));")
            elif op == 142:
                for k in range(n):
                    lines.append(f"v{rid}[{k}] = asu(asf({V(a[2])}[{k}]) * asf({V(a[3])}[0]));")
            elif op == 145:
                xt = m.types[m.type_of[a[2]]]
                lines.append(f"rt_mat_vec<{xt.count}, {m.types[xt.elem].count}>(v{rid}, {V(a[2])}, {V(a[3])});")
            elif op == 148:
                lines.append(f"rt_dot<{m.types[m.type_of[a[2]]].count}>(v{rid}, {V(a[2])}, {V(a[3])});")
            elif op in (166, 167):
                sym = "||" if op == 166 else "&&"
                for k in range(n):
                    lines.append(f"v{rid}[{k}] = ({V(a[2])}[{k}] {sym} {V(a[3])}[{k}]) ? 1u : 0u;")
            elif op == 169:
                cnt = m.types[rtype].count if m.types[rtype].kind == VEC else 1
                per = n // cnt
                scalar_cond = m.types[m.type_of[a[2]]].kind == BOOL
                for k in range(cnt):
                    for q in range(per):
                        c = f"{V(a[2])}[{0 if scalar_cond else k}]"
                        lines.append(f"t_[{k * per + q}] = {c} ? {V(a[3])}[{k * per + q}] : {V(a[4])}[{k * per + q}];")
                lines.append(f"std::memcpy(v{rid}, t_, {4 * n});")
            elif op in ICMP:
                sym, sign = ICMP[op]
                cast = "(int32_t)" if sign == "s" else ""
                for k in range(n):
                    lines.append(f"v{rid}[{k}] = ({cast}{V(a[2])}[{k}] {sym} {cast}{V(a[3])}[{k}]) ? 1u : 0u;")
            elif op in FCMP:
                xw = m.scalar_width(m.type_of[a[2]])
                for k in range(n):
                    if xw == 64:
                        lines.append(f"v{rid}[{k}] = (asd({V(a[2])} + {2 * k}) {FCMP[op]} asd({V(a[3])} + {2 * k})) ? 1u : 0u;")
                    else:
                        lines.append(f"v{rid}[{k}] = (asf({V(a[2])}[{k}]) {FCMP[op]} asf({V(a[3])}[{k}])) ? 1u : 0u;")
            else:
                raise NotImplementedError(f"opcode {op}")
        body = "\n    ".join(decl + ["uint32_t t_[64]; (void)t_;"] + lines)
        return f"{self.proto(fn)}\n{{\n    {body}\n}}\n"


def main():
    spv = Path(sys.argv[1]) if len(sys.argv) > 1 else DEFAULT_SPV
    out = Path(sys.argv[2]) if len(sys.argv) > 2 else HERE / "_ref" / "compute_pass_gen.cpp"
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(Gen(Module(spv)).generate())
    print(out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
