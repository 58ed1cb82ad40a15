"""A small SPIR-V reader (pure Python) for the parity tests: names, decorations, types and the
instruction stream of the reference's shipped compute_pass.comp.spv. Reads the binary; nothing
here executes it (oracle/spirv_vm.cpp does that)."""
from __future__ import annotations

import struct
from pathlib import Path

DEC_ARRAY_STRIDE, DEC_MATRIX_STRIDE, DEC_BUILTIN, DEC_BINDING, DEC_SET, DEC_OFFSET = 6, 7, 11, 33, 34, 35


def _string(words) -> str:
    raw = b"".join(struct.pack("<I", w) for w in words)
    return raw.split(b"\0")[0].decode()


class SpirvModule:
    def __init__(self, path: Path):
        data = Path(path).read_bytes()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        assert w[0] == 0x07230203, "not SPIR-V"
        self.version, self.bound = w[1], w[3]
        self.instructions: list[tuple[int, tuple[int, ...]]] = []
        i = 5
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            self.instructions.append((op, w[i + 1:i + wc]))
            i += wc
        self.names, self.member_names = {}, {}
        self.decorations, self.member_decorations = {}, {}
        self.types, self.constants = {}, {}
        for op, a in self.instructions:
            if op == 5:
                self.names[a[0]] = _string(a[1:])
            elif op == 6:
                self.member_names[(a[0], a[1])] = _string(a[2:])
            elif op == 71:
                self.decorations.setdefault(a[0], {})[a[1]] = a[2:]
            elif op == 72:
                self.member_decorations.setdefault((a[0], a[1]), {})[a[2]] = a[3:]
            elif op in (19, 20, 21, 22, 23, 24, 25, 28, 29, 30, 32, 33):
                self.types[a[0]] = (op, a[1:])
            elif op == 43:
                self.constants[a[1]] = (a[0], a[2:])

    def ids_named(self, name: str) -> list[int]:
        return [i for i, n in self.names.items() if n == name]

    def struct_layout(self, name: str) -> dict:
        """{member name: byte offset} of the struct type `name` that carries Offset decorations,
        plus '__stride__' when a runtime array of it is declared (ArrayStride)."""
        for sid in self.ids_named(name):
            op, members = self.types.get(sid, (0, ()))
            if op != 30 or (sid, 0) not in self.member_decorations:
                continue
            out = {}
            for k in range(len(members)):
                dec = self.member_decorations[(sid, k)]
                out[self.member_names[(sid, k)]] = dec[DEC_OFFSET][0]
                if DEC_MATRIX_STRIDE in dec:
                    out[self.member_names[(sid, k)] + "__matrix_stride__"] = dec[DEC_MATRIX_STRIDE][0]
            for tid, (top, targs) in self.types.items():
                if top in (28, 29) and targs[0] == sid and DEC_ARRAY_STRIDE in self.decorations.get(tid, {}):
                    out["__stride__"] = self.decorations[tid][DEC_ARRAY_STRIDE][0]
            return out
        raise KeyError(name)

    def binding_of(self, variable_name: str) -> int:
        (vid,) = self.ids_named(variable_name)
        return self.decorations[vid][DEC_BINDING][0]

    def bindings(self) -> dict:
        """{binding: (variable id, pointee type id)} of descriptor set 0."""
        out = {}
        for op, a in self.instructions:
            if op == 59 and DEC_BINDING in self.decorations.get(a[1], {}):
                out[self.decorations[a[1]][DEC_BINDING][0]] = (a[1], self.types[a[0]][1][1])
        return out

    def function_body(self, name_prefix: str) -> list[tuple[int, tuple[int, ...]]]:
        """Instructions of the function whose OpName starts with `name_prefix`."""
        fid = next(i for i, n in self.names.items() if n.startswith(name_prefix) and
                   any(op == 54 and a[1] == i for op, a in self.instructions))
        body, inside = [], False
        for op, a in self.instructions:
            if op == 54:
                inside = a[1] == fid
            elif op == 56 and inside:
                break
            elif inside:
                body.append((op, a))
        return body

    def float_constant(self, cid: int) -> tuple[int, int]:
        """(bit width, raw bits) of a float constant."""
        tid, words = self.constants[cid]
        width = self.types[tid][1][0]
        bits = words[0] if width == 32 else words[0] | (words[1] << 32)
        return width, bits
