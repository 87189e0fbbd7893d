"""Host-side tape front end (Python face of ``csrc/host``).

Mirrors the reference's host-side objects that stay on the CPU for every
backend: ``Context`` (fidget-core/src/context/mod.rs:49), ``Context::from_text``
(:878), ``VmData`` (fidget-core/src/vm/data.rs:65) and
``fidget_bytecode::Bytecode`` (fidget-bytecode/src/lib.rs:203).

The classes are parameterised by the shared library that provides the
``fh_*`` symbols so that the test oracle (which links its own copy of the
front end) can reuse them; the product always uses ``libfidget_cuda.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

# Opcode numbering == fidget_bytecode::BytecodeOp (fidget-bytecode/src/lib.rs:69-104)
OPCODES = [
    "Output", "Input", "Copy", "Neg", "Abs", "Recip", "Sqrt", "Square", "Floor",
    "Ceil", "Round", "Not", "Rand", "Sin", "Cos", "Tan", "Asin", "Acos", "Atan",
    "Exp", "Ln", "Add", "Sub", "Mul", "Div", "Atan2", "Compare", "Mix", "Mod",
    "Min", "Max", "And", "Or", "Mem",
]
OP = {name.lower(): i for i, name in enumerate(OPCODES)}
UNARY_OPS = [n.lower() for n in OPCODES[3:21]]
BINARY_OPS = [n.lower() for n in OPCODES[21:33]]


class FhTapeInfo(C.Structure):
    _fields_ = [
        ("ssa_len", C.c_uint32), ("asm_len", C.c_uint32), ("slot_count", C.c_uint32),
        ("choice_count", C.c_uint32), ("output_count", C.c_uint32), ("n_vars", C.c_uint32),
        ("n_regs", C.c_uint32), ("var_x", C.c_int32), ("var_y", C.c_int32), ("var_z", C.c_int32),
    ]


def bind_host_api(lib: C.CDLL) -> C.CDLL:
    """Declare the fh_* prototypes on ``lib`` (include/fidget_cuda.h, host section)."""
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
    P = C.POINTER
    lib.fh_last_error.restype = C.c_char_p
    lib.fh_context_new.argtypes = [P(vp)]
    lib.fh_context_free.argtypes = [vp]
    lib.fh_context_free.restype = None
    lib.fh_context_from_text.argtypes = [vp, C.c_char_p, P(u32)]
    lib.fh_constant.argtypes = [vp, C.c_float, P(u32)]
    lib.fh_var.argtypes = [vp, i32, P(u32), P(C.c_uint64)]
    lib.fh_unary.argtypes = [vp, C.c_uint8, u32, P(u32)]
    lib.fh_binary.argtypes = [vp, C.c_uint8, u32, u32, P(u32)]
    lib.fh_context_len.argtypes = [vp, P(u32)]
    lib.fh_tape_build.argtypes = [vp, P(u32), u32, u32, P(vp)]
    lib.fh_tape_free.argtypes = [vp]
    lib.fh_tape_free.restype = None
    lib.fh_tape_get_info.argtypes = [vp, P(FhTapeInfo)]
    lib.fh_tape_var.argtypes = [vp, u32, P(i32), P(C.c_uint64)]
    lib.fh_tape_bytecode.argtypes = [vp, i32, P(u32), C.c_size_t, P(C.c_size_t), P(C.c_uint8), P(u32)]
    lib.fh_tape_serialize.argtypes = [vp, vp, C.c_size_t, P(C.c_size_t)]
    lib.fh_tape_serialize.restype = i32
    lib.fh_tape_dump.argtypes = [vp, i32, C.c_char_p, C.c_size_t]
    lib.fh_tape_dump.restype = C.c_size_t
    for name in ("fh_context_new", "fh_context_from_text", "fh_constant", "fh_var", "fh_unary",
                 "fh_binary", "fh_context_len", "fh_tape_build", "fh_tape_get_info", "fh_tape_var",
                 "fh_tape_bytecode"):
        getattr(lib, name).restype = i32
    return lib


class HostError(RuntimeError):
    pass


def _check(lib, rc):
    if rc != 0:
        raise HostError(lib.fh_last_error().decode())


@dataclass
class Bytecode:
    words: np.ndarray  # uint32, including start/end markers
    reg_count: int
    mem_count: int


class Context:
    """Deduplicating expression arena; nodes are plain ints."""

    def __init__(self, lib: C.CDLL | None = None):
        if lib is None:
            from ._lib import load
            lib = load()
        self._lib = lib
        h = C.c_void_p()
        _check(lib, lib.fh_context_new(C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.fh_context_free(self._h)
            self._h = None

    @classmethod
    def from_text(cls, text: str, lib: C.CDLL | None = None):
        """``Context::from_text``: returns ``(ctx, root)``."""
        ctx = cls(lib)
        root = C.c_uint32()
        _check(ctx._lib, ctx._lib.fh_context_from_text(ctx._h, text.encode(), C.byref(root)))
        return ctx, root.value

    def __len__(self):
        n = C.c_uint32()
        _check(self._lib, self._lib.fh_context_len(self._h, C.byref(n)))
        return n.value

    def constant(self, v: float) -> int:
        n = C.c_uint32()
        _check(self._lib, self._lib.fh_constant(self._h, v, C.byref(n)))
        return n.value

    def _var(self, kind):
        n = C.c_uint32()
        vid = C.c_uint64()
        _check(self._lib, self._lib.fh_var(self._h, kind, C.byref(n), C.byref(vid)))
        return n.value, vid.value

    def x(self): return self._var(0)[0]
    def y(self): return self._var(1)[0]
    def z(self): return self._var(2)[0]

    def var(self):
        """Fresh anonymous variable (``Var::new``): returns ``(node, var_id)``."""
        return self._var(3)

    def _node(self, a):
        return self.constant(float(a)) if isinstance(a, float) else int(a)

    def unary(self, op: str, a) -> int:
        n = C.c_uint32()
        _check(self._lib, self._lib.fh_unary(self._h, OP[op], self._node(a), C.byref(n)))
        return n.value

    def binary(self, op: str, a, b) -> int:
        n = C.c_uint32()
        name = {"atan2": "atan2", "modulo": "mod"}.get(op, op)
        a = self._node(a)
        b = self._node(b)
        _check(self._lib, self._lib.fh_binary(self._h, OP[name], a, b, C.byref(n)))
        return n.value

    def __getattr__(self, name):
        # ctx.add(a, b), ctx.sqrt(a), ... like the reference's Context methods
        key = {"and_": "and", "or_": "or", "not_": "not", "modulo": "mod"}.get(name, name)
        if key in UNARY_OPS:
            return lambda a: self.unary(key, a)
        if key in BINARY_OPS:
            return lambda a, b: self.binary(key, a, b)
        raise AttributeError(name)

    def tape(self, roots, n_regs: int = 255) -> "TapeData":
        if isinstance(roots, int):
            roots = [roots]
        return TapeData(self, list(roots), n_regs)


class TapeData:
    """SSA tape + register tape + var map (``VmData<N>``)."""

    def __init__(self, ctx: Context, roots, n_regs=255):
        self._lib = ctx._lib
        arr = (C.c_uint32 * len(roots))(*roots)
        h = C.c_void_p()
        _check(self._lib, self._lib.fh_tape_build(ctx._h, arr, len(roots), n_regs, C.byref(h)))
        self._h = h
        info = FhTapeInfo()
        _check(self._lib, self._lib.fh_tape_get_info(h, C.byref(info)))
        self.info = info

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.fh_tape_free(self._h)
            self._h = None

    def __len__(self):
        return self.info.asm_len

    @property
    def choice_count(self): return self.info.choice_count
    @property
    def n_vars(self): return self.info.n_vars
    @property
    def output_count(self): return self.info.output_count

    def var_slots(self):
        """(x, y, z) input slots, -1 when the axis is unused."""
        return self.info.var_x, self.info.var_y, self.info.var_z

    def vars(self):
        out = []
        for i in range(self.info.n_vars):
            k, vid = C.c_int32(), C.c_uint64()
            _check(self._lib, self._lib.fh_tape_var(self._h, i, C.byref(k), C.byref(vid)))
            out.append(("xyzv"[k.value], vid.value))
        return out

    def bytecode(self, repack: bool = True) -> Bytecode:
        """``fidget_bytecode::Bytecode::new`` for this tape; built once and kept (treat ``words`` as read-only)."""
        cache = self.__dict__.setdefault("_bytecode", {})
        if repack in cache:
            return cache[repack]
        cache[repack] = self._build_bytecode(repack)
        return cache[repack]

    def _build_bytecode(self, repack: bool) -> Bytecode:
        n = C.c_size_t()
        rc, mc = C.c_uint8(), C.c_uint32()
        _check(self._lib, self._lib.fh_tape_bytecode(self._h, int(repack), None, 0, C.byref(n),
                                                     C.byref(rc), C.byref(mc)))
        words = np.zeros(n.value, dtype=np.uint32)
        _check(self._lib, self._lib.fh_tape_bytecode(
            self._h, int(repack), words.ctypes.data_as(C.POINTER(C.c_uint32)), n.value,
            C.byref(n), C.byref(rc), C.byref(mc)))
        return Bytecode(words, rc.value, mc.value)

    def serialize(self) -> bytes:
        """The tape's wire / on-disk blob ("FTAP", see include/fidget_cuda.h): what ``CudaShape.from_blob`` loads."""
        n = C.c_size_t()
        _check(self._lib, self._lib.fh_tape_serialize(self._h, None, 0, C.byref(n)))
        buf = (C.c_uint8 * n.value)()
        _check(self._lib, self._lib.fh_tape_serialize(self._h, buf, n.value, C.byref(n)))
        return bytes(buf)

    def dump(self, ssa: bool = False) -> str:
        n = self._lib.fh_tape_dump(self._h, int(ssa), None, 0)
        buf = C.create_string_buffer(n)
        self._lib.fh_tape_dump(self._h, int(ssa), buf, n)
        return buf.value.decode()
