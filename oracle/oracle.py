"""ORACLE -- test infrastructure, not product code.

ctypes face of ``oracle/liboracle.so`` (the CPU restatement of Fidget's
``VmShape`` path; see oracle/vm.h for the reference map).  Only tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from fidget_b200.host import Context as _Context, bind_host_api  # noqa: E402

_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = bind_host_api(C.CDLL(path))
    vp, u32, i32, P = C.c_void_p, C.c_uint32, C.c_int32, C.POINTER
    fp = P(C.c_float)
    L.orc_last_error.restype = C.c_char_p
    L.orc_tape_from_fh.argtypes = [vp, P(vp)]
    L.orc_tape_free.argtypes = [vp]
    L.orc_tape_free.restype = None
    L.orc_tape_info.argtypes = [vp] + [P(u32)] * 5
    L.orc_interval_eval.argtypes = [vp, fp, fp, P(C.c_uint8), P(C.c_uint8)]
    L.orc_point_eval.argtypes = [vp, fp, fp, P(C.c_uint8), P(C.c_uint8)]
    L.orc_float_slice_eval.argtypes = [vp, P(fp), P(fp), C.c_uint64]
    L.orc_grad_slice_eval.argtypes = [vp, P(fp), P(fp), C.c_uint64]
    L.orc_simplify.argtypes = [vp, P(C.c_uint8), C.c_size_t, P(vp)]
    L.orc_simplify_n.argtypes = [vp, P(C.c_uint8), C.c_size_t, u32, P(vp)]
    L.orc_tape_bytecode.argtypes = [vp, i32, P(u32), C.c_size_t, P(C.c_size_t), P(C.c_uint8), P(u32)]
    L.orc_screen_to_world_2d.argtypes = [u32, u32, fp]
    L.orc_screen_to_world_2d.restype = None
    L.orc_screen_to_world_3d.argtypes = [u32, u32, u32, fp]
    L.orc_screen_to_world_3d.restype = None
    L.orc_pixel_mat.argtypes = [u32, u32, fp, fp]
    L.orc_pixel_mat.restype = None
    L.orc_mat4_mul.argtypes = [fp, fp, fp]
    L.orc_mat4_mul.restype = None
    L.orc_transform_f32.argtypes = [fp, C.c_float, C.c_float, C.c_float, fp]
    L.orc_transform_f32.restype = None
    L.orc_render2d.argtypes = [vp, u32, u32, fp, C.c_float, i32, P(u32), u32, i32, u32, u32, fp, u32, fp, vp]
    L.orc_render3d.argtypes = [vp, u32, u32, u32, fp, P(u32), u32, i32, u32, u32, fp, u32, vp, vp]
    _LIB = L
    return L


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64 * 8) for n in
                ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified")] + \
               [("pixels", C.c_uint64)]

    def as_dict(self):
        d = {n: list(getattr(self, n)) for n in
             ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified")}
        d["pixels"] = int(self.pixels)
        return d


def _ck(rc):
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


GEOMETRY_PIXEL = np.dtype([("normal", np.float32, 3), ("depth", np.uint32)])


class Context(_Context):
    """Expression context bound to the oracle's copy of the host front end."""

    def __init__(self):
        super().__init__(lib())

    @classmethod
    def from_text(cls, text):
        ctx = cls()
        root = C.c_uint32()
        L = lib()
        if L.fh_context_from_text(ctx._h, text.encode(), C.byref(root)) != 0:
            raise RuntimeError(L.fh_last_error().decode())
        return ctx, root.value


class Tape:
    """A VmData-equivalent owned by the oracle."""

    def __init__(self, handle):
        self._h = handle
        vals = [C.c_uint32() for _ in range(5)]
        _ck(lib().orc_tape_info(handle, *[C.byref(v) for v in vals]))
        self.size, self.ssa_len, self.choice_count, self.slot_count, self.n_vars = [v.value for v in vals]

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:   # module globals are gone at interpreter shutdown
            lib().orc_tape_free(self._h)
            self._h = None

    @classmethod
    def from_data(cls, tape_data):
        h = C.c_void_p()
        _ck(lib().orc_tape_from_fh(tape_data._h, C.byref(h)))
        t = cls(h)
        t.data = tape_data
        return t

    @classmethod
    def from_vm(cls, text: str, n_regs: int = 255):
        ctx, root = Context.from_text(text)
        return cls.from_data(ctx.tape(root, n_regs))

    def interval_eval(self, vars_lo_hi):
        v = np.ascontiguousarray(vars_lo_hi, dtype=np.float32).reshape(-1, 2)
        assert v.shape[0] >= self.n_vars
        out = np.zeros(2, dtype=np.float32)
        choices = np.zeros(max(self.choice_count, 1), dtype=np.uint8)
        s = C.c_uint8()
        _ck(lib().orc_interval_eval(self._h, _fp(v), _fp(out),
                                    choices.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(s)))
        return out, choices[:self.choice_count], bool(s.value)

    def point_eval(self, vars_):
        v = np.ascontiguousarray(vars_, dtype=np.float32)
        out = np.zeros(1, dtype=np.float32)
        choices = np.zeros(max(self.choice_count, 1), dtype=np.uint8)
        s = C.c_uint8()
        _ck(lib().orc_point_eval(self._h, _fp(v), _fp(out),
                                 choices.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(s)))
        return out[0], choices[:self.choice_count], bool(s.value)

    def float_slice_eval(self, vars_):
        vs = [np.ascontiguousarray(v, dtype=np.float32) for v in vars_]
        n = len(vs[0]) if vs else 0
        out = np.zeros(n, dtype=np.float32)
        arr = (C.POINTER(C.c_float) * max(len(vs), 1))(*[_fp(v) for v in vs])
        oarr = (C.POINTER(C.c_float) * 1)(_fp(out))
        _ck(lib().orc_float_slice_eval(self._h, arr, oarr, n))
        return out

    def grad_slice_eval(self, vars_):
        """vars_: list of (n,4) arrays {v,dx,dy,dz}; returns (n,4)."""
        vs = [np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 4) for v in vars_]
        n = vs[0].shape[0] if vs else 0
        out = np.zeros((n, 4), dtype=np.float32)
        arr = (C.POINTER(C.c_float) * max(len(vs), 1))(*[_fp(v) for v in vs])
        oarr = (C.POINTER(C.c_float) * 1)(_fp(out))
        _ck(lib().orc_grad_slice_eval(self._h, arr, oarr, n))
        return out

    def simplify(self, choices, n_regs=None):
        c = np.ascontiguousarray(choices, dtype=np.uint8)
        h = C.c_void_p()
        cp = c.ctypes.data_as(C.POINTER(C.c_uint8))
        if n_regs is None:
            _ck(lib().orc_simplify(self._h, cp, len(c), C.byref(h)))
        else:
            _ck(lib().orc_simplify_n(self._h, cp, len(c), n_regs, C.byref(h)))
        return Tape(h)

    def bytecode(self, repack=True):
        from fidget_b200.host import Bytecode
        n = C.c_size_t()
        rc, mc = C.c_uint8(), C.c_uint32()
        _ck(lib().orc_tape_bytecode(self._h, int(repack), None, 0, C.byref(n), C.byref(rc), C.byref(mc)))
        words = np.zeros(n.value, dtype=np.uint32)
        _ck(lib().orc_tape_bytecode(self._h, int(repack), words.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    n.value, C.byref(n), C.byref(rc), C.byref(mc)))
        return Bytecode(words, rc.value, mc.value)


def screen_to_world_2d(w, h):
    m = np.zeros(16, dtype=np.float32)
    lib().orc_screen_to_world_2d(w, h, _fp(m))
    return m.reshape(4, 4)


def screen_to_world_3d(w, h, d):
    m = np.zeros(16, dtype=np.float32)
    lib().orc_screen_to_world_3d(w, h, d, _fp(m))
    return m.reshape(4, 4)


def pixel_mat(w, h, world_to_model=None):
    """pixel::RenderConfig::mat embedded as 4x4 (fidget-raster/src/pixel.rs:122-124,283-287)."""
    wm = np.eye(3, dtype=np.float32) if world_to_model is None else \
        np.ascontiguousarray(world_to_model, dtype=np.float32).reshape(3, 3)
    m = np.zeros(16, dtype=np.float32)
    lib().orc_pixel_mat(w, h, _fp(np.ascontiguousarray(wm)), _fp(m))
    return m.reshape(4, 4)


def voxel_mat(w, h, d, world_to_model=None):
    s = screen_to_world_3d(w, h, d)
    if world_to_model is None:
        return s
    wm = np.ascontiguousarray(world_to_model, dtype=np.float32).reshape(4, 4)
    out = np.zeros(16, dtype=np.float32)
    lib().orc_mat4_mul(_fp(wm), _fp(np.ascontiguousarray(s)), _fp(out))
    return out.reshape(4, 4)


def render2d(tape: Tape, width, height, mat=None, z=0.0, pixel_perfect=False, tile_sizes=(128, 32, 8),
             threads=1, first_root=0, n_roots=0, var_values=None):
    """Returns (image float32 [h,w] holding RawDistancePixel bits, stats dict)."""
    mat = pixel_mat(width, height) if mat is None else mat
    m = np.ascontiguousarray(mat, dtype=np.float32).reshape(16)
    ts = (C.c_uint32 * len(tile_sizes))(*tile_sizes)
    out = np.zeros((height, width), dtype=np.float32)
    st = OrcStats()
    vv = np.ascontiguousarray(var_values if var_values is not None else [], dtype=np.float32)
    _ck(lib().orc_render2d(tape._h, width, height, _fp(m), z, int(pixel_perfect), ts, len(tile_sizes),
                           threads, first_root, n_roots, _fp(vv), len(vv), _fp(out), C.byref(st)))
    return out, st.as_dict()


def render3d(tape: Tape, width, height, depth, mat=None, tile_sizes=(128, 64, 32, 16, 8), threads=1,
             first_root=0, n_roots=0, var_values=None):
    mat = voxel_mat(width, height, depth) if mat is None else mat
    m = np.ascontiguousarray(mat, dtype=np.float32).reshape(16)
    ts = (C.c_uint32 * len(tile_sizes))(*tile_sizes)
    out = np.zeros((height, width), dtype=GEOMETRY_PIXEL)
    st = OrcStats()
    vv = np.ascontiguousarray(var_values if var_values is not None else [], dtype=np.float32)
    _ck(lib().orc_render3d(tape._h, width, height, depth, _fp(m), ts, len(tile_sizes), threads,
                           first_root, n_roots, _fp(vv), len(vv), out.ctypes.data_as(C.c_void_p), C.byref(st)))
    return out, st.as_dict()


def pixel_inside(img):
    """RawDistancePixel::inside (fidget-raster/src/pixel.rs:177-183) on a float32 image."""
    bits = img.view(np.uint32)
    isnan = np.isnan(img)
    is_fill = isnan & ((bits & np.uint32(0xFF << 9)) == np.uint32(0xF6 << 9))
    return np.where(is_fill, (bits & 1) == 1, img < 0.0)


OCTREE_LEAF = np.dtype([("ix", np.uint16), ("iy", np.uint16), ("iz", np.uint16), ("mask", np.uint8),
                        ("n_edges", np.uint8), ("present", np.uint16), ("pad", np.uint16),
                        ("pos", np.float32, (12, 3)), ("grad", np.float32, (12, 4))])


def octree_sample(tape: Tape, depth: int, world_to_model=None, threads: int = 1):
    """Sampler half of fidget-mesh's Octree::build: returns (leaves sorted by (iz,iy,ix), stats dict).
    threads > 1 splits the tree into subtrees like Octree::build_inner_mt; the result is the serial one."""
    L = lib()
    L.orc_octree_sample_mt.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.c_int32, C.c_void_p, C.c_uint64,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    m = None if world_to_model is None else _fp(np.ascontiguousarray(world_to_model, dtype=np.float32).reshape(16))
    n = C.c_uint64()
    st = (C.c_uint64 * 69)()
    _ck(L.orc_octree_sample_mt(tape._h, depth, m, threads, None, 0, C.byref(n), st))
    leaves = np.zeros(n.value, dtype=OCTREE_LEAF)
    _ck(L.orc_octree_sample_mt(tape._h, depth, m, threads, leaves.ctypes.data_as(C.c_void_p), n.value, C.byref(n), st))
    s = list(st)
    stats = {"evaluated": s[0:16], "full": s[16:32], "empty": s[32:48], "ambiguous": s[48:64],
             "leaf_empty": s[64], "leaf_full": s[65], "leaf_surface": s[66], "float_points": s[67],
             "grad_points": s[68]}
    return leaves, stats


# ---- fidget-raster effects (oracle/effects.h) --------------------------------
def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _geo(image):
    image = np.ascontiguousarray(image, dtype=GEOMETRY_PIXEL)
    assert image.ndim == 2
    return image


def denoise_normals(image):
    image = _geo(image)
    out = np.zeros_like(image)
    lib().orc_denoise_normals(_vp(image), C.c_uint32(image.shape[1]), C.c_uint32(image.shape[0]), _vp(out))
    return out


def compute_ssao(image, depth, kernel, noise):
    image = _geo(image)
    k = np.ascontiguousarray(kernel, dtype=np.float32).reshape(-1, 3)
    nz = np.ascontiguousarray(noise, dtype=np.float32).reshape(-1, 2)
    out = np.zeros(image.shape, dtype=np.float32)
    lib().orc_compute_ssao(_vp(image), C.c_uint32(image.shape[1]), C.c_uint32(image.shape[0]), C.c_uint32(depth),
                           _fp(k), C.c_uint32(len(k)), _fp(nz), C.c_uint32(len(nz)), _fp(out))
    return out


def blur_ssao(ssao):
    ssao = np.ascontiguousarray(ssao, dtype=np.float32)
    out = np.zeros_like(ssao)
    lib().orc_blur_ssao(_fp(ssao), C.c_uint32(ssao.shape[1]), C.c_uint32(ssao.shape[0]), _fp(out))
    return out


def apply_shading(image, depth, ssao=None):
    """effects.rs:42-66 with the blurred occlusion map (or None) as an input."""
    image = _geo(image)
    out = np.zeros(image.shape + (3,), dtype=np.uint8)
    s = None if ssao is None else _fp(np.ascontiguousarray(ssao, dtype=np.float32))
    lib().orc_apply_shading(_vp(image), C.c_uint32(image.shape[1]), C.c_uint32(image.shape[0]), C.c_uint32(depth),
                            s, _u8p(out))
    return out


def normals_to_color(image):
    image = _geo(image)
    out = np.zeros(image.shape + (3,), dtype=np.uint8)
    lib().orc_normals_to_color(_vp(image), C.c_uint64(image.size), _u8p(out))
    return out


def _rgba(fn, image, *extra):
    image = np.ascontiguousarray(image, dtype=np.float32)
    out = np.zeros(image.shape + (4,), dtype=np.uint8)
    fn(_fp(image), C.c_uint64(image.size), *extra, _u8p(out))
    return out


def to_rgba_bitmap(image, transparent=False):
    return _rgba(lib().orc_to_rgba_bitmap, image, C.c_int32(int(transparent)))


def to_debug_bitmap(image):
    return _rgba(lib().orc_to_debug_bitmap, image)


def to_rgba_distance(image):
    return _rgba(lib().orc_to_rgba_distance, image)
