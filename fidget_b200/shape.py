"""Python face of the CUDA backend, shaped like the reference's API:

  CudaShape            <-> ``Shape<F>`` for a new ``F = CudaFunction``
                           (fidget-core/src/shape/mod.rs:51, eval/mod.rs:80-208)
  .interval_eval etc.  <-> the four evaluators (eval/tracing.rs, eval/bulk.rs)
  RenderConfig2D/3D    <-> pixel::RenderConfig / voxel::RenderConfig
                           (fidget-raster/src/pixel.rs:27-39, voxel.rs:26-36)
  render2d / render3d  <-> pixel::render / voxel::render (pixel.rs:452, voxel.rs:500)

Everything here calls through the C ABI in include/fidget_cuda.h; there is no
CPU implementation behind it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .host import Context, TapeData

GEOMETRY_PIXEL = np.dtype([("normal", np.float32, 3), ("depth", np.uint32)])


class CudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[fc_status {code}] {msg}")
        self.code = code


def _ck(rc):
    if rc != 0:
        raise CudaError(rc, _lib.load().fc_last_error().decode())


def _ptr(a):
    """Raw address of a numpy array or of anything with ``data_ptr()`` (torch)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def schedule_check(tape: TapeData) -> dict:
    """Host-only: build and symbolically replay the level-0 schedule (waves, chain runs, slot colouring) of
    ``tape``; raises CudaError if the schedule is inconsistent.  Needs no GPU."""
    lib = _lib.load()
    bc = tape.bytecode()
    info = _lib.FcScheduleInfo()
    _ck(lib.fc_schedule_check(bc.words.ctypes.data_as(C.POINTER(C.c_uint32)), len(bc.words), bc.reg_count, bc.mem_count,
                              tape.n_vars, tape.output_count, C.byref(info)))
    return {n: getattr(info, n) for n, _ in info._fields_}


class CudaContext:
    """One GPU: stream + scratch arenas (``fc_ctx``)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _ck(self._lib.fc_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device
        self._stream = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_stream(self, cuda_stream: int | None):
        """Run on the given cudaStream_t handle (0 = CUDA default stream); None = the context's own stream."""
        if cuda_stream is None:
            _ck(self._lib.fc_ctx_set_stream(self._h, None, 1))
        else:
            _ck(self._lib.fc_ctx_set_stream(self._h, C.c_void_p(cuda_stream), 0))
        self._stream = cuda_stream

    def on_stream(self, cuda_stream: int):
        """Context manager: enqueue on ``cuda_stream`` (e.g. ``torch.cuda.current_stream().cuda_stream``) inside
        the block, then go back to whatever stream was bound before.  A no-op when already bound to it."""
        ctx = self

        class _Bound:
            def __enter__(self_b):
                self_b.prev = getattr(ctx, "_stream", None)
                self_b.changed = self_b.prev != cuda_stream
                if self_b.changed:
                    ctx.set_stream(cuda_stream)
                return ctx

            def __exit__(self_b, *a):
                if self_b.changed:
                    ctx.set_stream(self_b.prev)
                return False

        return _Bound()

    def synchronize(self):
        _ck(self._lib.fc_ctx_synchronize(self._h))

    def set_arena_bytes(self, n: int):
        _ck(self._lib.fc_ctx_set_arena_bytes(self._h, n))


class CudaShape:
    """A tape resident on the GPU plus its evaluators."""

    def __init__(self, cuda: CudaContext, tape: TapeData | None = None, *, _handle=None, _axes=None):
        self._lib = cuda._lib
        self.cuda = cuda
        if _handle is None:
            bc = tape.bytecode()
            h = C.c_void_p()
            _ck(self._lib.fc_tape_create(
                cuda._h, bc.words.ctypes.data_as(C.POINTER(C.c_uint32)), len(bc.words), bc.reg_count,
                bc.mem_count, tape.n_vars, tape.output_count, tape.choice_count, C.byref(h)))
            self._h = h
            self._axes = tape.var_slots()
            _ck(self._lib.fc_tape_set_axes(h, *self._axes))
        else:
            self._h = _handle
            self._axes = _axes
        info = _lib.FcTapeInfo()
        _ck(self._lib.fc_tape_get_info(self._h, C.byref(info)))
        self.info = info
        self._eval = None

    @classmethod
    def from_vm(cls, cuda: CudaContext, text: str, n_regs: int = 255):
        ctx, root = Context.from_text(text)
        return cls(cuda, ctx.tape(root, n_regs))

    @classmethod
    def from_blob(cls, cuda: CudaContext, blob: bytes):
        """Loads a serialized tape (``TapeData.serialize`` / ``CudaShape.serialize``): fc_tape_deserialize."""
        h = C.c_void_p()
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _ck(cuda._lib.fc_tape_deserialize(cuda._h, buf, len(blob), C.byref(h)))
        ax = np.frombuffer(blob, dtype=np.int32, count=3, offset=28)
        return cls(cuda, _handle=h, _axes=tuple(int(a) for a in ax))

    def serialize(self) -> bytes:
        n = C.c_size_t()
        _ck(self._lib.fc_tape_serialize(self._h, None, 0, C.byref(n)))
        buf = (C.c_uint8 * n.value)()
        _ck(self._lib.fc_tape_serialize(self._h, buf, n.value, C.byref(n)))
        return bytes(buf)

    def __del__(self):
        if getattr(self, "_eval", None):
            self._lib.fc_eval_destroy(self._eval)
            self._eval = None
        if getattr(self, "_h", None):
            self._lib.fc_tape_release(self._h)
            self._h = None

    # Function::size / choice_count / vars
    def size(self): return self.info.ref_len
    @property
    def choice_count(self): return self.info.choice_count
    @property
    def n_vars(self): return self.info.n_vars

    def _ev(self):
        if self._eval is None:
            h = C.c_void_p()
            _ck(self._lib.fc_eval_create(self.cuda._h, C.byref(h)))
            self._eval = h
        return self._eval

    def bytecode_words(self):
        n = C.c_size_t()
        _ck(self._lib.fc_tape_read(self._h, None, 0, C.byref(n)))
        w = np.zeros(n.value, dtype=np.uint32)
        _ck(self._lib.fc_tape_read(self._h, w.ctypes.data_as(C.POINTER(C.c_uint32)), n.value, C.byref(n)))
        return w

    # ---- tracing evaluators ------------------------------------------------
    def interval_eval(self, vars_lo_hi):
        """-> (out [n_out,2], choices uint8[choice_count], simplify flag)"""
        v = np.ascontiguousarray(vars_lo_hi, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((self.info.n_outputs, 2), dtype=np.float32)
        ch = np.zeros(max(self.choice_count, 1), dtype=np.uint8)
        s = np.zeros(1, dtype=np.uint8)
        _ck(self._lib.fc_interval_eval(self._ev(), self._h, _ptr(v), _ptr(out), _ptr(ch), _ptr(s)))
        return out, ch[:self.choice_count], bool(s[0])

    def point_eval(self, vars_):
        v = np.ascontiguousarray(vars_, dtype=np.float32)
        out = np.zeros(self.info.n_outputs, dtype=np.float32)
        ch = np.zeros(max(self.choice_count, 1), dtype=np.uint8)
        s = np.zeros(1, dtype=np.uint8)
        _ck(self._lib.fc_point_eval(self._ev(), self._h, _ptr(v), _ptr(out), _ptr(ch), _ptr(s)))
        return out, ch[:self.choice_count], bool(s[0])

    def interval_eval_batch(self, boxes, want_choices=False):
        """boxes: [n, n_vars, 2] -> out [n, n_out, 2] (+ choices [n, choice_count], simplify [n])"""
        v = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, max(self.n_vars, 1), 2)
        n = v.shape[0]
        out = np.zeros((n, self.info.n_outputs, 2), dtype=np.float32)
        ch = np.zeros((n, max(self.choice_count, 1)), dtype=np.uint8) if want_choices else None
        s = np.zeros(n, dtype=np.uint8)
        _ck(self._lib.fc_interval_eval_batch(self._ev(), self._h, _ptr(v), n, _ptr(out), _ptr(ch), _ptr(s)))
        if want_choices:
            return out, ch[:, :self.choice_count], s.astype(bool)
        return out

    # ---- bulk evaluators ---------------------------------------------------
    def float_slice_eval(self, vars_, out=None):
        """vars_: n_vars arrays (numpy or torch, host or device) of n floats"""
        n = int(vars_[0].shape[0]) if len(vars_) else 0
        host = not (len(vars_) and hasattr(vars_[0], "data_ptr"))
        if host:
            vars_ = [np.ascontiguousarray(v, dtype=np.float32) for v in vars_]
        if out is None:
            if host:
                outs = [np.zeros(n, dtype=np.float32) for _ in range(self.info.n_outputs)]
            else:
                import torch
                outs = [torch.empty(n, dtype=torch.float32, device=vars_[0].device)
                        for _ in range(self.info.n_outputs)]
        else:
            outs = out if isinstance(out, (list, tuple)) else [out]
        va = (C.c_void_p * max(len(vars_), 1))(*[_ptr(v) for v in vars_])
        oa = (C.c_void_p * len(outs))(*[_ptr(o) for o in outs])
        _ck(self._lib.fc_float_slice_eval(self._ev(), self._h, va, oa, n))
        return outs[0] if self.info.n_outputs == 1 else outs

    def grad_slice_eval(self, vars_, out=None):
        """vars_: n_vars arrays [n,4] = {v,dx,dy,dz}"""
        n = int(vars_[0].shape[0]) if len(vars_) else 0
        host = not (len(vars_) and hasattr(vars_[0], "data_ptr"))
        if host:
            vars_ = [np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 4) for v in vars_]
        if out is None:
            if host:
                outs = [np.zeros((n, 4), dtype=np.float32) for _ in range(self.info.n_outputs)]
            else:
                import torch
                outs = [torch.empty((n, 4), dtype=torch.float32, device=vars_[0].device)
                        for _ in range(self.info.n_outputs)]
        else:
            outs = out if isinstance(out, (list, tuple)) else [out]
        va = (C.c_void_p * max(len(vars_), 1))(*[_ptr(v) for v in vars_])
        oa = (C.c_void_p * len(outs))(*[_ptr(o) for o in outs])
        _ck(self._lib.fc_grad_slice_eval(self._ev(), self._h, va, oa, n))
        return outs[0] if self.info.n_outputs == 1 else outs

    def simplify(self, choices) -> "CudaShape":
        c = np.ascontiguousarray(choices, dtype=np.uint8)
        h = C.c_void_p()
        _ck(self._lib.fc_simplify(self._ev(), self._h, _ptr(c), len(c), C.byref(h)))
        return CudaShape(self.cuda, _handle=h, _axes=self._axes)


# ---------------------------------------------------------------------------
# Transforms (f32 arithmetic, same operation order as the reference)
_f = np.float32


def screen_to_world_2d(w: int, h: int) -> np.ndarray:
    """RegionSize<2>::screen_to_world (fidget-core/src/render/region.rs:87-108) as a 4x4."""
    cx, cy = _f(w) / _f(2), _f(h) / _f(2) - _f(1)
    s = _f(2) / _f(min(w, h))
    sy = s * _f(-1)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 3] = s, -cx * s
    m[1, 1], m[1, 3] = sy, -cy * sy
    return m


def screen_to_world_3d(w: int, h: int, d: int) -> np.ndarray:
    cx, cy, cz = _f(w) / _f(2), _f(h) / _f(2) - _f(1), _f(d) / _f(2)
    s = _f(2) / _f(min(w, h, d))
    sy = s * _f(-1)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 3] = s, -cx * s
    m[1, 1], m[1, 3] = sy, -cy * sy
    m[2, 2], m[2, 3] = s, -cz * s
    return m


def _matmul_f32(a, b):
    n = a.shape[0]
    r = np.zeros((n, n), dtype=np.float32)
    for i in range(n):
        for j in range(n):
            acc = _f(0)
            for k in range(n):
                acc = _f(acc + _f(a[i, k] * b[k, j]))
            r[i, j] = acc
    return r


def pixel_mat(w: int, h: int, world_to_model=None) -> np.ndarray:
    """pixel::RenderConfig::mat embedded as 4x4 with Z preserved (pixel.rs:122-124,283-287)."""
    s = screen_to_world_2d(w, h)
    s3 = np.array([[s[0, 0], s[0, 1], s[0, 3]], [s[1, 0], s[1, 1], s[1, 3]], [0, 0, 1]], dtype=np.float32)
    wm = np.eye(3, dtype=np.float32) if world_to_model is None else \
        np.asarray(world_to_model, dtype=np.float32).reshape(3, 3)
    r = _matmul_f32(wm, s3)
    m = np.zeros((4, 4), dtype=np.float32)
    idx = [0, 1, 3]
    for i in range(3):
        for j in range(3):
            m[idx[i], idx[j]] = r[i, j]
    m[2, 2] = 1.0
    return m


def voxel_mat(w: int, h: int, d: int, world_to_model=None) -> np.ndarray:
    """voxel::RenderConfig::mat (voxel.rs:107-109)."""
    s = screen_to_world_3d(w, h, d)
    if world_to_model is None:
        return s
    return _matmul_f32(np.asarray(world_to_model, dtype=np.float32).reshape(4, 4), s)


@dataclass
class RenderConfig2D:
    width: int
    height: int
    world_to_model: np.ndarray | None = None   # 3x3
    pixel_perfect: bool = False
    z: float = 0.0
    tile_sizes: tuple = ()                      # () => backend default (128, 32, 8)
    mat: np.ndarray | None = None               # full 4x4 override
    root_rows: tuple = (0, 0)                   # band of root-tile rows (multi-GPU)
    timing: bool = False
    var_values: tuple = ()                      # ShapeVars: value per tape input slot (axis slots ignored)
    interleave: tuple = (0, 0)                  # (N, r): only the root tiles rank r of N owns (shard.tile_owner)
    out_format: str = "f32"                     # "f32" | "mask_u8" | "bitmap_1bit" | "rgba8"
    fused_tail: bool = False                    # experimental: levels 1.., leaf pixels, fills as one persistent launch

    def matrix(self):
        return self.mat if self.mat is not None else pixel_mat(self.width, self.height, self.world_to_model)


@dataclass
class RenderConfig3D:
    width: int
    height: int
    depth: int
    world_to_model: np.ndarray | None = None   # 4x4
    tile_sizes: tuple = ()
    mat: np.ndarray | None = None
    z_range: tuple = (0, 0)
    root_rows: tuple = (0, 0)                   # band of root-tile rows, full depth (multi-GPU)
    timing: bool = False
    var_values: tuple = ()
    clamp: bool = True                          # False for slab renders (fc_merge_slabs applies it)
    interleave: tuple = (0, 0)                  # (N, r): only the root-tile columns rank r of N owns
    exact_census: bool = False                  # stats = the reference's front-to-back census (voxel.rs:244-357)
    full_ladder: bool = False                   # default tile sizes: evaluate all of (128,64,32,16,8), not the device's (128,32,8)

    def matrix(self):
        return self.mat if self.mat is not None else voxel_mat(self.width, self.height, self.depth,
                                                               self.world_to_model)


OUT_FORMATS = {"f32": _lib.FC_OUT_F32, "mask_u8": _lib.FC_OUT_MASK_U8, "bitmap_1bit": _lib.FC_OUT_BITMAP_1BIT,
               "rgba8": _lib.FC_OUT_RGBA8}


def render2d(shape: CudaShape, cfg: RenderConfig2D, out=None, stats: bool = False, asynchronous: bool = False):
    """pixel::render.  ``out``: None (returns a numpy float32 [h,w] of
    RawDistancePixel bits), a numpy array, or a CUDA torch tensor."""
    lib = shape._lib
    c = _lib.FcRender2dCfg()
    c.width, c.height = cfg.width, cfg.height
    c.mat[:] = np.ascontiguousarray(cfg.matrix(), dtype=np.float32).reshape(16).tolist()
    c.z = cfg.z
    c.pixel_perfect = int(cfg.pixel_perfect)
    c.n_tile_sizes = len(cfg.tile_sizes)
    for i, t in enumerate(cfg.tile_sizes):
        c.tile_sizes[i] = t
    c.flags = (_lib.FC_FLAG_TIMING if cfg.timing else 0) | (_lib.FC_FLAG_ASYNC if asynchronous else 0) | \
        (_lib.FC_FLAG_FUSED_TAIL if cfg.fused_tail else 0)
    c.root_row_begin, c.root_row_end = cfg.root_rows
    c.root_stride, c.root_offset = cfg.interleave
    c.n_var_values = len(cfg.var_values)
    for i, v in enumerate(cfg.var_values):
        c.var_values[i] = float(v)
    c.out_format = OUT_FORMATS[cfg.out_format]
    if out is None:
        if cfg.out_format == "f32":
            out = np.zeros((cfg.height, cfg.width), dtype=np.float32)
        elif cfg.out_format == "mask_u8":
            out = np.zeros((cfg.height, cfg.width), dtype=np.uint8)
        elif cfg.out_format == "bitmap_1bit":
            out = np.zeros((cfg.height, (cfg.width + 7) // 8), dtype=np.uint8)
        else:
            out = np.zeros((cfg.height, cfg.width, 4), dtype=np.uint8)
    st = _lib.FcRenderStats() if stats else None
    _ck(lib.fc_render2d(shape.cuda._h, shape._h, C.byref(c), _ptr(out), C.byref(st) if stats else None))
    return (out, st.as_dict()) if stats else out


def render3d(shape: CudaShape, cfg: RenderConfig3D, out=None, stats: bool = False, asynchronous: bool = False):
    """voxel::render -> numpy structured array [h,w] of GEOMETRY_PIXEL (or fills ``out``)."""
    lib = shape._lib
    c = _lib.FcRender3dCfg()
    c.width, c.height, c.depth = cfg.width, cfg.height, cfg.depth
    c.mat[:] = np.ascontiguousarray(cfg.matrix(), dtype=np.float32).reshape(16).tolist()
    c.n_tile_sizes = len(cfg.tile_sizes)
    for i, t in enumerate(cfg.tile_sizes):
        c.tile_sizes[i] = t
    c.flags = (_lib.FC_FLAG_TIMING if cfg.timing else 0) | (_lib.FC_FLAG_ASYNC if asynchronous else 0) | \
        (0 if cfg.clamp else _lib.FC_FLAG_NO_CLAMP) | (_lib.FC_FLAG_EXACT_CENSUS if cfg.exact_census else 0) | \
        (_lib.FC_FLAG_FULL_LADDER if cfg.full_ladder else 0)
    c.z_begin, c.z_end = cfg.z_range
    c.root_row_begin, c.root_row_end = cfg.root_rows
    c.root_stride, c.root_offset = cfg.interleave
    c.n_var_values = len(cfg.var_values)
    for i, v in enumerate(cfg.var_values):
        c.var_values[i] = float(v)
    if out is None:
        out = np.zeros((cfg.height, cfg.width), dtype=GEOMETRY_PIXEL)
    st = _lib.FcRenderStats() if stats else None
    _ck(lib.fc_render3d(shape.cuda._h, shape._h, C.byref(c), _ptr(out), C.byref(st) if stats else None))
    return (out, st.as_dict()) if stats else out


OCTREE_LEAF = np.dtype([("ix", np.uint16), ("iy", np.uint16), ("iz", np.uint16), ("mask", np.uint8),
                        ("n_edges", np.uint8), ("present", np.uint16), ("pad", np.uint16),
                        ("pos", np.float32, (12, 3)), ("grad", np.float32, (12, 4))])


def octree_sample(shape: CudaShape, depth: int, world_to_model=None, capacity: int | None = None,
                  stats: bool = False, timing: bool = False, var_values=()):
    """Sampler half of ``fidget_mesh::Octree::build`` (octree.rs:521-808): surface leaves with their
    corner mask and per-edge Hermite data, sorted by (iz, iy, ix)."""
    lib = shape._lib
    c = _lib.FcOctreeCfg()
    c.depth = depth
    if world_to_model is not None:
        c.has_transform = 1
        c.world_to_model[:] = np.ascontiguousarray(world_to_model, dtype=np.float32).reshape(16).tolist()
    c.flags = _lib.FC_FLAG_TIMING if timing else 0
    c.n_var_values = len(var_values)
    for i, v in enumerate(var_values):
        c.var_values[i] = float(v)
    cap = capacity if capacity is not None else max(1024, min(8 ** depth, 6 * 4 ** depth))
    st = _lib.FcOctreeStats()
    while True:
        out = np.zeros(cap, dtype=OCTREE_LEAF)
        n = C.c_uint64()
        rc = lib.fc_octree_sample(shape.cuda._h, shape._h, C.byref(c), _ptr(out), cap, C.byref(n), C.byref(st))
        if rc != 0 and n.value > cap and capacity is None:
            cap = int(n.value)          # retry once with the exact count
            continue
        _ck(rc)
        break
    leaves = out[:n.value]
    leaves = leaves[np.lexsort((leaves["ix"], leaves["iy"], leaves["iz"]))]
    return (leaves, st.as_dict()) if stats else leaves


def mesh(shape: CudaShape, depth: int, world_to_model=None, var_values=(), stl: bool = False):
    """``Octree::build(...).walk_dual()`` without cell collapse (fidget-mesh): returns ``(vertices [n,3] float32,
    triangles [m,3] uint32, info dict)`` -- plus the binary STL bytes (``Mesh::write_stl``) when ``stl``."""
    lib = shape._lib
    c = _lib.FcOctreeCfg()
    c.depth = depth
    if world_to_model is not None:
        c.has_transform = 1
        c.world_to_model[:] = np.ascontiguousarray(world_to_model, dtype=np.float32).reshape(16).tolist()
    c.flags = _lib.FC_FLAG_TIMING
    c.n_var_values = len(var_values)
    for i, v in enumerate(var_values):
        c.var_values[i] = float(v)
    info = _lib.FcMeshInfo()
    _ck(lib.fc_mesh_build(shape.cuda._h, shape._h, C.byref(c), C.byref(info)))
    verts = np.zeros((info.n_vertices, 3), dtype=np.float32)
    tris = np.zeros((info.n_triangles, 3), dtype=np.uint32)
    _ck(lib.fc_mesh_read(shape.cuda._h, _ptr(verts), _ptr(tris)))
    d = {n: getattr(info, n) for n, _ in info._fields_}
    if not stl:
        return verts, tris, d
    n = C.c_size_t()
    _ck(lib.fc_mesh_write_stl(shape.cuda._h, None, 0, C.byref(n)))
    buf = np.zeros(n.value, dtype=np.uint8)
    _ck(lib.fc_mesh_write_stl(shape.cuda._h, _ptr(buf), n.value, C.byref(n)))
    return verts, tris, d, buf.tobytes()


def pixel_inside(img: np.ndarray) -> np.ndarray:
    """RawDistancePixel::inside (pixel.rs:177-183)."""
    bits = img.view(np.uint32)
    is_fill = np.isnan(img) & ((bits & np.uint32(0xFF << 9)) == np.uint32(0xF6 << 9))
    return np.where(is_fill, (bits & 1) == 1, img < 0.0)
