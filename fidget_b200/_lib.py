"""Loader for the in-tree ``libfidget_cuda.so`` (built by ``build.sh`` /
``__graft_entry__.build()``).  There is no CPU fallback: if the CUDA library is
missing this raises instead of silently routing anywhere else."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FIDGET_B200_LIB: another build of the same library (A/B measurements of a kernel change on one box)
LIB_PATH = os.environ.get("FIDGET_B200_LIB") or os.path.join(_HERE, "libfidget_cuda.so")
_LIB = None


class BackendMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BackendMissing(
                f"{LIB_PATH} not found: build it with ./build.sh (nvcc, sm_100a). "
                "fidget_b200 has no CPU fallback.")
        from .host import bind_host_api
        lib = C.CDLL(LIB_PATH)
        bind_host_api(lib)
        _bind_cuda_api(lib)
        _LIB = lib
    return _LIB


class FcTapeInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("n_ops", "ref_len", "choice_count", "reg_count", "mem_count", "n_vars", "n_outputs")]


class FcRender2dCfg(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("mat", C.c_float * 16), ("z", C.c_float),
                ("pixel_perfect", C.c_uint32), ("n_tile_sizes", C.c_uint32), ("tile_sizes", C.c_uint32 * 8),
                ("flags", C.c_uint32), ("root_row_begin", C.c_uint32), ("root_row_end", C.c_uint32),
                ("n_var_values", C.c_uint32), ("var_values", C.c_float * 16),
                ("root_stride", C.c_uint32), ("root_offset", C.c_uint32), ("out_format", C.c_uint32)]


class FcScheduleInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("suitable", "n_clauses", "n_waves", "widest_wave", "n_tail", "n_segments",
                                          "n_chain_clauses", "n_slots")]


class FcRender3dCfg(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32), ("mat", C.c_float * 16),
                ("n_tile_sizes", C.c_uint32), ("tile_sizes", C.c_uint32 * 8), ("flags", C.c_uint32),
                ("z_begin", C.c_uint32), ("z_end", C.c_uint32), ("n_var_values", C.c_uint32),
                ("var_values", C.c_float * 16), ("root_row_begin", C.c_uint32), ("root_row_end", C.c_uint32),
                ("root_stride", C.c_uint32), ("root_offset", C.c_uint32)]


class FcRenderStats(C.Structure):
    _fields_ = [(n, C.c_uint64 * 8) for n in
                ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified")] + \
               [("pixels", C.c_uint64), ("grads", C.c_uint64), ("arena_bytes_used", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("stage_ms", C.c_float * 16)]

    def as_dict(self):
        d = {n: list(getattr(self, n)) for n in
             ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified")}
        d.update(pixels=int(self.pixels), grads=int(self.grads), arena_bytes_used=int(self.arena_bytes_used),
                 kernel_launches=int(self.kernel_launches), stage_ms=list(self.stage_ms))
        return d


class FcOctreeCfg(C.Structure):
    _fields_ = [("depth", C.c_uint32), ("has_transform", C.c_uint32), ("world_to_model", C.c_float * 16),
                ("flags", C.c_uint32), ("n_var_values", C.c_uint32), ("var_values", C.c_float * 16)]


class FcOctreeStats(C.Structure):
    _fields_ = [(n, C.c_uint64 * 16) for n in ("evaluated", "full", "empty", "ambiguous")] + \
               [(n, C.c_uint64) for n in ("leaf_empty", "leaf_full", "leaf_surface", "float_points", "grad_points",
                                          "arena_bytes_used")] + \
               [("kernel_launches", C.c_uint32), ("total_ms", C.c_float)]

    def as_dict(self):
        d = {n: list(getattr(self, n)) for n in ("evaluated", "full", "empty", "ambiguous")}
        for n in ("leaf_empty", "leaf_full", "leaf_surface", "float_points", "grad_points", "arena_bytes_used",
                  "kernel_launches", "total_ms"):
            d[n] = getattr(self, n)
        return d


class FcMeshInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_leaves", "n_vertices", "n_triangles", "open_edges")] + \
               [("sampler_ms", C.c_float), ("mesh_ms", C.c_float)]


FC_FLAG_ASYNC = 1
FC_FLAG_TIMING = 2
FC_FLAG_NO_CLAMP = 4
FC_FLAG_FUSED_TAIL = 8
FC_FLAG_EXACT_CENSUS = 16
FC_FLAG_FULL_LADDER = 32
FC_OUT_F32, FC_OUT_MASK_U8, FC_OUT_BITMAP_1BIT, FC_OUT_RGBA8 = 0, 1, 2, 3

# name -> (restype, argtypes); mirrors include/fidget_cuda.h one to one
_vp, _u32, _i32, _u64, _u8 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint8
_P = C.POINTER
CUDA_API = {
    "fc_last_error": (C.c_char_p, []),
    "fc_abi_version": (_u32, []),
    "fc_ctx_create": (_i32, [_i32, _P(_vp)]),
    "fc_ctx_destroy": (None, [_vp]),
    "fc_ctx_set_stream": (_i32, [_vp, _vp, _i32]),
    "fc_ctx_synchronize": (_i32, [_vp]),
    "fc_ctx_set_arena_bytes": (_i32, [_vp, _u64]),
    "fc_tape_create": (_i32, [_vp, _P(_u32), C.c_size_t, _u8, _u32, _u32, _u32, _u32, _P(_vp)]),
    "fc_tape_retain": (_i32, [_vp]),
    "fc_tape_release": (_i32, [_vp]),
    "fc_tape_get_info": (_i32, [_vp, _P(FcTapeInfo)]),
    "fc_tape_set_axes": (_i32, [_vp, _i32, _i32, _i32]),
    "fc_tape_read": (_i32, [_vp, _P(_u32), C.c_size_t, _P(C.c_size_t)]),
    "fc_tape_serialize": (_i32, [_vp, _vp, C.c_size_t, _P(C.c_size_t)]),
    "fc_tape_deserialize": (_i32, [_vp, _vp, C.c_size_t, _P(_vp)]),
    "fc_eval_create": (_i32, [_vp, _P(_vp)]),
    "fc_eval_destroy": (None, [_vp]),
    "fc_interval_eval": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "fc_point_eval": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "fc_interval_eval_batch": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "fc_float_slice_eval": (_i32, [_vp, _vp, _P(_vp), _P(_vp), _u64]),
    "fc_grad_slice_eval": (_i32, [_vp, _vp, _P(_vp), _P(_vp), _u64]),
    "fc_simplify": (_i32, [_vp, _vp, _vp, C.c_size_t, _P(_vp)]),
    "fc_render2d": (_i32, [_vp, _vp, _P(FcRender2dCfg), _vp, _P(FcRenderStats)]),
    "fc_render3d": (_i32, [_vp, _vp, _P(FcRender3dCfg), _vp, _P(FcRenderStats)]),
    "fc_merge_slabs": (_i32, [_vp, _P(_vp), _u32, _u32, _u32, _u32, _vp]),
    "fc_tiles_per_rank": (_u32, [_u32, _u32, _u32, _u32]),
    "fc_tiles_pack": (_i32, [_vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp]),
    "fc_tiles_unpack": (_i32, [_vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp]),
    "fc_octree_sample": (_i32, [_vp, _vp, _P(FcOctreeCfg), _vp, _u64, _P(_u64), _P(FcOctreeStats)]),
    "fc_mesh_build": (_i32, [_vp, _vp, _P(FcOctreeCfg), _P(FcMeshInfo)]),
    "fc_mesh_read": (_i32, [_vp, _vp, _vp]),
    "fc_mesh_write_stl": (_i32, [_vp, _vp, C.c_size_t, _P(C.c_size_t)]),
    "fc_schedule_check": (_i32, [_P(_u32), C.c_size_t, _u8, _u32, _u32, _u32, _P(FcScheduleInfo)]),
    "fc_denoise_normals": (_i32, [_vp, _vp, _u32, _u32, _vp]),
    "fc_compute_ssao": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _u32, _vp, _u32, _vp]),
    "fc_blur_ssao": (_i32, [_vp, _vp, _u32, _u32, _vp]),
    "fc_apply_shading": (_i32, [_vp, _vp, _u32, _u32, _u32, _i32, _vp, _u32, _vp, _u32, _vp]),
    "fc_shade_with_occlusion": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "fc_normals_to_color": (_i32, [_vp, _vp, _u32, _u32, _vp]),
    "fc_to_rgba_bitmap": (_i32, [_vp, _vp, _u32, _u32, _i32, _vp]),
    "fc_to_debug_bitmap": (_i32, [_vp, _vp, _u32, _u32, _vp]),
    "fc_to_rgba_distance": (_i32, [_vp, _vp, _u32, _u32, _vp]),
}


def _bind_cuda_api(lib):
    for name, (res, args) in CUDA_API.items():
        f = getattr(lib, name)  # AttributeError here == header/library mismatch
        f.restype = res
        f.argtypes = args
