"""Post-processing effects of ``fidget_raster::effects`` (fidget-raster/src/effects.rs) on the GPU.

Same names and argument meaning as the reference: ``denoise_normals``, ``compute_ssao``,
``blur_ssao``, ``apply_shading``, ``to_rgba_bitmap``, ``to_debug_bitmap``, ``to_rgba_distance``
plus ``normals_to_color`` (``GeometryPixel::to_color``).  Images are numpy arrays (copied
through HBM) or torch CUDA tensors / anything with ``data_ptr()`` (used in place); ``out=``
lets the caller keep a result on the device.

The reference draws the SSAO sample tables from ``rand::rng()`` on every call
(effects.rs:385-440).  ``ssao_kernel`` / ``ssao_noise`` here follow the same rejection
sampling and scaling but take their uniform variates from Fidget's own deterministic hash
(fidget-core/src/rng/mod.rs:8-33), so that a render is reproducible.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .shape import GEOMETRY_PIXEL, CudaContext, _ck, _ptr


def _hash(v: int) -> int:
    state = (v * 747796405 + 2891336453) & 0xFFFFFFFF
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    return (word >> 22) ^ word


def _rand(seed: int) -> np.float32:
    """rng::rand: uniform in [0, 1) from an arbitrary integer seed."""
    bits = np.array([(_hash(seed & 0xFFFFFFFF) >> 9) | 0x3F800000], dtype=np.uint32)
    return bits.view(np.float32)[0] - np.float32(1.0)


def ssao_kernel(n: int = 64, seed: int = 0) -> np.ndarray:
    """effects::ssao_kernel (effects.rs:385-414): ``n`` points in the +Z unit hemisphere, scaled to
    cluster near the centre.  Returns float32 [n, 3]."""
    out = np.zeros((n, 3), dtype=np.float32)
    eps = np.finfo(np.float32).eps
    s = seed * 0x9E3779B1 + 1
    for i in range(n):
        while True:
            row = np.array([_rand(s) * 2 - 1, _rand(s + 1) * 2 - 1, _rand(s + 2)], dtype=np.float32)
            s += 3
            norm = np.float32(np.sqrt(np.float32((row * row).sum(dtype=np.float32))))
            if eps < norm < 1.0:
                t = np.float32(i) / np.float32(max(n - 1, 1))
                scale = t * t * np.float32(0.9) + np.float32(0.1)
                out[i] = row * scale / norm
                break
    return out


def ssao_noise(n: int = 256, seed: int = 0) -> np.ndarray:
    """effects::ssao_noise (effects.rs:420-440): ``n`` unit XY rotation vectors.  float32 [n, 2]."""
    out = np.zeros((n, 2), dtype=np.float32)
    eps = np.finfo(np.float32).eps
    s = seed * 0x85EBCA6B + 0x10000001
    for i in range(n):
        while True:
            row = np.array([_rand(s) * 2 - 1, _rand(s + 1) * 2 - 1], dtype=np.float32)
            s += 2
            norm = np.float32(np.sqrt(np.float32((row * row).sum(dtype=np.float32))))
            if eps < norm < 1.0:
                out[i] = row / norm
                break
    return out


def _shape2(image):
    shp = tuple(image.shape)
    if hasattr(image, "data_ptr") and len(shp) == 3:   # torch [h, w, 4] float view of GeometryPixel
        shp = shp[:2]
    if len(shp) != 2:
        raise ValueError("expected a [height, width] image")
    return shp


def _geo(image):
    if isinstance(image, np.ndarray):
        image = np.ascontiguousarray(image, dtype=GEOMETRY_PIXEL)
    return image, _shape2(image)


def _f32(image):
    if isinstance(image, np.ndarray):
        image = np.ascontiguousarray(image, dtype=np.float32)
    return image, _shape2(image)


def _tables(kernel, noise):
    k = np.ascontiguousarray(ssao_kernel(64) if kernel is None else kernel, dtype=np.float32).reshape(-1, 3)
    n = np.ascontiguousarray(ssao_noise(256) if noise is None else noise, dtype=np.float32).reshape(-1, 2)
    return k, n


def denoise_normals(cuda: CudaContext, image, out=None):
    image, (h, w) = _geo(image)
    if out is None:
        out = np.zeros((h, w), dtype=GEOMETRY_PIXEL)
    _ck(cuda._lib.fc_denoise_normals(cuda._h, _ptr(image), w, h, _ptr(out)))
    return out


def compute_ssao(cuda: CudaContext, image, depth: int, kernel=None, noise=None, out=None):
    image, (h, w) = _geo(image)
    k, n = _tables(kernel, noise)
    if out is None:
        out = np.zeros((h, w), dtype=np.float32)
    _ck(cuda._lib.fc_compute_ssao(cuda._h, _ptr(image), w, h, depth, _ptr(k), len(k), _ptr(n), len(n), _ptr(out)))
    return out


def blur_ssao(cuda: CudaContext, ssao, out=None):
    ssao, (h, w) = _f32(ssao)
    if out is None:
        out = np.zeros((h, w), dtype=np.float32)
    _ck(cuda._lib.fc_blur_ssao(cuda._h, _ptr(ssao), w, h, _ptr(out)))
    return out


def apply_shading(cuda: CudaContext, image, depth: int, ssao: bool = True, kernel=None, noise=None, out=None):
    """effects::apply_shading -> uint8 [h, w, 3]."""
    image, (h, w) = _geo(image)
    if out is None:
        out = np.zeros((h, w, 3), dtype=np.uint8)
    if ssao:
        k, n = _tables(kernel, noise)
        _ck(cuda._lib.fc_apply_shading(cuda._h, _ptr(image), w, h, depth, 1, _ptr(k), len(k), _ptr(n), len(n),
                                       _ptr(out)))
    else:
        _ck(cuda._lib.fc_apply_shading(cuda._h, _ptr(image), w, h, depth, 0, None, 0, None, 0, _ptr(out)))
    return out


def shade_with_occlusion(cuda: CudaContext, image, depth: int, blurred_ssao=None, out=None):
    image, (h, w) = _geo(image)
    if blurred_ssao is not None:
        blurred_ssao, _ = _f32(blurred_ssao)
    if out is None:
        out = np.zeros((h, w, 3), dtype=np.uint8)
    _ck(cuda._lib.fc_shade_with_occlusion(cuda._h, _ptr(image), w, h, depth, _ptr(blurred_ssao), _ptr(out)))
    return out


def normals_to_color(cuda: CudaContext, image, out=None):
    image, (h, w) = _geo(image)
    if out is None:
        out = np.zeros((h, w, 3), dtype=np.uint8)
    _ck(cuda._lib.fc_normals_to_color(cuda._h, _ptr(image), w, h, _ptr(out)))
    return out


def _rgba(fn, cuda, image, out, *extra):
    image, (h, w) = _f32(image)
    if out is None:
        out = np.zeros((h, w, 4), dtype=np.uint8)
    _ck(fn(cuda._h, _ptr(image), w, h, *extra, _ptr(out)))
    return out


def to_rgba_bitmap(cuda: CudaContext, image, transparent: bool = False, out=None):
    return _rgba(cuda._lib.fc_to_rgba_bitmap, cuda, image, out, int(transparent))


def to_debug_bitmap(cuda: CudaContext, image, out=None):
    return _rgba(cuda._lib.fc_to_debug_bitmap, cuda, image, out)


def to_rgba_distance(cuda: CudaContext, image, out=None):
    return _rgba(cuda._lib.fc_to_rgba_distance, cuda, image, out)
