"""GPU parity tests for the fidget-raster post-processing effects: the CUDA kernels (through the
C ABI) against the CPU oracle on identical images.  Bit-exact everywhere except to_rgba_distance
(exp / cos from libdevice vs glibc: within one 8-bit step)."""
import numpy as np
import pytest

import fidget_b200 as fb
from fidget_b200 import effects as fx
from conftest import model_text, same_f32

pytestmark = pytest.mark.gpu

GEO = fb.GEOMETRY_PIXEL


def random_geometry(w, h, depth, seed):
    """Adversarial GeometryPixel image: empty pixels, back-facing, zero, huge and NaN normals."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), dtype=GEO)
    img["depth"] = np.where(rng.random((h, w)) < 0.2, 0, rng.integers(1, depth + 1, (h, w))).astype(np.uint32)
    n = rng.normal(size=(h, w, 3)).astype(np.float32)
    n[rng.random((h, w)) < 0.05] = 0.0
    n[rng.random((h, w)) < 0.02] *= np.float32(1e20)
    n[rng.random((h, w)) < 0.01, 0] = np.nan
    img["normal"] = n
    img["normal"][img["depth"] == 0] = 0.0
    return img


@pytest.fixture(scope="module")
def bear_geometry(cuda):
    shape = fb.CudaShape.from_vm(cuda, model_text("bear.vm"))
    return fb.render3d(shape, fb.RenderConfig3D(256, 256, 256))


@pytest.fixture(scope="module")
def tables():
    return fx.ssao_kernel(64), fx.ssao_noise(256)


def same_geometry(a, b):
    """Equal depths and bitwise-equal normals, any NaN matching any NaN (payloads are unspecified)."""
    a = np.asarray(a).view(np.float32).reshape(-1, 4)
    b = np.asarray(b).view(np.float32).reshape(-1, 4)
    return bool(np.array_equal(a[:, 3].view(np.uint32), b[:, 3].view(np.uint32)) and same_f32(a[:, :3], b[:, :3]))


SIZES = [(1, 1, 8), (5, 3, 16), (67, 45, 33), (128, 96, 64), (33, 257, 512)]


@pytest.mark.parametrize("w,h,d", SIZES)
def test_denoise_random(orc, cuda, w, h, d):
    img = random_geometry(w, h, d, w * 7 + h)
    assert same_geometry(fx.denoise_normals(cuda, img), orc.denoise_normals(img))


@pytest.mark.parametrize("w,h,d", SIZES)
def test_ssao_blur_shading_random(orc, cuda, tables, w, h, d):
    k, n = tables
    img = random_geometry(w, h, d, w * 11 + h)
    s_gpu, s_cpu = fx.compute_ssao(cuda, img, d, k, n), orc.compute_ssao(img, d, k, n)
    assert same_f32(s_gpu, s_cpu)
    b_gpu, b_cpu = fx.blur_ssao(cuda, s_cpu), orc.blur_ssao(s_cpu)
    assert same_f32(b_gpu, b_cpu)
    assert np.array_equal(fx.shade_with_occlusion(cuda, img, d, b_cpu), orc.apply_shading(img, d, b_cpu))
    assert np.array_equal(fx.shade_with_occlusion(cuda, img, d, None), orc.apply_shading(img, d, None))
    # the fused pipeline (SSAO -> blur inside the shading kernel)
    assert np.array_equal(fx.apply_shading(cuda, img, d, True, k, n), orc.apply_shading(img, d, b_cpu))
    assert np.array_equal(fx.apply_shading(cuda, img, d, False), orc.apply_shading(img, d, None))
    assert np.array_equal(fx.normals_to_color(cuda, img), orc.normals_to_color(img))


def test_blur_adversarial(orc, cuda):
    rng = np.random.default_rng(9)
    s = rng.random((61, 83)).astype(np.float32)
    s[rng.random(s.shape) < 0.3] = np.nan
    s[10:20, 10:20] = 0.5          # ties between windows: min_by_key keeps the first
    assert same_f32(fx.blur_ssao(cuda, s), orc.blur_ssao(s))


def test_effects_on_rendered_bear(orc, cuda, tables, bear_geometry):
    """The reference's viewer pipeline: voxel::render -> denoise_normals -> apply_shading(ssao)."""
    k, n = tables
    img = bear_geometry
    assert (img["depth"] > 0).sum() > 5000
    den_gpu, den_cpu = fx.denoise_normals(cuda, img), orc.denoise_normals(img)
    assert den_gpu.tobytes() == den_cpu.tobytes()
    want = orc.apply_shading(den_cpu, 256, orc.blur_ssao(orc.compute_ssao(den_cpu, 256, k, n)))
    got = fx.apply_shading(cuda, den_gpu, 256, True, k, n)
    assert np.array_equal(got, want)
    assert got.max() > 100 and got[img["depth"] == 0].max() == 0
    assert np.array_equal(fx.normals_to_color(cuda, img), orc.normals_to_color(img))


def test_rgba_conversions_on_rendered_prospero(orc, cuda):
    shape = fb.CudaShape.from_vm(cuda, model_text("prospero.vm"))
    img = fb.render2d(shape, fb.RenderConfig2D(512, 512))
    assert np.array_equal(fx.to_rgba_bitmap(cuda, img), orc.to_rgba_bitmap(img))
    assert np.array_equal(fx.to_rgba_bitmap(cuda, img, transparent=True), orc.to_rgba_bitmap(img, transparent=True))
    assert np.array_equal(fx.to_debug_bitmap(cuda, img), orc.to_debug_bitmap(img))
    g, o = fx.to_rgba_distance(cuda, img).astype(int), orc.to_rgba_distance(img).astype(int)
    assert np.abs(g - o).max() <= 1 and (g != o).mean() < 0.01
    # inside() of the bitmap is the renderer's own notion of inside
    assert np.array_equal(fx.to_rgba_bitmap(cuda, img)[..., 0] == 255, fb.pixel_inside(img))


@pytest.mark.parametrize("n", [1, 2, 3, 5, 1023, 1025])
def test_rgba_ragged_sizes(orc, cuda, n):
    rng = np.random.default_rng(n)
    img = rng.normal(size=(1, n)).astype(np.float32)
    bits = img.view(np.uint32)
    fill = rng.random((1, n)) < 0.3
    bits[fill] = (0x7FC00000 | (0xF6 << 9) | rng.integers(0, 512, (1, n)).astype(np.uint32))[fill]
    img[rng.random((1, n)) < 0.05] = np.nan
    assert np.array_equal(fx.to_debug_bitmap(cuda, img), orc.to_debug_bitmap(img))
    assert np.array_equal(fx.to_rgba_bitmap(cuda, img, True), orc.to_rgba_bitmap(img, True))
    assert np.abs(fx.to_rgba_distance(cuda, img).astype(int) - orc.to_rgba_distance(img).astype(int)).max() <= 1


def test_device_resident_pipeline(orc, cuda, tables):
    """torch CUDA tensors go through the kernels in place: nothing is staged or copied back."""
    import torch
    k, n = tables
    img = random_geometry(200, 120, 64, 77)
    d_img = torch.from_numpy(img.view(np.float32).reshape(120, 200, 4).copy()).cuda()
    d_den = torch.empty_like(d_img)
    d_rgb = torch.empty((120, 200, 3), dtype=torch.uint8, device="cuda")
    fx.denoise_normals(cuda, d_img, out=d_den)
    fx.apply_shading(cuda, d_den, 64, True, k, n, out=d_rgb)
    den = orc.denoise_normals(img)
    want = orc.apply_shading(den, 64, orc.blur_ssao(orc.compute_ssao(den, 64, k, n)))
    assert same_geometry(d_den.cpu().numpy(), den)
    assert np.array_equal(d_rgb.cpu().numpy(), want)


def test_effect_argument_errors(cuda):
    img = random_geometry(8, 8, 8, 1)
    with pytest.raises(fb.CudaError):
        fx.compute_ssao(cuda, img, 8, np.zeros((0, 3), np.float32), np.zeros((4, 2), np.float32))
    d = fx.denoise_normals(cuda, np.zeros((0, 0), dtype=GEO))
    assert d.size == 0
