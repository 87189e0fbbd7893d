"""Pins the effects oracle (oracle/effects.cc) to fidget-raster/src/effects.rs.

The reference has no tests or fixtures for effects.rs, so the pins are (a) the literal colour
tables of the source, (b) closed-form cases, and (c) an independent float32 numpy transcription
of compute_pixel_ssao / shade_pixel written from the Rust source (not from the C++ oracle).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from fidget_b200.effects import ssao_kernel, ssao_noise

GEO = orc.GEOMETRY_PIXEL
f32 = np.float32


def fill_bits(depth, inside):
    return np.array([0x7FC00000 | (depth << 1) | int(inside) | (0xF6 << 9)], dtype=np.uint32).view(np.float32)[0]


def sphere_heightmap(n, depth, rng=None):
    """Analytic sphere of radius 0.8 rendered as a GeometryPixel image (plus a few back-facing normals)."""
    img = np.zeros((n, n), dtype=GEO)
    ys, xs = np.mgrid[0:n, 0:n]
    wx = (xs + 0.5) / n * 2 - 1
    wy = (ys + 0.5) / n * 2 - 1
    r2 = 0.64 - wx * wx - wy * wy
    hit = r2 > 0
    wz = np.sqrt(np.where(hit, r2, 0))
    img["depth"] = np.where(hit, ((wz + 1) / 2 * depth).astype(np.uint32), 0)
    nrm = np.stack([wx, wy, wz], axis=-1) / 0.8
    img["normal"] = np.where(hit[..., None], nrm, 0).astype(np.float32)
    if rng is not None:
        bad = hit & (rng.random((n, n)) < 0.05)
        img["normal"][bad] *= np.float32(-1.0)
    return img


def test_rgba_bitmap_table():
    # effects.rs:446-467: inside -> 255 x4; outside -> transparent or opaque black
    img = np.array([[-1.0, 0.0, 2.0, np.nan, fill_bits(0, True), fill_bits(3, False)]], dtype=np.float32)
    out = orc.to_rgba_bitmap(img, transparent=False)[0]
    assert out.tolist() == [[255] * 4, [0, 0, 0, 255], [0, 0, 0, 255], [0, 0, 0, 255], [255] * 4, [0, 0, 0, 255]]
    out = orc.to_rgba_bitmap(img, transparent=True)[0]
    assert out.tolist() == [[255] * 4, [0] * 4, [0] * 4, [0] * 4, [255] * 4, [0] * 4]


def test_debug_bitmap_table():
    # effects.rs:470-497
    cases = [(-0.5, [255] * 4), (0.5, [0, 0, 0, 255]), (fill_bits(0, True), [255, 0, 0, 255]),
             (fill_bits(0, False), [50, 0, 0, 255]), (fill_bits(1, True), [0, 255, 0, 255]),
             (fill_bits(1, False), [0, 50, 0, 255]), (fill_bits(2, True), [0, 0, 255, 255]),
             (fill_bits(2, False), [0, 0, 50, 255]), (fill_bits(3, True), [255, 255, 0, 255]),
             (fill_bits(7, False), [50, 50, 0, 255])]
    img = np.array([[c[0] for c in cases]], dtype=np.float32)
    assert orc.to_debug_bitmap(img)[0].tolist() == [c[1] for c in cases]


def test_rgba_distance():
    # effects.rs:504-547: fills, NaN and an independent numpy evaluation of the banded SDF shading
    img = np.array([[fill_bits(0, True), fill_bits(1, False), np.nan]], dtype=np.float32)
    assert orc.to_rgba_distance(img)[0].tolist() == [[184, 235, 255, 255], [217, 144, 72, 255], [255, 0, 0, 255]]
    f = np.linspace(-0.7, 0.7, 301, dtype=np.float32).reshape(1, -1)
    got = orc.to_rgba_distance(f)[0].astype(int)
    af = np.abs(f[0]).astype(np.float64)
    dim = 1 - np.exp(-4 * af)
    bands = 0.8 + 0.2 * np.cos(140 * f[0].astype(np.float64))

    def smooth(e1, x):
        t = np.clip(x / e1, 0, 1)
        return t * t * (3 - 2 * t)
    for c, base in enumerate([0.1, 0.4, 0.7]):
        v = (1 - np.copysign(base, f[0])) * dim * bands
        for e1 in (0.015, 0.005):
            a = 1 - smooth(e1, af)
            v = v * (1 - a) + a
        want = (np.clip(v, 0, 1) * 255).astype(int)
        assert np.abs(got[:, c] - want).max() <= 1
    assert (got[:, 3] == 255).all()


def test_normals_to_color():
    # voxel.rs:136-153
    img = np.zeros((1, 4), dtype=GEO)
    img["normal"][0] = [[0, 0, 0], [0, 0, 2], [3, -4, 0], [-1, 1, 1]]
    out = orc.normals_to_color(img)[0]
    s = np.sqrt(f32(3))
    k = int(f32(1) * (f32(255) / s))
    assert out.tolist() == [[0, 0, 0], [0, 0, 255], [153, 204, 0], [k, k, k]]


def test_denoise_keeps_front_facing_and_empty():
    # effects.rs:17-36, 262-265
    img = sphere_heightmap(32, 32)
    out = orc.denoise_normals(img)
    assert out.tobytes() == img.tobytes()


def test_denoise_replaces_back_facing():
    # one back-facing pixel in a constant field -> the neighbourhood mean (= the field)
    img = np.zeros((9, 9), dtype=GEO)
    img["depth"] = 5
    img["normal"] = np.array([0.25, 0.5, 0.75], dtype=np.float32)
    img["normal"][4, 4] = [0.0, 0.0, -1.0]
    out = orc.denoise_normals(img)
    assert out["normal"][4, 4].tolist() == [0.25, 0.5, 0.75]
    assert (out["depth"] == 5).all()
    # a back-facing pixel with no front-facing neighbour keeps its normal (unwrap_or, effects.rs:324)
    img["normal"] = np.array([0.0, 0.0, -1.0], dtype=np.float32)
    assert orc.denoise_normals(img).tobytes() == img.tobytes()
    # max_by_key picks the window with the highest score: the lower-right window (first in the list)
    # sees normals of length 2, the others length 1
    img["normal"] = np.array([0.0, 0.0, 1.0], dtype=np.float32)
    img["normal"][4:7, 4:7] = [0.0, 0.0, 2.0]
    img["normal"][4, 4] = [0.0, 0.0, -1.0]
    out = orc.denoise_normals(img)
    # window (0,0): front-facing mean = 2 (8 px), score = 8*4 + (-1)*2 = 30; the other windows score less
    assert out["normal"][4, 4].tolist() == [0.0, 0.0, 2.0]


def test_blur_ssao_closed_form():
    # effects.rs:98-115, 329-381
    s = np.full((8, 8), 0.5, dtype=np.float32)
    s[0, :] = np.nan
    out = orc.blur_ssao(s)
    assert np.isnan(out[0]).all() and (out[1:] == 0.5).all()
    # an edge: each pixel takes the mean of its least-varying 3x3 window -> the step is preserved
    s = np.zeros((8, 8), dtype=np.float32)
    s[:, 4:] = 1.0
    assert (orc.blur_ssao(s) == s).all()
    # a single NaN-surrounded value keeps itself
    s = np.full((5, 5), np.nan, dtype=np.float32)
    s[2, 2] = 0.25
    assert orc.blur_ssao(s)[2, 2] == np.float32(0.25)


def np_normalize(v):
    n = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2], dtype=np.float32)
    return v / n[..., None]


def np_dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def np_hash(v):
    v = np.asarray(v, dtype=np.uint32)
    with np.errstate(over="ignore"):
        state = v * np.uint32(747796405) + np.uint32(2891336453)
        word = ((state >> ((state >> np.uint32(28)) + np.uint32(4))) ^ state) * np.uint32(277803737)
    return (word >> np.uint32(22)) ^ word


def np_ssao(img, depth, kernel, noise):
    """float32 numpy transcription of compute_pixel_ssao (effects.rs:159-253)."""
    h, w = img.shape
    fw, fh, fd = f32(w), f32(h), f32(depth)
    ys, xs = np.mgrid[0:h, 0:w]
    smin = f32(min(w, h, depth))
    sx, sy, sz = smin / fw, smin / fh, smin / fd
    d = img["depth"]
    p = np.stack([(((xs.astype(f32) + f32(0.5)) / fw) - f32(0.5)) * f32(2),
                  (((ys.astype(f32) + f32(0.5)) / fh) - f32(0.5)) * f32(2),
                  ((d.astype(f32) / fd) - f32(0.5)) * f32(2)], axis=-1).astype(f32)
    with np.errstate(all="ignore"):
        n = np_normalize(img["normal"].astype(f32))
        with np.errstate(over="ignore"):
            ri = np_hash(ys.astype(np.uint32) + np_hash(xs.astype(np.uint32))) % np.uint32(len(noise))
        rvec = np.concatenate([noise[ri], np.zeros((h, w, 1), f32)], axis=-1)
        tangent = np_normalize(rvec - n * np_dot(rvec, n)[..., None])
        bit = np.stack([n[..., 1] * tangent[..., 2] - n[..., 2] * tangent[..., 1],
                        n[..., 2] * tangent[..., 0] - n[..., 0] * tangent[..., 2],
                        n[..., 0] * tangent[..., 1] - n[..., 1] * tangent[..., 0]], axis=-1)
        R = f32(0.1)
        occ = np.zeros((h, w), f32)
        for k in kernel:
            off = tangent * k[0]
            off = bit * k[1] + off
            off = n * k[2] + off
            off = off * R
            off = off * np.array([sx, sy, sz], f32)
            sp = off + p
            px = ((sp[..., 0] / f32(2)) + f32(0.5)) * fw
            py = ((sp[..., 1] / f32(2)) + f32(0.5)) * fh
            ok = (px < fw) & (py < fh) & (px > 0) & (py > 0)
            ix = np.where(ok, px, 0).astype(np.int64)
            iy = np.where(ok, py, 0).astype(np.int64)
            ah = np.where(ok, d[iy, ix], 0)
            az = ((ah.astype(f32) / fd) - f32(0.5)) * f32(2)
            dz = sp[..., 2] - az
            le = sp[..., 2] <= az
            t = (R - (dz - R)) / R
            occ = occ + np.where(dz < R, le.astype(f32), np.where((dz < R * f32(2)) & le, t * t, f32(0)))
        out = f32(1) - occ / f32(len(kernel))
    return np.where(d > 0, out, np.nan).astype(f32)


@pytest.mark.parametrize("shape", [(48, 48, 48), (40, 24, 64)])
def test_ssao_matches_numpy_transcription(shape):
    w, h, depth = shape
    img = sphere_heightmap(max(w, h), depth, np.random.default_rng(3))[:h, :w].copy()
    kernel, noise = ssao_kernel(64), ssao_noise(256)
    got = orc.compute_ssao(img, depth, kernel, noise)
    want = np_ssao(img, depth, kernel, noise)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert got.view(np.uint32).tolist() == want.view(np.uint32).tolist()
    m = ~np.isnan(got)
    assert got[m].min() >= 0.0 and got[m].max() <= 1.0 and got[m].std() > 0.01


def test_ssao_flat_plane_is_unoccluded():
    img = np.zeros((32, 32), dtype=GEO)
    img["depth"] = 16
    img["normal"] = np.array([0, 0, 1], dtype=np.float32)
    s = orc.compute_ssao(img, 32, ssao_kernel(64), ssao_noise(256))
    assert (s[4:-4, 4:-4] == 1.0).all()      # away from the border nothing is above the plane
    img["depth"][0, 0] = 0
    assert np.isnan(orc.compute_ssao(img, 32, ssao_kernel(64), ssao_noise(256))[0, 0])


def np_shade(img, depth, ssao):
    """float32 numpy transcription of shade_pixel (effects.rs:118-154)."""
    h, w = img.shape
    ys, xs = np.mgrid[0:h, 0:w]
    with np.errstate(all="ignore"):
        n = np_normalize(img["normal"].astype(f32))
        p = np.stack([f32(2) * (xs.astype(f32) / f32(w) - f32(0.5)), f32(2) * (ys.astype(f32) / f32(h) - f32(0.5)),
                      f32(2) * (img["depth"].astype(f32) / f32(depth) - f32(0.5))], axis=-1).astype(f32)
        acc = np.full((h, w), 0.2, f32)
        for lx, ly, lz, lw in [(5, -5, 10, 0.5), (-5, 0, 10, 0.15), (0, -5, 10, 0.15)]:
            dirv = np_normalize(np.array([lx, ly, lz], f32) - p)
            dd = np_dot(dirv, n)
            acc = acc + np.where(np.isnan(dd), f32(0), np.maximum(dd, f32(0))) * f32(lw)
        if ssao is not None:
            acc = acc * (ssao * f32(0.6) + f32(0.4))
        acc = np.where(acc < 0, f32(0), acc)
        acc = np.where(acc > 1, f32(1), acc)
        v = acc * f32(255)
        c = np.where(np.isnan(v), 0, np.clip(v, 0, 255)).astype(np.uint8)
    c = np.where(img["depth"] > 0, c, 0).astype(np.uint8)
    return np.repeat(c[..., None], 3, axis=-1)


@pytest.mark.parametrize("with_ssao", [False, True])
def test_shading_matches_numpy_transcription(with_ssao):
    img = sphere_heightmap(64, 64, np.random.default_rng(5))
    ssao = None
    if with_ssao:
        ssao = orc.blur_ssao(orc.compute_ssao(img, 64, ssao_kernel(64), ssao_noise(256)))
    got = orc.apply_shading(img, 64, ssao)
    want = np_shade(img, 64, ssao)
    assert np.array_equal(got, want)
    assert got[img["depth"] == 0].max() == 0 and got.max() > 150


def test_ssao_tables_follow_the_reference_construction():
    # effects.rs:385-440: hemisphere samples of radius (i/(n-1))^2*0.9+0.1, unit rotations
    k, n = ssao_kernel(64), ssao_noise(256)
    r = np.linalg.norm(k.astype(np.float64), axis=1)
    want = (np.arange(64) / 63.0) ** 2 * 0.9 + 0.1
    assert np.allclose(r, want, rtol=1e-5) and (k[:, 2] >= 0).all()
    assert np.allclose(np.linalg.norm(n.astype(np.float64), axis=1), 1.0, rtol=1e-6)
    assert abs(k[:, 0].mean()) < 0.2 and abs(n[:, 0].mean()) < 0.15       # no directional bias
    assert np.array_equal(k, ssao_kernel(64)) and not np.array_equal(k, ssao_kernel(64, seed=1))


def test_hand_derived_ssao_and_shading_cases():
    """Numbers worked out by hand from effects.rs:118-245 (no code in the loop):

    SSAO, 64^3 image, floor at depth 20 with normal (0,0,1), a wall of depth 40 from column 40 on; kernel = one
    sample (1, 0, 0.5), noise = one rotation (1, 0).  With n = (0,0,1): tangent = (1,0,0), bitangent = (0,1,0),
    so the sample sits at p + (0.1, 0, 0.05), i.e. 3.2 pixels to the right and 0.05 above the floor.
      * pixel (37, 32): 37.5 + 3.2 = 40.7 -> column 40, the wall, whose z is 0.625 above the floor: dz = -0.575 < RADIUS
        and sample.z <= wall z -> occlusion 1 of 1 -> SSAO 0.
      * pixel (30, 32): lands on the floor, dz = 0.05 < RADIUS but sample.z > floor z -> occlusion 0 -> SSAO 1.
      * an empty pixel (depth 0) -> NaN.
    Shading at the image centre (p = 0) with n = (0,0,1): 0.2 + 0.5 * 10/sqrt(150) + 2 * 0.15 * 10/sqrt(125)
      = 0.2 + 0.408248 + 0.268328 = 0.876576 -> (0.876576 * 255) as u8 = 223; with SSAO 0 the factor 0.4 gives
      0.350630 * 255 = 89.41 -> 89; with SSAO 1 the factor is 1.0 -> 223 again."""
    n = 64
    img = np.zeros((n, n), dtype=GEO)
    img["depth"] = 20
    img["depth"][:, 40:] = 40
    img["normal"] = (0.0, 0.0, 1.0)
    img["depth"][5, 5] = 0
    kernel = np.array([[1.0, 0.0, 0.5]], dtype=np.float32)
    noise = np.array([[1.0, 0.0]], dtype=np.float32)
    ssao = orc.compute_ssao(img, n, kernel, noise)
    assert ssao[32, 37] == 0.0
    assert ssao[32, 30] == 1.0
    assert np.isnan(ssao[5, 5])
    flat = np.zeros((n, n), dtype=GEO)
    flat["depth"] = n // 2
    flat["normal"] = (0.0, 0.0, 1.0)
    assert orc.apply_shading(flat, n)[n // 2, n // 2].tolist() == [223, 223, 223]
    assert orc.apply_shading(flat, n, ssao=np.zeros((n, n), dtype=np.float32))[n // 2, n // 2].tolist() == [89, 89, 89]
    assert orc.apply_shading(flat, n, ssao=np.ones((n, n), dtype=np.float32))[n // 2, n // 2].tolist() == [223, 223, 223]
