"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle
on identical tapes and inputs."""
import os

import numpy as np
import pytest

import fidget_b200 as fb
from conftest import MODELS as MODELS_DIR, model_text, same_f32

pytestmark = pytest.mark.gpu

EXACT_MODELS = ["prospero.vm", "hi.vm", "quarter.vm", "colonnade.vm", "tanglecube.vm"]


def _pair(orc, cuda, name, n_regs=255):
    text = model_text(name)
    return orc.Tape.from_vm(text, n_regs), fb.CudaShape.from_vm(cuda, text, n_regs)


def _points(n, nv, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(nv)]


@pytest.mark.parametrize("name", EXACT_MODELS)
def test_float_slice_bit_exact(orc, cuda, name):
    ot, gs = _pair(orc, cuda, name)
    pts = _points(4099, ot.n_vars)
    assert same_f32(gs.float_slice_eval(pts), ot.float_slice_eval(pts))


def test_float_slice_bear_tolerance(orc, cuda):
    ot, gs = _pair(orc, cuda, "bear.vm")
    pts = _points(4099, ot.n_vars, 1)
    g, o = gs.float_slice_eval(pts), ot.float_slice_eval(pts)
    assert np.array_equal(np.isnan(g), np.isnan(o))
    m = ~np.isnan(o)
    assert np.all(np.abs(g[m] - o[m]) <= 1e-5 * np.maximum(1.0, np.abs(o[m])))  # north_star tolerance


@pytest.mark.parametrize("n", [0, 1, 3, 4, 7, 8, 9, 31, 33])
def test_float_slice_sizes(orc, cuda, n):
    ot, gs = _pair(orc, cuda, "hi.vm")
    pts = _points(n, ot.n_vars, n)
    assert same_f32(gs.float_slice_eval(pts), ot.float_slice_eval(pts))


def _boxes(n, nv, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, nv)).astype(np.float32)
    w = (rng.uniform(0, 1, (n, nv)) ** 3 * scale).astype(np.float32)
    return np.stack([c - w, c + w], axis=-1).astype(np.float32)


@pytest.mark.parametrize("name", EXACT_MODELS)
def test_interval_bit_exact_with_choices(orc, cuda, name):
    ot, gs = _pair(orc, cuda, name)
    boxes = _boxes(257, ot.n_vars, 3)
    out, ch, simp = gs.interval_eval_batch(boxes, want_choices=True)
    for i in range(boxes.shape[0]):
        o, oc, os_ = ot.interval_eval(boxes[i])
        assert same_f32(out[i, 0], o), (name, i)
        assert np.array_equal(ch[i], oc), (name, i)
        assert bool(simp[i]) == os_


@pytest.mark.parametrize("name", ["prospero.vm", "hi.vm", "colonnade.vm", "bear.vm"])
def test_simplify_matches_reference_length_and_values(orc, cuda, name):
    ot, gs = _pair(orc, cuda, name)
    boxes = _boxes(24, ot.n_vars, 5, scale=0.3)
    for i in range(boxes.shape[0]):
        _, oc, os_ = ot.interval_eval(boxes[i])
        if not os_:
            continue
        child_o = ot.simplify(oc)
        child_g = gs.simplify(oc)
        assert child_g.size() == child_o.size, (name, i)          # Function::size() of the child
        assert child_g.choice_count == child_o.choice_count
        # value parity of the simplified tapes inside the box
        rng = np.random.default_rng(i)
        pts = [rng.uniform(boxes[i, v, 0], boxes[i, v, 1], 513).astype(np.float32) for v in range(ot.n_vars)]
        g, o = child_g.float_slice_eval(pts), child_o.float_slice_eval(pts)
        if name == "bear.vm":
            assert np.allclose(g, o, rtol=1e-5, atol=1e-5, equal_nan=True)
        else:
            assert same_f32(g, o)
        # second-generation simplification keeps agreeing
        sub = boxes[i].copy()
        mid = (sub[:, 0] + sub[:, 1]) / 2
        sub[:, 1] = mid
        o2, oc2, os2 = child_o.interval_eval(sub)
        g2, gc2, gs2 = child_g.interval_eval(sub)
        if name == "bear.vm":  # libdevice vs glibc transcendentals: tolerance, not bits
            assert np.allclose(g2[0], o2, rtol=1e-5, atol=1e-5, equal_nan=True)
            continue
        assert same_f32(g2[0], o2) and np.array_equal(gc2, oc2) and gs2 == os2
        if os2:
            assert child_g.simplify(gc2).size() == child_o.simplify(oc2).size


HI_32 = """
.................#..............
.................#..............
.................#..............
.................#..........##..
.................#..........##..
.................#..............
.................#..............
.................######.....##..
.................###..##....##..
.................##....##...##..
.................#......#...##..
.................#......#...##..
.................#......#...##..
.................#......#...##..
.................#......#...##..
""".strip().split("\n") + ["." * 32] * 17


def test_render2d_hi_golden(cuda):
    # fidget/tests/pixel_render.rs:70-106 (check_hi)
    gs = fb.CudaShape.from_vm(cuda, model_text("hi.vm"))
    img = fb.render2d(gs, fb.RenderConfig2D(32, 32))
    rows = ["".join("#" if b else "." for b in r) for r in fb.pixel_inside(img)]
    assert rows == HI_32


@pytest.mark.parametrize("name,size", [("hi.vm", 256), ("quarter.vm", 256), ("prospero.vm", 512),
                                       ("colonnade.vm", 256), ("prospero.vm", 1000)])
def test_render2d_matches_oracle(orc, cuda, name, size):
    ot, gs = _pair(orc, cuda, name)
    o_img, o_st = orc.render2d(ot, size, size, threads=8)
    g_img, g_st = fb.render2d(gs, fb.RenderConfig2D(size, size), stats=True)
    for k in ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified"):
        assert g_st[k] == o_st[k], k                             # tile masks, level by level
    assert g_st["pixels"] == o_st["pixels"]
    assert np.array_equal(g_img.view(np.uint32), o_img.view(np.uint32))   # pixel-exact incl. fill encodings


def test_render2d_prospero_4096_full(orc, cuda):
    """BASELINE config 2 at full size: bit-identical image and tile census."""
    ot, gs = _pair(orc, cuda, "prospero.vm")
    o_img, o_st = orc.render2d(ot, 4096, 4096, threads=0 or 8)
    g_img, g_st = fb.render2d(gs, fb.RenderConfig2D(4096, 4096), stats=True)
    for k in ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified"):
        assert g_st[k] == o_st[k], k
    assert np.array_equal(g_img.view(np.uint32), o_img.view(np.uint32))
    assert np.array_equal(fb.pixel_inside(g_img), orc.pixel_inside(o_img))


# ---------------------------------------------------------------------------
# 3D: voxel::render (heightmap + normals)
def _sphere_tape(Ctx, r=0.8):
    ctx = Ctx()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    d = ctx.sqrt(ctx.add(ctx.add(ctx.square(x), ctx.square(y)), ctx.square(z)))
    return ctx.tape(ctx.sub(d, ctx.constant(r)))


def _cmp3d(g_img, o_img, exact_normals):
    assert np.array_equal(g_img["depth"], o_img["depth"])
    if exact_normals:
        assert same_f32(g_img["normal"], o_img["normal"])
    else:
        assert np.allclose(g_img["normal"], o_img["normal"], rtol=1e-5, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("size", [64, 128, 200])
def test_render3d_sphere_matches_oracle(orc, cuda, size):
    ot = orc.Tape.from_data(_sphere_tape(orc.Context))
    gs = fb.CudaShape(cuda, _sphere_tape(fb.Context))
    o_img, _ = orc.render3d(ot, size, size, size, threads=8)
    g_img = fb.render3d(gs, fb.RenderConfig3D(size, size, size))
    _cmp3d(g_img, o_img, exact_normals=True)
    assert g_img["depth"].max() > size // 2        # the sphere is there


@pytest.mark.parametrize("name,size,exact", [("prospero.vm", 256, True), ("colonnade.vm", 256, True),
                                             ("hi.vm", 128, True), ("tanglecube.vm", 256, True)])
def test_render3d_models_match_oracle(orc, cuda, name, size, exact):
    ot, gs = _pair(orc, cuda, name)
    o_img, _ = orc.render3d(ot, size, size, size, threads=8)
    g_img = fb.render3d(gs, fb.RenderConfig3D(size, size, size))
    _cmp3d(g_img, o_img, exact)


def test_render3d_bear_within_tolerance(orc, cuda):
    """bear.vm uses exp/ln/sin/cos: libdevice vs glibc differ by ulps, so a razor-edge voxel may
    flip; count mismatching pixels instead of demanding zero (SURVEY.md §7 hard part 7)."""
    ot, gs = _pair(orc, cuda, "bear.vm")
    size = 256
    o_img, _ = orc.render3d(ot, size, size, size, threads=8)
    g_img = fb.render3d(gs, fb.RenderConfig3D(size, size, size))
    same = g_img["depth"] == o_img["depth"]
    assert same.mean() > 0.9995, f"{(~same).sum()} depth mismatches"
    n_ok = np.isclose(g_img["normal"], o_img["normal"], rtol=1e-4, atol=1e-4, equal_nan=True).all(axis=-1)
    assert (n_ok | ~same).mean() > 0.999


def test_render3d_slabs_merge_to_full_image(cuda):
    """Z-slab sharding: slabs rendered independently + fc_merge_slabs == one full render."""
    import ctypes as C
    import torch
    from fidget_b200 import _lib
    gs = fb.CudaShape.from_vm(cuda, model_text("colonnade.vm"))
    size = 256
    full = fb.render3d(gs, fb.RenderConfig3D(size, size, size))
    slabs = []
    for zb in range(0, size, 128):
        t = torch.zeros((size, size, 4), dtype=torch.float32, device="cuda")
        fb.render3d(gs, fb.RenderConfig3D(size, size, size, z_range=(zb, zb + 128), clamp=False), out=t)
        slabs.append(t)
    out = torch.zeros((size, size, 4), dtype=torch.float32, device="cuda")
    ptrs = (C.c_void_p * len(slabs))(*[s.data_ptr() for s in slabs])
    rc = _lib.load().fc_merge_slabs(cuda._h, ptrs, len(slabs), size, size, size, C.c_void_p(out.data_ptr()))
    assert rc == 0
    merged = out.cpu().numpy().view(fb.GEOMETRY_PIXEL).reshape(size, size)
    assert np.array_equal(merged["depth"], full["depth"])
    assert same_f32(merged["normal"], full["normal"])


# ---------------------------------------------------------------------------
# Reference golden images through the CUDA path (fidget/tests/pixel_render.rs:70-385)
import json
import os

_PIX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pixel_render.json")))


def _rows(img):
    return ["".join("#" if b else "." for b in r) for r in fb.pixel_inside(img)]


def _view2(center, scale):
    return np.array([[scale, 0, center[0]], [0, scale, center[1]], [0, 0, 1]], dtype=np.float32)


def test_render3d_ybands_assemble_to_full_image(cuda):
    """3D sharding by bands of root-tile rows (full depth): the bands are disjoint pieces of the full image."""
    shape = fb.CudaShape.from_vm(cuda, model_text("bear.vm"))
    n = 512
    full = fb.render3d(shape, fb.RenderConfig3D(n, n, n))
    out = np.zeros((n, n), dtype=fb.GEOMETRY_PIXEL)
    for r in range(4):
        fb.render3d(shape, fb.RenderConfig3D(n, n, n, root_rows=(r, r + 1)), out=out)
    assert out.tobytes() == full.tobytes()
    with pytest.raises(fb.CudaError):
        fb.render3d(shape, fb.RenderConfig3D(n, n, n, root_rows=(3, 9)))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_interleave_assembles_to_full_image(cuda, world):
    """The sharding bench.py measures at N > 1, on one GPU: every "rank" renders its interleaved root tiles,
    packs its chunk (fc_tiles_pack); the concatenated chunks (what the all-gather delivers) are unpacked
    (fc_tiles_unpack) into the very bytes of a plain full render -- 2D and 3D, ragged sizes included."""
    import ctypes as C
    import torch
    from fidget_b200 import _lib
    from fidget_b200.shard import tiles_per_rank
    lib = _lib.load()
    for dim, name, dims in ((2, "prospero.vm", (1000, 600)), (2, "hi.vm", (512, 512)), (3, "colonnade.vm", (300, 260, 256)),
                            (3, "bear.vm", (512, 512, 512))):
        shape = fb.CudaShape.from_vm(cuda, model_text(name))
        w, h = dims[0], dims[1]
        px = 1 if dim == 2 else 4
        if dim == 2:
            full = torch.zeros((h, w), dtype=torch.float32, device="cuda")
            fb.render2d(shape, fb.RenderConfig2D(w, h), out=full)
        else:
            full = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
            fb.render3d(shape, fb.RenderConfig3D(w, h, dims[2]), out=full)
        per = tiles_per_rank(world, w, h)
        gathered = torch.zeros((world * per, 128, 128, px), dtype=torch.float32, device="cuda")
        for r in range(world):
            img = torch.full_like(full, 7.0)
            if dim == 2:
                fb.render2d(shape, fb.RenderConfig2D(w, h, interleave=(world, r)), out=img)
            else:
                fb.render3d(shape, fb.RenderConfig3D(w, h, dims[2], interleave=(world, r)), out=img)
            assert lib.fc_tiles_pack(cuda._h, C.c_void_p(img.data_ptr()), w, h, 4 * px, 128, world, r,
                                     C.c_void_p(gathered[r * per:].data_ptr())) == 0
        out = torch.full_like(full, 9.0)
        assert lib.fc_tiles_unpack(cuda._h, C.c_void_p(gathered.data_ptr()), w, h, 4 * px, 128, world,
                                   C.c_void_p(out.data_ptr())) == 0
        cuda.synchronize()
        assert torch.equal(out.view(torch.int32), full.view(torch.int32)), (dim, name, world)
    with pytest.raises(fb.CudaError):      # a host image cannot take an interleaved render
        fb.render2d(shape, fb.RenderConfig2D(256, 256, interleave=(2, 0)))
    with pytest.raises(fb.CudaError):
        fb.render2d(shape, fb.RenderConfig2D(256, 256, interleave=(2, 2)), out=torch.zeros((256, 256), device="cuda"))


def test_golden_hi_variants(cuda):
    gs = fb.CudaShape.from_vm(cuda, model_text("hi.vm"))
    assert _rows(fb.render2d(gs, fb.RenderConfig2D(32, 32))) == _PIX["check_hi:EXPECTED"]["rows"]
    assert _rows(fb.render2d(gs, fb.RenderConfig2D(64, 32))) == _PIX["check_hi_wide:EXPECTED"]["rows"]
    cfg = fb.RenderConfig2D(32, 32, world_to_model=_view2((0.5, 0.5), 0.5))
    assert _rows(fb.render2d(gs, cfg)) == _PIX["check_hi_transformed:EXPECTED"]["rows"]
    assert _rows(fb.render2d(gs, cfg)) == _PIX["check_hi_bounded:EXPECTED"]["rows"]


def test_golden_quarter(cuda):
    gs = fb.CudaShape.from_vm(cuda, model_text("quarter.vm"))
    assert _rows(fb.render2d(gs, fb.RenderConfig2D(32, 32))) == _PIX["check_quarter:EXPECTED"]["rows"]


def test_golden_circle_with_bound_var(cuda):
    ctx = fb.Context()
    x, y = ctx.x(), ctx.y()
    r = ctx.sqrt(ctx.add(ctx.square(x), ctx.square(y)))
    c, _ = ctx.var()
    td = ctx.tape(ctx.sub(r, c))
    gs = fb.CudaShape(cuda, td)
    slot = [i for i, (k, _) in enumerate(td.vars()) if k == "v"][0]
    for radius, key in ((0.75, "check_circle_var:EXPECTED_075"), (0.5, "check_circle_var:EXPECTED_05")):
        vv = [0.0] * td.n_vars
        vv[slot] = radius
        assert _rows(fb.render2d(gs, fb.RenderConfig2D(32, 32, var_values=tuple(vv)))) == _PIX[key]["rows"]
    with pytest.raises(fb.CudaError):           # MissingVar
        fb.render2d(gs, fb.RenderConfig2D(32, 32))


def test_golden_neg_infinity_pixel_perfect(cuda):
    ctx = fb.Context()
    gs = fb.CudaShape(cuda, ctx.tape(ctx.constant(float("-inf"))))
    img = fb.render2d(gs, fb.RenderConfig2D(256, 256, pixel_perfect=True))
    assert fb.pixel_inside(img).all()


def test_host_transforms_match_oracle(orc):
    for w, h in ((32, 32), (64, 32), (1000, 500), (4096, 4096)):
        assert np.array_equal(fb.pixel_mat(w, h), orc.pixel_mat(w, h))
        m = _view2((0.5, 0.25), 0.3)
        assert np.array_equal(fb.pixel_mat(w, h, m), orc.pixel_mat(w, h, m))
    assert np.array_equal(fb.voxel_mat(128, 256, 64), orc.voxel_mat(128, 256, 64))


@pytest.mark.parametrize("name", ["hi.vm", "colonnade.vm", "prospero.vm"])
def test_spilled_tapes_evaluate_like_the_oracle(orc, cuda, name):
    """GenericVmFunction<3>-style tapes (Load/Store) through the trait-level evaluators."""
    text = model_text(name)
    ot, gs = orc.Tape.from_vm(text, 3), fb.CudaShape.from_vm(cuda, text, 3)
    assert gs.info.mem_count > 0
    pts = _points(1025, ot.n_vars, 7)
    assert same_f32(gs.float_slice_eval(pts), ot.float_slice_eval(pts))
    boxes = _boxes(65, ot.n_vars, 11)
    out, ch, simp = gs.interval_eval_batch(boxes, want_choices=True)
    for i in range(boxes.shape[0]):
        o, oc, os_ = ot.interval_eval(boxes[i])
        assert same_f32(out[i, 0], o) and np.array_equal(ch[i], oc) and bool(simp[i]) == os_
    with pytest.raises(fb.CudaError):           # renderers refuse spilled tapes loudly
        fb.render2d(gs, fb.RenderConfig2D(64, 64))


def test_tape_using_all_255_registers(orc, cuda):
    """reg_count == 255 is the reference's VmData<255> (registers 0..254): such tapes -- every spilling tape
    at the default register count -- must upload and evaluate like the oracle."""
    tapes = []
    for Ctx in (fb.Context, orc.Context):
        ctx = Ctx()
        x, y, z = ctx.x(), ctx.y(), ctx.z()
        s = ctx.constant(0.0)
        inputs = []
        for i in range(1, 301):                       # 300 values live across the sin: more than 255 registers
            d = ctx.mul(ctx.constant(float(i)), [x, y, z][i % 3])
            inputs.append(d)
            s = ctx.add(s, d)
        s = ctx.sin(s)
        for d in reversed(inputs):
            s = ctx.min(ctx.add(s, d), ctx.mul(d, 0.5))
        tapes.append(ctx.tape(s))
    gs, ot = fb.CudaShape(cuda, tapes[0]), orc.Tape.from_data(tapes[1])
    assert gs.info.reg_count == 255 and gs.info.mem_count > 0
    pts = _points(1025, 3, 5)
    g, o = np.asarray(gs.float_slice_eval(pts)), ot.float_slice_eval(pts)
    assert np.allclose(g, o, rtol=1e-5, atol=1e-5, equal_nan=True)          # one sin: libm tolerance
    boxes = _boxes(33, 3, 9, scale=0.05)
    out, ch, simp = gs.interval_eval_batch(boxes, want_choices=True)
    for i in range(boxes.shape[0]):
        o, oc, os_ = ot.interval_eval(boxes[i])
        assert np.allclose(out[i, 0], o, rtol=1e-5, atol=1e-4, equal_nan=True)
    with pytest.raises(fb.CudaError):
        fb.render2d(gs, fb.RenderConfig2D(64, 64))


def test_render2d_band_leaves_other_rows_untouched(orc, cuda):
    """A band render into a HOST image writes the rows of its band and nothing else."""
    ot, gs = _pair(orc, cuda, "hi.vm")
    want, _ = orc.render2d(ot, 512, 512)
    img = np.full((512, 512), 123.0, dtype=np.float32)
    fb.render2d(gs, fb.RenderConfig2D(512, 512, root_rows=(1, 3)), out=img)
    assert np.all(img[:128] == 123.0) and np.all(img[384:] == 123.0)
    assert np.array_equal(img[128:384].view(np.uint32), want[128:384].view(np.uint32))


def test_grad_slice_matches_oracle(orc, cuda):
    for name, exact in (("prospero.vm", True), ("colonnade.vm", True), ("bear.vm", False)):
        ot, gs = _pair(orc, cuda, name)
        rng = np.random.default_rng(2)
        n = 515
        vars_ = []
        for k in range(ot.n_vars):
            g = np.zeros((n, 4), dtype=np.float32)
            g[:, 0] = rng.uniform(-1, 1, n)
            g[:, 1 + k] = 1.0
            vars_.append(g)
        g, o = gs.grad_slice_eval(vars_), ot.grad_slice_eval(vars_)
        if exact:
            assert same_f32(g, o), name
        else:
            assert np.allclose(g, o, rtol=1e-5, atol=1e-5, equal_nan=True)


def test_point_eval_matches_oracle(orc, cuda):
    ot, gs = _pair(orc, cuda, "hi.vm")
    rng = np.random.default_rng(4)
    for _ in range(32):
        v = rng.uniform(-1, 1, ot.n_vars).astype(np.float32)
        g, gc, gsim = gs.point_eval(v)
        o, oc, osim = ot.point_eval(v)
        assert same_f32(g[0], o) and np.array_equal(gc, oc) and gsim == osim


# ---------------------------------------------------------------------------
# Octree sampler (fidget-mesh Octree::build, sampling half)
def _cmp_leaves(g, o, exact):
    assert len(g) == len(o)
    for k in ("ix", "iy", "iz", "mask", "n_edges", "present"):
        assert np.array_equal(g[k], o[k]), k
    present = ((o["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    if exact:
        assert same_f32(g["pos"][present], o["pos"][present])
        assert same_f32(g["grad"][present], o["grad"][present])
    else:
        assert np.allclose(g["pos"][present], o["pos"][present], rtol=1e-5, atol=1e-6)
        assert np.allclose(g["grad"][present], o["grad"][present], rtol=1e-4, atol=1e-4, equal_nan=True)


@pytest.mark.parametrize("depth", [3, 5])
def test_octree_sphere_matches_oracle(orc, cuda, depth):
    ot = orc.Tape.from_data(_sphere_tape(orc.Context, 0.6))
    gs = fb.CudaShape(cuda, _sphere_tape(fb.Context, 0.6))
    o, ost = orc.octree_sample(ot, depth)
    g, gst = fb.octree_sample(gs, depth, stats=True)
    _cmp_leaves(g, o, exact=True)
    for k in ("evaluated", "full", "empty", "ambiguous"):
        assert gst[k][:depth + 1] == ost[k][:depth + 1], k
    assert (gst["leaf_empty"], gst["leaf_full"], gst["leaf_surface"]) == (ost["leaf_empty"], ost["leaf_full"], ost["leaf_surface"])
    # analytic check (fidget/tests/octree.rs:9-30 style): intersections lie on the sphere, gradients are unit normals
    present = ((g["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    r = np.linalg.norm(g["pos"][present], axis=-1)
    assert np.all(np.abs(r - 0.6) < 2.0 / 2 ** depth / 16 ** 3 + 1e-5)
    n = g["grad"][present][:, :3]
    assert np.allclose(np.linalg.norm(n, axis=-1), 1.0, atol=1e-4)


def test_octree_colonnade_and_transform(orc, cuda):
    ot, gs = _pair(orc, cuda, "colonnade.vm")
    o, _ = orc.octree_sample(ot, 5)
    g = fb.octree_sample(gs, 5)
    _cmp_leaves(g, o, exact=True)
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = m[1, 1] = m[2, 2] = 0.75
    m[0, 3] = 0.125
    o, _ = orc.octree_sample(ot, 4, world_to_model=m)
    g = fb.octree_sample(gs, 4, world_to_model=m)
    _cmp_leaves(g, o, exact=True)


def test_octree_gyroid_sphere(orc, cuda):
    """BASELINE config 4's model at a depth the oracle finishes quickly; sin/cos => tolerance on values,
    and cells whose corner samples sit within an ulp of zero may flip, so compare the common leaves."""
    ot, gs = _pair(orc, cuda, "gyroid-sphere.vm")
    o, _ = orc.octree_sample(ot, 6)
    g = fb.octree_sample(gs, 6)
    key = lambda a: (a["iz"].astype(np.int64) << 32) | (a["iy"].astype(np.int64) << 16) | a["ix"]
    ko, kg = key(o), key(g)
    common = np.intersect1d(ko, kg)
    assert len(common) > 0.999 * max(len(o), len(g))
    oc, gc = o[np.isin(ko, common)], g[np.isin(kg, common)]
    same_mask = oc["mask"] == gc["mask"]
    assert same_mask.mean() > 0.999
    oc, gc = oc[same_mask], gc[same_mask]
    present = ((oc["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    cell = 2.0 / 2 ** 6
    assert np.all(np.abs(gc["pos"][present] - oc["pos"][present]) <= cell / 1000)   # the 16^4-ary search may land one bracket apart
    assert np.isclose(gc["grad"][present], oc["grad"][present], rtol=1e-2, atol=1e-2).mean() > 0.999


def test_render_into_pinned_host_memory(cuda, monkeypatch):
    """Host output buffers: staged + DMA copy by default, written directly by the kernels when
    FIDGET_B200_ZEROCOPY=1 (page-locked memory); both give the same image."""
    import torch
    for zc in ("0", "1"):
        monkeypatch.setenv("FIDGET_B200_ZEROCOPY", zc)
        _check_pinned(cuda)


def _check_pinned(cuda):
    import torch
    gs = fb.CudaShape.from_vm(cuda, model_text("prospero.vm"))
    ref = fb.render2d(gs, fb.RenderConfig2D(1024, 1024))
    pinned = torch.empty((1024, 1024), dtype=torch.float32).pin_memory()
    pinned.fill_(123.0)
    fb.render2d(gs, fb.RenderConfig2D(1024, 1024), out=pinned.numpy())
    assert np.array_equal(pinned.numpy().view(np.uint32), ref.view(np.uint32))
    col = fb.CudaShape.from_vm(cuda, model_text("colonnade.vm"))
    ref3 = fb.render3d(col, fb.RenderConfig3D(256, 256, 256))
    p3 = torch.zeros((256, 256, 4), dtype=torch.float32).pin_memory()
    fb.render3d(col, fb.RenderConfig3D(256, 256, 256), out=p3.numpy())
    assert np.array_equal(p3.numpy().view(np.uint32).reshape(256, 256, 4), ref3.view(np.uint32).reshape(256, 256, 4))


# ---------------------------------------------------------------------------
# Edge cases and error behaviour of the C ABI
@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (100, 37), (129, 257), (640, 480)])
def test_render2d_ragged_sizes(orc, cuda, w, h):
    ot, gs = _pair(orc, cuda, "hi.vm")
    o_img, o_st = orc.render2d(ot, w, h)
    g_img, g_st = fb.render2d(gs, fb.RenderConfig2D(w, h), stats=True)
    assert np.array_equal(g_img.view(np.uint32), o_img.view(np.uint32))
    assert g_st["evaluated"] == o_st["evaluated"] and g_st["pixels"] == o_st["pixels"]


@pytest.mark.parametrize("ts", [(128, 16), (64, 8), (256, 32, 8), (8,), (96, 24, 6), (48, 12), (100, 20, 5), (36, 6, 3)])
def test_render2d_custom_tile_sizes(orc, cuda, ts):
    # EvalConfig::tile_sizes (pixel.rs:42-57); (128, 16) is the JIT's default (fidget-jit/src/lib.rs:984)
    ot, gs = _pair(orc, cuda, "prospero.vm")
    o_img, o_st = orc.render2d(ot, 512, 512, tile_sizes=ts, threads=8)
    g_img, g_st = fb.render2d(gs, fb.RenderConfig2D(512, 512, tile_sizes=ts), stats=True)
    assert np.array_equal(g_img.view(np.uint32), o_img.view(np.uint32))
    assert g_st["evaluated"] == o_st["evaluated"]


def test_render2d_pixel_perfect_and_z(orc, cuda):
    ot, gs = _pair(orc, cuda, "colonnade.vm")
    for z in (0.0, 0.3):
        o_img, _ = orc.render2d(ot, 256, 256, z=z, pixel_perfect=True, threads=8)
        g_img = fb.render2d(gs, fb.RenderConfig2D(256, 256, z=z, pixel_perfect=True))
        assert np.array_equal(g_img.view(np.uint32), o_img.view(np.uint32))


def test_render3d_ragged_volume(orc, cuda):
    ot = orc.Tape.from_data(_sphere_tape(orc.Context, 0.7))
    gs = fb.CudaShape(cuda, _sphere_tape(fb.Context, 0.7))
    o_img, _ = orc.render3d(ot, 100, 60, 90, threads=8)
    g_img = fb.render3d(gs, fb.RenderConfig3D(100, 60, 90))
    _cmp3d(g_img, o_img, exact_normals=True)


def test_bad_inputs_fail_loudly(cuda):
    import ctypes as C
    from fidget_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    words = (C.c_uint32 * 4)(0xFFFFFFFF, 0, 0xFFFFFFFF, 0xFFFFFFFE)          # broken end marker
    assert lib.fc_tape_create(cuda._h, words, 4, 1, 0, 1, 1, 0, C.byref(h)) == -1
    assert b"marker" in lib.fc_last_error()
    words = (C.c_uint32 * 6)(0xFFFFFFFF, 0, 0x00FFFF63, 0, 0xFFFFFFFF, 0xFFFFFFFF)   # opcode 0x63
    assert lib.fc_tape_create(cuda._h, words, 6, 1, 0, 1, 1, 0, C.byref(h)) == -1
    gs = fb.CudaShape.from_vm(cuda, model_text("hi.vm"))
    with pytest.raises(fb.CudaError):
        gs.simplify(np.zeros(3, dtype=np.uint8))                               # BadChoiceSlice
    with pytest.raises(fb.CudaError):
        gs.simplify(np.zeros(gs.choice_count, dtype=np.uint8))                 # Choice::Unknown in a trace
    with pytest.raises(fb.CudaError):
        fb.render2d(gs, fb.RenderConfig2D(64, 64, tile_sizes=(8, 32)))         # TileSizeError::BadTileOrder
    with pytest.raises(fb.CudaError):
        fb.render2d(gs, fb.RenderConfig2D(0, 64))
    cuda.set_arena_bytes(1 << 20)                                              # arena far too small for prospero
    try:
        big = fb.CudaShape.from_vm(cuda, model_text("prospero.vm"))
        with pytest.raises(fb.CudaError) as e:
            fb.render2d(big, fb.RenderConfig2D(2048, 2048))
        assert e.value.code == -4                                              # FC_ERR_ARENA, not a silent fallback
    finally:
        cuda.set_arena_bytes(1 << 30)
    img = fb.render2d(fb.CudaShape.from_vm(cuda, model_text("prospero.vm")), fb.RenderConfig2D(512, 512))
    assert fb.pixel_inside(img).any()


def test_constant_and_single_axis_shapes(orc, cuda):
    for build in (lambda c: c.constant(1.5), lambda c: c.x(), lambda c: c.sub(c.y(), c.constant(0.25)),
                  lambda c: c.neg(c.z())):
        o_ctx, g_ctx = orc.Context(), fb.Context()
        ot = orc.Tape.from_data(o_ctx.tape(build(o_ctx)))
        gs = fb.CudaShape(cuda, g_ctx.tape(build(g_ctx)))
        o_img, _ = orc.render2d(ot, 96, 96)
        assert np.array_equal(fb.render2d(gs, fb.RenderConfig2D(96, 96)).view(np.uint32), o_img.view(np.uint32))
        o3, _ = orc.render3d(ot, 64, 64, 64)
        g3 = fb.render3d(gs, fb.RenderConfig3D(64, 64, 64))
        _cmp3d(g3, o3, exact_normals=True)


def test_c_client_renders_like_the_python_face(cuda, tmp_path):
    """examples/render2d.c (plain C over the C ABI: fh_* front end + fc_tape_create + fc_render2d) produces the
    same RawDistancePixel words as fidget_b200.render2d."""
    import subprocess
    from test_capi_symbols import build_c_example
    exe = build_c_example(tmp_path)
    for name, size in (("hi.vm", 64), ("prospero.vm", 512)):
        out = subprocess.check_output([str(exe), os.path.join(MODELS_DIR, name), str(size)], text=True)
        img = fb.render2d(fb.CudaShape.from_vm(cuda, model_text(name)), fb.RenderConfig2D(size, size))
        h = 1469598103934665603
        for b in img.view(np.uint32).ravel().tolist():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        assert f"inside {int(fb.pixel_inside(img).sum())} px" in out, out
        assert f"fnv1a {h:016x}" in out, out


@pytest.mark.parametrize("n", [4096, 4097, 5119, 8192, 100003])
def test_slice_tma_path_matches_per_thread_path_and_oracle(orc, cuda, monkeypatch, n):
    """fc_float_slice_eval / fc_grad_slice_eval take the TMA-fed persistent kernel (bulk.cu) for n >= 4096 on
    tapes without spills: same bits as the per-thread kernel (FIDGET_B200_NO_TMA=1) and as the oracle, full
    tiles and ragged tails, host and device slices."""
    import torch
    for name, exact in (("hi.vm", True), ("colonnade.vm", True), ("bear.vm", False)):
        ot, gs = _pair(orc, cuda, name)
        pts = _points(n, ot.n_vars, n)
        want = ot.float_slice_eval(pts)
        monkeypatch.setenv("FIDGET_B200_NO_TMA", "0")
        fast = np.asarray(gs.float_slice_eval(pts))
        dev_pts = [torch.from_numpy(p).cuda() for p in pts]
        fast_dev = gs.float_slice_eval(dev_pts).cpu().numpy()
        monkeypatch.setenv("FIDGET_B200_NO_TMA", "1")
        slow = np.asarray(gs.float_slice_eval(pts))
        assert same_f32(fast, slow) and same_f32(fast_dev, slow), name
        if exact:
            assert same_f32(fast, want), name
        else:
            assert np.allclose(fast, want, rtol=1e-5, atol=1e-5, equal_nan=True)
        # gradients
        rng = np.random.default_rng(n)
        vars_ = []
        for k in range(ot.n_vars):
            g = np.zeros((n, 4), dtype=np.float32)
            g[:, 0] = pts[k]
            g[:, 1 + (k % 3)] = 1.0
            g[:, 1:] += rng.uniform(-1, 1, (n, 3)).astype(np.float32)
            vars_.append(g)
        gwant = ot.grad_slice_eval(vars_)
        gslow = np.asarray(gs.grad_slice_eval(vars_))
        monkeypatch.setenv("FIDGET_B200_NO_TMA", "0")
        gfast = np.asarray(gs.grad_slice_eval(vars_))
        assert same_f32(gfast, gslow), name
        if exact:
            assert same_f32(gfast, gwant), name
        else:   # libm ulps amplified by bear's divisions: count the outliers instead of demanding none
            ok = np.isclose(gfast, gwant, rtol=1e-4, atol=1e-4, equal_nan=True).all(axis=-1)
            assert ok.mean() > 0.999, ok.mean()


@pytest.mark.parametrize("w,h", [(256, 256), (1000, 37), (33, 65), (4096, 4096)])
def test_render2d_output_formats(orc, cuda, w, h):
    """fc_render2d_cfg.out_format: the byte mask, the 1-bit bitmap and the RGBA8 bitmap are exactly
    RawDistancePixel::inside / effects::to_rgba_bitmap of the distance image -- host and device outputs."""
    import torch
    name = "prospero.vm" if w == 4096 else "hi.vm"
    ot, gs = _pair(orc, cuda, name)
    o_img, _ = orc.render2d(ot, w, h, threads=8)
    inside = orc.pixel_inside(o_img)
    f32 = fb.render2d(gs, fb.RenderConfig2D(w, h))
    assert np.array_equal(f32.view(np.uint32), o_img.view(np.uint32))
    mask = fb.render2d(gs, fb.RenderConfig2D(w, h, out_format="mask_u8"))
    assert np.array_equal(mask, np.where(inside, 255, 0).astype(np.uint8))
    bits = fb.render2d(gs, fb.RenderConfig2D(w, h, out_format="bitmap_1bit"))
    assert np.array_equal(bits, np.packbits(inside, axis=1, bitorder="little"))
    rgba = fb.render2d(gs, fb.RenderConfig2D(w, h, out_format="rgba8"))
    assert np.array_equal(rgba, orc.to_rgba_bitmap(o_img))
    dev = torch.zeros((h, (w + 7) // 8), dtype=torch.uint8, device="cuda")
    fb.render2d(gs, fb.RenderConfig2D(w, h, out_format="bitmap_1bit"), out=dev)
    assert np.array_equal(dev.cpu().numpy(), bits)
    with pytest.raises(fb.CudaError):
        fb.render2d(gs, fb.RenderConfig2D(w, h, out_format="mask_u8", root_rows=(0, 1)))


def test_tape_blob_round_trip(orc, cuda):
    """fc_tape_serialize / fc_tape_deserialize (the on-disk / wire form of a tape): a shape loaded from a blob
    written by the host front end, or by another shape, evaluates and renders identically; damaged blobs fail."""
    text = model_text("prospero.vm")
    ctx, root = fb.Context.from_text(text)
    td = ctx.tape(root)
    a = fb.CudaShape(cuda, td)
    blob = td.serialize()
    assert a.serialize() == blob
    b = fb.CudaShape.from_blob(cuda, blob)
    assert (b.size(), b.choice_count, b.n_vars, b._axes) == (a.size(), a.choice_count, a.n_vars, tuple(td.var_slots()))
    pts = _points(5000, a.n_vars, 3)
    assert same_f32(a.float_slice_eval(pts), b.float_slice_eval(pts))
    ia, ib = fb.render2d(a, fb.RenderConfig2D(512, 512)), fb.render2d(b, fb.RenderConfig2D(512, 512))
    assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32))
    # a simplified child survives the trip too (alias copies come back as plain copies: same values)
    box = _boxes(1, a.n_vars, 5, scale=0.1)[0]
    _, ch, simp = a.interval_eval(box)
    if simp:
        child = a.simplify(ch)
        again = fb.CudaShape.from_blob(cuda, child.serialize())
        sub = [np.random.default_rng(k).uniform(box[k, 0], box[k, 1], 4100).astype(np.float32) for k in range(a.n_vars)]
        assert same_f32(child.float_slice_eval(sub), again.float_slice_eval(sub))
    for bad in (b"", blob[:20], b"XTAP" + blob[4:], blob[:4] + b"\x02\x00\x00\x00" + blob[8:], blob[:-8]):
        with pytest.raises(fb.CudaError):
            fb.CudaShape.from_blob(cuda, bad)


def test_solver_style_gradient_batches(orc, cuda):
    """fidget-solver's use of the gradient evaluator (fidget-solver/src/lib.rs:30-36,191): a function of more
    than three free variables is differentiated three variables at a time over a batch of points -- every
    partial must match the oracle, and the value column must not depend on which triple carries the seeds."""
    tapes = []
    for Ctx in (fb.Context, orc.Context):
        ctx = Ctx()
        vs = [ctx.var()[0] for _ in range(5)] + [ctx.x()]
        e = ctx.sub(ctx.add(ctx.mul(vs[0], vs[1]), ctx.square(ctx.sub(vs[2], vs[5]))), ctx.mul(vs[3], 0.5))
        e = ctx.max(e, ctx.div(vs[4], ctx.add(ctx.square(vs[0]), 1.0)))
        tapes.append(ctx.tape(e))
    g, o = fb.CudaShape(cuda, tapes[0]), orc.Tape.from_data(tapes[1])
    nv, n = tapes[0].n_vars, 6000
    rng = np.random.default_rng(12)
    vals = rng.uniform(-2, 2, (nv, n)).astype(np.float32)
    values = None
    for first in range(0, nv, 3):
        vars_ = []
        for k in range(nv):
            a = np.zeros((n, 4), dtype=np.float32)
            a[:, 0] = vals[k]
            if first <= k < first + 3:
                a[:, 1 + k - first] = 1.0
            vars_.append(a)
        gg, og = np.asarray(g.grad_slice_eval(vars_)), o.grad_slice_eval(vars_)
        assert same_f32(gg, og), first
        values = gg[:, 0] if values is None else values
        assert same_f32(gg[:, 0], values)


@pytest.mark.parametrize("name,size,ts", [("prospero.vm", 1024, ()), ("prospero.vm", 512, (128, 16)), ("hi.vm", 300, ()),
                                          ("colonnade.vm", 512, (256, 64, 16, 4)), ("prospero.vm", 2048, (64, 8)),
                                          ("bear.vm", 256, (32, 16, 8, 4, 2))])
def test_fused_tail_equals_per_level_launches(orc, cuda, name, size, ts):
    """FC_FLAG_FUSED_TAIL (experimental) runs the levels after the root level, the leaf pixels and the fills as one
    persistent launch draining a dependency-ordered queue (tail2d.cu) instead of one launch per stage.  Same image,
    same census, and both equal the oracle (bear: libm, so the two CUDA paths are compared with each other)."""
    ot, gs = _pair(orc, cuda, name)
    a, sa = fb.render2d(gs, fb.RenderConfig2D(size, size, tile_sizes=ts, fused_tail=True), stats=True)
    b, sb = fb.render2d(gs, fb.RenderConfig2D(size, size, tile_sizes=ts), stats=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for k in ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified", "pixels"):
        assert sa[k] == sb[k], k
    assert sa["kernel_launches"] < sb["kernel_launches"]
    if name != "bear.vm":
        o, so = orc.render2d(ot, size, size, tile_sizes=ts or (128, 32, 8), threads=8)
        assert np.array_equal(a.view(np.uint32), o.view(np.uint32))
        assert sa["evaluated"] == so["evaluated"] and sa["simplified"] == so["simplified"]
    for _ in range(3):      # the queue is rebuilt per render: repeated frames stay identical
        c = fb.render2d(gs, fb.RenderConfig2D(size, size, tile_sizes=ts, fused_tail=True))
        assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


@pytest.mark.parametrize("name,size", [("prospero.vm", 256), ("colonnade.vm", 256), ("tanglecube.vm", 256), ("hi.vm", 128),
                                       ("colonnade.vm", 512), ("sphere", 128)])
def test_render3d_exact_census_matches_the_reference_walk(orc, cuda, name, size):
    """FC_FLAG_EXACT_CENSUS: the per-level tile census (evaluated / filled inside / filled outside / ambiguous /
    simplified) and the number of voxels evaluated are those of the reference's front-to-back, depth-first walk
    with its "every pixel already filled" early exit (voxel.rs:244-357) -- although the device evaluates the levels
    breadth first without that culling.  Equal to the oracle's counts, level by level."""
    if name == "sphere":
        ot = orc.Tape.from_data(_sphere_tape(orc.Context, 0.7))
        gs = fb.CudaShape(cuda, _sphere_tape(fb.Context, 0.7))
    else:
        ot, gs = _pair(orc, cuda, name)
    o_img, o_st = orc.render3d(ot, size, size, size, threads=8)
    g_img, g_st = fb.render3d(gs, fb.RenderConfig3D(size, size, size, exact_census=True), stats=True)
    _cmp3d(g_img, o_img, True)
    for k in ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified"):
        assert g_st[k] == o_st[k], (k, g_st[k], o_st[k])
    assert g_st["pixels"] == o_st["pixels"]
    # without the flag the census is what the device evaluated: a superset (on the same ladder of tile sizes)
    _, raw = fb.render3d(gs, fb.RenderConfig3D(size, size, size, full_ladder=True), stats=True)
    assert all(a >= b for a, b in zip(raw["evaluated"], o_st["evaluated"]))
    with pytest.raises(fb.CudaError):
        fb.render3d(gs, fb.RenderConfig3D(100, 100, 100, exact_census=True), stats=True)


@pytest.mark.parametrize("name,size", [("bear.vm", 512), ("prospero.vm", 512), ("colonnade.vm", 256)])
def test_render3d_occlusion_culling_changes_nothing_but_the_work(orc, cuda, name, size, monkeypatch):
    """Parents whose 16 x 16 pixel blocks are all finished in front of them (occlusion map raised by interval-proven
    tiles) are skipped, like the reference's front-to-back walk skips them (voxel.rs:283-293).  The image is the
    same bit for bit with and without; the device census shrinks (or stays) and still covers the reference's."""
    ot, gs = _pair(orc, cuda, name)
    cfg = fb.RenderConfig3D(size, size, size, full_ladder=True)
    monkeypatch.setenv("FIDGET_B200_NO_CULL", "1")
    plain, st_plain = fb.render3d(gs, cfg, stats=True)
    monkeypatch.setenv("FIDGET_B200_NO_CULL", "0")
    culled, st_cull = fb.render3d(gs, cfg, stats=True)
    assert np.array_equal(plain.view(np.uint32), culled.view(np.uint32))
    assert all(a <= b for a, b in zip(st_cull["evaluated"], st_plain["evaluated"]))
    _, o_st = orc.render3d(ot, size, size, size, threads=8)
    assert all(a >= b for a, b in zip(st_cull["evaluated"], o_st["evaluated"]))
    if name == "bear.vm":
        assert sum(st_cull["evaluated"]) < sum(st_plain["evaluated"])


@pytest.mark.parametrize("name,size", [("prospero.vm", 512), ("bear.vm", 512), ("colonnade.vm", 256), ("tanglecube.vm", 256),
                                       ("hi.vm", 128), ("prospero.vm", 200)])
def test_render3d_device_ladder_gives_the_reference_ladders_image(cuda, name, size, models):
    """With the default tile sizes the device evaluates (128, 32, 8) instead of the reference's (128, 64, 32, 16, 8):
    a 4 x 4 x 4 split on the parent's tape.  Depth and normals are the same bit for bit (a choice decided on a region
    is decided the same way on every sub-region, and the pruned branch never contributed to the value), also with an
    explicit mixed ladder; only the census describes different levels.  (The slab, band and interleave tests run on
    the device ladder; the full-size tests compare it with the oracle's reference ladder.)"""
    gs = fb.CudaShape.from_vm(cuda, models(name))
    full, st_full = fb.render3d(gs, fb.RenderConfig3D(size, size, size, full_ladder=True), stats=True)
    dev, st_dev = fb.render3d(gs, fb.RenderConfig3D(size, size, size), stats=True)
    assert np.array_equal(full.view(np.uint32), dev.view(np.uint32))
    assert st_dev["evaluated"][0] == st_full["evaluated"][0] and st_dev["pixels"] == pytest.approx(st_full["pixels"], rel=0.02)
    n_full = sum(1 for e in st_full["evaluated"] if e)
    n_dev = sum(1 for e in st_dev["evaluated"] if e)
    assert n_dev < n_full or n_full <= 2
    explicit = fb.render3d(gs, fb.RenderConfig3D(size, size, size, tile_sizes=(128, 64, 16, 8)))
    assert np.array_equal(full.view(np.uint32), explicit.view(np.uint32))


def _edge_slab_and_ball(Ctx):
    """Inside for x > 0.9 and z > 0.3 (a slab hugging the right edge, in front), plus a small ball at the left, behind."""
    ctx = Ctx()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    slab = ctx.max(ctx.sub(ctx.constant(0.9), x), ctx.sub(ctx.constant(0.3), z))
    dx, dy, dz = ctx.add(x, ctx.constant(0.85)), ctx.sub(y, ctx.constant(0.1)), ctx.add(z, ctx.constant(0.5))
    ball = ctx.sub(ctx.sqrt(ctx.add(ctx.add(ctx.square(dx), ctx.square(dy)), ctx.square(dz))), ctx.constant(0.12))
    return ctx.tape(ctx.min(slab, ball))


@pytest.mark.parametrize("size", [(100, 100, 100), (140, 100, 120), (172, 172, 100), (104, 104, 72)])
@pytest.mark.parametrize("full_ladder", [False, True])
def test_render3d_ragged_image_with_full_tiles_hanging_over_the_edge(orc, cuda, size, full_ladder):
    """Interval-proven tiles that overhang the right / bottom edge of a ragged image must not touch occlusion blocks
    outside it (an unchecked block column wraps into the next block row and would cull the ball at the left)."""
    w, h, d = size
    ot = orc.Tape.from_data(_edge_slab_and_ball(orc.Context))
    gs = fb.CudaShape(cuda, _edge_slab_and_ball(fb.Context))
    o_img, _ = orc.render3d(ot, w, h, d, threads=8)
    g_img = fb.render3d(gs, fb.RenderConfig3D(w, h, d, full_ladder=full_ladder))
    _cmp3d(g_img, o_img, exact_normals=True)
    assert (o_img["depth"][:, : w // 4] > 0).any()        # the ball is in the picture
