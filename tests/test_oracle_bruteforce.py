"""The oracle's tile renderers against brute force: every pixel / voxel evaluated one by one with the
(separately pinned) float and gradient evaluators.  This is what the tile recursion, the interval
proofs and the chain of tape simplifications must reproduce exactly (pixel.rs:316-440, voxel.rs:244-553)."""
import ctypes as C

import numpy as np
import pytest

from conftest import model_text, same_f32
from test_gpu_fuzz import random_shape


def model_points(orc, mat, xs, ys, zs):
    """transform_f32 of the oracle for every (x, y, z): the model-space point a pixel / voxel is evaluated at."""
    m = np.ascontiguousarray(mat, dtype=np.float32).reshape(16)
    fp = C.POINTER(C.c_float)
    out = np.zeros((len(xs), 3), dtype=np.float32)
    tmp = np.zeros(3, dtype=np.float32)
    f = orc.lib().orc_transform_f32
    for i, (x, y, z) in enumerate(zip(xs, ys, zs)):
        f(m.ctypes.data_as(fp), float(x), float(y), float(z), tmp.ctypes.data_as(fp))
        out[i] = tmp
    return out


def axis_columns(t, pts):
    """SoA inputs in the tape's own input-slot order (ShapeTape::vars, shape/mod.rs:355-376)."""
    slots = t.data.var_slots()
    cols = [np.zeros(len(pts), dtype=np.float32) for _ in range(max(t.n_vars, 1))]
    for axis, slot in enumerate(slots):
        if slot >= 0:
            cols[slot] = np.ascontiguousarray(pts[:, axis])
    return cols, slots


def tapes(orc):
    yield "hi.vm", orc.Tape.from_vm(model_text("hi.vm")), None
    yield "quarter.vm", orc.Tape.from_vm(model_text("quarter.vm")), None
    yield "prospero.vm", orc.Tape.from_vm(model_text("prospero.vm")), None
    for seed in range(4):
        ctx = orc.Context()
        td = ctx.tape(random_shape(ctx, np.random.default_rng(40 + seed), 20 + 10 * seed, use_z=False))
        yield f"csg{seed}", orc.Tape.from_data(td), td


@pytest.mark.parametrize("size,tile_sizes", [(64, (32, 8)), (96, (32, 16, 8))])
def test_render2d_equals_brute_force(orc, size, tile_sizes):
    for name, t, _ in tapes(orc):
        img, st = orc.render2d(t, size, size, tile_sizes=tile_sizes)
        ys, xs = np.mgrid[0:size, 0:size]
        pts = model_points(orc, orc.pixel_mat(size, size), xs.ravel(), ys.ravel(), np.zeros(size * size))
        brute = t.float_slice_eval(axis_columns(t, pts)[0]).reshape(size, size)
        bits = img.view(np.uint32)
        is_fill = np.isnan(img) & ((bits & np.uint32(0xFF << 9)) == np.uint32(0xF6 << 9))
        # interval-proven tiles: the proof is sound at every pixel; shaded pixels: the simplified leaf tape
        # returns what the root tape returns
        assert np.array_equal(orc.pixel_inside(img), brute < 0), name
        assert same_f32(img[~is_fill], brute[~is_fill]), name
        assert is_fill.sum() + st["pixels"] == size * size, name


@pytest.mark.parametrize("name", ["sphere", "bear.vm", "csg"])
def test_render3d_equals_brute_force(orc, name):
    size = 32
    if name == "sphere":
        ctx = orc.Context()
        x, y, z = ctx.x(), ctx.y(), ctx.z()
        td = ctx.tape(ctx.sub(ctx.sqrt(ctx.add(ctx.add(ctx.square(x), ctx.square(y)), ctx.square(z))), 0.6))
        t = orc.Tape.from_data(td)
    elif name == "csg":
        ctx = orc.Context()
        td = ctx.tape(random_shape(ctx, np.random.default_rng(7), 12, use_z=True))
        t = orc.Tape.from_data(td)
    else:
        t = orc.Tape.from_vm(model_text(name))
    img, _ = orc.render3d(t, size, size, size, tile_sizes=(16, 8))
    zs, ys, xs = np.mgrid[0:size, 0:size, 0:size]
    pts = model_points(orc, orc.voxel_mat(size, size, size), xs.ravel(), ys.ravel(), zs.ravel())
    vals = t.float_slice_eval(axis_columns(t, pts)[0]).reshape(size, size, size)
    inside = vals < 0                                                  # [z, y, x]
    top = np.where(inside.any(axis=0), size - np.argmax(inside[::-1], axis=0), 0)   # highest inside voxel + 1
    # voxel.rs:535-546: a column that reaches the top of the volume saturates to depth = size
    assert np.array_equal(np.minimum(img["depth"], size), np.minimum(top, size)), name
    # normals: gradient at the surface voxel (voxel.rs:449-481); saturated columns get [0, 0, 1]
    hit = (img["depth"] > 0) & (img["depth"] < size - 1)
    yy, xx = np.nonzero(hit)
    zz = img["depth"][hit].astype(np.int64) - 1
    p = model_points(orc, orc.voxel_mat(size, size, size), xx, yy, zz)
    # the renderer differentiates with respect to the VOXEL coordinates: the seeds are the rows of the (affine)
    # screen-to-model matrix (Transformable for Grad, shape/mod.rs:918-948)
    M = np.asarray(orc.voxel_mat(size, size, size), dtype=np.float32).reshape(4, 4)
    slots = t.data.var_slots()
    vars_ = [np.zeros((len(xx), 4), dtype=np.float32) for _ in range(max(t.n_vars, 1))]
    for axis, slot in enumerate(slots):
        if slot >= 0:
            vars_[slot][:, 0] = p[:, axis]
            vars_[slot][:, 1:4] = M[axis, :3]
    grads = t.grad_slice_eval(vars_)
    assert same_f32(img["normal"][hit], grads[:, 1:4]), name


@pytest.mark.parametrize("name", ["sphere", "colonnade.vm", "csg"])
def test_octree_leaves_equal_brute_force_corner_masks(orc, name):
    """OctreeBuilder::leaf (octree.rs:590-640): a depth-D cell is a surface leaf iff its 8 corner samples differ
    in sign; bit i of the mask is set when corner i (x = bit 0, y = bit 1, z = bit 2) is inside.  The interval
    descent may only discard cells whose corners all agree."""
    depth = 4
    if name == "sphere":
        ctx = orc.Context()
        x, y, z = ctx.x(), ctx.y(), ctx.z()
        t = orc.Tape.from_data(ctx.tape(ctx.sub(ctx.sqrt(ctx.add(ctx.add(ctx.square(x), ctx.square(y)), ctx.square(z))), 0.6)))
    elif name == "csg":
        ctx = orc.Context()
        t = orc.Tape.from_data(ctx.tape(random_shape(ctx, np.random.default_rng(11), 10, use_z=True)))
    else:
        t = orc.Tape.from_vm(model_text(name))
    leaves, _ = orc.octree_sample(t, depth)
    n = 1 << depth
    g = (np.arange(n + 1, dtype=np.float32) * np.float32(2.0 / n) - np.float32(1.0)).astype(np.float32)   # CellBounds: dyadic, exact
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=-1)
    inside = (t.float_slice_eval(axis_columns(t, pts)[0]) < 0).reshape(n + 1, n + 1, n + 1)     # [z, y, x]
    mask = np.zeros((n, n, n), dtype=np.uint16)
    for i in range(8):
        dx, dy, dz = i & 1, (i >> 1) & 1, (i >> 2) & 1
        mask |= inside[dz:dz + n, dy:dy + n, dx:dx + n].astype(np.uint16) << i
    surface = (mask != 0) & (mask != 255)
    want = {(int(x), int(y), int(z)): int(mask[z, y, x]) for z, y, x in zip(*np.nonzero(surface))}
    got = {(int(l["ix"]), int(l["iy"]), int(l["iz"])): int(l["mask"]) for l in leaves}
    assert got == want, (name, len(got), len(want))
    assert len(got) > 50


def test_octree_edge_intersections_bracket_the_surface(orc):
    """OctreeBuilder::leaf edge search (octree.rs:642-740): four rounds of 16-ary search leave a bracket of
    1/65536 of the edge, the intersection is its midpoint: it lies ON the cell edge, and the field changes sign
    (or vanishes) within half a bracket either side of it."""
    depth = 3
    ctx = orc.Context()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    r = ctx.sqrt(ctx.add(ctx.add(ctx.square(x), ctx.square(y)), ctx.square(z)))
    t = orc.Tape.from_data(ctx.tape(ctx.sub(r, 0.6)))
    leaves, _ = orc.octree_sample(t, depth)
    h = 2.0 / (1 << depth)
    lo_pts, hi_pts, n_edges = [], [], 0
    for l in leaves:
        origin = np.array([l["ix"], l["iy"], l["iz"]], dtype=np.float64) * h - 1.0
        for e in range(12):
            if not (int(l["present"]) >> e) & 1:
                continue
            ta, u, v = e >> 2, e & 1, (e >> 1) & 1
            ua, va = (ta + 1) % 3, (ta + 2) % 3
            base = origin.copy()
            base[ua] += u * h
            base[va] += v * h
            pos = l["pos"][e].astype(np.float64)
            assert abs(pos[ua] - base[ua]) < 1e-6 and abs(pos[va] - base[va]) < 1e-6, (e, pos, base)
            assert base[ta] - 1e-6 <= pos[ta] <= base[ta] + h + 1e-6
            d = np.zeros(3)
            d[ta] = h / 65536.0
            lo_pts.append(pos - d)
            hi_pts.append(pos + d)
            n_edges += 1
    assert n_edges > 100
    a = t.float_slice_eval(axis_columns(t, np.array(lo_pts, dtype=np.float32))[0])
    b = t.float_slice_eval(axis_columns(t, np.array(hi_pts, dtype=np.float32))[0])
    assert np.all(a * b <= 0.0)
    # the Hermite normal of a sphere at the intersection is radial
    for l in leaves[:50]:
        for e in range(12):
            if (int(l["present"]) >> e) & 1:
                g, p = l["grad"][e][:3].astype(np.float64), l["pos"][e].astype(np.float64)
                assert np.allclose(g / np.linalg.norm(g), p / np.linalg.norm(p), atol=1e-4)


def test_octree_threads_do_not_change_the_result(orc):
    """The multi-threaded oracle sampler (subtrees on worker threads, like Octree::build_inner_mt) returns
    the serial leaves and census, byte for byte."""
    t = orc.Tape.from_vm(model_text("gyroid-sphere.vm"))
    a, sa = orc.octree_sample(t, 5)
    b, sb = orc.octree_sample(t, 5, threads=4)
    assert sa == sb and a.tobytes() == b.tobytes()
    m = np.eye(4, dtype=np.float32)
    m[0, 3], m[1, 1] = 0.1, 1.2
    t = orc.Tape.from_vm(model_text("colonnade.vm"))
    a, sa = orc.octree_sample(t, 4, m)
    b, sb = orc.octree_sample(t, 4, m, threads=3)
    assert sa == sb and a.tobytes() == b.tobytes()
