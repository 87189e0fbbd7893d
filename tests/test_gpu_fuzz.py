"""Randomised parity: random CSG expressions (unions / intersections of primitives, i.e. long
min/max chains over deep arithmetic) rendered by the CUDA path and by the oracle must agree bit for
bit -- image, per-level tile census and simplification count.  Only IEEE-exact opcodes are drawn, so
there is no tolerance.  The shapes are big enough (hundreds of clauses) to take the cooperative
level-0 kernel with its dependency waves, chain scans and slot-coloured values."""
import numpy as np
import pytest

import fidget_b200 as fb

pytestmark = pytest.mark.gpu


def random_shape(ctx, rng, n_prims, use_z):
    x, y, z = ctx.x(), ctx.y(), ctx.z()

    def c(lo, hi):
        return float(np.float32(rng.uniform(lo, hi)))

    def coord(a):
        # random affine re-parameterisation of an axis
        return ctx.add(ctx.mul(a, c(0.5, 2.0)), c(-0.8, 0.8))

    def prim():
        kind = rng.integers(0, 6)
        px, py = coord(x), coord(y)
        terms = [ctx.square(px), ctx.square(py)]
        if use_z:
            terms.append(ctx.square(coord(z)))
        if kind == 0:      # sphere / circle
            s = terms[0]
            for t in terms[1:]:
                s = ctx.add(s, t)
            return ctx.sub(ctx.sqrt(s), c(0.05, 0.5))
        if kind == 1:      # box: max of |p| - h
            m = ctx.sub(ctx.abs(px), c(0.05, 0.4))
            m = ctx.max(m, ctx.sub(ctx.abs(py), c(0.05, 0.4)))
            if use_z:
                m = ctx.max(m, ctx.sub(ctx.abs(coord(z)), c(0.05, 0.4)))
            return m
        if kind == 2:      # half plane through a product (exercises mul of intervals spanning zero)
            return ctx.sub(ctx.mul(px, py), c(-0.3, 0.3))
        if kind == 3:      # ring: | |p| - r | - w
            s = ctx.add(terms[0], terms[1])
            return ctx.sub(ctx.abs(ctx.sub(ctx.sqrt(s), c(0.2, 0.6))), c(0.02, 0.1))
        if kind == 4:      # rational bump: division by a positive denominator
            den = ctx.add(ctx.add(terms[0], terms[1]), c(0.1, 0.5))
            return ctx.sub(ctx.div(c(0.05, 0.3), den), c(0.1, 0.8))
        # stepped field: floor / modulo
        return ctx.sub(ctx.abs(ctx.sub(ctx.modulo(ctx.mul(px, c(1.0, 4.0)), c(0.3, 1.0)), c(0.1, 0.4))),
                       ctx.mul(ctx.floor(ctx.mul(py, 3.0)), c(0.01, 0.05)))

    shape = prim()
    for _ in range(n_prims - 1):
        p = prim()
        r = rng.random()
        if r < 0.65:
            shape = ctx.min(shape, p)                  # union
        elif r < 0.85:
            shape = ctx.max(shape, ctx.neg(p))         # difference
        else:
            shape = ctx.max(shape, p)                  # intersection
    return shape


def build(orc, cuda, seed, n_prims, use_z):
    roots = []
    tapes = []
    for Ctx in (fb.Context, orc.Context):
        ctx = Ctx()
        root = random_shape(ctx, np.random.default_rng(seed), n_prims, use_z)
        roots.append(root)
        tapes.append(ctx.tape(root))
    g = fb.CudaShape(cuda, tapes[0])
    o = orc.Tape.from_data(tapes[1])
    return g, o, tapes[0]


@pytest.mark.parametrize("seed", range(32))
def test_random_csg_2d(orc, cuda, seed):
    rng = np.random.default_rng(1000 + seed)
    n_prims = int(rng.integers(12, 90))
    g, o, tape = build(orc, cuda, seed, n_prims, use_z=False)
    size = int(rng.choice([256, 384, 512]))
    z = float(np.float32(rng.uniform(-0.2, 0.2)))
    o_img, o_st = orc.render2d(o, size, size, z=z, threads=8)
    g_img, g_st = fb.render2d(g, fb.RenderConfig2D(size, size, z=z), stats=True)
    for k in ("evaluated", "filled_inside", "filled_outside", "ambiguous", "simplified"):
        assert g_st[k] == o_st[k], (k, len(tape))
    assert np.array_equal(g_img.view(np.uint32), o_img.view(np.uint32)), len(tape)
    inside = fb.pixel_inside(g_img).mean()
    assert 0.0 <= inside <= 1.0


@pytest.mark.parametrize("seed", range(10))
def test_random_csg_3d(orc, cuda, seed):
    rng = np.random.default_rng(2000 + seed)
    n_prims = int(rng.integers(10, 40))
    g, o, tape = build(orc, cuda, 100 + seed, n_prims, use_z=True)
    size = 128
    o_img, _ = orc.render3d(o, size, size, size, threads=8)
    g_img = fb.render3d(g, fb.RenderConfig3D(size, size, size))
    assert np.array_equal(g_img["depth"], o_img["depth"]), len(tape)
    a, b = g_img["normal"], o_img["normal"]
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb)
    assert np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]), len(tape)


@pytest.mark.parametrize("seed", range(4))
def test_random_csg_octree(orc, cuda, seed):
    rng = np.random.default_rng(3000 + seed)
    n_prims = int(rng.integers(6, 24))
    g, o, tape = build(orc, cuda, 200 + seed, n_prims, use_z=True)
    o_leaves, o_st = orc.octree_sample(o, 5)
    g_leaves, g_st = fb.octree_sample(g, 5, stats=True)
    assert len(g_leaves) == len(o_leaves)
    for f in ("ix", "iy", "iz", "mask", "n_edges", "present"):
        assert np.array_equal(g_leaves[f], o_leaves[f]), f
    present = ((o_leaves["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)   # entries of absent edges are unspecified
    for f in ("pos", "grad"):
        a, b = g_leaves[f][present], o_leaves[f][present]
        na, nb = np.isnan(a), np.isnan(b)
        assert np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]), f
    for k in ("evaluated", "full", "empty", "ambiguous"):
        assert g_st[k][:6] == o_st[k][:6], k
