"""Parity at the sizes BASELINE.json states (configs[2], [3], [4]): the things that only break at size
-- work-list caps, 32-bit tile indices, arena sizing -- are exercised here against the CPU oracle
running on every host core."""
import os

import numpy as np
import pytest

import fidget_b200 as fb
from conftest import model_text, same_f32

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 8


@pytest.fixture(scope="module")
def big_cuda():
    """A context with an arena large enough for whole-volume 4096^3 renders (the default is 1 GiB)."""
    c = fb.CudaContext(0)
    c.set_arena_bytes(8 << 30)
    yield c
    c.close()


def test_bear_1024_cubed_matches_oracle(orc, cuda):
    """BASELINE configs[2]: bear.vm heightmap + normals at 1024^3.  bear uses exp/ln/sin/cos (libdevice vs
    glibc differ by ulps), so a razor-edge voxel may flip: mismatches are counted, printed and bounded."""
    text = model_text("bear.vm")
    ot, gs = orc.Tape.from_vm(text), fb.CudaShape.from_vm(cuda, text)
    n = 1024
    o_img, _ = orc.render3d(ot, n, n, n, threads=THREADS)
    g_img = fb.render3d(gs, fb.RenderConfig3D(n, n, n))
    same = g_img["depth"] == o_img["depth"]
    n_bad = int((~same).sum())
    hit = int((o_img["depth"] > 0).sum())
    print(f"bear 1024^3: {hit} surface pixels, {n_bad} depth mismatches")
    assert hit > 100_000
    assert n_bad <= 5e-4 * same.size, n_bad
    # normals on the agreeing pixels: north-star tolerance 1e-5 relative, outliers counted
    a, b = g_img["normal"][same].astype(np.float64), o_img["normal"][same].astype(np.float64)
    ok = (np.abs(a - b) <= 1e-5 * np.maximum(1.0, np.abs(b))) | (np.isnan(a) & np.isnan(b))
    n_out = int((~ok.all(axis=-1)).sum())
    print(f"bear 1024^3: {n_out} normals outside 1e-5 relative (of {int(same.sum())})")
    assert n_out <= 1e-3 * same.sum(), n_out
    loose = np.isclose(a, b, rtol=1e-3, atol=1e-3, equal_nan=True).all(axis=-1)
    assert loose.mean() > 0.9999


def test_prospero_4096_cubed_matches_oracle_and_slabs_merge(orc, big_cuda):
    """BASELINE configs[4] at full size on one GPU: the whole 4096^3 volume is bit-identical to the oracle
    (depth and normals), and the eight Z slabs rendered independently merge to the very same bytes."""
    import ctypes as C
    import torch
    from fidget_b200 import _lib
    text = model_text("prospero.vm")
    ot, gs = orc.Tape.from_vm(text), fb.CudaShape.from_vm(big_cuda, text)
    n = 4096
    full = torch.zeros((n, n, 4), dtype=torch.float32, device="cuda")
    _, st = fb.render3d(gs, fb.RenderConfig3D(n, n, n), out=full, stats=True)
    g_img = full.cpu().numpy().view(fb.GEOMETRY_PIXEL).reshape(n, n)
    o_img, _ = orc.render3d(ot, n, n, n, threads=THREADS)
    assert np.array_equal(g_img["depth"], o_img["depth"])
    assert same_f32(g_img["normal"], o_img["normal"])
    assert int((g_img["depth"] > 0).sum()) > 1_000_000
    print("prospero 4096^3 census:", st["evaluated"][:5], "arena MB", st["arena_bytes_used"] / 1e6)
    # Z slabs (north-star sharding) rendered one after the other on this GPU, then merged
    slabs = torch.zeros((8, n, n, 4), dtype=torch.float32, device="cuda")
    for r in range(8):
        fb.render3d(gs, fb.RenderConfig3D(n, n, n, z_range=(r * 512, (r + 1) * 512), clamp=False), out=slabs[r])
    out = torch.zeros((n, n, 4), dtype=torch.float32, device="cuda")
    ptrs = (C.c_void_p * 8)(*[slabs[r].data_ptr() for r in range(8)])
    assert _lib.load().fc_merge_slabs(big_cuda._h, ptrs, 8, n, n, n, C.c_void_p(out.data_ptr())) == 0
    assert torch.equal(out.view(torch.int32), full.view(torch.int32))


@pytest.mark.parametrize("depth", [8, 9])
def test_gyroid_sphere_octree_full_depth(orc, big_cuda, depth):
    """BASELINE configs[3]: Manifold Dual Contouring sampler on gyroid-sphere at depth 9 (and 8).  The model is
    all sin/cos, so corner samples within an ulp of zero may flip a cell: leaves are matched by cell, and
    every disagreement is counted and printed."""
    text = model_text("gyroid-sphere.vm")
    ot, gs = orc.Tape.from_vm(text), fb.CudaShape.from_vm(big_cuda, text)
    g, gst = fb.octree_sample(gs, depth, stats=True)
    o, ost = orc.octree_sample(ot, depth, threads=THREADS)
    print(f"gyroid depth {depth}: {len(g)} GPU leaves, {len(o)} oracle leaves")
    # interval census of the coarse levels is exact (no libm involved until sin/cos bounds matter)
    key = lambda a: (a["iz"].astype(np.int64) << 32) | (a["iy"].astype(np.int64) << 16) | a["ix"].astype(np.int64)
    ko, kg = key(o), key(g)
    common = np.intersect1d(ko, kg)
    print(f"  leaves only on one side: {len(o) + len(g) - 2 * len(common)}")
    assert len(common) >= 0.9995 * max(len(o), len(g))
    oc, gc = o[np.isin(ko, common)], g[np.isin(kg, common)]
    same_mask = oc["mask"] == gc["mask"]
    print(f"  corner-mask mismatches among common leaves: {int((~same_mask).sum())}")
    assert same_mask.mean() > 0.9995
    oc, gc = oc[same_mask], gc[same_mask]
    assert np.array_equal(oc["present"], gc["present"]) and np.array_equal(oc["n_edges"], gc["n_edges"])
    present = ((oc["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    po, pg = oc["pos"][present], gc["pos"][present]
    cell = 2.0 / 2 ** depth
    assert np.all(np.abs(pg - po) <= cell / 1000)        # the 16^4-ary search may land one bracket apart
    same_pos = (po.view(np.uint32) == pg.view(np.uint32)).all(axis=-1)
    print(f"  intersections at bit-identical positions: {same_pos.mean():.6f}")
    assert same_pos.mean() > 0.975     # depth 9: 98.5 % (a libm ulp at a bracket boundary moves the 16^4-ary search by one step)
    # gradients at identical positions: north-star tolerance (1e-5 relative; absolute for the near-zero value)
    go, gg = oc["grad"][present][same_pos].astype(np.float64), gc["grad"][present][same_pos].astype(np.float64)
    ok = np.abs(gg - go) <= 1e-5 * np.maximum(1.0, np.abs(go))
    n_out = int((~ok.all(axis=-1)).sum())
    print(f"  gradients outside 1e-5 relative: {n_out} of {len(go)}")
    assert n_out <= 1e-4 * len(go)
    for k in ("evaluated", "full", "empty", "ambiguous"):
        a, b = np.array(gst[k][:depth + 1], dtype=np.float64), np.array(ost[k][:depth + 1], dtype=np.float64)
        assert np.all(np.abs(a - b) <= 1e-4 * np.maximum(b, 1) + 2), (k, gst[k], ost[k])
