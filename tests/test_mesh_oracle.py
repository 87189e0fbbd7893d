"""Pins of the meshing oracle (oracle/mesh.py) that need no GPU: the vertex-group rule against the counts of the
reference's table (fidget-mesh/build.rs), the QEF solver against the reference's own unit tests
(fidget-mesh/src/qef.rs:171-..., octree.rs test_cube_verts / test_sphere_verts properties), and manifoldness of the
dual walk (octree.rs:1216-1232 test_sphere_manifold: every edge is shared by exactly two triangles, once in each
direction)."""
import numpy as np
import pytest

from oracle import mesh as om


def test_vertex_groups_table_properties():
    # build.rs: a cell has between 0 and 4 vertices; empty and full cells have none; a single inside corner gives one
    # vertex with three transitions; two diagonal corners of a face give two vertices (the ambiguous face)
    counts = [om.corner_groups(m)[1] for m in range(256)]
    assert counts[0] == 0 and max(counts) == 4
    assert counts[0b00000001] == 1 and counts[0b00001001] == 2 and counts[0b10010110] == 4
    assert counts[255] == 1           # one group, but no transition: the caller emits no vertex edges
    for m in range(256):
        g, n = om.corner_groups(m)
        assert sorted(set(g.values())) == list(range(n))
        # complementary masks: the transitions are the same edges seen from the other side
        trans = lambda mask: {(min(s, s ^ t), max(s, s ^ t)) for s in range(8) if (mask >> s) & 1
                              for t in (1, 2, 4) if not (mask >> (s ^ t)) & 1}
        assert trans(m) == trans(255 ^ m)


def test_qef_plane_edge_corner():
    f = np.float32
    # one plane z = 0.25 sampled at three points: rank 1, the vertex is the mass point projected onto the plane
    pts = [np.array(p, dtype=f) for p in ((0, 0, .25), (1, 0, .25), (0, 1, .25))]
    g = [np.array([0, 0, 1, 0], dtype=f)] * 3
    v = om.qef_vertex(pts, g)
    assert np.allclose(v, [1 / 3, 1 / 3, .25], atol=1e-6)
    # two planes x = 0.5 and y = -0.25: rank 2, the vertex lies on their intersection line at the mass point's z
    pts = [np.array(p, dtype=f) for p in ((.5, 0, 0), (.5, 1, 1), (0, -.25, .5), (1, -.25, .5))]
    g = [np.array(q, dtype=f) for q in ((2, 0, 0, 0), (1, 0, 0, 0), (0, 3, 0, 0), (0, 1, 0, 0))]
    v = om.qef_vertex(pts, g)
    assert np.allclose(v, [.5, -.25, .5], atol=1e-6)
    # three planes meeting in a corner (test_cube_verts, octree.rs:1235-1276): rank 3, the corner itself
    pts = [np.array(p, dtype=f) for p in ((.4, .1, .2), (.3, .4, .1), (.2, .3, .4))]
    g = [np.array(q, dtype=f) for q in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0))]
    assert np.allclose(om.qef_vertex(pts, g), [.4, .4, .4], atol=1e-6)
    # a NaN gradient snaps to that intersection (octree.rs:793-801)
    g[1] = np.array([np.nan, 0, 0, 0], dtype=f)
    assert np.array_equal(om.qef_vertex(pts, g), pts[1])


def _mesh(orc, tape, depth):
    leaves, _ = orc.octree_sample(tape, depth)
    return leaves, om.build(leaves)


def test_sphere_mesh_is_manifold_and_round(orc):
    ctx = orc.Context()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    r = 0.6
    t = orc.Tape.from_data(ctx.tape(ctx.sub(ctx.sqrt(ctx.add(ctx.add(ctx.square(x), ctx.square(y)), ctx.square(z))), r)))
    leaves, (verts, tris, open_edges) = _mesh(orc, t, 4)
    assert open_edges == 0 and len(tris) > 0 and len(tris) % 4 == 0
    # test_sphere_verts (octree.rs:1180-1214): every cell vertex is close to the sphere
    rad = np.array([np.linalg.norm(v) for v in verts.values()])
    assert np.all(np.abs(rad - r) < 0.02)
    # test_sphere_manifold: directed edges pair up exactly
    key = lambda p: p.tobytes()
    edges = {}
    for a, b, c in tris:
        for p, q in ((a, b), (b, c), (c, a)):
            edges[(key(p), key(q))] = edges.get((key(p), key(q)), 0) + 1
    assert all(n == 1 for n in edges.values())
    assert all((q, p) in edges for (p, q) in edges)
    # outward orientation: the signed volume is the ball's, within the faceting error
    vol = sum(np.dot(a.astype(np.float64), np.cross(b.astype(np.float64), c.astype(np.float64))) for a, b, c in tris) / 6
    assert abs(vol - 4 / 3 * np.pi * r ** 3) < 0.03
    # STL framing
    stl = om.write_stl(tris)
    assert len(stl) == 84 + 50 * len(tris) and stl[:44] == b"This is a binary STL file exported by Fidget"
    assert int.from_bytes(stl[80:84], "little") == len(tris)
