"""Pins the CPU oracle to the reference's own golden vectors and known-answer
tests.  CPU only.  Fixtures under tests/golden/ were
transcribed from the reference by the committed extract_*.py scripts."""
import json
import math
import os

import numpy as np
import pytest

from conftest import model_text, same_f32

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIX = json.load(open(os.path.join(GOLD, "pixel_render.json")))
IVL = json.load(open(os.path.join(GOLD, "interval_known_answers.json")))
IVL.update(json.load(open(os.path.join(GOLD, "interval_known_answers_manual.json"))))   # transcribed by hand
POINT = json.load(open(os.path.join(GOLD, "point_known_answers.json")))
CHOICE = {"Left": 1, "Right": 2, "Both": 3}


def ascii_rows(orc, img):
    return ["".join("#" if b else "." for b in r) for r in orc.pixel_inside(img)]


def view2(center, scale):
    # View2::world_to_model = translation(center) * scaling(scale)  (fidget-gui/src/lib.rs:91-104)
    return np.array([[scale, 0, center[0]], [0, scale, center[1]], [0, 0, 1]], dtype=np.float32)


@pytest.mark.parametrize("n_regs", [255, 3])        # render_tests!(vm, ..) and render_tests!(vm3, ..)
def test_hi_goldens(orc, n_regs):
    t = orc.Tape.from_vm(model_text("hi.vm"), n_regs)
    img, _ = orc.render2d(t, 32, 32)
    assert ascii_rows(orc, img) == PIX["check_hi:EXPECTED"]["rows"]
    img, _ = orc.render2d(t, 64, 32)
    assert ascii_rows(orc, img) == PIX["check_hi_wide:EXPECTED"]["rows"]
    m = view2((0.5, 0.5), 0.5)                       # prepend_translation(.5,.5); prepend_scaling(.5)
    img, _ = orc.render2d(t, 32, 32, mat=orc.pixel_mat(32, 32, m))
    assert ascii_rows(orc, img) == PIX["check_hi_transformed:EXPECTED"]["rows"]
    assert ascii_rows(orc, img) == PIX["check_hi_bounded:EXPECTED"]["rows"]   # same matrix via View2


@pytest.mark.parametrize("n_regs", [255, 3])
def test_quarter_golden(orc, n_regs):
    t = orc.Tape.from_vm(model_text("quarter.vm"), n_regs)
    img, _ = orc.render2d(t, 32, 32)
    assert ascii_rows(orc, img) == PIX["check_quarter:EXPECTED"]["rows"]


@pytest.mark.parametrize("n_regs", [255, 3])
def test_circle_with_bound_var(orc, n_regs):
    # fidget/tests/pixel_render.rs:277-364
    ctx = orc.Context()
    x, y = ctx.x(), ctx.y()
    r = ctx.sqrt(ctx.add(ctx.square(x), ctx.square(y)))
    c, _ = ctx.var()
    td = ctx.tape(ctx.sub(r, c), n_regs)
    t = orc.Tape.from_data(td)
    slot = [i for i, (k, _) in enumerate(td.vars()) if k == "v"][0]
    for radius, key in ((0.75, "check_circle_var:EXPECTED_075"), (0.5, "check_circle_var:EXPECTED_05")):
        vv = np.zeros(td.n_vars, dtype=np.float32)
        vv[slot] = radius
        img, _ = orc.render2d(t, 32, 32, var_values=vv)
        assert ascii_rows(orc, img) == PIX[key]["rows"]


@pytest.mark.parametrize("n_regs", [255, 3])
def test_voxel_sphere_with_bound_var(orc, n_regs):
    """fidget/tests/voxel_render.rs:13-75 (sphere_var + check_sphere): a sphere of variable radius rendered
    at 32^3 through View3 cameras of scale 1 and 0.5 lands within two voxels of the analytic surface."""
    ctx = orc.Context()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    r = ctx.sqrt(ctx.add(ctx.add(ctx.square(x), ctx.square(y)), ctx.square(z)))
    c, _ = ctx.var()
    td = ctx.tape(ctx.sub(r, c), n_regs)
    t = orc.Tape.from_data(td)
    slot = [i for i, (k, _) in enumerate(td.vars()) if k == "v"][0]
    size = 32
    m = orc.screen_to_world_3d(size, size, size).astype(np.float64)
    for scale in (1.0, 0.5):
        wm = np.diag([scale, scale, scale, 1.0]).astype(np.float32)      # View3::from_center_and_scale(0, scale)
        for radius in (0.5, 0.75):
            vv = np.zeros(td.n_vars, dtype=np.float32)
            vv[slot] = radius
            img, _ = orc.render3d(t, size, size, size, mat=orc.voxel_mat(size, size, size, wm), var_values=vv)
            eps = 2.0 / size / scale * 2.0
            depth = img["depth"].astype(np.int64)
            ys, xs = np.mgrid[0:size, 0:size]
            pts = np.stack([xs, ys, depth, np.ones_like(xs)], axis=-1).astype(np.float64) @ m.T
            pos = pts[..., :3] / pts[..., 3:4] * scale
            empty, hit = depth == 0, (depth != 0) & (depth != size)       # saturated voxels are skipped
            assert hit.sum() > 20
            assert (np.hypot(pos[..., 0], pos[..., 1])[empty] + eps > radius).all()
            assert (np.abs(radius - np.linalg.norm(pos, axis=-1))[hit] < eps).all()


def test_neg_infinity_pixel_perfect(orc):
    # pixel_render.rs:366-377
    ctx = orc.Context()
    t = orc.Tape.from_data(ctx.tape(ctx.constant(float("-inf"))))
    img, _ = orc.render2d(t, 256, 256, pixel_perfect=True)
    assert orc.pixel_inside(img).all()


def test_screen_to_world(orc):
    # fidget-core/src/render/region.rs:204-228
    m = orc.screen_to_world_2d(1000, 500)

    def tp(x, y):
        out = np.zeros(3, dtype=np.float32)
        import ctypes as C
        orc.lib().orc_transform_f32(m.ctypes.data_as(C.POINTER(C.c_float)), x, y, 0.0,
                                    out.ctypes.data_as(C.POINTER(C.c_float)))
        return float(out[0]), float(out[1])
    assert tp(500.0, 249.0) == (0.0, 0.0)
    assert tp(500.0, -1.0) == (0.0, 1.0)
    assert tp(500.0, 499.0) == (0.0, -1.0)
    assert tp(0.0, 249.0) == (-2.0, 0.0)
    assert tp(1000.0, 249.0) == (2.0, 0.0)


def test_camera_render_config(orc):
    # pixel_render.rs:429-474
    import ctypes as C
    for scale, exp in ((0.5, [(0.0, 1.0), (1.0, 1.0), (1.0, 0.0)]), (0.25, [(0.25, 0.75), (0.75, 0.75), (0.75, 0.25)])):
        m = np.ascontiguousarray(orc.pixel_mat(512, 512, view2((0.5, 0.5), scale)))
        for (px, py), e in zip([(0.0, -1.0), (512.0, -1.0), (512.0, 511.0)], exp):
            out = np.zeros(3, dtype=np.float32)
            orc.lib().orc_transform_f32(m.ctypes.data_as(C.POINTER(C.c_float)), px, py, 0.0,
                                        out.ctypes.data_as(C.POINTER(C.c_float)))
            assert (float(out[0]), float(out[1])) == e


def _f(v):
    return {"nan": math.nan, "inf": math.inf, "-inf": -math.inf}.get(v, v) if isinstance(v, str) else v


def _build(ctx, nodes):
    env = {}
    for name, op, args in nodes:
        if op == "var":
            env[name] = getattr(ctx, args[0])()
        elif op == "const":
            env[name] = ctx.constant(float(args[0]))
        else:
            a = [env[x] if isinstance(x, str) else float(x) for x in args]
            env[name] = getattr(ctx, {"and": "and_", "or": "or_", "not": "not_"}.get(op, op))(*a)
    return env


@pytest.mark.parametrize("name", sorted(IVL))
def test_interval_known_answers(orc, name):
    """fidget-core/src/eval/test/interval.rs (exact bounds and exact Choice traces)."""
    spec = IVL[name]
    ctx = orc.Context()
    env = _build(ctx, spec["nodes"])
    tapes = {}
    for case in spec["cases"]:
        root = case["root"]
        if root not in tapes:
            td = ctx.tape(env[root])
            tapes[root] = (td, orc.Tape.from_data(td))
        td, t = tapes[root]
        vx, vy, vz = td.var_slots()
        ins = [[_f(a), _f(b)] for a, b in case["inputs"]]
        vars_ = np.zeros((max(td.n_vars, 1), 2), dtype=np.float32)
        for slot, iv in zip([s for s in (vx, vy, vz)], ins + [None] * 3):
            if slot >= 0 and iv is not None:
                vars_[slot] = iv
        if td.n_vars == 1:                       # single-variable tests pass `[[a, b].into()]`
            vars_[0] = ins[0]
        out, choices, simplify = t.interval_eval(vars_)
        exp = [_f(v) for v in case["expect"]]
        if any(isinstance(v, float) and math.isnan(v) for v in exp):
            assert np.isnan(out).all(), (name, case)
        else:
            assert out.tolist() == [np.float32(exp[0]), np.float32(exp[1])], (name, case, out)
        if "trace" in case:
            if case["trace"] is None:
                assert not simplify, (name, case)
            else:
                assert simplify and choices.tolist() == [CHOICE[c] for c in case["trace"]], (name, case)


@pytest.mark.parametrize("name", sorted(POINT))
def test_point_known_answers(orc, name):
    """fidget-core/src/eval/test/point.rs: exact values, exact Choice traces, simplified size."""
    spec = POINT[name]
    ctx = orc.Context()
    env = _build(ctx, spec["nodes"])
    td = ctx.tape(env[spec["root"]])
    t = orc.Tape.from_data(td)
    slots = [s for s in td.var_slots()[:2]]

    def run(tape, ins):
        vars_ = np.zeros(max(td.n_vars, 1), dtype=np.float32)
        if td.n_vars == 1 and ins:
            vars_[0] = _f(ins[0])
        else:
            for slot, v in zip(slots, ins):
                if slot >= 0:
                    vars_[slot] = _f(v)
        return tape.point_eval(vars_)

    for case in spec["cases"]:
        out, choices, simplify = run(t, case["inputs"])
        exp = _f(case["expect"])
        if isinstance(exp, float) and math.isnan(exp):
            assert np.isnan(out), (name, case)
        else:
            assert out == np.float32(exp), (name, case, out)
        if "trace" in case:
            if case["trace"] is None:
                assert not simplify, (name, case)
            else:
                assert simplify and choices.tolist() == [CHOICE[c] for c in case["trace"]], (name, case)
        if "child_size" in case:
            child = t.simplify(choices)
            assert child.size == case["child_size"], (name, child.size)
            for cc in case["child_cases"]:
                assert run(child, cc["inputs"])[0] == np.float32(cc["expect"])


def test_interval_contains_point_samples(orc):
    """Property test in the spirit of interval.rs:1087-1170: for every op, interval results contain
    point samples (or are the NaN interval)."""
    from fidget_b200.host import UNARY_OPS, BINARY_OPS
    rng = np.random.default_rng(0)
    args = [np.float32(np.pi * 2 * i / 8) for i in range(-8, 9)] + [np.float32(v) for v in (1, 5, .5, 1.5, 10)]
    for op in UNARY_OPS + BINARY_OPS:
        if op in ("rand", "mix"):
            continue
        ctx = orc.Context()
        x, y = ctx.x(), ctx.y()
        node = ctx.unary(op, x) if op in UNARY_OPS else ctx.binary(op, x, y)
        td = ctx.tape(node)
        t = orc.Tape.from_data(td)
        vx, vy, _ = td.var_slots()
        for _ in range(60):
            a, b = sorted(rng.choice(args, 2))
            c, d = sorted(rng.choice(args, 2))
            box = np.zeros((td.n_vars, 2), dtype=np.float32)
            box[vx] = (a, b)
            if vy >= 0:
                box[vy] = (c, d)
            out, _, _ = t.interval_eval(box)
            if np.isnan(out).any():
                continue
            px = rng.uniform(a, b, 16).astype(np.float32)
            py = rng.uniform(c, d, 16).astype(np.float32)
            pts = [None] * td.n_vars
            pts[vx] = px
            if vy >= 0:
                pts[vy] = py
            v = t.float_slice_eval(pts)
            ok = np.isnan(v) | ((v >= out[0] - 1e-5 * max(1, abs(out[0]))) & (v <= out[1] + 1e-5 * max(1, abs(out[1]))))
            assert ok.all(), (op, box, out, v[~ok])


GRD = json.load(open(os.path.join(GOLD, "grad_known_answers.json")))


@pytest.mark.parametrize("name", sorted(GRD))
def test_grad_known_answers(orc, name):
    """fidget-core/src/eval/test/grad_slice.rs:54-445 (exact Grad values)."""
    spec = GRD[name]
    ctx = orc.Context()
    env = _build(ctx, spec["nodes"])
    for case in spec["cases"]:
        td = ctx.tape(env[case["root"]])
        t = orc.Tape.from_data(td)
        vx, vy, vz = td.var_slots()
        vars_ = [np.zeros((1, 4), dtype=np.float32) for _ in range(max(td.n_vars, 1))]
        for axis, slot in enumerate((vx, vy, vz)):
            if slot >= 0:
                vars_[slot][0, 0] = _f(case["xyz"][axis])
                vars_[slot][0, 1 + axis] = 1.0
        out = t.grad_slice_eval(vars_)[0]
        exp = np.array([_f(v) for v in case["expect"]], dtype=np.float32)
        assert np.array_equal(out, exp) or (np.isnan(exp).any() and np.array_equal(np.isnan(out), np.isnan(exp))), \
            (name, case, out)


def test_float_slice_vectorized(orc):
    """fidget-core/src/eval/test/float_slice.rs:46-90 (test_vectorized): ragged slice lengths."""
    ctx = orc.Context()
    x, y = ctx.x(), ctx.y()
    t = orc.Tape.from_data(ctx.tape(x))
    for n in (4, 8, 9):
        v = np.arange(n, dtype=np.float32)
        assert t.float_slice_eval([v]).tolist() == v.tolist()
    t = orc.Tape.from_data(ctx.tape(ctx.mul(y, 2.0)))
    for ins, exp in (([3.0, 2.0, 1.0, 0.0], [6.0, 4.0, 2.0, 0.0]), ([1.0, 4.0, 8.0], [2.0, 8.0, 16.0]),
                     ([1.0, 4.0, 4.0, -1.0, -2.0, -3.0, 0.0], [2.0, 8.0, 8.0, -2.0, -4.0, -6.0, 0.0])):
        assert t.float_slice_eval([np.array(ins, dtype=np.float32)]).tolist() == exp


def test_grad_add_and_modulo(orc):
    """grad_slice.rs test_g_add (two-element slices) and test_g_modulo (y - x mod 1)."""
    ctx = orc.Context()
    x, y = ctx.x(), ctx.y()
    td = ctx.tape(ctx.add(x, y))
    t = orc.Tape.from_data(td)
    vx, vy, _ = td.var_slots()
    vars_ = [np.zeros((2, 4), dtype=np.float32) for _ in range(2)]
    vars_[vx][:, 0] = [0.0, 1.0]          # Grad::from(f32): value only, zero derivative
    vars_[vy][:, 0] = [2.0, 3.0]
    assert t.grad_slice_eval(vars_).tolist() == [[2.0, 0.0, 0.0, 0.0], [4.0, 0.0, 0.0, 0.0]]

    td = ctx.tape(ctx.sub(y, ctx.modulo(x, 1.0)))
    t = orc.Tape.from_data(td)
    vx, vy, _ = td.var_slots()
    for xv, exp in ((0.0, [0.5, -1.0, 1.0, 0.0]), (-0.01, [-0.49, -1.0, 1.0, 0.0]), (0.01, [0.49, -1.0, 1.0, 0.0])):
        vars_ = [np.zeros((1, 4), dtype=np.float32) for _ in range(2)]
        vars_[vx][0, 0], vars_[vx][0, 1] = xv, 1.0
        vars_[vy][0, 0], vars_[vy][0, 2] = 0.5, 1.0
        assert t.grad_slice_eval(vars_)[0].tolist() == [float(np.float32(v)) for v in exp], xv


def _test_args():
    # eval/test/mod.rs:48-63 (test_args)
    a = [np.float32(np.pi) * np.float32(2.0) * np.float32(i) / np.float32(32) for i in range(-32, 33)]
    a += [1.0, 5.0, 0.5, 1.5, 10.0, np.pi, np.pi / 2, 1 / np.pi, np.sqrt(2.0), np.nan]
    return np.array(a, dtype=np.float32)


def _canon(op, a, b=None):
    """Canonical f32 definitions of eval/test/mod.rs:188-246, evaluated with numpy float32
    (only the operations IEEE-754 fixes bit for bit are listed)."""
    f = np.float32
    with np.errstate(all="ignore"):
        if op == "neg": return -a
        if op == "abs": return np.abs(a)
        if op == "recip": return f(1.0) / a
        if op == "sqrt": return np.sqrt(a)
        if op == "square": return a * a
        if op == "floor": return np.floor(a)
        if op == "ceil": return np.ceil(a)
        if op == "round": return np.where(np.isnan(a), a, np.copysign(np.floor(np.abs(a) + f(0.5)), a)).astype(f)
        if op == "not": return (a == 0).astype(f)
        if op == "add": return a + b
        if op == "sub": return a - b
        if op == "mul": return a * b
        if op == "div": return a / b
        if op == "min": return np.where(np.isnan(a) | np.isnan(b), f(np.nan), np.where(a < b, a, b))
        if op == "max": return np.where(np.isnan(a) | np.isnan(b), f(np.nan), np.where(a > b, a, b))
        if op == "compare": return np.where(np.isnan(a) | np.isnan(b), f(np.nan),
                                            np.where(a < b, f(-1), np.where(a > b, f(1), f(0))))
        if op == "and": return np.where(a == 0, a, b)
        if op == "or": return np.where(a != 0, a, b)
        if op == "mod":
            r = np.fmod(a, b)
            return np.where(r < 0, r + np.abs(b), r)
    raise KeyError(op)


@pytest.mark.parametrize("op", ["neg", "abs", "recip", "sqrt", "square", "floor", "ceil", "round", "not"])
def test_float_slice_unary_canonical(orc, op):
    # eval/test/float_slice.rs:391-432: every op equals its canonical f32 definition exactly
    ctx = orc.Context()
    t = orc.Tape.from_data(ctx.tape(ctx.unary(op, ctx.x())))
    a = _test_args()
    got, want = t.float_slice_eval([a]), _canon(op, a).astype(np.float32)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "min", "max", "compare", "and", "or", "mod"])
def test_float_slice_binary_canonical(orc, op):
    args = _test_args()
    a, b = [g.ravel().astype(np.float32) for g in np.meshgrid(args, args)]
    # reg-reg, reg-imm and imm-reg forms (float_slice.rs:434-560)
    ctx = orc.Context()
    x, y = ctx.x(), ctx.y()
    td = ctx.tape(ctx.binary(op, x, y))
    t = orc.Tape.from_data(td)
    vx, vy, _ = td.var_slots()
    vars_ = [None, None]
    vars_[vx], vars_[vy] = a, b
    got, want = t.float_slice_eval(vars_), _canon(op, a, b).astype(np.float32)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    for k in args[::7]:
        if np.isnan(k):
            continue
        for imm_first in (False, True):
            ctx = orc.Context()
            x = ctx.x()
            node = ctx.binary(op, float(k), x) if imm_first else ctx.binary(op, x, float(k))
            t = orc.Tape.from_data(ctx.tape(node))
            if t.n_vars == 0:
                continue                     # folded to a constant
            got = t.float_slice_eval([args])
            kk = np.full_like(args, k)
            want = (_canon(op, kk, args) if imm_first else _canon(op, args, kk)).astype(np.float32)
            assert np.array_equal(np.isnan(got), np.isnan(want)) and \
                np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), (op, k, imm_first)


# ---------------------------------------------------------------------------
# Octree sampler: the reference's mesh tests that only depend on the sampled edge crossings
def _mesh_sphere(ctx, center, radius):
    # fidget-mesh/src/octree.rs:1075-1082
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    sx = ctx.square(ctx.sub(x, ctx.constant(center[0])))
    sy = ctx.square(ctx.sub(y, ctx.constant(center[1])))
    sz = ctx.square(ctx.sub(z, ctx.constant(center[2])))
    return ctx.sub(ctx.sqrt(ctx.add(ctx.add(sx, sy), sz)), ctx.constant(radius))


def _mesh_cube(ctx, bx, by, bz):
    # octree.rs:1084-1090
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    xb = ctx.max(ctx.sub(ctx.constant(bx[0]), x), ctx.sub(x, ctx.constant(bx[1])))
    yb = ctx.max(ctx.sub(ctx.constant(by[0]), y), ctx.sub(y, ctx.constant(by[1])))
    zb = ctx.max(ctx.sub(ctx.constant(bz[0]), z), ctx.sub(z, ctx.constant(bz[1])))
    return ctx.max(ctx.max(xb, yb), zb)


def _edge_points(leaves):
    present = ((leaves["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    return leaves["pos"][present]


def test_octree_sphere_edge_vertices(orc):
    # octree.rs:1181-1214 (test_sphere_verts): depth 1, radius 0.2: the 6 edge vertices sit on the axes
    ctx = orc.Context()
    t = orc.Tape.from_data(ctx.tape(_mesh_sphere(ctx, (0, 0, 0), 0.2)))
    leaves, st = orc.octree_sample(t, 1)
    assert st["leaf_surface"] == 8
    pts = _edge_points(leaves)
    assert np.all((pts != 0).sum(axis=1) == 1)
    assert np.all(np.abs(np.linalg.norm(pts, axis=1) - 0.2) < 2.0 / 65535)
    assert len(np.unique(np.round(pts, 5), axis=0)) == 6


def test_octree_cube_edge_positions(orc):
    # octree.rs:1235-1276 (test_cube_verts)
    ctx = orc.Context()
    bx, by, bz = (-0.1, 0.6), (-0.2, 0.75), (-0.3, 0.4)
    t = orc.Tape.from_data(ctx.tape(_mesh_cube(ctx, bx, by, bz)))
    leaves, _ = orc.octree_sample(t, 1)
    pts = _edge_points(leaves)
    assert len(pts) > 0
    eps = 2.0 / 65535
    for v in pts:
        nz = v != 0
        assert nz.sum() == 1
        a = int(np.argmax(nz))
        lo, hi = (bx, by, bz)[a]
        assert abs(v[a] - lo) < eps or abs(v[a] - hi) < eps, v


def test_octree_cube_single_edge(orc):
    # octree.rs:1093-1107 (test_cube_edge): depth 0, 4 crossings (+ 1 QEF vertex = the 5 verts of the reference)
    ctx = orc.Context()
    f = 2.0
    t = orc.Tape.from_data(ctx.tape(_mesh_cube(ctx, (-f, f), (-f, 0.3), (-f, 0.6))))
    leaves, _ = orc.octree_sample(t, 0)
    assert len(leaves) == 1 and leaves[0]["n_edges"] == 4
    pts = _edge_points(leaves)
    for v in pts:
        assert abs(v[1] - 0.3) < 1e-3 or abs(v[2] - 0.6) < 1e-3
    # gradients at the crossings are the face normals
    present = ((leaves["present"][:, None] >> np.arange(12)[None, :]) & 1).astype(bool)
    g = leaves["grad"][present][:, :3]
    assert np.all((np.abs(g) == 1).sum(axis=1) == 1)


@pytest.mark.parametrize("seed", range(12))
def test_simplified_tapes_agree_with_their_parent_inside_the_box(orc, seed):
    """The contract of VmData::simplify (vm/data.rs:123-318) that every renderer relies on: a tape simplified
    with the trace of an interval evaluation computes the same values as its parent everywhere inside that
    interval -- point results bit for bit, and the interval result over the same box."""
    from test_gpu_fuzz import random_shape
    rng = np.random.default_rng(500 + seed)
    ctx = orc.Context()
    td = ctx.tape(random_shape(ctx, rng, int(rng.integers(6, 60)), use_z=bool(seed % 2)))
    parent = orc.Tape.from_data(td)
    nv = max(td.n_vars, 1)
    checked = 0
    for _ in range(40):
        c = rng.uniform(-1, 1, nv).astype(np.float32)
        w = (rng.uniform(0, 1, nv) ** 3 * 0.6).astype(np.float32)
        box = np.stack([c - w, c + w], axis=-1).astype(np.float32)
        out, choices, simplify = parent.interval_eval(box)
        if not simplify:
            continue
        child = parent.simplify(choices)
        assert child.size <= parent.size
        o2, _, _ = child.interval_eval(box)
        assert same_f32(o2, out)
        for _ in range(16):
            pt = (box[:, 0] + (box[:, 1] - box[:, 0]) * rng.random(nv).astype(np.float32)).astype(np.float32)
            pt = np.minimum(np.maximum(pt, box[:, 0]), box[:, 1])
            a, b = parent.point_eval(pt)[0], child.point_eval(pt)[0]
            assert same_f32(a, b), (seed, pt, a, b)
        checked += 1
    assert checked > 0
