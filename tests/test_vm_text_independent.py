"""An evaluator that shares NOTHING with the tape front end: it reads the `.vm` text line by line and computes every
node with numpy float32 arrays, in source order.  The oracle and the product both go text -> Context -> SSA ->
register allocation -> bytecode through the same C++ front end (csrc/host), so a bug there would be invisible to the
oracle-vs-CUDA parity tests; here the register tape's results (oracle VM, all register budgets, spilling included) are
held against the plain reading of the text.  IEEE opcodes must agree bit for bit; libm opcodes within 1e-5 relative."""

import numpy as np
import pytest

from conftest import model_text

IEEE_MODELS = ["hi.vm", "quarter.vm", "prospero.vm", "colonnade.vm", "tanglecube.vm"]
LIBM_MODELS = ["bear.vm", "gyroid-sphere.vm"]


def f32_min(a, b):      # the VM's rule (vm/mod.rs min: NaN if either is NaN); ties carry the same value
    return np.where(np.isnan(a) | np.isnan(b), np.float32(np.nan), np.minimum(a, b)).astype(np.float32)


def f32_max(a, b):
    return np.where(np.isnan(a) | np.isnan(b), np.float32(np.nan), np.maximum(a, b)).astype(np.float32)


UNARY = {
    "neg": lambda a: -a, "abs": np.abs, "sqrt": np.sqrt, "square": lambda a: a * a,
    "recip": lambda a: np.float32(1.0) / a, "exp": np.exp, "ln": np.log, "sin": np.sin, "cos": np.cos, "tan": np.tan,
    "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "floor": np.floor, "ceil": np.ceil,
}
BINARY = {
    "add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b, "div": lambda a, b: a / b,
    "min": f32_min, "max": f32_max, "atan2": np.arctan2,
}


def eval_vm_text(text, x, y, z):
    """Value of the last node of a .vm listing at the points (x, y, z), all arithmetic in float32."""
    env, last = {}, None
    n = len(x)
    with np.errstate(all="ignore"):
        for line in text.splitlines():
            line = line.split("#", 1)[0].split()
            if not line:
                continue
            name, op, args = line[0], line[1], line[2:]
            if op == "const":
                v = np.full(n, np.float32(float(args[0])), dtype=np.float32)
            elif op in ("var-x", "var-y", "var-z"):
                v = {"var-x": x, "var-y": y, "var-z": z}[op]
            elif op in UNARY:
                v = UNARY[op](env[args[0]])
            elif op in BINARY:
                v = BINARY[op](env[args[0]], env[args[1]])
            else:
                raise AssertionError(f"opcode {op} not known to the independent evaluator")
            env[name] = np.asarray(v, dtype=np.float32)
            last = name
    return env[last]


def _points(seed, n):
    rng = np.random.default_rng(seed)
    return [rng.uniform(-1.2, 1.2, n).astype(np.float32) for _ in range(3)]


def _tape_eval(orc, text, n_regs, x, y, z):
    t = orc.Tape.from_vm(text, n_regs)
    slots = t.data.var_slots()                      # input slot of each axis, -1 when unused
    inputs = [None] * t.n_vars
    for axis, slot in zip((x, y, z), slots):
        if slot >= 0:
            inputs[slot] = axis
    assert all(v is not None for v in inputs)
    return t.float_slice_eval(inputs)


@pytest.mark.parametrize("name", IEEE_MODELS)
@pytest.mark.parametrize("n_regs", [255, 24, 6])
def test_register_tape_equals_the_plain_reading_of_the_text(orc, name, n_regs):
    text = model_text(name)
    x, y, z = _points(11, 4096 if name == "prospero.vm" else 20000)
    want = eval_vm_text(text, x, y, z)
    got = _tape_eval(orc, text, n_regs, x, y, z)
    both_nan = np.isnan(want) & np.isnan(got)
    assert np.array_equal(want.view(np.uint32)[~both_nan], got.view(np.uint32)[~both_nan])
    assert np.isfinite(want).mean() > 0.99


@pytest.mark.parametrize("name", LIBM_MODELS)
@pytest.mark.parametrize("n_regs", [255, 12])
def test_register_tape_matches_the_text_within_libm_tolerance(orc, name, n_regs):
    text = model_text(name)
    x, y, z = _points(5, 20000)
    want = eval_vm_text(text, x, y, z)
    got = _tape_eval(orc, text, n_regs, x, y, z)
    ok = np.isfinite(want)
    assert ok.mean() > 0.9
    assert np.array_equal(np.isnan(want), np.isnan(got))
    err = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-3)
    # numpy's float32 exp / ln / sin / cos differ from glibc's by an ulp here and there, and the models subtract such
    # values: nearly every point agrees to 1e-5, the worst cancellation stays below 1e-3
    assert err.max() < 1e-3 and np.quantile(err, 0.995) < 1e-5


def test_the_independent_evaluator_knows_every_opcode_the_models_use():
    used = set()
    for name in IEEE_MODELS + LIBM_MODELS:
        for line in model_text(name).splitlines():
            parts = line.split("#", 1)[0].split()
            if len(parts) >= 2:
                used.add(parts[1])
    assert used <= set(UNARY) | set(BINARY) | {"const", "var-x", "var-y", "var-z"}


@pytest.mark.parametrize("name", IEEE_MODELS + LIBM_MODELS)
def test_interval_results_enclose_the_text_evaluated_inside_the_box(orc, name):
    """Soundness of the interval path of the register tape against the same independent reading: for random boxes the
    interval the VM returns contains the function's value at points sampled inside the box (small slack for the VM's
    round-to-nearest interval arithmetic)."""
    text = model_text(name)
    t = orc.Tape.from_vm(text)
    slots = t.data.var_slots()
    rng = np.random.default_rng(3)
    checked = 0
    for _ in range(60 if name == "prospero.vm" else 150):
        centre = rng.uniform(-1, 1, 3)
        half = rng.choice([0.5, 0.1, 0.02]) * rng.uniform(0.2, 1.0, 3)
        lo, hi = (centre - half).astype(np.float32), (centre + half).astype(np.float32)
        box = np.zeros((t.n_vars, 2), dtype=np.float32)
        for axis, slot in enumerate(slots):
            if slot >= 0:
                box[slot] = (lo[axis], hi[axis])
        (out_lo, out_hi), _, _ = t.interval_eval(box)
        if np.isnan(out_lo) or np.isnan(out_hi):
            continue                                   # the VM gave up on this box (e.g. sqrt of a straddling interval)
        pts = [rng.uniform(lo[k], hi[k], 64).astype(np.float32) for k in range(3)]
        vals = eval_vm_text(text, *pts)
        vals = vals[np.isfinite(vals)]
        slack = 1e-5 * max(1.0, abs(float(out_lo)), abs(float(out_hi)))
        assert (vals >= out_lo - slack).all() and (vals <= out_hi + slack).all(), (name, lo, hi, out_lo, out_hi)
        checked += 1
    assert checked >= 30


@pytest.mark.parametrize("name", IEEE_MODELS)
def test_simplified_tapes_still_compute_the_text_inside_their_box(orc, name):
    """VmData::simplify through the independent reading: a tape simplified with the choices of a box (and once more
    with those of a sub-box) returns, at points inside, exactly what the full text returns -- bit for bit -- and never
    grows."""
    text = model_text(name)
    t = orc.Tape.from_vm(text)
    slots = t.data.var_slots()
    rng = np.random.default_rng(9)

    def box_of(lo, hi):
        b = np.zeros((t.n_vars, 2), dtype=np.float32)
        for axis, slot in enumerate(slots):
            if slot >= 0:
                b[slot] = (lo[axis], hi[axis])
        return b

    def inputs_of(pts):
        ins = [None] * t.n_vars
        for axis, slot in enumerate(slots):
            if slot >= 0:
                ins[slot] = pts[axis]
        return ins

    shrunk = 0
    for _ in range(25 if name == "prospero.vm" else 60):
        centre = rng.uniform(-0.9, 0.9, 3)
        half = rng.choice([0.25, 0.06]) * rng.uniform(0.3, 1.0, 3)
        lo, hi = (centre - half).astype(np.float32), (centre + half).astype(np.float32)
        _, choices, can = t.interval_eval(box_of(lo, hi))
        child = t.simplify(choices) if can else t
        assert child.size <= t.size
        # a sub-box, simplified from the child
        lo2, hi2 = (centre - half / 4).astype(np.float32), (centre + half / 4).astype(np.float32)
        _, choices2, can2 = child.interval_eval(box_of(lo2, hi2))
        grandchild = child.simplify(choices2) if can2 else child
        assert grandchild.size <= child.size
        shrunk += grandchild.size < t.size
        pts = [rng.uniform(lo2[k], hi2[k], 256).astype(np.float32) for k in range(3)]
        want = eval_vm_text(text, *pts)
        for tape in (child, grandchild):
            got = tape.float_slice_eval(inputs_of(pts))
            both_nan = np.isnan(want) & np.isnan(got)
            assert np.array_equal(want.view(np.uint32)[~both_nan], got.view(np.uint32)[~both_nan])
    assert shrunk > 0 or t.choice_count == 0          # tanglecube has no min / max: nothing to prune


@pytest.mark.parametrize("name,size", [("hi.vm", 96), ("quarter.vm", 96), ("prospero.vm", 128)])
def test_rendered_image_is_the_text_evaluated_at_every_pixel(orc, name, size):
    """The whole chain -- text -> front end -> register tape -> tile recursion, interval proofs, chained simplification,
    leaf evaluation -- against the independent reading at every pixel's model-space point: same sign everywhere, same
    bits wherever the renderer shaded a pixel instead of filling a proven tile."""
    from test_oracle_bruteforce import model_points
    text = model_text(name)
    t = orc.Tape.from_vm(text)
    img, st = orc.render2d(t, size, size, tile_sizes=(32, 8))
    ys, xs = np.mgrid[0:size, 0:size]
    pts = model_points(orc, orc.pixel_mat(size, size), xs.ravel(), ys.ravel(), np.zeros(size * size))
    want = eval_vm_text(text, np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1]),
                        np.ascontiguousarray(pts[:, 2])).reshape(size, size)
    bits = img.view(np.uint32)
    is_fill = np.isnan(img) & ((bits & np.uint32(0xFF << 9)) == np.uint32(0xF6 << 9))
    assert np.array_equal(orc.pixel_inside(img), want < 0)
    assert np.array_equal(img.view(np.uint32)[~is_fill], want.view(np.uint32)[~is_fill])
    assert is_fill.sum() + st["pixels"] == size * size and st["pixels"] > 0
