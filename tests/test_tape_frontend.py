"""Host tape front end (fidget_b200/csrc/host) against the reference's own
tape-shape and bytecode unit tests.  CPU only."""
import numpy as np
import pytest

from conftest import model_text
from fidget_b200.host import OP


@pytest.fixture()
def Ctx(orc):
    return orc.Context


def test_ssa_ring(Ctx):
    # fidget-core/src/compiler/ssa_tape.rs:426-443
    ctx = Ctx()
    c0 = ctx.constant(0.5)
    x, y = ctx.x(), ctx.y()
    r = ctx.add(ctx.square(x), ctx.square(y))
    c6 = ctx.sub(r, c0)
    c8 = ctx.sub(ctx.constant(0.25), r)
    c9 = ctx.max(c8, c6)
    t = ctx.tape(c9)
    assert t.info.ssa_len == 9 and t.n_vars == 2


def test_ssa_dupe_and_constant(Ctx):
    # ssa_tape.rs:445-464
    ctx = Ctx()
    x = ctx.x()
    t = ctx.tape(ctx.mul(x, x))
    assert t.info.ssa_len == 3 and t.n_vars == 1      # x, square, output
    ctx = Ctx()
    t = ctx.tape(ctx.constant(1.5))
    assert t.info.ssa_len == 2 and t.n_vars == 0      # CopyImm, output


def test_context_ring_and_dupe_vmdata_len(Ctx):
    # fidget-core/src/context/mod.rs:1609-1637 (VmData::len counts register-tape clauses)
    ctx = Ctx()
    c0 = ctx.constant(0.5)
    x, y = ctx.x(), ctx.y()
    r = ctx.add(ctx.square(x), ctx.square(y))
    c9 = ctx.max(ctx.sub(ctx.constant(0.25), r), ctx.sub(r, c0))
    t = ctx.tape(c9)
    assert len(t) == 9 and t.n_vars == 2
    ctx = Ctx()
    x = ctx.x()
    t = ctx.tape(ctx.mul(x, x))
    assert len(t) == 3 and t.n_vars == 1
    # import_optimization (mod.rs:1665-1671): x + 0 folds to x
    ctx = Ctx()
    x = ctx.x()
    assert ctx.add(x, 0.0) == x


def test_vmdata_doc_example(Ctx):
    # fidget-core/src/vm/data.rs:46-58
    ctx = Ctx()
    s = ctx.add(ctx.x(), ctx.y())
    t = ctx.tape(s)
    assert len(t) == 4
    lines = t.dump().strip().split("\n")
    vx, vy, _ = t.var_slots()
    assert lines[0] == f"Input 0 <- in[{vx}]"
    assert lines[1] == f"Input 1 <- in[{vy}]"
    assert lines[2] == "Add 0 <- 0, 1"


def test_simplify_reg_count_change(orc, Ctx):
    # fidget-core/src/vm/data.rs:411-437
    ctx = Ctx()
    x, y, z = ctx.x(), ctx.y(), ctx.z()        # node order matters: commutative operands are sorted by id
    xyz = ctx.add(ctx.add(x, y), z)
    d3 = orc.Tape.from_data(ctx.tape(xyz, 3))
    assert d3.size == 6                                  # 3x input, 2x add, 1x output
    assert d3.simplify([], n_regs=2).size == 8           # extra load + store
    d2 = orc.Tape.from_data(ctx.tape(xyz, 2))
    assert d2.size == 8
    assert d2.simplify([], n_regs=3).size == 6


def _word(op, b1=0xFF, b2=0xFF, b3=0xFF):
    return OP[op] | b1 << 8 | b2 << 16 | b3 << 24


def test_bytecode_simple(Ctx):
    # fidget-bytecode/src/lib.rs:352-379
    ctx = Ctx()
    out = ctx.add(ctx.x(), ctx.constant(1.0))
    bc = ctx.tape(out).bytecode()
    one = int(np.float32(1.0).view(np.uint32))
    assert bc.words.tolist() == [
        0xFFFFFFFF, 0,
        _word("input", 0), 0,
        _word("add", 0, 0), one,
        _word("output", 0), 0,
        0xFFFFFFFF, 0xFFFFFFFF]


def test_bytecode_load_store(Ctx):
    # fidget-bytecode/src/lib.rs:381-450: two registers force a spill
    ctx = Ctx()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    out = ctx.max(ctx.max(x, y), z)
    bc = ctx.tape(out, 2).bytecode()
    assert bc.reg_count == 2 and bc.mem_count == 1
    assert bc.words.tolist() == [
        0xFFFFFFFF, 0,
        _word("input", 1), 2,            # Input(1, Z)
        _word("mem", 0xFF, 1), 0,        # reg[1] -> mem[0]
        _word("input", 1), 1,            # Input(1, Y)
        _word("input", 0), 0,            # Input(0, X)
        _word("max", 1, 1, 0), 0xFF000000,
        _word("mem", 0), 0,              # mem[0] -> reg[0]
        _word("max", 0, 0, 1), 0xFF000000,
        _word("output", 0), 0,
        0xFFFFFFFF, 0xFFFFFFFF]


@pytest.mark.parametrize("name,clauses,choices", [
    # SURVEY.md §2 row 21 (counts obtained by replaying the reference's Context rules)
    ("prospero.vm", 6363, 2878), ("hi.vm", 46, 18), ("bear.vm", 541, 27), ("colonnade.vm", 680, 332)])
def test_model_tape_shapes(orc, name, clauses, choices):
    t = orc.Tape.from_vm(model_text(name))
    assert t.ssa_len == clauses and t.choice_count == choices
    assert t.size == clauses                     # 255 registers: no spills on these models


def test_context_folding_rules(Ctx):
    # fidget-core/src/context/mod.rs:234-322,586-623
    ctx = Ctx()
    x, y = ctx.x(), ctx.y()
    assert ctx.add(x, ctx.constant(0.0)) == x
    assert ctx.mul(x, ctx.constant(1.0)) == x
    zero = ctx.constant(0.0)
    assert ctx.mul(x, zero) == zero
    assert ctx.mul(x, x) == ctx.square(x)
    assert ctx.add(x, x) == ctx.mul(x, ctx.constant(2.0))
    assert ctx.min(x, x) == x and ctx.max(y, y) == y
    assert ctx.add(x, y) == ctx.add(y, x)                       # commutative dedup
    assert ctx.sub(zero, x) == ctx.neg(x)
    assert ctx.div(x, ctx.constant(1.0)) == x
    assert ctx.add(ctx.constant(1.0), ctx.constant(2.0)) == ctx.constant(3.0)   # constant folding
    n = len(ctx)
    ctx.x()
    assert len(ctx) == n                                        # deduplicated


def test_from_text_errors(Ctx):
    with pytest.raises(Exception):
        Ctx.from_text("")
    with pytest.raises(Exception):
        Ctx.from_text("a frobnicate\n")
    with pytest.raises(Exception):
        Ctx.from_text("a neg b\n")


def test_register_spill_stress_matches_unspilled(orc, Ctx):
    # eval/test/mod.rs:20-45 + float_slice.rs:296-315: GenericVmFunction<3> must equal VmFunction exactly
    ctx = Ctx()
    x, y, z = ctx.x(), ctx.y(), ctx.z()
    s = ctx.constant(0.0)
    inputs = []
    for i in range(1, 65):
        d = ctx.mul(ctx.constant(float(i)), [x, y, z][i % 3])
        inputs.append(d)
        s = ctx.add(s, d)
    s = ctx.sin(s)
    for d in reversed(inputs):
        s = ctx.add(s, d)
    big, small = orc.Tape.from_data(ctx.tape(s)), orc.Tape.from_data(ctx.tape(s, 3))
    assert small.size > big.size
    rng = np.random.default_rng(0)
    pts = [rng.uniform(-1, 1, 33).astype(np.float32) for _ in range(3)]
    assert np.array_equal(big.float_slice_eval(pts).view(np.uint32), small.float_slice_eval(pts).view(np.uint32))


def test_tape_blob_layout(Ctx):
    """The wire / on-disk form ("FTAP", include/fidget_cuda.h) carries the bytecode words of
    fidget_bytecode::Bytecode::new plus exactly the metadata fc_tape_create needs."""
    import struct
    ctx, root = Ctx.from_text(model_text("hi.vm"))
    td = ctx.tape(root)
    blob, bc = td.serialize(), td.bytecode()
    assert blob[:4] == b"FTAP"
    version, regs, mem, n_vars, n_out, n_choice = struct.unpack("<6I", blob[4:28])
    assert (version, regs, mem, n_vars, n_out, n_choice) == (1, bc.reg_count, bc.mem_count, td.n_vars, 1, td.choice_count)
    assert struct.unpack("<3i", blob[28:40]) == td.var_slots()
    (n_words,) = struct.unpack("<Q", blob[40:48])
    assert n_words == len(bc.words) and len(blob) == 48 + 4 * n_words
    assert np.array_equal(np.frombuffer(blob, dtype=np.uint32, offset=48), bc.words)
    spilled = ctx.tape(root, 3)
    assert struct.unpack("<2I", spilled.serialize()[8:16]) == (3, spilled.bytecode().mem_count)
