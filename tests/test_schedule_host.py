"""Host logic of the cooperative level-0 kernel, checked without a GPU: fc_schedule_check builds the
schedule fc_tape_create would upload (dependency waves, serial / chain tail segments, slot colouring)
and replays it symbolically -- every operand slot must hold the defining clause's value when it is read,
and no clause of a concurrent step may overwrite a slot the step still reads."""
import glob
import os

import numpy as np
import pytest

import fidget_b200 as fb
from conftest import MODELS
from test_gpu_fuzz import random_shape


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(MODELS, "*.vm"))))
def test_models_schedule_is_consistent(path):
    ctx, root = fb.Context.from_text(open(path).read())
    tape = ctx.tape(root)
    info = fb.schedule_check(tape)
    assert info["n_clauses"] == len(tape)
    if info["suitable"]:
        assert 0 < info["n_slots"] <= info["n_clauses"]
        assert info["n_waves"] >= 1 and info["widest_wave"] >= 1
        assert info["n_chain_clauses"] <= info["n_tail"]


def test_prospero_schedule_shape():
    # the numbers DESIGN.md quotes: 18 waves, the 657-clause min chain in the tail, 2687 slots for 6363 values
    ctx, root = fb.Context.from_text(open(os.path.join(MODELS, "prospero.vm")).read())
    info = fb.schedule_check(ctx.tape(root))
    assert info == {"suitable": 1, "n_clauses": 6363, "n_waves": 18, "widest_wave": 1337, "n_tail": 660,
                    "n_segments": 5, "n_chain_clauses": 657, "n_slots": 2687}


@pytest.mark.parametrize("seed", range(40))
def test_random_csg_schedule_is_consistent(seed):
    rng = np.random.default_rng(seed)
    ctx = fb.Context()
    root = random_shape(ctx, rng, int(rng.integers(4, 120)), use_z=bool(seed % 2))
    tape = ctx.tape(root)
    info = fb.schedule_check(tape)          # raises CudaError on any inconsistency
    assert info["n_clauses"] == len(tape)
    if len(tape) >= 64:
        assert info["suitable"] == 1 and info["n_slots"] < info["n_clauses"]


def test_few_registers_and_spills():
    # tapes that spill to memory slots are left to the per-lane kernel
    ctx, root = fb.Context.from_text(open(os.path.join(MODELS, "prospero.vm")).read())
    info = fb.schedule_check(ctx.tape(root, n_regs=8))
    assert info["suitable"] == 0
