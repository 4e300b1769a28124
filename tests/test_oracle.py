"""CPU tests: the plain-C oracle (oracle/cvtt_oracle.c) against the golden vectors recorded from the unmodified
reference, and against the reference itself where oracle/_ref is available."""
import struct

import numpy as np
import pytest

from conftest import golden_names, load_golden, first_mismatch
from convectionkernels_b200 import synth


@pytest.mark.parametrize("name", golden_names("bc7_"))
def test_oracle_matches_golden(oracle, name):
    g = load_golden(name)
    if struct.unpack("<I", bytes(g["options"][0:4]))[0] & 0x20:
        pytest.skip("Flags::BC7_RespectPunchThrough needs the 8-lane lock-step model; the scalar C restatement rejects it (oracle/cvtt_oracle.c), "
                    "the unmodified reference build (oracle/_ref) is the checker for these fixtures")
    oracle.set_rcp_table(g["rcp"])
    try:
        got = oracle.encode_bc7(g["blocks"], g["options"], g["plan"])
    finally:
        oracle.set_rcp_table(None)
    assert (got == g["expected"]).all(), first_mismatch(g["expected"], got)


def test_golden_bc1_plumbing(reference):
    """BASELINE.json configs[0]: EncodeBC1 on the 256x256 gradient through the reference CPU path, 8 blocks per call."""
    g = load_golden("bc1_gradient256")
    blocks = synth.image_to_blocks(synth.gradient_rgba8(256, 256))
    assert (blocks == g["blocks"]).all()
    got = reference.encode("BC1", blocks, g["options"])
    assert got.shape == (4096, 8)
    assert (got == g["expected"]).all()


@pytest.mark.parametrize("seed", [101, 202])
def test_oracle_matches_reference_live(oracle, reference, seed):
    blocks = synth.random_blocks_rgba8(128, seed=seed)
    opt = reference.default_options()
    for plan in (reference.plan_from_quality(100), reference.default_plan(), reference.plan_from_quality(17)):
        want = reference.encode("BC7", blocks, opt, plan)
        got = oracle.encode_bc7(blocks, opt, plan)
        assert (got == want).all(), first_mismatch(want, got)


def test_oracle_plans_match_reference(oracle, reference):
    for q in (1, 2, 10, 33, 50, 77, 99, 100, 1000, -5):
        assert (oracle.plan_from_quality(q) == reference.plan_from_quality(q)).all(), q
    rng = np.random.default_rng(5)
    for _ in range(20):
        ft = rng.integers(0, 5, size=285, dtype=np.uint8)
        ft[rng.random(285) < 0.3] = 0
        assert (oracle.plan_from_finetune(ft) == reference.plan_from_finetune(ft)).all()


def test_group_coupling_is_reproduced(oracle, reference):
    """The same block gives different bytes depending on its 7 neighbours only through the two group votes
    (reference BC67.cpp:1069-1072); the oracle must track the reference when neighbours change."""
    rng = np.random.default_rng(9)
    probe = synth.random_blocks_rgba8(8, seed=4)[3].copy()
    probe[:, 3] = 255
    probe[5, 3] = 252                 # minAlpha in (250, 255): both flags matter
    opt, plan = reference.default_options(), reference.default_plan()
    for trial in range(6):
        grp = synth.random_blocks_rgba8(8, seed=50 + trial)
        if trial % 2 == 0:
            grp[:, :, 3] = 255
        else:
            grp[:, :, 3] = rng.integers(0, 200, size=(8, 16))
        grp[trial % 8] = probe
        want = reference.encode("BC7", grp, opt, plan)
        got = oracle.encode_bc7(grp, opt, plan)
        assert (got == want).all(), first_mismatch(want, got)


def test_oracle_rejects_unrestated_flags(oracle):
    blocks = np.zeros((8, 16, 4), np.uint8)
    opt = np.zeros(44, np.uint8)
    opt[0:4] = np.frombuffer(struct.pack("<I", 0x020), np.uint8)
    with pytest.raises(ValueError):
        oracle.encode_bc7(blocks, opt, oracle.plan_from_quality(100))
