"""The long-pair wavefront kernel (csrc/kern_wave32.cuh) on the CPU SIMT emulation vs the oracle.
The emulation runs blocks one after another, so the inter-strip waits are always already
satisfied; what is checked here is the strip hand-over and the end-cell bookkeeping.  The
concurrent schedule is exercised on the GPU (tests/test_gpu_parity.py::test_long_pair_wavefront)."""
import numpy as np
import pytest

import emu_harness
import psb_data
from test_oracle_properties import SG_FLAGS


@pytest.mark.parametrize("v2", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("K", [1, 2])
def test_wave_matches_oracle(oracle, mode, K, v2):
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    r = psb_data.random_seq(5001, 0, 170, protein=False)
    q = psb_data.mutate(r, 5001, 1, 0.10, 0.02, protein=False)[:150]
    for o, e in ((5, 2), (3, 3), (0, 0)):
        exp = oracle.align(q, r, mat, mode=mode, open=o, gap=e)
        got = emu_harness.wave32(q, r, mat, K, mode, o, e, v2=v2)
        assert got == (exp["score"], exp["end_query"], exp["end_ref"]), (o, e)


@pytest.mark.parametrize("v2", [False, True])
@pytest.mark.parametrize("flags", SG_FLAGS[1:])
def test_wave_sg_flags(oracle, flags, v2):
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    q = psb_data.random_seq(5002, 0, 100, protein=False)
    r = psb_data.random_seq(5002, 1, 90, protein=False)
    exp = oracle.align(q, r, mat, mode=1, open=5, gap=2, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2], s2_end=flags[3])
    got = emu_harness.wave32(q, r, mat, 1, 1, 5, 2, flags, v2=v2)
    assert got == (exp["score"], exp["end_query"], exp["end_ref"])


def test_wave_multi_pair(oracle, blosum62):
    # database-scan form: several subjects, strips of all of them in one work queue
    q = psb_data.random_seq(5003, 0, 100)
    subs = [psb_data.random_seq(5004, i, 40 + 25 * i) for i in range(4)]
    subs[2] = np.concatenate([subs[2][:20], psb_data.mutate(q, 5005, 0, 0.1, 0.02)])
    for mode, v2 in ((0, False), (2, False), (0, True), (2, True)):
        got = emu_harness.wave32_multi(q, subs, blosum62, 1, mode, 10, 1, v2=v2)
        for i, s in enumerate(subs):
            exp = oracle.align(q, s, blosum62, mode=mode, open=10, gap=1)
            assert (got[0][i], got[1][i], got[2][i]) == (exp["score"], exp["end_query"], exp["end_ref"]), (mode, i)


@pytest.mark.parametrize("K", [4, 8, 16])
def test_wave_gen3_matches_oracle(oracle, K):
    # column-blocked generation (local alignment): several strips, reference lengths around the
    # 4-column block and the 32-column staging boundaries, gap penalties incl. open == extend and 0/0
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    lq = {4: 300, 8: 300, 16: 600}[K]
    for lr in (1, 3, 4, 5, 31, 32, 33, 64, 97, 130, 257):
        r = psb_data.random_seq(5101, lr, lr, protein=False)
        base = np.concatenate([r] * (lq // max(lr, 1) + 2))[: lq + 40]
        q = psb_data.mutate(base, 5102, lr, 0.10, 0.02, protein=False)[:lq]
        for o, e in ((5, 2), (3, 3), (0, 0)):
            exp = oracle.align(q, r, mat, mode=2, open=o, gap=e)
            got = emu_harness.wave32(q, r, mat, K, 2, o, e, v2=2)
            assert got == (exp["score"], exp["end_query"], exp["end_ref"]), (K, lr, o, e)


def test_wave_gen3_protein_and_ties(oracle, blosum62):
    # protein scores, a query shorter than one strip, repeated sequence (many equal maxima: the
    # smallest end_ref, then the smallest end_query must win), and an all-mismatch pair (score 0)
    q = psb_data.random_seq(5103, 0, 90)
    r = np.concatenate([q[10:60]] * 5)
    for o, e in ((10, 1), (11, 11)):
        exp = oracle.align(q, r, blosum62, mode=2, open=o, gap=e)
        got = emu_harness.wave32(q, r, blosum62, 4, 2, o, e, v2=2)
        assert got == (exp["score"], exp["end_query"], exp["end_ref"])
    dna = oracle.Matrix.create(b"ACGT", 2, -3)
    a = np.frombuffer(b"A" * 200, dtype=np.uint8)
    c = np.frombuffer(b"C" * 77, dtype=np.uint8)
    assert emu_harness.wave32(a, c, dna, 4, 2, 5, 2, v2=2) == (0, 0, 0)
    exp = oracle.align(a, a[:77], dna, mode=2, open=5, gap=2)
    assert emu_harness.wave32(a, a[:77], dna, 4, 2, 5, 2, v2=2) == (exp["score"], exp["end_query"], exp["end_ref"])


def test_wave_gen3_multi_pair(oracle, blosum62):
    q = psb_data.random_seq(5003, 0, 300)
    subs = [psb_data.random_seq(5004, i, 40 + 25 * i) for i in range(4)]
    subs[2] = np.concatenate([subs[2][:20], psb_data.mutate(q, 5005, 0, 0.1, 0.02)])
    got = emu_harness.wave32_multi(q, subs, blosum62, 4, 2, 10, 1, v2=2)
    for i, s in enumerate(subs):
        exp = oracle.align(q, s, blosum62, mode=2, open=10, gap=1)
        assert (got[0][i], got[1][i], got[2][i]) == (exp["score"], exp["end_query"], exp["end_ref"]), i


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K", [4, 8])
def test_wave_gen3_global_modes(oracle, mode, K):
    # the column-blocked generation for nw / sg: several strips, reference lengths around the block size
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    for lr in (1, 4, 5, 33, 130, 259):
        r = psb_data.random_seq(5301, lr, lr, protein=False)
        base = np.concatenate([r] * (300 // lr + 2))[:340]
        q = psb_data.mutate(base, 5302, lr, 0.10, 0.02, protein=False)[:300]
        for o, e in ((5, 2), (3, 3), (0, 0)):
            exp = oracle.align(q, r, mat, mode=mode, open=o, gap=e)
            got = emu_harness.wave32(q, r, mat, K, mode, o, e, v2=2)
            assert got == (exp["score"], exp["end_query"], exp["end_ref"]), (mode, K, lr, o, e)


@pytest.mark.parametrize("flags", SG_FLAGS[1:])
def test_wave_gen3_sg_flags(oracle, blosum62, flags):
    q = psb_data.random_seq(5303, 0, 290)
    r = psb_data.mutate(q, 5304, 0, 0.2, 0.04)[40:231]
    exp = oracle.align(q, r, blosum62, mode=1, open=10, gap=1, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2], s2_end=flags[3])
    got = emu_harness.wave32(q, r, blosum62, 4, 1, 10, 1, flags, v2=2)
    assert got == (exp["score"], exp["end_query"], exp["end_ref"])
