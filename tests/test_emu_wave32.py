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


# ---- long pairs WITH traceback / statistics: TRACE instantiation of generation 3 + walk32_kernel -----------------
def _check_trace(oracle, mat, q, r, mode, o, e, flags=(1, 1, 1, 1)):
    exp = oracle.align(q, r, mat, mode=mode, open=o, gap=e, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2], s2_end=flags[3], trace=True)
    got = emu_harness.wave32_trace(q, r, mat, mode, o, e, flags, what=1)
    tag = (len(q), len(r), mode, o, e, flags)
    assert (got["score"], got["end_query"], got["end_ref"]) == (exp["score"], exp["end_query"], exp["end_ref"]), tag
    assert np.array_equal(got["cigar_ops"], exp["cigar_ops"]), (tag, oracle.decode_cigar(got["cigar_ops"]), exp["cigar"])
    assert (got["beg_query"], got["beg_ref"]) == (exp["beg_query"], exp["beg_ref"]), tag
    exs = exp   # (the oracle's statistics recurrences run with every call)
    gs = emu_harness.wave32_trace(q, r, mat, mode, o, e, flags, what=2)
    assert (gs["matches"], gs["similar"], gs["length"]) == (exs["matches"], exs["similar"], exs["length"]), tag


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_wave_trace_matches_oracle(oracle, mode):
    # several strips (256 rows each), reference lengths around the 4-column block, indels on the path,
    # open == extend and 0/0 penalties (every decision is a tie)
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    for lr in (5, 33, 130, 259):
        r = psb_data.random_seq(5401, lr, lr, protein=False)
        base = np.concatenate([r] * (600 // lr + 2))[:640]
        q = psb_data.mutate(base, 5402, lr, 0.10, 0.04, protein=False)[:600]
        for o, e in ((5, 2), (3, 3), (0, 0)):
            _check_trace(oracle, mat, q, r, mode, o, e)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_wave_trace_protein_long_gaps(oracle, blosum62, mode):
    # protein scores; the reference lacks two stretches of the query and carries an insert, so the path has gaps
    # that cross lane (8 rows) and strip (256 rows) boundaries
    q = psb_data.random_seq(5403, 0, 560)
    r = np.concatenate([q[:100], q[130:250], psb_data.random_seq(5404, 0, 37), q[250:270], q[300:]])
    r = psb_data.mutate(r, 5405, 0, 0.08, 0.01)
    for o, e in ((10, 1), (11, 11)):
        _check_trace(oracle, blosum62, q, r, mode, o, e)


@pytest.mark.parametrize("flags", SG_FLAGS[1:])
def test_wave_trace_sg_flags(oracle, blosum62, flags):
    q = psb_data.random_seq(5406, 0, 290)
    r = psb_data.mutate(q, 5407, 0, 0.2, 0.04)[40:231]
    _check_trace(oracle, blosum62, q, r, 1, 10, 1, flags)


def test_wave_trace_ties_and_empty(oracle):
    # repeats (equal maxima, equal-score paths) and an all-mismatch local pair (score 0, empty CIGAR)
    dna = oracle.Matrix.create(b"ACGT", 2, -3)
    unit = np.frombuffer(b"ACGTTGCAAC", dtype=np.uint8)
    q = np.concatenate([unit] * 30)
    r = np.concatenate([unit[:7]] * 20)
    for mode in (0, 1, 2):
        for o, e in ((5, 2), (1, 1), (0, 0), (2, 0)):
            _check_trace(oracle, dna, q, r, mode, o, e)
    a = np.frombuffer(b"A" * 300, dtype=np.uint8)
    c = np.frombuffer(b"C" * 77, dtype=np.uint8)
    for mode in (0, 1, 2):
        _check_trace(oracle, dna, a, c, mode, 5, 2)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_wave_gen3_steady_state_steps(oracle, mode):
    # references long enough that most steps take the all-lanes-active form of the step (31 <= s <= blocks - 4), with
    # lengths around the 4-column block, three strips (K = 8), score-only and traced launches
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    for lr in (689, 690, 691, 692):
        r = psb_data.random_seq(5501, lr, lr, protein=False)
        q = psb_data.mutate(r, 5502, lr, 0.10, 0.03, protein=False)[:600]
        exp = oracle.align(q, r, mat, mode=mode, open=5, gap=2)
        assert emu_harness.wave32(q, r, mat, 8, mode, 5, 2, v2=2) == (exp["score"], exp["end_query"], exp["end_ref"]), lr
    _check_trace(oracle, mat, q, r, mode, 5, 2)
    if mode == 1:
        for flags in SG_FLAGS[1:]:
            exp = oracle.align(q, r, mat, mode=1, open=5, gap=2, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2], s2_end=flags[3])
            assert emu_harness.wave32(q, r, mat, 8, 1, 5, 2, flags, v2=2) == (exp["score"], exp["end_query"], exp["end_ref"]), flags
