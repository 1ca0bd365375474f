"""The packed 16-bit scan kernel source (csrc/kern_sw16.cuh) on the CPU SIMT emulation vs the
oracle: score, end_query, end_ref bit-exact; overflowing subjects must land on the retry list."""
import numpy as np
import pytest

import emu_harness
import psb_data


def check(oracle, mat, query, subjects, o, e, bits=5):
    outs, retry = emu_harness.sw16(query, subjects, mat, o, e, bits)
    for i, s in enumerate(subjects):
        if i in retry:
            continue
        exp = oracle.align(query, s, mat, mode=2, open=o, gap=e)
        got = (outs["score"][i], outs["end_query"][i], outs["end_ref"][i])
        assert got == (exp["score"], exp["end_query"], exp["end_ref"]), (i, len(query), len(s))
    return outs, retry


@pytest.mark.parametrize("lq", [40, 100, 400])
def test_protein_scan(oracle, blosum62, lq):
    q = psb_data.random_seq(7, 0, lq)
    subs = []
    for i in range(7):
        L = [35, 90, 91, 150, 17, 64, 33][i]
        if i % 2 == 0:
            seg = psb_data.mutate(q[: min(lq, L)], 7, 100 + i, 0.2, 0.05)[:L]
            s = np.concatenate([seg, psb_data.random_seq(8, i, max(0, L - len(seg)))])[:L]
        else:
            s = psb_data.random_seq(8, i, L)
        subs.append(s)
    _, retry = check(oracle, blosum62, q, subs, 10, 1)
    assert retry == []


def test_ties_and_zero(oracle):
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    q = np.frombuffer(b"ACTACGGG", dtype=np.uint8)
    subs = [np.frombuffer(s, dtype=np.uint8) for s in (b"ACTTAC", b"AC", b"TTTT", b"ACGGGACGGG", b"GG")]
    check(oracle, mat, q, subs, 5, 2)
    check(oracle, mat, q, subs, 0, 0)
    check(oracle, mat, q, subs, 5, 2, bits=5)


def test_overflow_goes_to_retry(oracle):
    # 400 matches at +100 exceed the 16-bit range: both subjects of that item are re-run at 32 bit,
    # everything else stays exact
    mat = oracle.Matrix.create(b"ACGT", 100, -90)
    q = psb_data.random_seq(9, 0, 400, protein=False)
    subs = [q.copy(), psb_data.random_seq(9, 1, 380, protein=False), psb_data.random_seq(9, 2, 60, protein=False),
            psb_data.random_seq(9, 3, 50, protein=False)]
    _, retry = check(oracle, mat, q, subs, 5, 2)
    assert retry == [0, 1]


def test_high_scores_stay_in_16_bit(oracle, blosum62):
    # an identical 400-aa protein scores ~2100: no re-run needed any more
    q = psb_data.random_seq(9, 0, 400)
    subs = [q.copy(), psb_data.random_seq(9, 1, 380), np.concatenate([psb_data.random_seq(9, 2, 60), q[50:300]])]
    outs, retry = check(oracle, blosum62, q, subs, 10, 1)
    assert retry == [] and outs["score"][0] > 2000


@pytest.mark.parametrize("lq,rows", [(100, 64), (130, 64), (300, 128), (400, 192)])
def test_strip_wise_scan_of_long_queries(oracle, blosum62, lq, rows):
    # queries longer than one strip: one sweep per strip, bottom rows handed over, per-strip bests merged with
    # (score, smaller end_ref, smaller end_query).  Every strip but the last is exactly 16*K rows high (the
    # bottom row of lane 15 must be a real query row)
    q = psb_data.random_seq(17, 0, lq)
    subs = []
    for i, L in enumerate([35, 90, 91, 150, 17, 64, 33, 200]):
        if i % 2 == 0:
            a = (i * 13) % max(1, lq - 30)
            seg = psb_data.mutate(q[a: a + L], 17, 100 + i, 0.2, 0.05)[:L]
            s = np.concatenate([seg, psb_data.random_seq(18, i, max(0, L - len(seg)))])[:L]
        else:
            s = psb_data.random_seq(18, i, L)
        subs.append(s)
    subs.append(np.concatenate([q[rows - 5: rows + 25], q[rows - 5: rows + 25]]))   # a match that straddles the strip boundary, twice
    outs, retry = emu_harness.sw16_strips(q, subs, blosum62, 10, 1, rows)
    assert retry == []
    for i, s in enumerate(subs):
        exp = oracle.align(q, s, blosum62, mode=2, open=10, gap=1)
        assert (outs["score"][i], outs["end_query"][i], outs["end_ref"][i]) == (exp["score"], exp["end_query"], exp["end_ref"]), (i, lq, len(s))


def test_strip_wise_scan_ties_across_strips(oracle):
    # the same maximum in two strips: the smaller end_ref must win whichever strip finds it
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    unit = np.frombuffer(b"ACGTTGCA", dtype=np.uint8)
    q = np.concatenate([unit, psb_data.random_seq(19, 0, 60, protein=False), unit, psb_data.random_seq(19, 1, 40, protein=False)])
    subs = [np.concatenate([psb_data.random_seq(19, 2, 11, protein=False), unit, psb_data.random_seq(19, 3, 9, protein=False)]),
            unit.copy(), np.concatenate([unit, unit]), psb_data.random_seq(19, 4, 50, protein=False)]
    outs, _ = emu_harness.sw16_strips(q, subs, mat, 5, 2, 64, bits=3)
    for i, s in enumerate(subs):
        exp = oracle.align(q, s, mat, mode=2, open=5, gap=2)
        assert (outs["score"][i], outs["end_query"][i], outs["end_ref"][i]) == (exp["score"], exp["end_query"], exp["end_ref"]), i
