"""Oracle self-consistency: cross-check against an independently written pure-Python Gotoh,
hand-derived tie-break cases and oracle-free properties of CIGARs."""
import numpy as np
import pytest

import psb_data
import py_gotoh

MODES = {0: "nw", 1: "sg", 2: "sw"}
SG_FLAGS = [(1, 1, 1, 1), (1, 0, 0, 0), (0, 1, 0, 0), (1, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (0, 0, 1, 1),
            (1, 0, 0, 1), (0, 1, 1, 0), (1, 0, 1, 0), (0, 1, 0, 1)]


def rand_pair(seed, i, lq, lr, protein):
    q = psb_data.random_seq(seed, 2 * i, lq, protein)
    if i % 3 == 0:
        r = psb_data.mutate(q, seed, 2 * i + 1, 0.2, 0.1, protein)
        r = r[:lr] if len(r) >= lr else np.concatenate([r, psb_data.random_seq(seed + 7, i, lr - len(r), protein)])
    else:
        r = psb_data.random_seq(seed, 2 * i + 1, lr, protein)
    return q, r


@pytest.mark.parametrize("protein", [False, True])
@pytest.mark.parametrize("gaps", [(0, 0), (5, 2), (10, 1), (3, 3), (1, 4)])
def test_cross_check_python(oracle, dna_default, blosum62, protein, gaps):
    mat = blosum62 if protein else oracle.Matrix.create(b"ACGT", 2, -3)
    o, e = gaps
    rng = np.random.default_rng(123)
    for i in range(12):
        lq, lr = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        q, r = rand_pair(17, i, lq, lr, protein)
        for mode in (0, 2):
            res = oracle.align(q, r, mat, mode=mode, open=o, gap=e)
            exp = py_gotoh.gotoh(q, r, mat.table, mat.mapper, MODES[mode], o, e)
            assert (res["score"], res["end_query"], res["end_ref"]) == exp, (mode, i)
        for fl in SG_FLAGS:
            res = oracle.align(q, r, mat, mode=1, open=o, gap=e, s1_beg=fl[0], s1_end=fl[1], s2_beg=fl[2],
                               s2_end=fl[3])
            exp = py_gotoh.gotoh(q, r, mat.table, mat.mapper, "sg", o, e, *fl)
            assert (res["score"], res["end_query"], res["end_ref"]) == exp, (fl, i)


def test_sw_tie_two_columns(oracle):
    # equal maxima in two columns: the smaller end_ref wins (A.4)
    m = oracle.Matrix.create(b"ACGT", 2, -3)
    res = oracle.align(b"AC", b"ACTTAC", m, mode=2, open=5, gap=2)
    assert (res["score"], res["end_query"], res["end_ref"]) == (4, 1, 1)


def test_sw_tie_two_rows_same_column(oracle):
    # equal maxima in one column, two rows: the smaller end_query wins (A.4)
    m = oracle.Matrix.create(b"ACGT", 2, -3)
    # query ACTAC, ref AC: "AC" ends at column 1 for rows 1 and 4
    res = oracle.align(b"ACTAC", b"AC", m, mode=2, open=5, gap=2)
    assert (res["score"], res["end_query"], res["end_ref"]) == (4, 1, 1)


def test_sg_row_beats_column_on_tie(oracle):
    # last row and last column tie: the last row (smaller end_ref) keeps the hit (A.4)
    m = oracle.Matrix.create(b"ACGT", 2, -3)
    res = oracle.align(b"TTAC", b"ACGG", m, mode=1, open=5, gap=2)
    # last row best: "AC" of the query end vs ref[0:2] -> (3,1) score 4; last column best is lower
    assert (res["score"], res["end_query"], res["end_ref"]) == (4, 3, 1)
    res = oracle.align(b"ACGG", b"TTAC", m, mode=1, open=5, gap=2)
    assert (res["score"], res["end_query"], res["end_ref"]) == (4, 1, 3)


def test_sg_corner_tied_with_earlier_last_column_cell(oracle):
    # corner (found while scanning the last row) is kept when an earlier last-column cell ties (A.4)
    m = oracle.Matrix.create(b"ACGT", 1, -1)
    res = oracle.align(b"AA", b"A", m, mode=1, open=0, gap=0, rowcol=True)
    assert list(res["score_col"]) == [1, 1]
    assert (res["score"], res["end_query"], res["end_ref"]) == (1, 1, 0)


def test_priority_diag_over_gaps(oracle):
    # H_dag == F: diagonal wins (A.5) -> CIGAR has no gap when a gap-free path ties
    m = oracle.Matrix.create(b"ACGT", 1, -1)
    res = oracle.align(b"AAT", b"AT", m, mode=0, open=0, gap=0, trace=True, tables=True)
    assert res["score"] == 2
    assert res["cigar"].count("I") == 1 and res["cigar"].endswith("=")


def test_e_tie_extends(oracle):
    # H - o == E - e  => extension (strict > needed to open), flag INS_E not DIAG_E (A.5)
    m = oracle.Matrix.create(b"ACGT", 1, -1)
    res = oracle.align(b"A", b"AAAA", m, mode=0, open=0, gap=0, trace=True)
    t = res["trace"].astype(int)
    # row 0: E at column j>=2 ties between open (H[0][j-1]-0) and extend (E-0): INS_E (16) must be set
    assert all((t[0, j] & 16) for j in range(2, 4))


def test_sw_zero_diag_stops_trace(oracle):
    # SW path through an exact-zero cell: ZERO wins, the walk stops before it (A.5/A.7)
    m = oracle.Matrix.create(b"ACGT", 2, -2)
    res = oracle.align(b"ATGG", b"ACGG", m, mode=2, open=5, gap=2, trace=True, tables=True)
    # A=A (2), T/C (-2) -> 0 -> ZERO; then GG = 4
    assert res["score"] == 4 and res["cigar"] == "2=" and (res["beg_query"], res["beg_ref"]) == (2, 2)


def cigar_rescore(cigar_ops, q, r, bq, br, mat, o, e):
    i, j, score, m_, s_, l_ = bq, br, 0, 0, 0, 0
    for op in cigar_ops:
        n, c = int(op) >> 4, int(op) & 15
        if c in (7, 8):
            for _ in range(n):
                a, b = mat.mapper[q[i]], mat.mapper[r[j]]
                sc = int(mat.table[a][b])
                score += sc
                m_ += int(a == b)
                s_ += int(sc > 0)
                i += 1
                j += 1
        elif c == 1:
            score -= o + (n - 1) * e
            i += n
        elif c == 2:
            score -= o + (n - 1) * e
            j += n
        l_ += n
    return score, i - 1, j - 1, m_, s_, l_


@pytest.mark.parametrize("protein", [False, True])
def test_cigar_rescore_and_recount(oracle, blosum62, protein):
    # oracle-free properties: re-scoring the CIGAR gives the score, recounting gives the stats
    mat = blosum62 if protein else oracle.Matrix.create(b"ACGT", 2, -3)
    o, e = (10, 1) if protein else (5, 2)
    for i in range(25):
        q, r = rand_pair(99, i, 30 + i, 45 - i, protein)
        for mode in (0, 1, 2):
            res = oracle.align(q, r, mat, mode=mode, open=o, gap=e, trace=True)
            sc, ei, ej, m_, s_, l_ = cigar_rescore(res["cigar_ops"], q, r, res["beg_query"], res["beg_ref"], mat, o, e)
            if mode == 2 and res["score"] == 0:
                continue
            assert (ei, ej) == (res["end_query"], res["end_ref"])
            if mode == 0:
                assert sc == res["score"]
                assert (m_, s_) == (res["matches"], res["similar"])
            else:
                # sg / sw: a walk that runs off the top or left edge emits the rest of the other
                # sequence as one leading I/D run (A.7 edge rule); it is free in the score and
                # is not part of the stats length
                first = int(res["cigar_ops"][0])
                if (first & 15) in (1, 2):
                    n = first >> 4
                    sc += o + (n - 1) * e
                    l_ -= n
                assert sc == res["score"]
                if mode == 2:
                    assert (m_, s_, l_) == (res["matches"], res["similar"], res["length"])


def test_mode_ordering(oracle, blosum62):
    # SW >= SG >= NW on the same inputs
    for i in range(20):
        q, r = rand_pair(5, i, 25 + i, 40, True)
        s = [oracle.align(q, r, blosum62, mode=m, open=10, gap=1)["score"] for m in (0, 1, 2)]
        assert s[2] >= s[1] >= s[0]


def test_nw_symmetry(oracle, blosum62):
    for i in range(10):
        q, r = rand_pair(6, i, 20 + i, 33, True)
        a = oracle.align(q, r, blosum62, mode=0, open=10, gap=1)["score"]
        b = oracle.align(r, q, blosum62, mode=0, open=10, gap=1)["score"]
        assert a == b
