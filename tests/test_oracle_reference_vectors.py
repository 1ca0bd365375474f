"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for
the alignment path.  Each test cites the Rust test it restates."""
import numpy as np
import pytest


def run(oracle, mat, q, r, mode, **kw):
    return oracle.align(q, r, mat, mode=mode, **kw)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_acgt_identity(oracle, dna_default, mode):
    # [REF tests/test_parasail.rs:64-122] global/semi_global/local_alignment: 4, 3, 3
    res = run(oracle, dna_default, b"ACGT", b"ACGT", mode)
    assert (res["score"], res["end_query"], res["end_ref"]) == (4, 3, 3)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_acgt_stats(oracle, dna_default, mode):
    # [REF tests/test_parasail.rs:124-173] *_with_stats: matches 4, length 4
    res = run(oracle, dna_default, b"ACGT", b"ACGT", mode)
    assert res["matches"] == 4 and res["length"] == 4 and res["similar"] == 4


def test_twelve_mer_free_gaps(oracle, dna_default):
    # [REF tests/test_parasail.rs:175-253] global_8/16/32/64bit: gaps 0/0 => the mismatch is
    # skipped by two free gaps: score 11, ends 11/11
    res = run(oracle, dna_default, b"ACTGACTGACTG", b"ACTGTCTGACTG", 0)
    assert (res["score"], res["end_query"], res["end_ref"]) == (11, 11, 11)


def test_score_table(oracle, dna_default):
    # [REF tests/test_parasail.rs:255-325] dims 4x4, last() == 4; custom matrix (3,-2) => 12
    res = run(oracle, dna_default, b"ACGT", b"ACGT", 0, tables=True)
    assert res["score_table"].shape == (4, 4) and res["score_table"][-1, -1] == 4
    m = oracle.Matrix.create(b"ACGT", 3, -2)
    res = run(oracle, m, b"ACGT", b"ACGT", 0, tables=True)
    assert res["score_table"][-1, -1] == 12


def test_stats_tables(oracle, dna_default):
    # [REF tests/test_parasail.rs:327-383] matches_table ACGT/ACGTT last()==4; dims of the others
    res = run(oracle, dna_default, b"ACGT", b"ACGTT", 0, tables=True)
    assert res["matches_table"].shape == (4, 5) and res["matches_table"][-1, -1] == 4
    res = run(oracle, dna_default, b"ACGT", b"ACGTTT", 0, tables=True)
    assert res["length_table"].shape == (4, 6)
    res = run(oracle, dna_default, b"ACGT", b"ACGT", 0, tables=True)
    assert res["similar_table"].shape == (4, 4)


def test_last_rows(oracle, dna_default):
    # [REF tests/test_parasail.rs:385-463] query ACGT, ref ACG, nw_stats_rowcol
    res = run(oracle, dna_default, b"ACGT", b"ACG", 0, rowcol=True)
    assert list(res["score_row"]) == [1, 2, 3]
    assert list(res["matches_row"]) == [1, 2, 3]
    assert list(res["similar_row"]) == [1, 2, 3]
    assert list(res["length_row"]) == [4, 4, 4]


def test_last_cols(oracle, dna_default):
    # [REF tests/test_parasail.rs:465-543] query ACG, ref ACGT
    res = run(oracle, dna_default, b"ACG", b"ACGT", 0, rowcol=True)
    assert list(res["score_col"]) == [1, 2, 3]
    assert list(res["matches_col"]) == [1, 2, 3]
    assert list(res["similar_col"]) == [1, 2, 3]
    assert list(res["length_col"]) == [4, 4, 4]


def test_trace_table_flags_valid(oracle, dna_default):
    # [REF tests/test_parasail.rs:545-578] nw_trace: 4x4, 16 bytes, every cell decodes to
    # known TraceFlags bits [REF src/alignment/table.rs:127-142]
    res = run(oracle, dna_default, b"ACGT", b"ACGT", 0, trace=True)
    t = res["trace"]
    assert t.shape == (4, 4) and t.size == 16
    assert np.all((t.astype(np.int32) & ~127) == 0)
    assert res["cigar"] == "4="  # ACGT vs ACGT => "4="
    assert res["traceback"] == ("ACGT", "||||", "ACGT")


def test_multithread_vector(oracle, dna_default):
    # [REF tests/test_parasail.rs:689-723] nw with a stats profile: score 4 (per thread)
    res = run(oracle, dna_default, b"ACGT", b"ACGT", 0)
    assert res["score"] == 4
