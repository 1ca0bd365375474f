"""The [UP] rule switches of the oracle (oracle/UP_ASSUMPTIONS.md): every switch must change the result of
at least one crafted case -- otherwise tools/diff_vs_parasail.py could not tell the alternatives apart -- and
oracle.set_rules() must restore the documented defaults."""
import numpy as np
import pytest

import psb_data


def s(b):
    return np.frombuffer(b, dtype=np.uint8)


def run(oracle, mat, q, r, mode, o, e, **kw):
    x = oracle.align(s(q), s(r), mat, mode=mode, open=o, gap=e, trace=True, **kw)
    return (x["score"], x["end_query"], x["end_ref"], x["matches"], x["similar"], x["length"], x["cigar"], x["beg_query"], x["beg_ref"])


@pytest.fixture()
def dna(oracle):
    return oracle.Matrix.create(b"ACGT", 2, -3)


CASES = [
    # (rule, value, query, reference, mode, open, gap)
    ("sw_end_tie", 1, b"GACGTCG", b"CGGGATT", 2, 0, 0),             # the same maximum in several cells
    ("sw_end_tie", 2, b"ACGT", b"ACGTTTTTACGT", 2, 5, 2),
    ("sg_col_wins_tie", 1, b"GC", b"TGTGGGA", 1, 0, 0),
    ("sg_row_last_wins", 1, b"ACG", b"ACGTTACG", 1, 5, 2),
    ("h_priority", 1, b"TTACT", b"CCCAATA", 0, 1, 1),
    ("h_priority", 2, b"ACGT", b"AGGT", 0, 0, 0),
    ("h_priority", 3, b"TCCT", b"TTACTTT", 1, 0, 0),
    ("open_on_tie", 1, b"CATGCTCC", b"CT", 0, 2, 2),
    ("match_raw_bytes", 1, b"acgt", b"ACGT", 0, 5, 2),
    ("count_boundary_gaps", 1, b"TTACGT", b"ACGT", 0, 1, 1),
    ("cigar_edge_stop", 1, b"ACGT", b"TTACGT", 1, 5, 2),
    ("cigar_swap_id", 1, b"ACGTACGT", b"ACGTTACGT", 0, 1, 1),
    ("sg_flag_swap", 1, b"ACGT", b"TTACGTTT", 1, 5, 2),
    ("zero_beats_diag", 0, b"ATCAGAA", b"ACCAGTTG", 2, 2, 2),
]


@pytest.mark.parametrize("rule,value,q,r,mode,o,e", CASES)
def test_switch_is_observable(oracle, dna, rule, value, q, r, mode, o, e):
    kw = dict(s1_beg=True, s1_end=True, s2_beg=False, s2_end=False) if rule == "sg_flag_swap" else {}
    if rule == "zero_beats_diag":
        # a diagonal that sums to exactly zero inside a local alignment: +3 / -3 scores
        mat = type(dna).create(b"ACGT", 3, -3)
    else:
        mat = dna
    base = run(oracle, mat, q, r, mode, o, e, **kw)
    try:
        oracle.set_rules(**{rule: value})
        flipped = run(oracle, mat, q, r, mode, o, e, **kw)
    finally:
        oracle.set_rules()
    assert run(oracle, mat, q, r, mode, o, e, **kw) == base, "set_rules() did not restore the defaults"
    assert flipped != base, f"{rule}={value} is not observable on this case: {base}"


def test_band_rule(oracle, dna):
    q, r = psb_data.random_seq(1, 0, 40, False), psb_data.random_seq(1, 1, 70, False)
    a = oracle.align(q, r, dna, mode=0, open=5, gap=2, band=3)["score"]
    try:
        oracle.set_rules(band_rule=1)
        b = oracle.align(q, r, dna, mode=0, open=5, gap=2, band=3)["score"]
    finally:
        oracle.set_rules()
    assert a != b   # with the plain band the corner of a 40 x 70 table is unreachable


def test_unknown_rule_is_rejected(oracle):
    with pytest.raises(KeyError):
        oracle.set_rules(no_such_rule=1)
