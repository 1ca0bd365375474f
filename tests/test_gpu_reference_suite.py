"""GPU: the reference's own integration tests [REF tests/test_parasail.rs], restated 1:1 on the
Python mirror of the Rust API, which calls the same C symbols the crate binds."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ps():
    import __graft_entry__ as g
    g.build()
    import parasail_rs_b200 as ps
    return ps


@pytest.mark.parametrize("mode", ["global_", "semi_global", "local"])
def test_alignment_modes(ps, mode):
    # global_alignment / semi_global_alignment / local_alignment [REF 64-122]
    aligner = getattr(ps.Aligner.new(), mode)().striped().build()
    result = aligner.align(b"ACGT", b"ACGT")
    assert result.get_score() == 4 and result.get_end_query() == 3 and result.get_end_ref() == 3
    assert result.is_global() == (mode == "global_")
    assert result.is_local() == (mode == "local")
    assert result.is_semi_global() == (mode == "semi_global")
    assert result.is_striped() and not result.is_saturated()


@pytest.mark.parametrize("mode", ["global_", "semi_global", "local"])
def test_alignment_with_stats(ps, mode):
    # [REF 124-173]
    aligner = getattr(ps.Aligner.new(), mode)().use_stats().striped().build()
    result = aligner.align(b"ACGT", b"ACGT")
    assert result.get_matches() == 4 and result.get_length() == 4 and result.get_similar() == 4
    plain = getattr(ps.Aligner.new(), mode)().build().align(b"ACGT", b"ACGT")
    with pytest.raises(ps.NoStats):
        plain.get_matches()


@pytest.mark.parametrize("width", [8, 16, 32, 64])
def test_global_explicit_widths(ps, width):
    # global_8bit..64bit [REF 175-253]
    aligner = ps.Aligner.new().solution_width(width).build()
    result = aligner.align(b"ACTGACTGACTG", b"ACTGTCTGACTG")
    assert (result.get_score(), result.get_end_query(), result.get_end_ref()) == (11, 11, 11)


def test_score_table(ps):
    # [REF 255-325]
    q = r = b"ACGT"
    result = ps.Aligner.new().use_table().striped().build().align(q, r)
    assert result.is_table() and not result.is_stats() and not result.is_stats_table()
    t = result.get_score_table()
    assert t.shape == (4, 4) and t[-1, -1] == 4
    result = ps.Aligner.new().use_stats().use_table().striped().build().align(q, r)
    assert result.is_stats() and result.is_stats_table() and result.is_table()
    assert result.get_score_table().shape == (4, 4)
    matrix = ps.Matrix.create(b"ACGT", 3, -2)
    profile = ps.Profile.new(q, False, matrix)
    res = ps.Aligner.new().profile(profile).use_table().striped().build().align(None, r)
    assert res.is_table() and not res.is_stats() and not res.is_stats_table()
    assert res.get_score_table()[-1, -1] == 12
    profile = ps.Profile.new(q, True, matrix)
    res = ps.Aligner.new().profile(profile).use_stats().use_table().striped().build().align(None, r)
    assert res.is_stats() and res.is_stats_table() and res.is_table()
    assert res.get_score_table()[-1, -1] == 12


def test_stats_tables(ps):
    # matches_table / similar_table / length_table [REF 327-383]
    a = ps.Aligner.new().use_table().use_stats().striped().build()
    res = a.align(b"ACGT", b"ACGTT")
    assert res.is_table() and res.is_stats() and res.is_stats_table()
    t = res.get_matches_table()
    assert t.shape == (4, 5) and t[-1, -1] == 4
    assert a.align(b"ACGT", b"ACGT").get_similar_table().shape == (4, 4)
    assert a.align(b"ACGT", b"ACGTTT").get_length_table().shape == (4, 6)


def test_last_rows_and_cols(ps):
    # score_row .. length_col [REF 385-543]
    a = ps.Aligner.new().use_last_rowcol().use_stats().striped().build()
    res = a.align(b"ACGT", b"ACG")
    assert res.is_stats_rowcol() and res.is_stats() and not res.is_stats_table()
    assert list(res.get_score_row()) == [1, 2, 3]
    assert list(res.get_matches_row()) == [1, 2, 3]
    assert list(res.get_similar_row()) == [1, 2, 3]
    assert list(res.get_length_row()) == [4, 4, 4]
    res = a.align(b"ACG", b"ACGT")
    assert list(res.get_score_col()) == [1, 2, 3]
    assert list(res.get_matches_col()) == [1, 2, 3]
    assert list(res.get_similar_col()) == [1, 2, 3]
    assert list(res.get_length_col()) == [4, 4, 4]


def test_trace_table_strings_cigar(ps):
    # trace_table / get_traceback_strings / print_traceback / get_cigar [REF 545-616]
    a = ps.Aligner.new().use_trace().build()
    res = a.align(b"ACGT", b"ACGT")
    t = res.get_trace_table()
    assert t.shape == (4, 4) and t.size == 16
    assert np.all((t.astype(np.int32) & ~ps.TraceFlags.ALL) == 0)
    tb = res.get_traceback_strings(b"ACGT", b"ACGT")
    assert (tb.query, tb.comparison, tb.reference) == ("ACGT", "||||", "ACGT")
    res.print_traceback(b"ACGT", b"ACGT")
    assert res.get_cigar(b"ACGT", b"ACGT") == "4="
    with pytest.raises(ps.NoTrace):
        ps.Aligner.new().build().align(b"ACGT", b"ACGT").get_cigar(b"ACGT", b"ACGT")


@pytest.mark.parametrize("mode", ["global_", "semi_global", "local"])
def test_alignment_with_profile(ps, mode):
    # *_with_profile [REF 618-687]
    profile = ps.Profile.new(b"ACGT", True, ps.Matrix.default())
    aligner = getattr(ps.Aligner.new(), mode)().profile(profile).use_stats().striped().build()
    result = aligner.align(None, b"ACGT")
    assert result.is_striped() and result.is_stats()
    assert result.is_global() == (mode == "global_")
    assert result.get_score() == 4 and result.get_matches() == 4


def test_multithread_global_alignment(ps):
    # [REF 689-723]: an aligner holding a stats profile shared by two threads
    profile = ps.Profile.new(b"ACGT", True, ps.Matrix.default())
    aligner = ps.Aligner.new().profile(profile).use_stats().build()
    out = []

    def work():
        out.append(aligner.align(None, b"ACGT").get_score())
    th = [threading.Thread(target=work) for _ in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert out == [4, 4]


def test_banded_nw(ps):
    # [REF 725-736]
    res = ps.Aligner.new().bandwidth(4).build().banded_nw(b"ACGT", b"ACGT")
    assert res.get_score() == 4 and res.is_banded()
