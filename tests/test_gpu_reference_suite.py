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


def test_banded_nw_bandwidth_2(ps):
    # the reference's own case: bandwidth(2), ACGT vs ACGT, score = len [REF 726-736]
    res = ps.Aligner.new().bandwidth(2).build().banded_nw(b"ACGT", b"ACGT")
    assert res.get_score() == 4


def test_ssw_alignment(ps):
    # [REF 738-757]
    result = ps.Aligner.new().build().ssw(b"ACGT", b"ACGT")
    assert result.score() == 4
    assert result.query_end() == 3 and result.ref_end() == 3
    assert result.query_start() == 0 and result.ref_start() == 0
    assert result.cigar_len() == 1 and int(result.cigar()[0]) == (4 << 4 | 7)


def test_ssw_init(ps):
    # [REF 759-766]
    p = ps.Profile.new_ssw(b"ACGT", ps.Matrix.default(), 2)
    assert not p.is_null()


def test_ssw_matches_sw_trace(ps, oracle, blosum62):
    # SSW = sw_trace + begin coordinates + CIGAR words, against the oracle's local alignment
    import psb_data
    b62 = ps.Matrix.from_name("blosum62")
    a = ps.Aligner.new().matrix(b62).gap_open(10).gap_extend(1).build()
    for i in range(6):
        q = psb_data.random_seq(301, 2 * i, 80 + 17 * i)
        r = np.concatenate([psb_data.random_seq(301, 2 * i + 1, 20), psb_data.mutate(q[10:70], 301, 50 + i, 0.15, 0.05),
                            psb_data.random_seq(302, i, 15)])
        exp = oracle.align(q, r, blosum62, mode=2, open=10, gap=1, trace=True)
        got = a.ssw(q, r)
        assert (got.score(), got.query_end(), got.ref_end()) == (exp["score"], exp["end_query"], exp["end_ref"])
        assert (got.query_start(), got.ref_start()) == (exp["beg_query"], exp["beg_ref"])
        assert np.array_equal(got.cigar(), exp["cigar_ops"])


def test_banded_nw_band_excludes_optimum(ps, oracle):
    # a 12-residue insertion: a band of 4 cannot follow the optimal path, so the banded score is lower than
    # the full one, and equal to the oracle's banded fill for every k
    import psb_data
    dna, odna = ps.Matrix.create(b"ACGT", 2, -3), oracle.Matrix.create(b"ACGT", 2, -3)
    q = psb_data.random_seq(303, 0, 120, protein=False)
    r = np.concatenate([q[:60], psb_data.random_seq(303, 1, 12, protein=False), q[60:]])[:120]
    full = oracle.align(q, r, odna, mode=0, open=5, gap=2)["score"]
    seen_lower = False
    for k in (1, 2, 4, 8, 16, 64, 200):
        exp = oracle.align(q, r, odna, mode=0, open=5, gap=2, band=k)
        got = ps.Aligner.new().matrix(dna).gap_open(5).gap_extend(2).bandwidth(k).build().banded_nw(q, r)
        assert got.get_score() == exp["score"], k
        assert (got.get_end_query(), got.get_end_ref()) == (len(q) - 1, len(r) - 1)
        seen_lower |= exp["score"] < full
    assert seen_lower and exp["score"] == full
    # unequal lengths: the band is widened by the length difference so the corner stays reachable
    r2 = np.concatenate([r, psb_data.random_seq(303, 2, 25, protein=False)])
    for k in (1, 3, 10):
        exp = oracle.align(q, r2, odna, mode=0, open=5, gap=2, band=k)
        got = ps.Aligner.new().matrix(dna).gap_open(5).gap_extend(2).bandwidth(k).build().banded_nw(q, r2)
        assert got.get_score() == exp["score"], k


def test_coarse_family_stats_multistrip(ps, oracle, blosum62):
    # `_stats` with a PSSM / large-valued matrix and a query of several strips used to size
    # the strip-boundary scratch for the narrow statistics word while the coarse kernels carry the wide one
    import psb_data
    big = oracle.Matrix.create(b"ACGT", 200, -150)       # S + open does not fit a byte: coarse family
    pbig = ps.Matrix.create(b"ACGT", 200, -150)
    qs = [psb_data.random_seq(304, i, L, protein=False) for i, L in enumerate((600, 900, 1500))]
    rs = [psb_data.mutate(q, 304, 10 + i, 0.1, 0.04, protein=False)[:700] for i, q in enumerate(qs)]
    for mode, name in ((0, "global_"), (1, "semi_global"), (2, "local")):
        got = getattr(ps.Aligner.new(), name)().matrix(pbig).gap_open(30).gap_extend(4).use_stats().build().align_batch(qs, rs)
        qc, qo = psb_data.concat(qs)
        rc, ro = psb_data.concat(rs)
        exp = oracle.align_batch(qc, qo, rc, ro, big, mode=mode, open=30, gap=4, stats=True)
        for k in ("score", "end_query", "end_ref", "matches", "similar", "length"):
            assert np.array_equal(getattr(got, k), exp[k]), (name, k)
    # the same through a PSSM (query 700 rows)
    q = psb_data.random_seq(305, 0, 700)
    pssm = ps.Matrix.from_name("blosum62").to_pssm(q)
    r = psb_data.mutate(q, 305, 1, 0.2, 0.05)[:650]
    for mode, name in ((0, "global_"), (2, "local")):
        a = getattr(ps.Aligner.new(), name)().matrix(pssm).gap_open(10).gap_extend(1).use_stats().build().align(q, r)
        exp = oracle.align(q, r, blosum62, mode=mode, open=10, gap=1)
        assert (a.get_score(), a.get_matches(), a.get_similar(), a.get_length()) == (exp["score"], exp["matches"], exp["similar"], exp["length"])
