"""GPU parity proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs -- bit-exact score, end_query, end_ref, matches/similar/length, CIGAR, beg_*."""
import numpy as np
import pytest

import psb_data
from test_oracle_properties import SG_FLAGS, cigar_rescore

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ps():
    import __graft_entry__ as g
    g.build()
    import parasail_rs_b200 as ps
    return ps


MODE_NAME = {0: "global_", 1: "semi_global", 2: "local"}
KEYS3 = ("score", "end_query", "end_ref")
KEYS6 = KEYS3 + ("matches", "similar", "length")


def builder(ps, mode, mat, o, e):
    return getattr(ps.Aligner.new(), MODE_NAME[mode])().matrix(mat).gap_open(o).gap_extend(e)


def mixed_pairs(seed, n, lq_rng, lr_rng, protein):
    rng = np.random.default_rng(seed)
    qs, rs = [], []
    for i in range(n):
        lq, lr = int(rng.integers(*lq_rng)), int(rng.integers(*lr_rng))
        q = psb_data.random_seq(seed, 2 * i, lq, protein)
        if i % 3 == 0:
            r = psb_data.mutate(q, seed, 2 * i + 1, 0.15, 0.06, protein)
            r = r[:lr] if len(r) >= lr else np.concatenate([r, psb_data.random_seq(seed + 7, i, lr - len(r), protein)])
        else:
            r = psb_data.random_seq(seed, 2 * i + 1, lr, protein)
        qs.append(q); rs.append(r)
    return qs, rs


def oracle_batch(oracle, qs, rs, omat, mode, o, e, flags=(1, 1, 1, 1), **kw):
    qc, qo = psb_data.concat(qs)
    rc, ro = psb_data.concat(rs)
    return oracle.align_batch(qc, qo, rc, ro, omat, mode=mode, open=o, gap=e, s1_beg=flags[0], s1_end=flags[1],
                              s2_beg=flags[2], s2_end=flags[3], **kw)


def assert_same(got, exp, keys, tag=""):
    for k in keys:
        g, x = getattr(got, k), exp[k]
        if not np.array_equal(g, x):
            bad = np.nonzero(g != x)[0]
            raise AssertionError(f"{tag} {k}: {len(bad)} mismatches, first at {bad[0]}: got {g[bad[0]]} expected {x[bad[0]]}")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pairs_protein_all_length_classes(ps, oracle, blosum62, mode):
    # query lengths 1..700 exercise every K class and the multi-strip path
    qs, rs = mixed_pairs(101, 160, (1, 700), (1, 500), True)
    got = builder(ps, mode, ps.Matrix.from_name("blosum62"), 10, 1).build().align_batch(qs, rs)
    assert_same(got, oracle_batch(oracle, qs, rs, blosum62, mode, 10, 1), KEYS3, f"mode {mode}")


@pytest.mark.parametrize("flags", SG_FLAGS)
def test_pairs_sg_flags(ps, oracle, flags):
    qs, rs = mixed_pairs(102, 60, (1, 200), (1, 200), False)
    b = ps.Aligner.new().semi_global().matrix(ps.Matrix.create(b"ACGT", 2, -3)).gap_open(5).gap_extend(2)
    qg = [g for g, f in (("prefix", flags[0]), ("suffix", flags[1])) if f]
    dg = [g for g, f in (("prefix", flags[2]), ("suffix", flags[3])) if f]
    got = b.allow_query_gaps(qg).allow_ref_gaps(dg).build().align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, oracle.Matrix.create(b"ACGT", 2, -3), 1, 5, 2, flags)
    assert_same(got, exp, KEYS3, str(flags))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("gaps", [(0, 0), (5, 2), (1, 4)])
def test_pairs_stats(ps, oracle, mode, gaps):
    qs, rs = mixed_pairs(103, 80, (1, 400), (1, 400), False)
    got = builder(ps, mode, ps.Matrix.create(b"ACGT", 2, -3), *gaps).use_stats().build().align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, oracle.Matrix.create(b"ACGT", 2, -3), mode, *gaps, stats=True)
    assert_same(got, exp, KEYS6, f"mode {mode} gaps {gaps}")


def test_pairs_stats_wide_counters(ps, oracle, blosum62):
    # lengths beyond the 10/10/12-bit packed counters switch to the 64-bit statistics word
    qs, rs = mixed_pairs(104, 6, (1100, 1300), (2900, 3100), True)
    got = builder(ps, 2, ps.Matrix.from_name("blosum62"), 10, 1).use_stats().build().align_batch(qs, rs)
    assert_same(got, oracle_batch(oracle, qs, rs, blosum62, 2, 10, 1, stats=True), KEYS6)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pairs_trace_cigar(ps, oracle, blosum62, mode):
    qs, rs = mixed_pairs(105, 90, (1, 420), (1, 300), True)
    got = builder(ps, mode, ps.Matrix.from_name("blosum62"), 10, 1).use_trace().build().align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, blosum62, mode, 10, 1, cigar=True)
    assert_same(got, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"), f"mode {mode}")


def test_single_pair_api_matches_batch(ps, oracle, blosum62):
    # the per-pair C entry points (what parasail-rs calls) against the oracle, incl. tables and trace bytes
    b62 = ps.Matrix.from_name("blosum62")
    qs, rs = mixed_pairs(106, 6, (1, 150), (1, 150), True)
    for q, r in zip(qs, rs):
        for mode in (0, 1, 2):
            exp = oracle.align(q, r, blosum62, mode=mode, open=10, gap=1, tables=True, rowcol=True, trace=True)
            a = builder(ps, mode, b62, 10, 1).use_stats().use_table().build().align(q, r)
            assert (a.get_score(), a.get_end_query(), a.get_end_ref()) == (exp["score"], exp["end_query"], exp["end_ref"])
            assert (a.get_matches(), a.get_similar(), a.get_length()) == (exp["matches"], exp["similar"], exp["length"])
            for nm in ("score_table", "matches_table", "similar_table", "length_table"):
                assert np.array_equal(getattr(a, "get_" + nm)(), exp[nm]), nm
            a = builder(ps, mode, b62, 10, 1).use_stats().use_last_rowcol().build().align(q, r)
            for nm in ("score_row", "matches_row", "similar_row", "length_row", "score_col", "matches_col", "similar_col", "length_col"):
                assert np.array_equal(getattr(a, "get_" + nm)(), exp[nm]), nm
            a = builder(ps, mode, b62, 10, 1).use_trace().build().align(q, r)
            assert np.array_equal(a.get_trace_table(), exp["trace"])
            assert a.get_cigar(q, r) == exp["cigar"] and a.cigar_beg == (exp["beg_query"], exp["beg_ref"])
            tb = a.get_traceback_strings(q, r)
            assert (tb.query, tb.comparison, tb.reference) == exp["traceback"]


def test_pssm_matrix(ps, oracle, blosum62):
    q = psb_data.random_seq(107, 0, 60)
    pssm = ps.Matrix.from_name("blosum62").to_pssm(q)
    rs = [psb_data.random_seq(107, i + 1, 40 + 3 * i) for i in range(8)]
    for mode in (0, 2):
        a = builder(ps, mode, pssm, 10, 1).build()
        for r in rs:
            exp = oracle.align(q, r, blosum62, mode=mode, open=10, gap=1)
            got = a.align(q, r)
            assert (got.get_score(), got.get_end_query(), got.get_end_ref()) == (exp["score"], exp["end_query"], exp["end_ref"])


# ---- database scan (config C2, reduced) ------------------------------------------------------------
def scan_case(ps, oracle, blosum62, query, cat, off, o=10, e=1, mode="local", stats=False):
    b62 = ps.Matrix.from_name("blosum62")
    db = ps.Database((cat, off), b62)
    prof = ps.Profile.new(query, stats, b62)
    a = getattr(ps.Aligner.new(), mode)().gap_open(o).gap_extend(e).profile(prof).build()
    got = a.scan(db)
    omode = {"local": 2, "global_": 0, "semi_global": 1}[mode]
    exp = oracle.align_batch(query, np.array([0, len(query)]), cat, off, blosum62, mode=omode, open=o, gap=e,
                             shared_query=True, stats=stats)
    assert_same(got, exp, KEYS6 if stats else KEYS3, f"scan {mode}")
    return got


def test_scan_c2_reduced(ps, oracle, blosum62):
    query = psb_data.random_seq(2001, 0, 400)
    cat, off = psb_data.protein_db(2002, 2003, 6000, query=query, planted_frac=0.02)
    got = scan_case(ps, oracle, blosum62, query, cat, off)
    assert got.score.max() > 200  # planted segments are found
    idx, sc = got.topk(10)
    order = np.lexsort((np.arange(got.n), -got.score))[:10]
    assert np.array_equal(idx, order) and np.array_equal(sc, got.score[order])


@pytest.mark.parametrize("lq", [17, 64, 129, 300, 512])
def test_scan_query_lengths(ps, oracle, blosum62, lq):
    query = psb_data.random_seq(2101, 0, lq)
    cat, off = psb_data.protein_db(2102, 2103, 700, query=query if lq >= 60 else None, planted_frac=0.05)
    scan_case(ps, oracle, blosum62, query, cat, off)


def test_scan_overflow_and_long_subjects(ps, oracle, blosum62):
    # identical copies of the query score > 2000 (still inside 16 bit); a 70 kb subject exceeds the
    # 16-bit column range and is re-run by the 32-bit kernel; a 700-aa query takes the multi-strip path
    query = psb_data.random_seq(2201, 0, 400)
    subs = [psb_data.random_seq(2202, i, 30 + 11 * i) for i in range(40)]
    subs[3] = query.copy()
    subs[17] = np.concatenate([psb_data.random_seq(2203, 0, 100), query, psb_data.random_seq(2203, 1, 50)])
    subs.append(psb_data.random_seq(2204, 0, 70000))
    cat, off = psb_data.concat(subs)
    got = scan_case(ps, oracle, blosum62, query, cat, off)
    assert got.n_retried >= 1 and got.score[3] > 2048
    long_q = psb_data.random_seq(2205, 0, 700)
    scan_case(ps, oracle, blosum62, long_q, *psb_data.concat(subs[:20]))


def test_scan_true_16bit_overflow(ps, oracle):
    # +100 per match over 400 columns leaves the 16-bit range: those subjects must come back exact
    # from the 32-bit re-run, their item partners too
    om = oracle.Matrix.create(b"ACGT", 100, -90)
    m = ps.Matrix.create(b"ACGT", 100, -90)
    q = psb_data.random_seq(2601, 0, 400, protein=False)
    subs = [psb_data.random_seq(2602, i, 100 + 7 * i, protein=False) for i in range(30)]
    subs[5] = q.copy()
    subs[11] = np.concatenate([subs[11][:50], q[20:390]])
    cat, off = psb_data.concat(subs)
    db = ps.Database((cat, off), m)
    a = ps.Aligner.new().local().gap_open(5).gap_extend(2).profile(ps.Profile.new(q, False, m)).build()
    got = a.scan(db)
    exp = oracle.align_batch(q, np.array([0, len(q)]), cat, off, om, mode=2, open=5, gap=2, shared_query=True)
    assert_same(got, exp, KEYS3, "overflow scan")
    assert got.n_retried >= 2 and got.score[5] == 40000


@pytest.mark.parametrize("mode", ["global_", "semi_global"])
def test_scan_other_modes_and_stats(ps, oracle, blosum62, mode):
    query = psb_data.random_seq(2301, 0, 120)
    cat, off = psb_data.protein_db(2302, 2303, 300, query=query, planted_frac=0.1)
    scan_case(ps, oracle, blosum62, query, cat, off, mode=mode)
    scan_case(ps, oracle, blosum62, query, cat, off, mode=mode, stats=True)
    scan_case(ps, oracle, blosum62, query, cat, off, mode="local", stats=True)


def test_scan_host_equals_resident_scan(ps, oracle, blosum62):
    # psb_scan_host (piecewise upload pipelined with the scan) returns exactly what a resident scan returns
    query = psb_data.random_seq(2501, 0, 400)
    cat, off = psb_data.protein_db(2502, 2503, 400000, query=query, planted_frac=0.01)   # ~145 MB: several pieces
    b62 = ps.Matrix.from_name("blosum62")
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(query, False, b62)).build()
    base = a.scan(ps.Database((cat, off), b62))
    host = a.scan_host((cat, off))
    for k in KEYS3:
        assert np.array_equal(getattr(host, k), getattr(base, k)), k
    assert host.cells == base.cells
    # the arrays above are pageable (uploads staged through the library's bounce buffers); the same from
    # page-locked memory (plain asynchronous copies, uploads queued ahead of the scan launches), twice so that
    # the second call runs with the piece plan from measured rates
    import torch
    pc = torch.empty(len(cat), dtype=torch.uint8, pin_memory=True); pc.numpy()[:] = cat
    po = torch.empty(len(off), dtype=torch.int64, pin_memory=True); po.numpy()[:] = off
    for _ in range(2):
        pinned = a.scan_host((pc.numpy(), po.numpy()))
        for k in KEYS3:
            assert np.array_equal(getattr(pinned, k), getattr(base, k)), ("pinned", k)
    # small database, other mode, stats: single piece through the 32-bit path, against the oracle
    cat2, off2 = psb_data.protein_db(2504, 2505, 300, query=query[:120], planted_frac=0.1)
    a2 = ps.Aligner.new().semi_global().gap_open(10).gap_extend(1).profile(ps.Profile.new(query[:120], True, b62)).build()
    got = a2.scan_host((cat2, off2))
    exp = oracle.align_batch(query[:120], np.array([0, 120]), cat2, off2, blosum62, mode=1, open=10, gap=1, shared_query=True, stats=True)
    assert_same(got, exp, KEYS6, "scan_host sg stats")


def test_scan_permutation_invariance_large(ps):
    # size-independent property at a size the scalar oracle would need minutes for: shuffling the
    # database permutes the results and nothing else
    query = psb_data.random_seq(2401, 0, 400)
    cat, off = psb_data.protein_db(2402, 2403, 100000, query=query, planted_frac=0.01)
    b62 = ps.Matrix.from_name("blosum62")
    prof = ps.Profile.new(query, False, b62)
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(prof).build()
    base = a.scan(ps.Database((cat, off), b62))
    perm = np.random.default_rng(5).permutation(len(off) - 1)
    lens = np.diff(off)
    off2 = np.zeros_like(off); off2[1:] = np.cumsum(lens[perm])
    cat2 = np.concatenate([cat[off[i]:off[i + 1]] for i in perm])
    shuf = a.scan(ps.Database((cat2, off2), b62))
    for k in KEYS3:
        assert np.array_equal(getattr(shuf, k), getattr(base, k)[perm]), k
    assert np.all(base.end_ref < lens) and np.all(base.end_query < 400) and np.all(base.score >= 0)


# ---- the other BASELINE configs, reduced ---------------------------------------------------------------
def test_config_c1_reduced(ps, oracle, blosum62):
    qs, rs = psb_data.protein_pairs(1001, 400, 300, related_frac=0.1)
    got = ps.Aligner.new().matrix(ps.Matrix.from_name("blosum62")).gap_open(10).gap_extend(1).build().align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, blosum62, 0, 10, 1)
    assert_same(got, exp, KEYS3)
    assert np.all(got.end_query == 299) and np.all(got.end_ref == 299)


def test_config_c3_reduced(ps, oracle):
    qs, rs = psb_data.dna_read_pairs(3001, 600)
    a = ps.Aligner.new().semi_global().matrix(ps.Matrix.create(b"ACGT", 2, -3)).gap_open(5).gap_extend(2).use_stats().build()
    assert a.fn_name == "sg_stats_striped_sat"
    got = a.align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, oracle.Matrix.create(b"ACGT", 2, -3), 1, 5, 2, stats=True)
    assert_same(got, exp, KEYS6)


def test_config_c4_reduced(ps, oracle, blosum62):
    qs, rs = psb_data.protein_pairs(4001, 300, 250, related_frac=0.8, p_sub=0.2, p_indel=0.02, geometric_mean=2.0)
    a = ps.Aligner.new().local().matrix(ps.Matrix.from_name("blosum62")).gap_open(10).gap_extend(1).use_trace().build()
    assert a.fn_name == "sw_trace_striped_sat"
    got = a.align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, blosum62, 2, 10, 1, cigar=True)
    assert_same(got, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"))
    # oracle-free: re-scoring each CIGAR reproduces the score
    for i in range(0, 300, 17):
        ops = got.cigar_ops[got.cigar_off[i]: got.cigar_off[i + 1]]
        if got.score[i] == 0:
            continue
        sc, ei, ej, *_ = cigar_rescore(ops, qs[i], rs[i], got.beg_query[i], got.beg_ref[i], blosum62, 10, 1)
        first = int(ops[0])
        if (first & 15) in (1, 2):
            sc += 10 + ((first >> 4) - 1)
        assert (sc, ei, ej) == (got.score[i], got.end_query[i], got.end_ref[i])


def test_config_c5_reduced(ps, oracle):
    # long DNA pair through the 32-bit path (multi-strip): 6 kb x 6 kb against the oracle
    r = psb_data.random_seq(5001, 0, 6000, protein=False)
    q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)[:6000]
    a = ps.Aligner.new().local().matrix(ps.Matrix.create(b"ACGT", 2, -3)).gap_open(5).gap_extend(2).solution_width(32).build()
    got = a.align(q, r)
    exp = oracle.align(q, r, oracle.Matrix.create(b"ACGT", 2, -3), mode=2, open=5, gap=2)
    assert (got.get_score(), got.get_end_query(), got.get_end_ref()) == (exp["score"], exp["end_query"], exp["end_ref"])
    assert got.get_score() > 3000


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_long_pair_wavefront(ps, oracle, mode):
    # queries >= 2048 rows go to the multi-warp wavefront kernel: strips run concurrently on
    # different SMs and hand their bottom rows over through global memory
    m = oracle.Matrix.create(b"ACGT", 2, -3)
    r = psb_data.random_seq(5101, 0, 9000, protein=False)
    q = psb_data.mutate(r, 5101, 1, 0.10, 0.01, protein=False)[:8000]
    got = builder(ps, mode, ps.Matrix.create(b"ACGT", 2, -3), 5, 2).solution_width(32).build().align(q, r)
    exp = oracle.align(q, r, m, mode=mode, open=5, gap=2)
    assert (got.get_score(), got.get_end_query(), got.get_end_ref()) == (exp["score"], exp["end_query"], exp["end_ref"])
    # mixed batch: long pairs next to short ones
    qs = [q, q[:100], q[:3000]]
    rs = [r[:500], r[:300], r[:2500]]
    gb = builder(ps, mode, ps.Matrix.create(b"ACGT", 2, -3), 5, 2).build().align_batch(qs, rs)
    assert_same(gb, oracle_batch(oracle, qs, rs, m, mode, 5, 2), KEYS3, f"wave batch mode {mode}")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_long_pairs_with_trace_and_stats_on_the_wavefront(ps, oracle, mode, monkeypatch):
    # batches with traceback or statistics whose queries have >= 2048 rows: the traced launch of the column-blocked
    # wavefront kernel (H low bytes + gap open/extend bits, 1.25 B per cell) + walk32_kernel, strips running
    # concurrently on different SMs.  Compared with the oracle and with the one-warp-per-pair kernels
    # (PSB_NO_WAVE_TRACE), in a batch that also holds short pairs.
    m = oracle.Matrix.create(b"ACGT", 2, -3)
    r = psb_data.random_seq(5601, 0, 5200, protein=False)
    q = psb_data.mutate(r, 5601, 1, 0.10, 0.02, protein=False)[:5000]
    # a query that lacks two stretches of the reference and carries an insert: long gaps across lanes and strips
    q2 = np.concatenate([q[:900], q[1300:2500], psb_data.random_seq(5602, 0, 333, protein=False), q[2500:4100]])
    qs = [q, q[:100], q2, q[:2048], q[1000:3100]]
    rs = [r, r[:300], r[:4300], r[:67], psb_data.random_seq(5603, 0, 2500, protein=False)]
    exp = oracle_batch(oracle, qs, rs, m, mode, 5, 2, cigar=True, stats=True)
    dm = ps.Matrix.create(b"ACGT", 2, -3)
    got = builder(ps, mode, dm, 5, 2).use_trace().build().align_batch(qs, rs)
    assert_same(got, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"), f"wave trace mode {mode}")
    gs = builder(ps, mode, dm, 5, 2).use_stats().build().align_batch(qs, rs)
    assert_same(gs, exp, KEYS6, f"wave stats mode {mode}")
    # single-pair API with statistics takes the same path
    a = builder(ps, mode, dm, 5, 2).use_stats().build().align(q2, rs[2])
    assert (a.get_score(), a.get_matches(), a.get_similar(), a.get_length()) == (exp["score"][2], exp["matches"][2], exp["similar"][2], exp["length"][2])
    monkeypatch.setenv("PSB_NO_WAVE_TRACE", "1")
    old = builder(ps, mode, dm, 5, 2).use_trace().build().align_batch(qs, rs)
    assert_same(old, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"), f"one-warp trace mode {mode}")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_single_long_pair_trace_api(ps, oracle, mode):
    # Aligner::align of ONE long pair with use_trace(): score, CIGAR and traceback strings come from the traced wavefront
    # launch; the TraceFlags table is fetched only when asked for (the pair once more on the flag-byte kernel) and
    # equals the oracle's
    m = oracle.Matrix.create(b"ACGT", 2, -3)
    r = psb_data.random_seq(5701, 0, 2400, protein=False)
    q = psb_data.mutate(r, 5701, 1, 0.10, 0.02, protein=False)[:2200]
    exp = oracle.align(q, r, m, mode=mode, open=5, gap=2, trace=True)
    a = builder(ps, mode, ps.Matrix.create(b"ACGT", 2, -3), 5, 2).use_trace().build().align(q, r)
    assert (a.get_score(), a.get_end_query(), a.get_end_ref()) == (exp["score"], exp["end_query"], exp["end_ref"])
    assert a.get_cigar(q, r) == exp["cigar"] and a.cigar_beg == (exp["beg_query"], exp["beg_ref"])
    tb = a.get_traceback_strings(q, r)
    assert (tb.query, tb.comparison, tb.reference) == exp["traceback"]
    assert np.array_equal(a.get_trace_table(), exp["trace"])
    assert np.array_equal(a.get_trace_table(), exp["trace"])   # the second call hands out the same table


def test_long_pair_trace_ties_and_protein(ps, oracle, blosum62):
    # equal-score paths (repeats, open == extend, 0/0 penalties) and protein scores on the traced wavefront path
    b62 = ps.Matrix.from_name("blosum62")
    q = psb_data.random_seq(5604, 0, 2300)
    rep = np.concatenate([q[100:400]] * 8)
    pr = psb_data.mutate(np.concatenate([q[:700], q[760:1500], q[1500:]]), 5605, 0, 0.15, 0.02)
    for o, e in ((10, 1), (11, 11)):
        for mode in (0, 1, 2):
            exp = oracle_batch(oracle, [q, q], [rep, pr], blosum62, mode, o, e, cigar=True, stats=True)
            got = builder(ps, mode, b62, o, e).use_trace().build().align_batch([q, q], [rep, pr])
            assert_same(got, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"), f"protein {mode} {o}/{e}")
            gs = builder(ps, mode, b62, o, e).use_stats().build().align_batch([q, q], [rep, pr])
            assert_same(gs, exp, KEYS6, f"protein stats {mode} {o}/{e}")
    dm, om = ps.Matrix.create(b"ACGT", 2, -3), oracle.Matrix.create(b"ACGT", 2, -3)
    unit = np.frombuffer(b"ACGTTGCAAC", dtype=np.uint8)
    dq, dr = np.concatenate([unit] * 230), np.concatenate([unit[:7]] * 150)
    for o, e in ((0, 0), (2, 0), (1, 1)):
        for mode in (0, 1, 2):
            exp = oracle_batch(oracle, [dq], [dr], om, mode, o, e, cigar=True, stats=True)
            got = builder(ps, mode, dm, o, e).use_trace().build().align_batch([dq], [dr])
            assert_same(got, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"), f"ties {mode} {o}/{e}")


@pytest.mark.parametrize("name", ["sw", "nw", "sg"])
def test_long_pair_50kb_against_cached_oracle(ps, name):
    # 50 kb x 50 kb (2.5e9 cells, the scalar oracle needs about a minute per mode): compared with
    # tests/golden/long_pair_50k.json, the oracle's results for exactly these seeded inputs
    # (tests/bench_configs.py C5 recipe).  Scores above 32767 -> true 32-bit lanes all the way.
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "long_pair_50k.json")))
    case = {c["name"]: c for c in g["cases"]}[name]
    L = g["L"]
    r = psb_data.random_seq(5001, 0, L, protein=False)
    q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)
    q = q[:L] if len(q) >= L else np.concatenate([q, psb_data.random_seq(5002, 0, L - len(q), protein=False)])
    got = builder(ps, case["mode"], ps.Matrix.create(b"ACGT", 2, -3), 5, 2).solution_width(32).build().align(q, r)
    assert (got.get_score(), got.get_end_query(), got.get_end_ref()) == (case["score"], case["end_query"], case["end_ref"])


def test_long_local_pair_column_blocked(ps, oracle, blosum62, monkeypatch):
    # local long pairs take the column-blocked generation of the wavefront kernel (4 columns per step,
    # strips hand over through self-validating 64-bit words): reference lengths around the block size,
    # protein scores, repeats (many equal maxima), open == extend, and a pair without any match
    b62 = ps.Matrix.from_name("blosum62")
    q = psb_data.random_seq(5201, 0, 2600)
    rep = np.concatenate([q[100:400]] * 9)
    cases = [(q, rep), (q, rep[:2049]), (q, rep[:2050]), (q, rep[:2051]), (q, psb_data.mutate(q, 5202, 0, 0.2, 0.03)[:1999]),
             (q, q[:5]), (q, psb_data.random_seq(5203, 1, 700))]
    for o, e in ((10, 1), (4, 4)):
        a = ps.Aligner.new().local().matrix(b62).gap_open(o).gap_extend(e).build()
        gb = a.align_batch([c[0] for c in cases], [c[1] for c in cases])
        assert_same(gb, oracle_batch(oracle, [c[0] for c in cases], [c[1] for c in cases], blosum62, 2, o, e), KEYS3,
                    f"column-blocked wavefront open {o} ext {e}")
    dna = ps.Matrix.create(b"ACGT", 2, -3)
    a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
    zero = a.align(np.frombuffer(b"A" * 4000, dtype=np.uint8), np.frombuffer(b"C" * 3001, dtype=np.uint8))
    assert (zero.get_score(), zero.get_end_query(), zero.get_end_ref()) == (0, 0, 0)
    # the two generations agree on a bigger pair (24 strips)
    r = psb_data.random_seq(5204, 0, 30001, protein=False)
    qq = psb_data.mutate(r, 5204, 1, 0.08, 0.01, protein=False)[:6100]
    g3 = a.align(qq, r)
    monkeypatch.setenv("PSB_WAVE_GEN", "2")
    g2 = a.align(qq, r)
    monkeypatch.delenv("PSB_WAVE_GEN")
    assert (g3.get_score(), g3.get_end_query(), g3.get_end_ref()) == (g2.get_score(), g2.get_end_query(), g2.get_end_ref())
    exp = oracle.align(qq, r, oracle.Matrix.create(b"ACGT", 2, -3), mode=2, open=5, gap=2)
    assert (g3.get_score(), g3.get_end_query(), g3.get_end_ref()) == (exp["score"], exp["end_query"], exp["end_ref"])


def test_concurrent_host_threads(ps, oracle, blosum62):
    # Aligner / Profile / Matrix handles are Send + Sync in the reference [REF src/aligner/mod.rs:532-535]:
    # four host threads (each gets its own stream) share one profile, one database and one aligner
    import threading
    b62 = ps.Matrix.from_name("blosum62")
    query = psb_data.random_seq(2701, 0, 200)
    cat, off = psb_data.protein_db(2702, 2703, 3000, query=query, planted_frac=0.05)
    scan_aligner = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(query, False, b62)).build()
    db = ps.Database((cat, off), b62)
    pair_aligner = ps.Aligner.new().semi_global().matrix(b62).gap_open(10).gap_extend(1).use_stats().build()
    qs, rs = mixed_pairs(2704, 60, (1, 300), (1, 300), True)
    exp_scan = oracle.align_batch(query, np.array([0, len(query)]), cat, off, blosum62, mode=2, open=10, gap=1, shared_query=True)
    exp_pairs = oracle_batch(oracle, qs, rs, blosum62, 1, 10, 1, stats=True)
    errors = []

    def work(tid):
        try:
            for _ in range(3):
                if tid % 2 == 0:
                    assert_same(scan_aligner.scan(db), exp_scan, KEYS3, f"thread {tid} scan")
                else:
                    assert_same(pair_aligner.align_batch(qs, rs), exp_pairs, KEYS6, f"thread {tid} pairs")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors


def test_edge_cases(ps, oracle, blosum62):
    b62 = ps.Matrix.from_name("blosum62")
    # single residues, unknown letters (mapped to '*'), lower case
    qs = [b"A", b"W", b"acgt", b"XXXX", b"ARNDJOU", b"M" * 33]
    rs = [b"A", b"A", b"ACGT", b"ARND", b"ARND??", b"M"]
    for mode in (0, 1, 2):
        got = builder(ps, mode, b62, 10, 1).use_stats().build().align_batch(qs, rs)
        exp = oracle_batch(oracle, [np.frombuffer(x, dtype=np.uint8) for x in qs], [np.frombuffer(x, dtype=np.uint8) for x in rs],
                           blosum62, mode, 10, 1, stats=True)
        assert_same(got, exp, KEYS6, f"edge mode {mode}")
    with pytest.raises(ps.DeviceError):
        builder(ps, 0, b62, 10, 1).build().align_batch([b"ACGT", b""], [b"ACGT", b"ACGT"])
    # explicit narrow width that cannot hold the result is reported saturated, not wrong
    q = psb_data.random_seq(9, 0, 400)
    res = builder(ps, 2, b62, 10, 1).solution_width(8).build().align(q, q)
    assert res.is_saturated()
    res = builder(ps, 2, b62, 10, 1).build().align(q, q)
    assert not res.is_saturated() and res.get_score() > 2000


# ---- packed 16-bit many-pairs kernels (csrc/kern_pairs16.cuh) -----------------------------------------
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("wide", ["0", "1"])
def test_pairs16_all_classes_vs_oracle(ps, oracle, blosum62, mode, wide, monkeypatch):
    # query lengths 1..512 walk through every (G, K) class; ragged lengths share words
    monkeypatch.setenv("PSB_P16_WIDE", wide)
    qs, rs = mixed_pairs(201, 240, (1, 513), (1, 400), True)
    got = builder(ps, mode, ps.Matrix.from_name("blosum62"), 10, 1).build().align_batch(qs, rs)
    assert_same(got, oracle_batch(oracle, qs, rs, blosum62, mode, 10, 1), KEYS3, f"mode {mode} wide {wide}")
    got = builder(ps, mode, ps.Matrix.from_name("blosum62"), 10, 1).use_trace().build().align_batch(qs, rs)
    exp = oracle_batch(oracle, qs, rs, blosum62, mode, 10, 1, cigar=True)
    assert_same(got, exp, KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops"), f"trace mode {mode} wide {wide}")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pairs16_stats_by_walk_vs_oracle(ps, oracle, mode):
    qs, rs = mixed_pairs(202, 300, (1, 400), (1, 600), False)
    for gaps in ((5, 2), (0, 0), (4, 4)):
        got = builder(ps, mode, ps.Matrix.create(b"ACGT", 2, -3), *gaps).use_stats().build().align_batch(qs, rs)
        exp = oracle_batch(oracle, qs, rs, oracle.Matrix.create(b"ACGT", 2, -3), mode, *gaps, stats=True)
        assert_same(got, exp, KEYS6, f"mode {mode} gaps {gaps}")


def test_pairs16_equals_32bit_path(ps, monkeypatch):
    # the same batch through the packed kernels and (PSB_NO_P16=1) through the 32-bit kernels
    qs, rs = mixed_pairs(203, 500, (1, 513), (1, 300), True)
    b62 = ps.Matrix.from_name("blosum62")
    for mode in (0, 1, 2):
        for kind in ("score", "stats", "trace"):
            bl = builder(ps, mode, b62, 11, 1)
            bl = bl.use_stats() if kind == "stats" else (bl.use_trace() if kind == "trace" else bl)
            a = bl.build()
            monkeypatch.delenv("PSB_NO_P16", raising=False)
            g16 = a.align_batch(qs, rs)
            monkeypatch.setenv("PSB_NO_P16", "1")
            g32 = a.align_batch(qs, rs)
            monkeypatch.delenv("PSB_NO_P16", raising=False)
            keys = KEYS6 if kind == "stats" else (KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops") if kind == "trace" else KEYS3)
            for k in keys:
                assert np.array_equal(getattr(g16, k), getattr(g32, k)), (mode, kind, k)


def test_pairs16_sixteen_bit_bound_routes_long_pairs_away(ps, oracle):
    # +100 matches: 300 matching residues exceed the static 16-bit bound, so these pairs must take the 32-bit path
    mat, omat = ps.Matrix.create(b"ACGT", 100, -90), oracle.Matrix.create(b"ACGT", 100, -90)
    qs = [psb_data.random_seq(204, i, 380, protein=False) for i in range(6)]
    rs = [q.copy() if i % 2 == 0 else psb_data.mutate(q, 204, 10 + i, 0.1, 0.02, protein=False) for i, q in enumerate(qs)]
    for mode in (0, 2):
        got = builder(ps, mode, mat, 20, 2).use_stats().build().align_batch(qs, rs)
        exp = oracle_batch(oracle, qs, rs, omat, mode, 20, 2, stats=True)
        assert_same(got, exp, KEYS6, f"mode {mode}")
        assert int(got.score[0]) == 38000


# ---- loader, on-disk format, packing widths, whole-box scan -----------------------------------------------
def test_dna_database_packs_at_2_and_3_bit(ps, oracle):
    # ACGT only -> 2 bit although the matrix has a wildcard column; one N anywhere -> 3 bit; both scan bit-exact
    dna, odna = ps.Matrix.create(b"ACGT", 2, -3), oracle.Matrix.create(b"ACGT", 2, -3)
    query = psb_data.random_seq(401, 0, 150, protein=False)
    subs = [psb_data.random_seq(401, 1 + i, 40 + 13 * i, protein=False) for i in range(40)]
    subs[3] = np.concatenate([subs[3][:20], psb_data.mutate(query[20:120], 401, 99, 0.05, 0.02, protein=False)])
    for with_n, bits in ((False, 2), (True, 3)):
        ss = [s.copy() for s in subs]
        if with_n:
            ss[7][5] = ord("N"); ss[3][30] = ord("n")
        cat, off = psb_data.concat(ss)
        db = ps.Database((cat, off), dna)
        assert db.bits == bits
        prof = ps.Profile.new(query, False, dna)
        for mode, om in (("local", 2), ("semi_global", 1), ("global_", 0)):
            a = getattr(ps.Aligner.new(), mode)().gap_open(5).gap_extend(2).profile(prof).build()
            got = a.scan(db)
            exp = oracle.align_batch(query, np.array([0, len(query)]), cat, off, odna, mode=om, open=5, gap=2, shared_query=True)
            assert_same(got, exp, KEYS3, f"{bits}-bit {mode}")


def test_fasta_and_packed_file_round_trip(ps, oracle, blosum62, tmp_path):
    b62 = ps.Matrix.from_name("blosum62")
    query = psb_data.random_seq(402, 0, 200)
    cat, off = psb_data.protein_db(403, 404, 300, query=query, planted_frac=0.05)
    fa = tmp_path / "db.fa"
    with open(fa, "wb") as f:
        for i in range(len(off) - 1):
            seq = bytes(cat[off[i]:off[i + 1]])
            f.write(b">seq%d some description\n" % i)
            for a in range(0, len(seq), 60):
                f.write(seq[a:a + 60] + b"\n")
    db1 = ps.Database.from_fasta(fa, b62)
    assert db1.n == len(off) - 1 and db1.residues == int(off[-1]) and db1.bits == 5
    packed = tmp_path / "db.psbdb"
    db1.save(packed)
    db2 = ps.Database.load(packed, b62)
    assert (db2.n, db2.residues, db2.bits) == (db1.n, db1.residues, db1.bits)
    prof = ps.Profile.new(query, False, b62)
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(prof).build()
    exp = oracle.align_batch(query, np.array([0, len(query)]), cat, off, blosum62, mode=2, open=10, gap=1, shared_query=True)
    assert_same(a.scan(db1), exp, KEYS3, "from_fasta")
    assert_same(a.scan(db2), exp, KEYS3, "load")
    with pytest.raises(ps.DeviceError):
        ps.Database.load(packed, ps.Matrix.create(b"ACGT", 2, -3))     # packed with another alphabet
    with pytest.raises(ps.DeviceError):
        ps.Database.load(fa, b62)                                      # not a packed file


def test_scan_box_equals_resident_scan(ps, oracle, blosum62):
    # the single-process whole-box entry point (here: however many GPUs the box has) returns caller-order results
    b62 = ps.Matrix.from_name("blosum62")
    query = psb_data.random_seq(405, 0, 400)
    cat, off = psb_data.protein_db(406, 407, 3000, query=query, planted_frac=0.02)
    prof = ps.Profile.new(query, False, b62)
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(prof).build()
    ref = a.scan(ps.Database((cat, off), b62))
    for ng in (0, 1):
        got = a.scan_box((cat, off), ng)
        for k in KEYS3:
            assert np.array_equal(getattr(got, k), getattr(ref, k)), (ng, k)
    exp = oracle.align_batch(query, np.array([0, len(query)]), cat, off[:301], blosum62, mode=2, open=10, gap=1, shared_query=True)
    for k in KEYS3:
        assert np.array_equal(getattr(ref, k)[:300], exp[k]), k


@pytest.mark.parametrize("lq", [401, 512, 700, 1000, 1601])
def test_scan_long_query_strip_wise(ps, oracle, blosum62, lq):
    # queries beyond one strip (400 rows) are swept strip by strip in packed 16-bit lanes
    query = psb_data.random_seq(501, lq, lq)
    cat, off = psb_data.protein_db(502, 503 + lq, 400, query=query[: min(lq, 400)], planted_frac=0.05)
    # plant segments from every part of the query, some straddling strip boundaries
    segs = [query[a: a + 120] for a in range(0, lq - 60, max(60, lq // 7))]
    extra = [np.concatenate([psb_data.random_seq(504, i, 30), psb_data.mutate(s, 504, 50 + i, 0.1, 0.03), psb_data.random_seq(505, i, 20)])
             for i, s in enumerate(segs)]
    ecat, eoff = psb_data.concat(extra)
    cat = np.concatenate([cat, ecat]); off = np.concatenate([off, eoff[1:] + off[-1]])
    scan_case(ps, oracle, blosum62, query, cat, off)


def test_pairs16_edge_shapes(ps, oracle):
    # 1 x 1, 1 x n, n x 1, zero gap penalties with a reference beyond 65 000 columns (takes the 32-bit path), lower case, N
    dna, odna = ps.Matrix.create(b"ACGT", 2, -3), oracle.Matrix.create(b"ACGT", 2, -3)
    s = lambda b: np.frombuffer(b, dtype=np.uint8)
    qs = [s(b"A"), s(b"A"), s(b"ACGTACGTAC"), s(b"acgtnACGT"), psb_data.random_seq(601, 0, 40, protein=False)]
    rs = [s(b"A"), s(b"TTTTACGT"), s(b"G"), s(b"ACGTNNACGT"), psb_data.random_seq(601, 1, 70000, protein=False)]
    for mode in (0, 1, 2):
        for gaps in ((5, 2), (0, 0)):
            for kind in ("score", "stats", "trace"):
                bl = builder(ps, mode, dna, *gaps)
                bl = bl.use_stats() if kind == "stats" else (bl.use_trace() if kind == "trace" else bl)
                got = bl.build().align_batch(qs, rs)
                exp = oracle_batch(oracle, qs, rs, odna, mode, *gaps, stats=(kind == "stats"), cigar=(kind == "trace"))
                keys = KEYS6 if kind == "stats" else (KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops") if kind == "trace" else KEYS3)
                assert_same(got, exp, keys, f"mode {mode} gaps {gaps} {kind}")


@pytest.mark.parametrize("kind", ["score", "stats", "trace"])
@pytest.mark.parametrize("mode", [0, 2])
def test_pairs_passes_on_two_lanes_equal_single_pass(ps, oracle, blosum62, mode, kind, monkeypatch):
    # a large batch is cut into passes that two host threads work through side by side (run_pairs_lanes); forced
    # here with 1 MB passes on a modest batch: every array, CIGAR CSR included, must equal the single-pass result
    # and the oracle's
    qs, rs = mixed_pairs(141, 3000, (1, 420), (1, 500), True)
    if kind != "score":   # (score-only batches holding a pair for the whole-GPU wavefront kernel run their passes serially)
        qs += [psb_data.random_seq(142, 0, 2500), psb_data.random_seq(142, 1, 40)]     # a long pair next to a tiny one
        rs += [psb_data.random_seq(142, 2, 2400), psb_data.random_seq(142, 3, 3000)]
    bld = builder(ps, mode, ps.Matrix.from_name("blosum62"), 10, 1)
    keys = KEYS3
    if kind == "stats":
        bld, keys = bld.use_stats(), KEYS6
    if kind == "trace":
        bld, keys = bld.use_trace(), KEYS3 + ("beg_query", "beg_ref", "cigar_off", "cigar_ops")
    a = bld.build()
    one = a.align_batch(qs, rs)
    monkeypatch.setenv("PSB_PAIRS_PASS_MB", "1")
    two = a.align_batch(qs, rs)
    monkeypatch.setenv("PSB_PAIRS_LANES", "1")
    serial = a.align_batch(qs, rs)
    for other, tag in ((two, "two lanes"), (serial, "serial passes")):
        for k in keys:
            assert np.array_equal(getattr(one, k), getattr(other, k)), (tag, k)
        assert other.cells == one.cells
    exp = oracle_batch(oracle, qs, rs, blosum62, mode, 10, 1, stats=(kind == "stats"), cigar=(kind == "trace"))
    assert_same(two, exp, keys, f"{kind} mode {mode}")


def test_pairs_from_pinned_and_pageable_memory(ps):
    # psb_align_pairs takes any host memory: pageable arrays are staged through bounce buffers, page-locked ones
    # are copied asynchronously; batches above 64 MB of residues run as passes on two lanes.  Same results.
    import torch
    n, lq, lr = 150000, 200, 260          # 69 MB of residues
    rng = np.random.default_rng(77)
    qc = rng.integers(0, 20, n * lq, dtype=np.uint8); rc = rng.integers(0, 20, n * lr, dtype=np.uint8)
    alpha = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
    qc, rc = alpha[qc], alpha[rc]
    rc.reshape(n, lr)[::3, :lq] = qc.reshape(n, lq)[::3]      # a third of the pairs related
    qo = np.arange(n + 1, dtype=np.int64) * lq; ro = np.arange(n + 1, dtype=np.int64) * lr
    a = ps.Aligner.new().local().matrix(ps.Matrix.from_name("blosum62")).gap_open(10).gap_extend(1).use_stats().build()
    pageable = a.align_batch((qc, qo), (rc, ro))
    keep = []
    for x in (qc, qo, rc, ro):
        t = torch.empty(x.shape, dtype=torch.from_numpy(x[:0].copy()).dtype, pin_memory=True); t.numpy()[...] = x; keep.append(t)
    pinned = a.align_batch((keep[0].numpy(), keep[1].numpy()), (keep[2].numpy(), keep[3].numpy()))
    for k in KEYS6:
        assert np.array_equal(getattr(pageable, k), getattr(pinned, k)), k
    assert pageable.score[0] > 900 and pageable.matches[0] == lq      # a related pair: the query is found whole
    # psb_trim gives the lanes' kept buffers, the recycled blocks and the idle part of the pool back; the next call
    # simply takes them again
    ps.trim()
    again = a.align_batch((keep[0].numpy(), keep[1].numpy()), (keep[2].numpy(), keep[3].numpy()))
    for k in KEYS6:
        assert np.array_equal(getattr(again, k), getattr(pinned, k)), k
