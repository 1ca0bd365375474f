"""The packed 16-bit many-pairs kernel source (csrc/kern_pairs16.cuh) and its trace walk on the CPU SIMT
emulation vs the oracle: scores, end cells, CIGARs, begin coordinates and statistics, bit-exact, in
every mode, with ragged lengths inside a word (two different pairs share each register)."""
import numpy as np
import pytest

import emu_harness
import psb_data
from test_oracle_properties import SG_FLAGS


def make_pairs(seed, n, lq_rng, lr_rng, protein):
    rng = np.random.default_rng(seed)
    qs, rs = [], []
    for i in range(n):
        lq, lr = int(rng.integers(*lq_rng)), int(rng.integers(*lr_rng))
        q = psb_data.random_seq(seed, 2 * i, lq, protein)
        if i % 3 != 2:
            r = psb_data.mutate(q, seed, 2 * i + 1, 0.15, 0.08, protein)
            r = r[:lr] if len(r) >= lr else np.concatenate([r, psb_data.random_seq(seed + 7, i, lr - len(r), protein)])
        else:
            r = psb_data.random_seq(seed, 2 * i + 1, lr, protein)
        qs.append(q); rs.append(r)
    return qs, rs


def check(oracle, mat, qs, rs, G, K, mode, o, e, flags=(1, 1, 1, 1), what=0):
    got = emu_harness.pairs16(qs, rs, mat, G, K, mode, o, e, flags, what)
    assert got is not None
    for i, (q, r) in enumerate(zip(qs, rs)):
        exp = oracle.align(q, r, mat, mode=mode, open=o, gap=e, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2],
                           s2_end=flags[3], trace=(what == 1))
        tag = (i, len(q), len(r), mode, flags)
        assert (got["score"][i], got["end_query"][i], got["end_ref"][i]) == (exp["score"], exp["end_query"], exp["end_ref"]), tag
        if what == 1:
            assert np.array_equal(got["cigar_ops"][i], exp["cigar_ops"]), (tag, oracle.decode_cigar(got["cigar_ops"][i]), exp["cigar"])
            assert (got["beg_query"][i], got["beg_ref"][i]) == (exp["beg_query"], exp["beg_ref"]), tag
        if what == 2:
            assert (got["matches"][i], got["similar"][i], got["length"][i]) == (exp["matches"], exp["similar"], exp["length"]), tag


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("GK", [(16, 4), (16, 10), (32, 2)])
def test_scores_dna(oracle, mode, GK):
    G, K = GK
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    qs, rs = make_pairs(11 + K, 6, (1, G * K + 1), (1, 90), False)
    check(oracle, mat, qs, rs, G, K, mode, 5, 2)


@pytest.mark.parametrize("flags", SG_FLAGS)
def test_sg_flags(oracle, flags):
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    qs, rs = make_pairs(21, 6, (1, 49), (1, 60), False)
    check(oracle, mat, qs, rs, 16, 3, 1, 5, 2, flags)
    check(oracle, mat, qs, rs, 16, 3, 1, 5, 2, flags, what=2)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_protein_scores_and_cigar(oracle, blosum62, mode):
    qs, rs = make_pairs(31, 5, (100, 161), (60, 170), True)
    check(oracle, blosum62, qs, rs, 16, 10, mode, 10, 1)
    check(oracle, blosum62, qs, rs, 16, 10, mode, 10, 1, what=1)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("gaps", [(5, 2), (0, 0), (3, 3)])
def test_stats_by_walk(oracle, mode, gaps):
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    qs, rs = make_pairs(41, 6, (1, 129), (1, 140), False)
    check(oracle, mat, qs, rs, 16, 8, mode, *gaps, what=2)
    check(oracle, mat, qs, rs, 16, 8, mode, *gaps, what=1)


def test_wide_group_and_odd_rows(oracle, blosum62):
    # G = 32 (one pair of pairs per warp) and odd K: the last pair of the batch has no partner
    qs, rs = make_pairs(51, 3, (200, 289), (50, 120), True)
    check(oracle, blosum62, qs, rs, 32, 9, 2, 10, 1, what=1)
    check(oracle, blosum62, qs, rs, 32, 9, 0, 10, 1, what=2)
    qs, rs = make_pairs(52, 3, (250, 305), (40, 80), True)
    check(oracle, blosum62, qs, rs, 16, 19, 1, 10, 1)


def test_ties(oracle):
    # hand-made tie cases of SURVEY A.7: equal maxima in two columns / two rows, diag == F, F == E > diag
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    s = lambda b: np.frombuffer(b, dtype=np.uint8)
    qs = [s(b"ACTACGGG"), s(b"ACGTACGT"), s(b"AAAA"), s(b"ACGT"), s(b"GGAACCTT"), s(b"TTTT")]
    rs = [s(b"ACTTACG"), s(b"ACGTTTACGT"), s(b"AAAAAAAA"), s(b"TGCA"), s(b"GGTTAACC"), s(b"AAAA")]
    for mode in (0, 1, 2):
        for gaps in ((5, 2), (1, 1), (0, 0), (2, 0)):
            check(oracle, mat, qs, rs, 16, 1, mode, *gaps, what=1)
            check(oracle, mat, qs, rs, 16, 1, mode, *gaps, what=2)


def test_walk_row_to_lane_division_trick():
    # walk16_kernel turns a row index into (lane, row in lane) with a multiply: exact over the whole frame
    for K in range(1, 33):
        M = 65536 // K + 1
        assert all((il * M) >> 16 == il // K for il in range(1024)), K
