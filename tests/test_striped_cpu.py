"""The CPU baseline (striped AVX2 restatement of parasail's sw_striped_profile_sat) must agree with
the scalar oracle before bench.py is allowed to time it."""
import numpy as np
import pytest

import psb_data


@pytest.mark.parametrize("gaps", [(10, 1), (5, 2), (3, 3), (0, 0), (12, 4), (1, 3)])
def test_striped_matches_oracle_protein(oracle, blosum62, gaps):
    query = psb_data.random_seq(2001, 0, 400)
    cat, off = psb_data.protein_db(2002, 2003, 1500, query=query, planted_frac=0.05)
    # add exact copies so that the 8 -> 16 -> 32 bit escalation is exercised
    seqs = [cat[off[i]:off[i + 1]] for i in range(len(off) - 1)] + [query.copy(), np.concatenate([query, query])]
    for rep in range(8):
        seqs.append(np.concatenate([query] * 9))  # score > 16-bit limit at open 0
    cat, off = psb_data.concat(seqs)
    got, secs = oracle.striped_sw_scan(query, cat, off, blosum62, gaps[0], gaps[1], threads=4)
    exp = oracle.align_batch(query, np.array([0, len(query)]), cat, off, blosum62, mode=oracle.SW, open=gaps[0], gap=gaps[1],
                             shared_query=True)
    for k in ("score", "end_query", "end_ref"):
        bad = np.nonzero(got[k] != exp[k])[0]
        assert len(bad) == 0, (k, bad[:5], got[k][bad[:5]], exp[k][bad[:5]])
    assert gaps[0] < gaps[1] or set(np.unique(got["width"])) >= {8, 16}
    assert secs > 0


def test_striped_matches_oracle_dna(oracle):
    mat = oracle.Matrix.create(b"ACGT", 2, -3)
    qs, rs = psb_data.dna_read_pairs(3001, 300)
    query = qs[0]
    cat, off = psb_data.concat(rs)
    got, _ = oracle.striped_sw_scan(query, cat, off, mat, 5, 2, threads=2)
    exp = oracle.align_batch(query, np.array([0, len(query)]), cat, off, mat, mode=oracle.SW, open=5, gap=2, shared_query=True)
    for k in ("score", "end_query", "end_ref"):
        assert np.array_equal(got[k], exp[k]), k
