"""Frozen vectors (tests/golden/oracle_vectors.json, written by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))["cases"]


def oracle_matrix(oracle, name):
    import psb_data
    if name == "blosum62":
        return oracle.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())
    if name == "acgt_2_-3":
        return oracle.Matrix.create(b"ACGT", 2, -3)
    return oracle.Matrix.create(b"ACGTA", 1, -1)


def test_oracle_reproduces_golden(oracle):
    assert len(CASES) >= 200
    for c in CASES:
        f = c["flags"]
        r = oracle.align(c["query"].encode(), c["ref"].encode(), oracle_matrix(oracle, c["matrix"]), mode=c["mode"],
                         open=c["open"], gap=c["gap"], s1_beg=f[0], s1_end=f[1], s2_beg=f[2], s2_end=f[3], trace=True)
        for k in ("score", "end_query", "end_ref", "matches", "similar", "length", "cigar", "beg_query", "beg_ref"):
            assert r[k] == c[k], (k, c["query"], c["ref"], c["mode"])


@pytest.mark.gpu
def test_gpu_matches_golden():
    import __graft_entry__ as g
    g.build()
    import parasail_rs_b200 as ps
    mats = {"blosum62": ps.Matrix.from_name("blosum62"), "acgt_2_-3": ps.Matrix.create(b"ACGT", 2, -3),
            "default_acgta_1_-1": ps.Matrix.default()}
    groups = {}
    for c in CASES:
        groups.setdefault((c["matrix"], c["mode"], c["open"], c["gap"], tuple(c["flags"])), []).append(c)
    for (mname, mode, o, e, fl), cs in groups.items():
        for variant in ("stats", "trace"):
            b = getattr(ps.Aligner.new(), {0: "global_", 1: "semi_global", 2: "local"}[mode])().matrix(mats[mname]).gap_open(o).gap_extend(e)
            if mode == 1:
                b = b.allow_query_gaps([g_ for g_, f in (("prefix", fl[0]), ("suffix", fl[1])) if f]) \
                     .allow_ref_gaps([g_ for g_, f in (("prefix", fl[2]), ("suffix", fl[3])) if f])
            b = b.use_stats() if variant == "stats" else b.use_trace()
            res = b.build().align_batch([c["query"].encode() for c in cs], [c["ref"].encode() for c in cs])
            for i, c in enumerate(cs):
                assert (res.score[i], res.end_query[i], res.end_ref[i]) == (c["score"], c["end_query"], c["end_ref"]), (c, variant)
                if variant == "stats":
                    assert (res.matches[i], res.similar[i], res.length[i]) == (c["matches"], c["similar"], c["length"]), c
                else:
                    assert res.cigar(i) == c["cigar"] and (res.beg_query[i], res.beg_ref[i]) == (c["beg_query"], c["beg_ref"]), c
