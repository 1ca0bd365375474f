#!/usr/bin/env python
"""Freezes CPU-oracle outputs on small seeded corpora into tests/golden/oracle_vectors.json.

The reference's own tests pin only trivial vectors, and no parasail binary exists
in this environment, so these vectors are NOT parasail outputs: they are the oracle's results,
frozen after it passed its self-consistency properties (tests/test_oracle_properties.py), so that
(a) any later change to the oracle or to a tie-break rule is visible as a diff of this file, and
(b) the GPU path can be checked on a box where the oracle library failed to build.
Inputs are stored explicitly (not only as seeds) so the file is self-contained.

    python tests/golden/make_golden.py        # rewrites oracle_vectors.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import psb_data  # noqa: E402
from oracle import oracle as orc  # noqa: E402

SG_FLAGS = [(1, 1, 1, 1), (1, 0, 0, 0), (0, 1, 0, 0), (1, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (0, 0, 1, 1),
            (1, 0, 0, 1), (0, 1, 1, 0), (1, 0, 1, 0), (0, 1, 0, 1)]


def main():
    b62 = orc.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())
    dna = orc.Matrix.create(b"ACGT", 2, -3)
    ident = orc.Matrix.create(b"ACGTA", 1, -1)
    cases = []
    rng = np.random.default_rng(20261018)

    def add(matname, mat, q, r, mode, o, e, flags=(1, 1, 1, 1)):
        res = orc.align(q, r, mat, mode=mode, open=o, gap=e, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2],
                        s2_end=flags[3], trace=True)
        cases.append({"matrix": matname, "mode": mode, "open": o, "gap": e, "flags": list(flags),
                      "query": bytes(q).decode(), "ref": bytes(r).decode(),
                      "score": res["score"], "end_query": res["end_query"], "end_ref": res["end_ref"],
                      "matches": res["matches"], "similar": res["similar"], "length": res["length"],
                      "cigar": res["cigar"], "beg_query": res["beg_query"], "beg_ref": res["beg_ref"]})

    for i in range(40):   # proteins, BLOSUM62 10/1 (C1, C2, C4 style)
        lq, lr = int(rng.integers(5, 90)), int(rng.integers(5, 90))
        q = psb_data.random_seq(7001, 2 * i, lq)
        r = psb_data.mutate(q, 7001, 2 * i + 1, 0.2, 0.06)[:lr] if i % 2 == 0 else psb_data.random_seq(7001, 2 * i + 1, lr)
        if len(r) == 0:
            r = psb_data.random_seq(7002, i, lr)
        for mode in (0, 1, 2):
            add("blosum62", b62, q, r, mode, 10, 1)
    for i in range(22):   # DNA +2/-3 5/2 (C3, C5 style), every semi-global flag set
        lq, lr = int(rng.integers(8, 60)), int(rng.integers(20, 120))
        r = psb_data.random_seq(7003, i, lr, protein=False)
        s = int(rng.integers(0, max(1, lr - lq)))
        q = psb_data.mutate(r[s:s + lq], 7003, 100 + i, 0.05, 0.03, protein=False)
        add("acgt_2_-3", dna, q, r, 2, 5, 2)
        add("acgt_2_-3", dna, q, r, 0, 5, 2)
        add("acgt_2_-3", dna, q, r, 1, 5, 2, SG_FLAGS[i % len(SG_FLAGS)])
    for q, r in ((b"ACGT", b"ACGT"), (b"ACTGACTGACTG", b"ACTGTCTGACTG"), (b"ACGT", b"ACG"), (b"ACG", b"ACGT"),
                 (b"AC", b"ACTTAC"), (b"ACTAC", b"AC"), (b"TTAC", b"ACGG"), (b"AA", b"A")):
        for mode in (0, 1, 2):   # the reference's own inputs and the hand-derived tie cases, gaps 0/0
            add("default_acgta_1_-1", ident, np.frombuffer(q, dtype=np.uint8), np.frombuffer(r, dtype=np.uint8), mode, 0, 0)
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump({"note": "frozen oracle outputs (not parasail outputs); see make_golden.py", "cases": cases}, f, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
