"""The threaded C generator (tests/gen/psb_gen.c) must reproduce the numpy generator of tests/psb_data.py."""
import numpy as np

import psb_data
import psb_gen


def test_random_matches_numpy():
    for protein in (True, False):
        idx = np.arange(5000, dtype=np.uint64)
        u = psb_data.rnd(77, 3, idx)
        exp = psb_data.protein_letters(u) if protein else psb_data.dna_letters(u)
        assert np.array_equal(psb_gen.random(77, 3, 5000, protein), exp)


def test_substitute_and_gather_match_numpy():
    src = psb_gen.random(5, 3, 6000, False)
    u = psb_data.rnd(9, 4, np.arange(6000, dtype=np.uint64))
    hit = (u % np.uint64(10000)) < np.uint64(400)
    letters = np.frombuffer(psb_data.DNA, dtype=np.uint8)
    exp = np.where(hit, letters[((u >> np.uint64(20)) % np.uint64(4)).astype(np.int64)], src)
    assert np.array_equal(psb_gen.substitute(src, 9, 4, 0.04, False), exp)
    st = psb_gen.starts(3002, 6, 12, 340)
    assert np.array_equal(st, (psb_data.rnd(3002, 6, np.arange(12, dtype=np.uint64)) % np.uint64(340)).astype(np.int64))
    g = psb_gen.gather(src, 12, 500, st, 150)
    for p in range(12):
        assert np.array_equal(g[p * 150:(p + 1) * 150], src[p * 500 + st[p]: p * 500 + st[p] + 150])
