"""Randomised (hypothesis) runs of the shipped kernel sources on the CPU SIMT emulation against the
oracle: the packed 16-bit scan kernel and the column-blocked long-pair kernel.  Sizes are small (every
lane is an OS thread), the value is in the odd shapes: query lengths across the K classes, subjects
shorter than a lane group, uneven word partners, gap penalties with open < extend, zero penalties."""
import numpy as np
from hypothesis import given, settings, strategies as st

import emu_harness
import psb_data


def seq(seed, idx, n, protein):
    return psb_data.random_seq(seed, idx, n, protein=protein)


@settings(max_examples=150, deadline=None)
@given(seed=st.integers(1, 10**6), lq=st.integers(1, 420), nsub=st.integers(1, 7), o=st.integers(0, 14), e=st.integers(0, 9),
       protein=st.booleans(), related=st.booleans(), lens=st.lists(st.integers(1, 140), min_size=7, max_size=7))
def test_sw16_scan_random(oracle, blosum62, seed, lq, nsub, o, e, protein, related, lens):
    # the packed kernels carry the vertical gap in a one-instruction chain that needs open >= extend; the
    # engine sends every other penalty pair to the 32-bit kernel (sw16_supported), so the emulation does too
    e = min(e, o)
    mat = blosum62 if protein else oracle.Matrix.create(b"ACGT", 2, -3)
    q = seq(seed, 0, lq, protein)
    subs = []
    for i in range(nsub):
        L = lens[i]
        s = seq(seed, 1 + i, L, protein)
        if related and i % 2 == 0:
            frag = psb_data.mutate(q, seed, 50 + i, 0.15, 0.04, protein=protein)[:L]
            s = np.concatenate([frag, s])[:L]
        subs.append(s)
    outs, retry = emu_harness.sw16(q, subs, mat, o, e, 5 if protein else 2)
    assert retry == []
    for i, s in enumerate(subs):
        exp = oracle.align(q, s, mat, mode=2, open=o, gap=e)
        got = (outs["score"][i], outs["end_query"][i], outs["end_ref"][i])
        assert got == (exp["score"], exp["end_query"], exp["end_ref"]), (i, lq, len(s), o, e)


@settings(max_examples=100, deadline=None)
@given(seed=st.integers(1, 10**6), K=st.sampled_from([4, 8, 16]), nstrips=st.integers(1, 3), extra=st.integers(0, 100),
       lr=st.integers(1, 300), o=st.integers(0, 12), de=st.integers(0, 12), protein=st.booleans(), related=st.booleans())
def test_wave_gen3_random(oracle, blosum62, seed, K, nstrips, extra, lr, o, de, protein, related):
    # generation 3 needs open >= extend (the engine only routes such configurations to it)
    e = max(0, o - de)
    mat = blosum62 if protein else oracle.Matrix.create(b"ACGT", 2, -3)
    lq = max(1, 32 * K * (nstrips - 1) + 1 + extra % (32 * K))
    r = seq(seed, 0, lr, protein)
    if related:
        base = np.concatenate([r] * (lq // lr + 2))[: lq + 30]
        q = psb_data.mutate(base, seed, 1, 0.12, 0.03, protein=protein)[:lq]
        if len(q) < lq:
            q = np.concatenate([q, seq(seed, 2, lq - len(q), protein)])
    else:
        q = seq(seed, 3, lq, protein)
    exp = oracle.align(q, r, mat, mode=2, open=o, gap=e)
    got = emu_harness.wave32(q, r, mat, K, 2, o, e, v2=2)
    assert got == (exp["score"], exp["end_query"], exp["end_ref"]), (K, lq, lr, o, e)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(1, 10**6), K=st.sampled_from([1, 2, 3, 4, 8]), mode=st.sampled_from([0, 1, 2]),
       flags=st.tuples(st.integers(0, 1), st.integers(0, 1), st.integers(0, 1), st.integers(0, 1)),
       o=st.integers(0, 12), e=st.integers(0, 8), protein=st.booleans(),
       variant=st.sampled_from(["score", "stats", "stats_wide", "trace"]), profile=st.booleans(),
       lq=st.integers(1, 150), lr=st.integers(1, 90), related=st.booleans())
def test_gotoh32_random(oracle, blosum62, seed, K, mode, flags, o, e, protein, variant, profile, lq, lr, related):
    # the general 32-bit kernel: every mode / free-end combination / output variant, matrix reads or
    # the per-warp int8 profile, one to several strips (lq up to 150 rows against 32*K rows per strip)
    from test_emu_gotoh32 import compare
    mat = blosum62 if protein else oracle.Matrix.create(b"ACGT", 2, -3)
    if mode != 1:
        flags = (1, 1, 1, 1)
    r = seq(seed, 0, lr, protein)
    if related:
        base = np.concatenate([r] * (lq // lr + 2))[: lq + 20]
        q = psb_data.mutate(base, seed, 1, 0.15, 0.04, protein=protein)[:lq]
        if len(q) == 0:
            q = seq(seed, 2, lq, protein)
    else:
        q = seq(seed, 3, lq, protein)
    q2, r2 = seq(seed, 4, max(1, lq // 2), protein), seq(seed, 5, max(1, lr // 3), protein)
    compare(oracle, mat, [q, q2], [r, r2], K, mode, o, e, flags, stats=variant.startswith("stats"), trace=variant == "trace",
            wide=variant == "stats_wide", profile=profile)
