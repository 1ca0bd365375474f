#!/usr/bin/env python
"""Throughput of the BASELINE configs other than the headline (C1, C3, C4, C5) on one GPU, each
checked against the CPU oracle on a sample.  Sizes are the config sizes where host-side synthetic
generation allows it, otherwise a stated fraction.  Writes gpurun_out/configs.json.

  python tests/bench_configs.py [--quick]      (test infrastructure: it runs the oracle as the checker)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import psb_data  # noqa: E402


def vec_random(seed, n, length, protein):
    """n i.i.d. sequences of equal length as one (n*length) uint8 array (vectorised)"""
    total = n * length
    out = np.empty(total, dtype=np.uint8)
    step = 1 << 24
    for a in range(0, total, step):
        b = min(total, a + step)
        u = psb_data.rnd(seed, 3, np.arange(a, b, dtype=np.uint64))
        out[a:b] = psb_data.protein_letters(u) if protein else psb_data.dna_letters(u)
    return out


def vec_substitute(src, seed, rate, protein):
    """per-residue substitution with probability `rate` (vectorised, indel-free)"""
    n = len(src)
    out = src.copy()
    step = 1 << 24
    letters = np.frombuffer(psb_data.PROTEIN if protein else psb_data.DNA, dtype=np.uint8)
    for a in range(0, n, step):
        b = min(n, a + step)
        u = psb_data.rnd(seed, 4, np.arange(a, b, dtype=np.uint64))
        hit = (u % np.uint64(10000)) < np.uint64(int(rate * 10000))
        new = letters[((u >> np.uint64(20)) % np.uint64(len(letters))).astype(np.int64)]
        out[a:b] = np.where(hit, new, out[a:b])
    return out


def equal_offsets(n, length):
    return np.arange(n + 1, dtype=np.int64) * length


def check_sample(orc, omat, mode, o, e, qc, qo, rc, ro, got, keys, sample, **kw):
    idx = np.unique(np.concatenate([np.arange(min(40, len(ro) - 1)), np.random.default_rng(1).integers(0, len(ro) - 1, sample)]))
    qs = [qc[qo[i]:qo[i + 1]] for i in idx]
    rs = [rc[ro[i]:ro[i + 1]] for i in idx]
    sqc, sqo = psb_data.concat(qs)
    src, sro = psb_data.concat(rs)
    exp = orc.align_batch(sqc, sqo, src, sro, omat, mode=mode, open=o, gap=e, **kw)
    for k in keys:
        if k in ("cigar_off", "cigar_ops"):
            continue
        assert np.array_equal(getattr(got, k)[idx], exp[k]), f"mismatch in {k}"
    if "cigar_ops" in keys:
        for t, i in enumerate(idx):
            a = got.cigar_ops[got.cigar_off[i]: got.cigar_off[i + 1]]
            b = exp["cigar_ops"][exp["cigar_off"][t]: exp["cigar_off"][t + 1]]
            assert np.array_equal(a, b), f"CIGAR mismatch at pair {i}"
    return len(idx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    import parasail_rs_b200 as ps
    from oracle import oracle as orc
    out = {}
    b62 = ps.Matrix.from_name("blosum62")
    ob62 = orc.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())
    dna, odna = ps.Matrix.create(b"ACGT", 2, -3), orc.Matrix.create(b"ACGT", 2, -3)
    sm_peak32 = 148 * 64 * 1.965 / 5

    def run(name, aligner, qc, qo, rc, ro, reps=3):
        aligner.align_batch((qc, qo), (rc, ro))  # warm-up (also grows the memory pool)
        best, kms = 1e30, 0.0
        res = None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = aligner.align_batch((qc, qo), (rc, ro))
            dt = time.perf_counter() - t0
            if dt < best:
                best, kms = dt, ps.kernel_ms()
        cells = res.cells
        return res, {"function": aligner.fn_name, "pairs": len(ro) - 1, "cells": cells, "e2e_s": best, "e2e_gcups": cells / best / 1e9,
                     "kernel_ms": kms, "kernel_gcups": cells / (kms * 1e-3) / 1e9, "frac_of_s32_roofline": cells / (kms * 1e-3) / 1e9 / sm_peak32}

    want = lambda c: not args.only or c in args.only.split(",")

    if want("C1"):
        n = 10000
        qc = vec_random(1001, n, 300, True)
        rc = vec_random(1002, n, 300, True)
        rel = np.arange(0, n, 10)  # 10% related pairs
        rsub = vec_substitute(qc, 1003, 0.15, True)
        for i in rel:
            rc[i * 300:(i + 1) * 300] = rsub[i * 300:(i + 1) * 300]
        qo = ro = equal_offsets(n, 300)
        a = ps.Aligner.new().matrix(b62).gap_open(10).gap_extend(1).build()
        res, info = run("C1", a, qc, qo, rc, ro)
        info["verified_pairs"] = check_sample(orc, ob62, orc.NW, 10, 1, qc, qo, rc, ro, res, ("score", "end_query", "end_ref"), 400)
        info["config"] = "C1 full size: 10k protein pairs 300x300, nw_striped_sat, BLOSUM62 10/1"
        out["C1"] = info
        print("C1", json.dumps(info), flush=True)

    if want("C3"):
        n = 100000 if args.quick else 1000000
        wins = vec_random(3001, n, 500, False)
        starts = (psb_data.rnd(3002, 6, np.arange(n, dtype=np.uint64)) % np.uint64(340)).astype(np.int64)
        idx = (np.arange(n, dtype=np.int64) * 500 + starts)[:, None] + np.arange(150, dtype=np.int64)[None, :]
        reads = vec_substitute(wins[idx.reshape(-1)], 3003, 0.04, False)
        # a slice with real indels (sequential mutate) and 5% unrelated reads
        for i in range(0, min(n, 20000), 7):
            m = psb_data.mutate(wins[i * 500 + starts[i]: i * 500 + starts[i] + 160], 3004, i, 0.04, 0.01, protein=False)[:150]
            if len(m) == 150:
                reads[i * 150:(i + 1) * 150] = m
        unrel = np.arange(0, n, 20)
        rnd_reads = vec_random(3005, len(unrel), 150, False)
        for t, i in enumerate(unrel):
            reads[i * 150:(i + 1) * 150] = rnd_reads[t * 150:(t + 1) * 150]
        qo, ro = equal_offsets(n, 150), equal_offsets(n, 500)
        a = ps.Aligner.new().semi_global().matrix(dna).gap_open(5).gap_extend(2).use_stats().build()
        res, info = run("C3", a, reads, qo, wins, ro)
        info["verified_pairs"] = check_sample(orc, odna, orc.SG, 5, 2, reads, qo, wins, ro, res,
                                              ("score", "end_query", "end_ref", "matches", "similar", "length"), 1500, stats=True)
        info["config"] = f"C3 at {n} pairs ({n / 1e7:.0%} of the config's 10M; host-side generation bounds the size): 150 bp reads vs 500 bp windows, sg_stats_striped_sat, +2/-3, 5/2"
        out["C3"] = info
        print("C3", json.dumps(info), flush=True)

    if want("C4"):
        n = 50000 if args.quick else 200000
        qc = vec_random(4001, n, 250, True)
        rc = vec_substitute(qc, 4002, 0.20, True)
        unrel = np.arange(0, n, 5)   # 20% unrelated
        rr = vec_random(4003, len(unrel), 250, True)
        for t, i in enumerate(unrel):
            rc[i * 250:(i + 1) * 250] = rr[t * 250:(t + 1) * 250]
        for i in range(1, min(n, 15000), 6):   # real indels on a slice
            m = psb_data.mutate(qc[i * 250:(i + 1) * 250], 4004, i, 0.2, 0.02, True, 2.0)
            m = m[:250] if len(m) >= 250 else np.concatenate([m, psb_data.random_seq(4005, i, 250 - len(m))])
            rc[i * 250:(i + 1) * 250] = m
        qo = ro = equal_offsets(n, 250)
        a = ps.Aligner.new().local().matrix(b62).gap_open(10).gap_extend(1).use_trace().build()
        res, info = run("C4", a, qc, qo, rc, ro, reps=2)
        info["verified_pairs"] = check_sample(orc, ob62, orc.SW, 10, 1, qc, qo, rc, ro, res,
                                              ("score", "end_query", "end_ref", "beg_query", "beg_ref", "cigar_off", "cigar_ops"), 600, cigar=True)
        info["cigar_ops_total"] = int(res.cigar_off[-1])
        info["config"] = f"C4 at {n} pairs ({n / 1e6:.0%} of the config's 1M): protein 250x250, sw_trace_striped_sat + CIGAR, BLOSUM62 10/1"
        out["C4"] = info
        print("C4", json.dumps(info), flush=True)

    if want("C5"):
        L = 20000 if args.quick else 100000
        r = psb_data.random_seq(5001, 0, L, protein=False)
        q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)
        q = q[:L] if len(q) >= L else np.concatenate([q, psb_data.random_seq(5002, 0, L - len(q), protein=False)])
        a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
        qo, ro = np.array([0, L], dtype=np.int64), np.array([0, L], dtype=np.int64)
        res, info = run("C5", a, q, qo, r, ro, reps=2)
        t0 = time.perf_counter()
        exp = orc.align(q, r, odna, mode=orc.SW, open=5, gap=2)
        info["oracle_seconds"] = time.perf_counter() - t0
        assert (int(res.score[0]), int(res.end_query[0]), int(res.end_ref[0])) == (exp["score"], exp["end_query"], exp["end_ref"]), "C5 mismatch"
        info["score"] = int(res.score[0])
        info["verified_pairs"] = 1
        info["config"] = f"C5: one {L} x {L} DNA pair, sw_striped_32, +2/-3, 5/2, multi-warp wavefront on one GPU"
        out["C5"] = info
        print("C5", json.dumps(info), flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
