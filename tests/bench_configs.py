#!/usr/bin/env python
"""Throughput AND verification of the BASELINE configs other than the headline (C1, C3, C4, C5) on one
GPU, at the sizes SURVEY.md 8d asks for:

  C1  all 10 000 pairs checked against the scalar oracle
  C3  the full 10^7 pairs run; a fixed 10^5-pair sample + every pair whose score is in the top / bottom
      0.1 % checked against the oracle (score, ends, matches, similar, length)
  C4  the full 10^6 pairs run; a fixed 10^5-pair sample checked against the oracle (score, ends, CIGAR
      words, beg_query, beg_ref) and the CIGAR re-score / recount property on ALL pairs
  C5  the full 100 kb x 100 kb pair against the scalar oracle
  +   single-pair Aligner::align latency (p50 / p99 over 10^4 calls)

Writes gpurun_out/r2_configs.json (copied to profiles/ when it is to be judged).

  python tests/bench_configs.py [--quick] [--only C1,C3]     (test infrastructure: the oracle is the checker)
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
import psb_gen  # noqa: E402


def equal_offsets(n, length):
    return np.arange(n + 1, dtype=np.int64) * length


def subset(cat, off, idx):
    """(cat, off) of the chosen sequences of an equal-length corpus"""
    L = int(off[1] - off[0])
    rows = cat.reshape(-1, L)[idx]
    return np.ascontiguousarray(rows).reshape(-1), equal_offsets(len(idx), L)


def check_sample(orc, omat, mode, o, e, qc, qo, rc, ro, got, keys, idx, **kw):
    sqc, sqo = subset(qc, qo, idx)
    src, sro = subset(rc, ro, idx)
    t0 = time.perf_counter()
    exp = orc.align_batch(sqc, sqo, src, sro, omat, mode=mode, open=o, gap=e, threads=0, **kw)
    secs = time.perf_counter() - t0
    for k in keys:
        if k in ("cigar_off", "cigar_ops"):
            continue
        g = getattr(got, k)[idx]
        if not np.array_equal(g, exp[k]):
            bad = np.nonzero(g != exp[k])[0]
            raise AssertionError(f"mismatch in {k}: {len(bad)} of {len(idx)} pairs, first pair {idx[bad[0]]}: got {g[bad[0]]} expected {exp[k][bad[0]]}")
    if "cigar_ops" in keys:
        glen = np.diff(got.cigar_off)[idx]
        elen = np.diff(exp["cigar_off"])
        assert np.array_equal(glen, elen), "CIGAR lengths differ"
        starts = got.cigar_off[idx]
        take = np.repeat(starts, glen) + (np.arange(int(glen.sum())) - np.repeat(np.cumsum(glen) - glen, glen))
        assert np.array_equal(got.cigar_ops[take], exp["cigar_ops"]), "CIGAR words differ"
    return len(idx), secs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="a tenth of C3 / C4, 20 kb C5")
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_configs.json"))
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
    os.makedirs(os.path.dirname(args.out), exist_ok=True)

    import torch
    def pinned(a):
        """a page-locked copy of a numpy array (what a caller who cares about upload speed hands the library)"""
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    def run(aligner, qc, qo, rc, ro, reps=2):
        # end to end: pinned host buffers in, pinned result arrays out, the library's own pass pipeline (two lanes)
        keep = [pinned(x) for x in (qc, qo, rc, ro)]
        pq, pqo, pr, pro = [t.numpy() for t in keep]
        for _ in range(2):
            aligner.align_batch((pq, pqo), (pr, pro))  # warm-up (also grows the memory pool)
        best = 1e30
        res = None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = aligner.align_batch((pq, pqo), (pr, pro))
            best = min(best, time.perf_counter() - t0)
        # the same call from pageable memory (plain numpy arrays; the library stages them through its bounce buffers)
        aligner.align_batch((qc, qo), (rc, ro))
        pageable = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            aligner.align_batch((qc, qo), (rc, ro))
            pageable = min(pageable, time.perf_counter() - t0)
        # kernel time: the passes one after the other on one stream, summed timed regions (uploads excluded)
        os.environ["PSB_PAIRS_LANES"] = "1"
        aligner.align_batch((pq, pqo), (pr, pro))
        kms = 1e30
        for _ in range(reps):
            aligner.align_batch((pq, pqo), (pr, pro))
            kms = min(kms, ps.kernel_ms())
        launches = ps.launches()
        del os.environ["PSB_PAIRS_LANES"]
        cells = res.cells
        gc = cells / (kms * 1e-3) / 1e9
        return res, {"function": aligner.fn_name, "pairs": len(ro) - 1, "cells": cells, "e2e_s": best, "e2e_gcups": cells / best / 1e9,
                     "e2e_pageable_s": pageable, "e2e_pageable_gcups": cells / pageable / 1e9,
                     "kernel_ms": kms, "kernel_gcups": gc, "frac_of_s32_roofline": gc / sm_peak32, "frac_of_s16x2_roofline": gc / (2 * sm_peak32),
                     "launches": launches}

    def save():
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)

    want = lambda c: not args.only or c in args.only.split(",")

    if want("C1"):
        n = 10000
        qc = psb_gen.random(1001, 3, n * 300, True)
        rc = psb_gen.random(1002, 3, n * 300, True)
        rsub = psb_gen.substitute(qc, 1003, 4, 0.15, True)
        rel = np.arange(0, n, 10)  # 10% related pairs
        rc.reshape(n, 300)[rel] = rsub.reshape(n, 300)[rel]
        qo = ro = equal_offsets(n, 300)
        a = ps.Aligner.new().matrix(b62).gap_open(10).gap_extend(1).build()
        res, info = run(a, qc, qo, rc, ro, reps=3)
        info["verified_pairs"], info["oracle_seconds"] = check_sample(orc, ob62, orc.NW, 10, 1, qc, qo, rc, ro, res, ("score", "end_query", "end_ref"), np.arange(n))
        info["verified"] = "all pairs vs the scalar oracle"
        info["config"] = "C1 full size: 10k protein pairs 300x300, nw_striped_sat, BLOSUM62 10/1"
        out["C1"] = info
        print("C1", json.dumps(info), flush=True)
        save()

    if want("C3"):
        n = 1000000 if args.quick else 10000000
        t0 = time.perf_counter()
        wins = psb_gen.random(3001, 3, n * 500, False)
        st = psb_gen.starts(3002, 6, n, 340)
        reads = psb_gen.substitute(psb_gen.gather(wins, n, 500, st, 150), 3003, 4, 0.04, False)
        # a slice with real indels (sequential mutate) and 5% unrelated reads
        for i in range(0, min(n, 20000), 7):
            m = psb_data.mutate(wins[i * 500 + st[i]: i * 500 + st[i] + 160], 3004, i, 0.04, 0.01, protein=False)[:150]
            if len(m) == 150:
                reads[i * 150:(i + 1) * 150] = m
        unrel = np.arange(0, n, 20)
        reads.reshape(n, 150)[unrel] = psb_gen.random(3005, 3, len(unrel) * 150, False).reshape(-1, 150)
        gen_s = time.perf_counter() - t0
        qo, ro = equal_offsets(n, 150), equal_offsets(n, 500)
        a = ps.Aligner.new().semi_global().matrix(dna).gap_open(5).gap_extend(2).use_stats().build()
        res, info = run(a, reads, qo, wins, ro, reps=1 if not args.quick else 2)
        ns = 10000 if args.quick else 100000
        sample = np.unique(np.concatenate([np.arange(min(20000, n)), np.random.default_rng(1).integers(0, n, ns)]))[:ns + 20000]
        order = np.argsort(res.score, kind="stable")
        k = max(1, n // 1000)
        extremes = np.concatenate([order[:k], order[-k:]])
        idx = np.unique(np.concatenate([sample, extremes]))
        info["verified_pairs"], info["oracle_seconds"] = check_sample(orc, odna, orc.SG, 5, 2, reads, qo, wins, ro, res,
                                                                        ("score", "end_query", "end_ref", "matches", "similar", "length"), idx, stats=True)
        info["verified"] = f"fixed sample of {len(sample)} pairs (incl. the {min(20000, n) // 7} pairs with real indels) + the {2 * k} pairs with the lowest / highest scores, vs the scalar oracle"
        info["generation_s"] = gen_s
        info["config"] = f"C3 at {n} pairs ({n / 1e7:.0%} of the config's 10M): 150 bp reads vs 500 bp windows, sg_stats_striped_sat, +2/-3, 5/2"
        out["C3"] = info
        print("C3", json.dumps(info), flush=True)
        save()
        del wins, reads, res

    if want("C4"):
        n = 100000 if args.quick else 1000000
        qc = psb_gen.random(4001, 3, n * 250, True)
        rc = psb_gen.substitute(qc, 4002, 4, 0.20, True)
        unrel = np.arange(0, n, 5)   # 20% unrelated
        rc.reshape(n, 250)[unrel] = psb_gen.random(4003, 3, len(unrel) * 250, True).reshape(-1, 250)
        for i in range(1, min(n, 15000), 6):   # real indels on a slice
            m = psb_data.mutate(qc[i * 250:(i + 1) * 250], 4004, i, 0.2, 0.02, True, 2.0)
            m = m[:250] if len(m) >= 250 else np.concatenate([m, psb_data.random_seq(4005, i, 250 - len(m))])
            rc[i * 250:(i + 1) * 250] = m
        qo = ro = equal_offsets(n, 250)
        a = ps.Aligner.new().local().matrix(b62).gap_open(10).gap_extend(1).use_trace().build()
        res, info = run(a, qc, qo, rc, ro, reps=2)
        ns = 10000 if args.quick else 100000
        idx = np.unique(np.concatenate([np.arange(min(15000, n)), np.random.default_rng(2).integers(0, n, ns)]))
        info["verified_pairs"], info["oracle_seconds"] = check_sample(orc, ob62, orc.SW, 10, 1, qc, qo, rc, ro, res,
                                                                        ("score", "end_query", "end_ref", "beg_query", "beg_ref", "cigar_off", "cigar_ops"), idx, cigar=True)
        t0 = time.perf_counter()
        bad, first = orc.cigar_check_batch(qc, qo, rc, ro, ob62, orc.SW, 10, 1, res.cigar_ops, res.cigar_off, res.beg_query, res.beg_ref,
                                           res.end_query, res.end_ref, res.score)
        info["cigar_property_all_pairs"] = {"pairs": n, "failing": bad, "first_failing": first, "seconds": time.perf_counter() - t0,
                                            "what": "every CIGAR re-walked on the host: consumes exactly beg..end, '='/'X' agree with the residues, substitution scores minus affine gap costs equal the reported score"}
        assert bad == 0, f"CIGAR property fails for {bad} pairs, first {first}"
        info["verified"] = f"{len(idx)} pairs (incl. the {min(15000, n) // 6} pairs with real indels) vs the scalar oracle: score, ends, CIGAR words, begins; CIGAR re-score property on all {n}"
        info["cigar_ops_total"] = int(res.cigar_off[-1])
        info["config"] = f"C4 at {n} pairs ({n / 1e6:.0%} of the config's 1M): protein 250x250, sw_trace_striped_sat + CIGAR, BLOSUM62 10/1"
        out["C4"] = info
        print("C4", json.dumps(info), flush=True)
        save()
        del qc, rc, res

    if want("C5"):
        L = 20000 if args.quick else 100000
        r = psb_data.random_seq(5001, 0, L, protein=False)
        q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)
        q = q[:L] if len(q) >= L else np.concatenate([q, psb_data.random_seq(5002, 0, L - len(q), protein=False)])
        a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
        qo, ro = np.array([0, L], dtype=np.int64), np.array([0, L], dtype=np.int64)
        res, info = run(a, q, qo, r, ro, reps=3)
        t0 = time.perf_counter()
        exp = orc.align(q, r, odna, mode=orc.SW, open=5, gap=2)
        info["oracle_seconds"] = time.perf_counter() - t0
        assert (int(res.score[0]), int(res.end_query[0]), int(res.end_ref[0])) == (exp["score"], exp["end_query"], exp["end_ref"]), "C5 mismatch"
        info["score"] = int(res.score[0])
        info["verified_pairs"] = 1
        info["config"] = f"C5: one {L} x {L} DNA pair, sw_striped_32, +2/-3, 5/2, multi-warp wavefront on one GPU"
        out["C5"] = info
        print("C5", json.dumps(info), flush=True)
        save()

    if want("latency"):
        # single-pair Aligner::align through the 7-argument C entry point the crate calls [REF src/aligner/mod.rs:411-429]
        q = psb_data.random_seq(6001, 0, 300)
        r = psb_data.mutate(q, 6001, 1, 0.2, 0.03)
        qb, rb = bytes(q), bytes(r)
        info = {}
        for name, al in (("nw_striped_sat", ps.Aligner.new().matrix(b62).gap_open(10).gap_extend(1).build()),
                         ("sw_trace_striped_sat", ps.Aligner.new().local().matrix(b62).gap_open(10).gap_extend(1).use_trace().build())):
            for _ in range(200):
                al.align(qb, rb)
            ts = np.empty(2000 if args.quick else 10000)
            for i in range(len(ts)):
                t0 = time.perf_counter()
                al.align(qb, rb)
                ts[i] = time.perf_counter() - t0
            info[name] = {"calls": len(ts), "p50_us": float(np.percentile(ts, 50) * 1e6), "p99_us": float(np.percentile(ts, 99) * 1e6),
                          "mean_us": float(ts.mean() * 1e6), "pair": "300 x ~300 protein"}
        out["single_pair_latency"] = info
        print("latency", json.dumps(info), flush=True)
        save()


if __name__ == "__main__":
    main()
