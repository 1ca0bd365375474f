#!/usr/bin/env python
"""Generates tests/golden/c2_block_hashes.json: the CPU result of config C2 for ALL 10^6 subjects, as one
SHA-1 per block of 1000 subjects over (score, end_query, end_ref).  The results come from the striped AVX2
restatement (oracle/striped_cpu.cpp, itself checked against the scalar oracle), run once on the host cores
(~10^11 cells: tens of minutes); the GPU run then only hashes its own results block by block
(bench.py --verify-all, tests/test_gpu_parity.py::test_scan_c2_full_golden).  SURVEY 8d: "all 10^6 of C2
(cached golden)".

  python tools/make_c2_golden.py [--db 1000000]
"""
import argparse
import hashlib
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

BLOCK = 1000


def block_hashes(score, end_query, end_ref, block=BLOCK):
    out = []
    n = len(score)
    for a in range(0, n, block):
        b = min(n, a + block)
        h = hashlib.sha1()
        for arr in (score, end_query, end_ref):
            h.update(np.ascontiguousarray(arr[a:b], dtype=np.int32).tobytes())
        out.append(h.hexdigest()[:16])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", type=int, default=1000000)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    import bench
    from oracle import oracle as orc
    query, cat, off = bench.make_inputs(args.db)
    omat = orc.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())
    t0 = time.time()
    res, secs = orc.striped_sw_scan(query, cat, off, omat, bench.OPEN, bench.GAP, threads=args.threads)
    # a slice against the scalar oracle, so the golden file does not rest on the striped port alone
    ns = min(args.db, 10000)
    exp = orc.align_batch(query, np.array([0, len(query)]), cat, off[: ns + 1], omat, mode=orc.SW, open=bench.OPEN, gap=bench.GAP,
                          shared_query=True, threads=0)
    for k in ("score", "end_query", "end_ref"):
        assert np.array_equal(res[k][:ns], exp[k]), f"striped port disagrees with the scalar oracle in {k}"
    out = {"what": "config C2: sw_striped_profile_sat, 400-aa query vs the synthetic protein database of bench.py; SHA-1[:16] of the int32 "
                   "(score, end_query, end_ref) arrays per block of subjects",
           "generated_by": "tools/make_c2_golden.py (oracle/striped_cpu.cpp, first %d subjects also equal to the scalar oracle)" % ns,
           "db": args.db, "block": BLOCK, "seeds": [bench.QUERY_SEED, bench.LEN_SEED, bench.RES_SEED], "open": bench.OPEN, "gap": bench.GAP,
           "cells": float(len(query)) * float(off[-1]), "cpu_seconds": secs, "threads": int(res["threads"]),
           "score_sum": int(res["score"].astype(np.int64).sum()), "hashes": block_hashes(res["score"], res["end_query"], res["end_ref"])}
    path = os.path.join(ROOT, "tests", "golden", "c2_block_hashes.json" if args.db == 1000000 else f"c2_block_hashes_{args.db}.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, f"{secs:.0f} s of striped CPU scan, total {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
