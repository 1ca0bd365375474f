#!/usr/bin/env python
"""Regenerates tests/golden/long_pair_50k.json: the scalar oracle's results (oracle/gotoh_oracle.c) for the seeded
50 kb x 50 kb DNA pair of tests/test_gpu_parity.py::test_long_pair_50kb_against_cached_oracle, in the three modes.
About a minute of CPU per mode; run with --check to compare against the committed file instead of writing it.

The file caches THIS repository's own CPU restatement (it is not parasail output): it lets the GPU suite check a
pair whose oracle run would take minutes on the GPU box."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import psb_data
from oracle import oracle as orc

L = 50000
PATH = os.path.join(ROOT, "tests", "golden", "long_pair_50k.json")


def inputs():
    r = psb_data.random_seq(5001, 0, L, protein=False)
    q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)
    q = q[:L] if len(q) >= L else np.concatenate([q, psb_data.random_seq(5002, 0, L - len(q), protein=False)])
    return q, r


def main():
    check = "--check" in sys.argv
    q, r = inputs()
    m = orc.Matrix.create(b"ACGT", 2, -3)
    cases = []
    for mode, name in ((orc.SW, "sw"), (orc.NW, "nw"), (orc.SG, "sg")):
        t0 = time.perf_counter()
        e = orc.align(q, r, m, mode=mode, open=5, gap=2)
        cases.append({"mode": int(mode), "name": name, "score": int(e["score"]), "end_query": int(e["end_query"]), "end_ref": int(e["end_ref"]),
                      "oracle_seconds": round(time.perf_counter() - t0, 1)})
        print(cases[-1], flush=True)
    doc = {"what": "long-pair golden results: oracle (oracle/gotoh_oracle.c) outputs for 50 kb x 50 kb DNA pairs generated as in tests/bench_configs.py C5 "
                   "(seed 5001/5002, mutate 10 % subst / 1 % indel), +2/-3, open 5, extend 2; a cache of this repository's own CPU result, not parasail output "
                   "(tools/make_long_pair_golden.py)", "L": L, "cases": cases}
    if check:
        old = json.load(open(PATH))
        key = lambda c: (c["mode"], c["score"], c["end_query"], c["end_ref"])
        assert sorted(map(key, old["cases"])) == sorted(map(key, cases)), "committed golden file differs from the oracle"
        print("golden file reproduced")
    else:
        json.dump(doc, open(PATH, "w"))
        print("wrote", PATH)


if __name__ == "__main__":
    main()
