#!/usr/bin/env python
"""Adjudicates the oracle's [UP] assumptions against a REAL libparasail the day one is available.

No parasail binary or source exists in this environment, so every rule of
oracle/UP_ASSUMPTIONS.md is recalled, not read.  This tool makes settling them one command:

    python tools/diff_vs_parasail.py [--lib /path/to/libparasail.so] [--pairs 2000]

It looks for libparasail.so under baseline/_ref/ (or takes --lib), binds the plain parasail C API through
ctypes (parasail_matrix_lookup / parasail_matrix_create, the kernels by name through
parasail_lookup_function, the result getters, parasail_result_get_cigar / parasail_cigar_decode), runs
reduced samples of the five BASELINE configs plus the hand-made tie cases through it, and compares every
field with the oracle.  For each field that differs it re-runs the oracle with each rule switch flipped
(oracle.set_rules) and reports which single flip -- or which pair of flips -- makes the differences vanish.

Any library exporting parasail's C ABI can be examined, including this repository's own
libparasail_b200.so (on a GPU box), which is how the tool itself is tested:

    python tools/diff_vs_parasail.py --lib parasail_rs_b200/libparasail_b200.so
"""
import argparse
import ctypes as C
import glob
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import psb_data  # noqa: E402
from oracle import oracle as orc  # noqa: E402


class Result(C.Structure):
    _fields_ = [("score", C.c_int), ("end_query", C.c_int), ("end_ref", C.c_int), ("flag", C.c_int), ("extra", C.c_void_p)]


class Cigar(C.Structure):
    _fields_ = [("seq", C.POINTER(C.c_uint32)), ("len", C.c_int), ("beg_query", C.c_int), ("beg_ref", C.c_int)]


class Parasail:
    """the handful of upstream entry points the comparison needs"""

    def __init__(self, path):
        L = self.L = C.CDLL(path)
        L.parasail_matrix_lookup.restype = C.c_void_p
        L.parasail_matrix_lookup.argtypes = [C.c_char_p]
        L.parasail_matrix_create.restype = C.c_void_p
        L.parasail_matrix_create.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.parasail_lookup_function.restype = C.c_void_p
        L.parasail_lookup_function.argtypes = [C.c_char_p]
        L.parasail_result_free.argtypes = [C.POINTER(Result)]
        for g in ("score", "end_query", "end_ref", "matches", "similar", "length"):
            f = getattr(L, "parasail_result_get_" + g)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(Result)]
        L.parasail_result_get_cigar.restype = C.POINTER(Cigar)
        L.parasail_result_get_cigar.argtypes = [C.POINTER(Result), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        L.parasail_cigar_free.argtypes = [C.POINTER(Cigar)]
        self.FN = C.CFUNCTYPE(C.POINTER(Result), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p)

    def matrix(self, spec):
        m = self.L.parasail_matrix_lookup(spec.encode()) if isinstance(spec, str) else self.L.parasail_matrix_create(spec[0], spec[1], spec[2])
        if not m:
            raise RuntimeError(f"matrix {spec!r} not available in this library")
        return m

    def align(self, fn_name, q, r, o, e, matrix, stats=False, cigar=False):
        ptr = self.L.parasail_lookup_function(fn_name.encode())
        if not ptr:
            raise RuntimeError(f"{fn_name} not found")
        qb, rb = bytes(q), bytes(r)
        res = self.FN(ptr)(qb, len(qb), rb, len(rb), o, e, matrix)
        out = {k: getattr(self.L, "parasail_result_get_" + k)(res) for k in ("score", "end_query", "end_ref")}
        if stats:
            out.update({k: getattr(self.L, "parasail_result_get_" + k)(res) for k in ("matches", "similar", "length")})
        if cigar:
            c = self.L.parasail_result_get_cigar(res, qb, len(qb), rb, len(rb), matrix)
            if c:
                out["cigar_ops"] = [int(c.contents.seq[i]) for i in range(c.contents.len)]
                out["beg_query"], out["beg_ref"] = c.contents.beg_query, c.contents.beg_ref
                self.L.parasail_cigar_free(c)
        self.L.parasail_result_free(res)
        return out


def corpora(npairs):
    """(name, fn_name, matrix spec, oracle matrix, mode, flags, open, gap, stats, cigar, pairs)"""
    ob62 = orc.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())
    odna = orc.Matrix.create(b"ACGT", 2, -3)
    s = lambda b: np.frombuffer(b, dtype=np.uint8)
    ties_q = [s(b"ACTACGGG"), s(b"ACGTACGT"), s(b"AAAA"), s(b"ACGT"), s(b"GGAACCTT"), s(b"TTTT"), s(b"ACGTTGCA")]
    ties_r = [s(b"ACTTACG"), s(b"ACGTTTACGT"), s(b"AAAAAAAA"), s(b"TGCA"), s(b"GGTTAACC"), s(b"AAAA"), s(b"ACGTACGTTGCATGCA")]
    out = []
    qs, rs = psb_data.protein_pairs(1001, min(npairs, 400), 300, related_frac=0.3)
    out.append(("C1", "nw_striped_sat", "blosum62", ob62, orc.NW, (1, 1, 1, 1), 10, 1, False, False, qs, rs))
    q = psb_data.random_seq(2001, 0, 400)
    cat, off = psb_data.protein_db(2002, 2003, min(npairs, 1500), query=q, planted_frac=0.1)
    out.append(("C2", "sw_striped_sat", "blosum62", ob62, orc.SW, (1, 1, 1, 1), 10, 1, False, False,
                [q] * (len(off) - 1), [cat[off[i]:off[i + 1]] for i in range(len(off) - 1)]))
    qs, rs = psb_data.dna_read_pairs(3001, min(npairs, 2000))
    out.append(("C3", "sg_stats_striped_sat", (b"ACGT", 2, -3), odna, orc.SG, (1, 1, 1, 1), 5, 2, True, False, qs, rs))
    qs, rs = psb_data.protein_pairs(4001, min(npairs, 600), 250, related_frac=0.8, p_sub=0.2, p_indel=0.04, geometric_mean=2.0)
    out.append(("C4", "sw_trace_striped_sat", "blosum62", ob62, orc.SW, (1, 1, 1, 1), 10, 1, False, True, qs, rs))
    r = psb_data.random_seq(5001, 0, 6000, protein=False)
    qq = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)[:6000]
    out.append(("C5", "sw_striped_32", (b"ACGT", 2, -3), odna, orc.SW, (1, 1, 1, 1), 5, 2, False, False, [qq], [r]))
    for mode, name in ((orc.NW, "nw"), (orc.SG, "sg"), (orc.SW, "sw")):
        for o, e in ((5, 2), (1, 1), (0, 0)):
            out.append((f"ties/{name}/{o}-{e}", f"{name}_trace_striped_sat", (b"ACGT", 2, -3), odna, mode, (1, 1, 1, 1), o, e, False, True, ties_q, ties_r))
            out.append((f"ties-stats/{name}/{o}-{e}", f"{name}_stats_striped_sat", (b"ACGT", 2, -3), odna, mode, (1, 1, 1, 1), o, e, True, False, ties_q, ties_r))
    sgq, sgr = psb_data.dna_read_pairs(3101, 200, 60, 90)
    for suffix, flags in (("_qb", (1, 0, 0, 0)), ("_qe", (0, 1, 0, 0)), ("_db", (0, 0, 1, 0)), ("_de", (0, 0, 0, 1)), ("_qx", (1, 1, 0, 0)),
                          ("_dx", (0, 0, 1, 1)), ("_qb_de", (1, 0, 0, 1)), ("_qe_db", (0, 1, 1, 0))):
        out.append((f"sg{suffix}", f"sg{suffix}_striped_sat", (b"ACGT", 2, -3), odna, orc.SG, flags, 5, 2, False, False, sgq, sgr))
    return out


def oracle_fields(case, rules):
    name, fn, mspec, omat, mode, flags, o, e, stats, cigar, qs, rs = case
    orc.set_rules(**rules)
    qc, qo = psb_data.concat(qs)
    rc, ro = psb_data.concat(rs)
    exp = orc.align_batch(qc, qo, rc, ro, omat, mode=mode, open=o, gap=e, s1_beg=flags[0], s1_end=flags[1], s2_beg=flags[2],
                          s2_end=flags[3], stats=stats, cigar=cigar, threads=0)
    orc.set_rules()
    return exp


def compare(case, got, exp):
    """{field: number of differing pairs}"""
    name, fn, mspec, omat, mode, flags, o, e, stats, cigar, qs, rs = case
    keys = ["score", "end_query", "end_ref"] + (["matches", "similar", "length"] if stats else []) + (["beg_query", "beg_ref"] if cigar else [])
    diff = {}
    for k in keys:
        d = sum(1 for i, g in enumerate(got) if k in g and int(g[k]) != int(exp[k][i]))
        if d:
            diff[k] = d
    if cigar:
        d = 0
        for i, g in enumerate(got):
            e_ops = [int(x) for x in exp["cigar_ops"][exp["cigar_off"][i]: exp["cigar_off"][i + 1]]]
            if "cigar_ops" in g and g["cigar_ops"] != e_ops:
                d += 1
        if d:
            diff["cigar"] = d
    return diff


ALTERNATIVES = {"sw_end_tie": (1, 2), "sg_col_wins_tie": (1,), "sg_row_last_wins": (1,), "h_priority": (1, 2, 3), "open_on_tie": (1,),
                "match_raw_bytes": (1,), "count_boundary_gaps": (1,), "cigar_edge_stop": (1,), "cigar_swap_id": (1,), "sg_flag_swap": (1,),
                "zero_beats_diag": (0,), "band_rule": (1,)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="")
    ap.add_argument("--pairs", type=int, default=2000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "diff_vs_parasail.json"))
    args = ap.parse_args()
    path = args.lib
    if not path:
        hits = sorted(glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", "libparasail*.so*"), recursive=True))
        if not hits:
            print("no libparasail.so under baseline/_ref/ and no --lib given: nothing to adjudicate (parity stays unpinned)")
            return 2
        path = hits[0]
    lib = Parasail(os.path.abspath(path))
    report = {"library": path, "cases": {}, "verdict": {}}
    flips_needed = {}
    for case in corpora(args.pairs):
        name, fn, mspec, omat, mode, flags, o, e, stats, cigar, qs, rs = case
        try:
            m = lib.matrix(mspec)
            got = [lib.align(fn, q, r, o, e, m, stats, cigar) for q, r in zip(qs, rs)]
        except RuntimeError as ex:
            report["cases"][name] = {"skipped": str(ex)}
            continue
        base = compare(case, got, oracle_fields(case, {}))
        entry = {"function": fn, "pairs": len(qs), "differences_with_default_rules": base}
        if base:
            fixes = []
            singles = [(k, v) for k, vs in ALTERNATIVES.items() for v in vs]
            for k, v in singles:
                if not compare(case, got, oracle_fields(case, {k: v})):
                    fixes.append({k: v})
            if not fixes:
                for (k1, v1), (k2, v2) in itertools.combinations(singles, 2):
                    if k1 != k2 and not compare(case, got, oracle_fields(case, {k1: v1, k2: v2})):
                        fixes.append({k1: v1, k2: v2})
            entry["rule_flips_that_remove_all_differences"] = fixes
            for f in fixes:
                flips_needed[json.dumps(f, sort_keys=True)] = flips_needed.get(json.dumps(f, sort_keys=True), 0) + 1
        report["cases"][name] = entry
        print(name, fn, "differences:", base or "none", "fixes:", entry.get("rule_flips_that_remove_all_differences", "-"), flush=True)
    ndiff = sum(1 for c in report["cases"].values() if c.get("differences_with_default_rules"))
    report["verdict"] = {"cases_that_differ": ndiff, "flips": flips_needed,
                         "meaning": "0 differing cases = the default rules of oracle/UP_ASSUMPTIONS.md are what this library does"}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report["verdict"]))
    return 0 if ndiff == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
