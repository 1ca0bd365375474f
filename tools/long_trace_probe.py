"""probe: one long DNA pair WITH traceback / statistics through psb_align_pairs (the traced launch of the column-blocked
wavefront kernel + walk32_kernel): kernel time of fill + walk, and oracle-free checks at sizes the scalar oracle needs
minutes for -- the CIGAR re-scores to the reported score, consumes exactly [beg, end] of both sequences, and its
=, X, I, D counts are the `_stats` result of the same pair.
usage: python tools/long_trace_probe.py [L ...]      (default 50000 100000)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, psb_data
import parasail_rs_b200 as ps

OPEN, EXT, MATCH, MISMATCH = 5, 2, 2, -3


def check(q, r, res, st, free_begin, i=0):
    ops = res.cigar_ops[res.cigar_off[i]:res.cigar_off[i + 1]]
    ln, op = (ops >> 4).astype(np.int64), ops & 15
    n_eq, n_x = int(ln[op == 7].sum()), int(ln[op == 8].sum())
    n_i, n_d = int(ln[op == 1].sum()), int(ln[op == 2].sum())
    isgap = (op == 1) | (op == 2)
    # a leading gap run is the walk running off the table: free when the begin is free, and not part of the statistics
    lead = int(ln[0]) if len(ops) and isgap[0] else 0
    gaps = ln[isgap]
    rescore = MATCH * n_eq + MISMATCH * n_x - int((OPEN + (gaps - 1) * EXT).sum()) + ((OPEN + (lead - 1) * EXT) if lead and free_begin else 0)
    bq, br, eq, er = int(res.beg_query[i]), int(res.beg_ref[i]), int(res.end_query[i]), int(res.end_ref[i])
    # walk the CIGAR over the residues: '=' runs must be equal residues, 'X' runs different ones
    qi, ri, ok = bq, br, True
    for l, o in zip(ln.tolist(), op.tolist()):
        if o in (7, 8):
            same = q[qi:qi + l] == r[ri:ri + l]
            ok = ok and bool(same.all() if o == 7 else (~same).all())
            qi += l; ri += l
        elif o == 1: qi += l
        else: ri += l
    return {"score": int(res.score[i]), "rescore": rescore, "consumes_query": qi - 1 == eq, "consumes_ref": ri - 1 == er, "runs_match_residues": ok,
            "cigar_runs": int(len(ops)), "eq": n_eq, "x": n_x, "i": n_i, "d": n_d,
            "stats": [int(st.matches[i]), int(st.similar[i]), int(st.length[i])], "stats_agree": [n_eq, n_eq, n_eq + n_x + n_i + n_d - lead] == [int(st.matches[i]), int(st.similar[i]), int(st.length[i])]}


def main():
    dna = ps.Matrix.create(b"ACGT", MATCH, MISMATCH)
    out = []
    for L in [int(x) for x in sys.argv[1:]] or [50000, 100000]:
        r = psb_data.random_seq(5001, 0, L, protein=False)
        q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)
        q = q[:L] if len(q) >= L else np.concatenate([q, psb_data.random_seq(5002, 0, L - len(q), protein=False)])
        for mode in [m for m in ("local", "global_", "semi_global") if m in os.environ.get("PROBE_MODES", "local global_ semi_global").split()]:
            base = getattr(ps.Aligner.new(), mode)().matrix(dna).gap_open(OPEN).gap_extend(EXT)
            score_only = getattr(ps.Aligner.new(), mode)().matrix(dna).gap_open(OPEN).gap_extend(EXT).solution_width(32).build()
            tr, st = base.use_trace().build(), getattr(ps.Aligner.new(), mode)().matrix(dna).gap_open(OPEN).gap_extend(EXT).use_stats().build()
            rec = {"L": L, "mode": mode}
            for name, al in (("score_only", score_only), ("trace", tr), ("stats", st)):
                al.align_batch([q], [r])
                ts, ws = [], []
                for _ in range(2):
                    t0 = time.perf_counter(); res = al.align_batch([q], [r]); ws.append((time.perf_counter() - t0) * 1e3); ts.append(ps.kernel_ms())
                rec[name + "_kernel_ms"] = round(min(ts), 3); rec[name + "_call_ms"] = round(min(ws), 3)
                rec[name + "_result"] = [int(res.score[0]), int(res.end_query[0]), int(res.end_ref[0])]
                if name == "trace": res_t = res
                if name == "stats": res_s = res
            rec.update(check(q, r, res_t, res_s, mode != "global_"))
            rec["same_end_cell"] = rec["score_only_result"] == rec["trace_result"] == rec["stats_result"]
            rec["ok"] = bool(rec["same_end_cell"] and rec["score"] == rec["rescore"] and rec["consumes_query"] and rec["consumes_ref"] and rec["runs_match_residues"] and rec["stats_agree"])
            print(json.dumps(rec), flush=True)
            out.append(rec)
    return 0 if all(x["ok"] for x in out) else 1


if __name__ == "__main__":
    sys.exit(main())
