"""probe: one long global / semi-global pair through wavefront generations 2 and 3 (PSB_WAVE_GEN)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psb_data
import parasail_rs_b200 as ps
L = int(os.environ.get("NW_LEN", "50000"))
dna = ps.Matrix.create(b"ACGT", 2, -3)
r_ = psb_data.random_seq(5001, 0, L, protein=False)
q_ = psb_data.mutate(r_, 5001, 1, 0.10, 0.01, protein=False)[:L]
for name, b in (("nw", ps.Aligner.new().global_()), ("sg", ps.Aligner.new().semi_global())):
    a = b.matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
    outs = {}
    for gen in ("2", "3"):
        os.environ["PSB_WAVE_GEN"] = gen
        a.align_batch([q_], [r_])
        res = a.align_batch([q_], [r_])
        outs[gen] = (int(res.score[0]), int(res.end_query[0]), int(res.end_ref[0]), round(ps.kernel_ms(), 3))
    print(name, L, "gen2", outs["2"], "gen3", outs["3"], "same", outs["2"][:3] == outs["3"][:3], flush=True)
