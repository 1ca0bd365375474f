"""probe for ncu captures of the long-pair kernels: one DNA pair of length L in one mode, score only or traced
usage: python tools/wave_ncu_probe.py L {local|global_|semi_global} {score|trace|stats}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psb_data
import parasail_rs_b200 as ps
L, mode, what = int(sys.argv[1]), sys.argv[2], sys.argv[3]
dna = ps.Matrix.create(b"ACGT", 2, -3)
r_ = psb_data.random_seq(5001, 0, L, protein=False)
q_ = psb_data.mutate(r_, 5001, 1, 0.10, 0.01, protein=False)[:L]
b = getattr(ps.Aligner.new(), mode)().matrix(dna).gap_open(5).gap_extend(2)
b = b.use_trace() if what == "trace" else (b.use_stats() if what == "stats" else b.solution_width(32))
a = b.build()
res = a.align_batch([q_], [r_])
print(mode, what, L, int(res.score[0]), int(res.end_query[0]), int(res.end_ref[0]), round(ps.kernel_ms(), 3), flush=True)
