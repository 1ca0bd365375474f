"""one long pair (2048 x 100k DNA, sw_striped_32) through the wavefront kernel -- the ncu target"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psb_data
import parasail_rs_b200 as ps
dna = ps.Matrix.create(b"ACGT", 2, -3)
r_ = psb_data.random_seq(5001, 0, 100000, protein=False)
q_ = psb_data.random_seq(5001, 1, int(os.environ.get("WAVE_LQ", "2048")), protein=False)
a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
for _ in range(2):
    res = a.align_batch([q_], [r_])
print("kernel ms", ps.kernel_ms(), int(res.score[0]))
