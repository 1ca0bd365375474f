"""probe: per-step time of the long-pair wavefront kernel as the number of strips grows
(one pair, Lq rows x 100k columns, K=8 -> 256 rows per strip)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psb_data
import parasail_rs_b200 as ps

LR = 100000
dna = ps.Matrix.create(b"ACGT", 2, -3)
r_ = psb_data.random_seq(5001, 0, LR, protein=False)
a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
K = int(os.environ.get("PSB_WAVE_K", "8"))
os.environ["PSB_WAVE_K"] = str(K)
for nstrips in (8, 9, 16, 64, 148, 391):
    lq = 32 * K * nstrips
    q_ = psb_data.random_seq(5001, 1, lq, protein=False)
    a.align_batch([q_], [r_])
    ts = []
    for _ in range(3):
        a.align_batch([q_], [r_]); ts.append(ps.kernel_ms())
    t = min(ts)
    print(f"K {K} strips {nstrips:4d} (Lq {lq:6d}): {t:8.3f} ms, {t * 1e3 / (LR + 31):.4f} us per column step, {lq * LR / t / 1e6:7.1f} GCUPS", flush=True)
