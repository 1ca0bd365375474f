"""probe: config C5 (one 100 kb x 100 kb DNA pair, sw_striped_32) over the wavefront launch knobs
usage: python tools/c5_probe.py [K:warps ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, psb_data
import parasail_rs_b200 as ps

L = int(os.environ.get("C5_LEN", "100000"))
dna = ps.Matrix.create(b"ACGT", 2, -3)
r_ = psb_data.random_seq(5001, 0, L, protein=False)
q_ = psb_data.mutate(r_, 5001, 1, 0.10, 0.01, protein=False)[:L]
a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
settings = [tuple(int(x) for x in s.split(":")) for s in sys.argv[1:]] or [(8, 8), (8, 4), (8, 2), (4, 8), (4, 4), (16, 8), (16, 2), (2, 8)]
ref = None
for K, W in settings:
    os.environ["PSB_WAVE_K"] = str(K); os.environ["PSB_WAVE_WARPS"] = str(W)
    a.align_batch([q_], [r_])
    ts = []
    for _ in range(3):
        res = a.align_batch([q_], [r_]); ts.append(ps.kernel_ms())
    out = (int(res.score[0]), int(res.end_query[0]), int(res.end_ref[0]))
    ref = ref or out
    print(f"K {K:2d} warps/CTA {W:2d}: {min(ts):8.3f} ms  {len(q_) * len(r_) / min(ts) / 1e6:7.1f} GCUPS  per step {min(ts) * 1e3 / (L + 31):.3f} us  result {out} same={out == ref}", flush=True)
