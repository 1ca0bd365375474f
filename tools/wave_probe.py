"""probe: per-step time of the wavefront kernel (routed long subjects of a scan; one long pair)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, psb_data
import parasail_rs_b200 as ps
b62 = ps.Matrix.from_name("blosum62")
q = psb_data.random_seq(1, 0, 400)
for nsub, L in ((1, 5000), (3, 5000), (24, 5000), (3, 20000)):
    subs = [psb_data.random_seq(2, i, L) for i in range(nsub)]
    db = ps.Database(subs, b62)
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(q, False, b62)).build()
    a.scan(db)
    t0 = time.perf_counter(); r = a.scan(db); dt = time.perf_counter() - t0
    print(f"scan route: {nsub} x {L}: kernel {ps.kernel_ms():.3f} ms, wall {dt*1e3:.3f} ms, retried {r.n_retried}, per step {ps.kernel_ms()*1e3/(L+31):.3f} us")
dna = ps.Matrix.create(b"ACGT", 2, -3)
for L in (4000, 20000):
    r_ = psb_data.random_seq(3, 0, L, protein=False); q_ = psb_data.random_seq(3, 1, L, protein=False)
    a = ps.Aligner.new().local().matrix(dna).gap_open(5).gap_extend(2).solution_width(32).build()
    a.align_batch([q_], [r_])
    a.align_batch([q_], [r_])
    print(f"single pair {L}x{L}: kernel {ps.kernel_ms():.3f} ms, per step {ps.kernel_ms()*1e3/(L+31):.3f} us")
