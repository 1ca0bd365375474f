"""probe: resident scan throughput for several query lengths (K classes of the packed 16-bit kernel)
usage: python tools/scan_lq_probe.py [lq ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, psb_data
import parasail_rs_b200 as ps
b62 = ps.Matrix.from_name("blosum62")
cat, off = psb_data.protein_db(2002, 2003, 200000)
db = ps.Database((cat, off), b62)
for lq in [int(x) for x in sys.argv[1:]] or [100, 256, 400, 448, 512]:
    q = psb_data.random_seq(2001, 0, lq)
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(q, False, b62)).build()
    for _ in range(2):
        a.scan(db)
    ts = []
    for _ in range(4):
        a.scan(db); ts.append(ps.kernel_ms())
    print(f"lq {lq:4d}: {min(ts):7.3f} ms  {lq * float(off[-1]) / min(ts) / 1e6:7.0f} GCUPS", flush=True)
