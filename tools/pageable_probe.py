#!/usr/bin/env python
"""probe: psb_align_pairs / psb_scan_host out of PAGEABLE host memory -- the library's bounce-buffer staging
(default) against the driver's own pageable path (PSB_NO_BOUNCE=1, read once at first use: run the script twice).
usage: [PSB_NO_BOUNCE=1] python tools/pageable_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, bench, psb_gen
import parasail_rs_b200 as ps

tag = "driver path" if os.environ.get("PSB_NO_BOUNCE") else "bounce buffers"
n, L = 1000000, 250
qc = psb_gen.random(11, 3, n * L, True); rc = psb_gen.random(12, 3, n * L, True)
off = np.arange(n + 1, dtype=np.int64) * L
a = ps.Aligner.new().matrix(ps.Matrix.from_name("blosum62")).gap_open(10).gap_extend(1).build()
for _ in range(2): a.align_batch((qc, off), (rc, off))
ts = []
for _ in range(5):
    t0 = time.perf_counter(); r = a.align_batch((qc, off), (rc, off)); ts.append(time.perf_counter() - t0)
print(f"{tag}: nw 250x250 x 1e6 from pageable arrays: min {min(ts)*1e3:.1f} median {np.median(ts)*1e3:.1f} ms, {r.cells/np.median(ts)/1e9:.0f} GCUPS", flush=True)
query, cat, offs = bench.make_inputs(1000000)
sa = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(query, False, ps.Matrix.from_name("blosum62"))).build()
for _ in range(3): sa.scan_host((cat, offs))
ts = []
for _ in range(5):
    t0 = time.perf_counter(); sa.scan_host((cat, offs)); ts.append(time.perf_counter() - t0)
print(f"{tag}: C2 scan_host from pageable arrays: min {min(ts)*1e3:.1f} median {np.median(ts)*1e3:.1f} ms, {400.0*float(offs[-1])/np.median(ts)/1e9:.0f} GCUPS", flush=True)
