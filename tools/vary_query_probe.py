#!/usr/bin/env python
"""probe: psb_scan_host with a DIFFERENT query (length) on every call, as a search service sees it: does the piece plan's
dependence on the query length (scan time per residue) cost anything through changing allocation sizes?
usage: python tools/vary_query_probe.py [nshards]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, bench, psb_data
import parasail_rs_b200 as ps

nsh = int(sys.argv[1]) if len(sys.argv) > 1 else 1
query, cat, off = bench.make_inputs(1000000)
total = int(off[-1])
cut = int(np.searchsorted(off, total // nsh))
sc, so = cat[: off[cut]], off[: cut + 1]
pc = torch.empty(len(sc), dtype=torch.uint8, pin_memory=True); pc.numpy()[:] = sc
po = torch.empty(len(so), dtype=torch.int64, pin_memory=True); po.numpy()[:] = so
b62 = ps.Matrix.from_name("blosum62")
lens = [400, 150, 320, 90, 250, 400, 200, 380, 120, 300] * 4
rows = []
for i, lq in enumerate(lens):
    q = psb_data.random_seq(77, i, lq)
    t0 = time.perf_counter()
    a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(q, False, b62)).build()
    a.scan_host((pc.numpy(), po.numpy()))
    dt = (time.perf_counter() - t0) * 1e3
    rows.append((lq, dt, ps.kernel_ms()))
for rnd in range(4):
    print(f"round {rnd}: " + "  ".join(f"{lq}:{dt:.2f}/{k:.2f}" for lq, dt, k in rows[rnd * 10:(rnd + 1) * 10]), flush=True)
last = rows[30:]
print(f"last round: call / kernel time, mean overhead {np.mean([dt - k for _, dt, k in last]):.3f} ms, worst {max(dt - k for _, dt, k in last):.3f} ms")
