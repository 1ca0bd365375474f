#!/usr/bin/env python
"""probe: where the host time of psb_align_pairs goes (PSB_DEBUG_TIMING) on C4-like and C3-like batches.
usage: python tools/pairs_host_probe.py [n_pairs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, psb_gen
import parasail_rs_b200 as ps

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True); t.numpy()[...] = a; return t
def offs(n, L): return np.arange(n + 1, dtype=np.int64) * L
b62 = ps.Matrix.from_name("blosum62")
cases = [("C4 sw_trace 250x250", ps.Aligner.new().local().matrix(b62).gap_open(10).gap_extend(1).use_trace().build(), 250, 250, True),
         ("score-only nw 250x250", ps.Aligner.new().matrix(b62).gap_open(10).gap_extend(1).build(), 250, 250, True),
         ("C3 sg_stats 150x500", ps.Aligner.new().semi_global().matrix(ps.Matrix.create(b"ACGT", 2, -3)).gap_open(5).gap_extend(2).use_stats().build(), 150, 500, False)]
for name, a, lq, lr, prot in cases:
    qc = psb_gen.random(11, 3, n * lq, prot); rc = psb_gen.random(12, 3, n * lr, prot)
    keep = [pinned(x) for x in (qc, offs(n, lq), rc, offs(n, lr))]
    args = ((keep[0].numpy(), keep[1].numpy()), (keep[2].numpy(), keep[3].numpy()))
    os.environ.pop("PSB_DEBUG_TIMING", None)
    for lanes in (2, 2):
        os.environ["PSB_PAIRS_LANES"] = str(lanes)
        for _ in range(3): a.align_batch(*args)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); r = a.align_batch(*args); ts.append(time.perf_counter() - t0)
        print(f"{name}: {n} pairs, {lanes} lanes: e2e min {min(ts) * 1e3:.2f} median {np.median(ts) * 1e3:.2f} ms, {r.cells / np.median(ts) / 1e9:.0f} GCUPS", flush=True)
    os.environ["PSB_PAIRS_LANES"] = "1"
    a.align_batch(*args); t0 = time.perf_counter(); a.align_batch(*args); t1 = time.perf_counter() - t0
    print(f"   one lane: e2e {t1 * 1e3:.2f} ms, kernels {ps.kernel_ms():.2f} ms", flush=True)
    del os.environ["PSB_PAIRS_LANES"]
    os.environ["PSB_DEBUG_TIMING"] = "1"
    sys.stderr.write(f"---- {name}\n"); sys.stderr.flush()
    a.align_batch(*args)
    sys.stderr.flush()
