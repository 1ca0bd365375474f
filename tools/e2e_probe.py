#!/usr/bin/env python
"""Sweep psb_scan_host's piece sizes on the C2 workload (host buffers in, host results out).

usage: python tools/e2e_probe.py [first_mb:piece_mb ...]       (default: a small grid)
Prints one line per setting with the median and best end-to-end GCUPS over a few calls and checks
that every setting returns the same scores as the resident-database scan.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import parasail_rs_b200 as ps
    settings = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [
        (192, 192), (48, 160), (32, 176), (24, 120), (16, 96), (64, 160), (32, 96), (96, 96)]
    query, cat, off = bench.make_inputs(1_000_000)
    cells = float(len(query)) * float(off[-1])
    pin_cat = torch.empty(len(cat), dtype=torch.uint8, pin_memory=True)
    pin_cat.numpy()[:] = cat
    pin_off = torch.empty(len(off), dtype=torch.int64, pin_memory=True)
    pin_off.numpy()[:] = off
    blosum = ps.Matrix.from_name("blosum62")

    def step():
        p = ps.Profile.new(query, False, blosum)
        a = ps.Aligner.new().local().gap_open(bench.OPEN).gap_extend(bench.GAP).profile(p).build()
        return a.scan_host((pin_cat.numpy(), pin_off.numpy()))

    prof = ps.Profile.new(query, False, blosum)
    al = ps.Aligner.new().local().gap_open(bench.OPEN).gap_extend(bench.GAP).profile(prof).build()
    db = ps.Database((pin_cat.numpy(), pin_off.numpy()), blosum)
    want = al.scan(db)
    want = (want.score.copy(), want.end_query.copy(), want.end_ref.copy())
    # resident scan of 1/8 and 1/4 shards (what each GPU of an 8- or 4-GPU run sees)
    for world in (8, 4):
        from parasail_rs_b200 import sharding
        shard = ps.shard_plan(off, world)
        c8, o8, _ = sharding.local_shard(cat, off, shard, 0)
        d8 = ps.Database((c8, o8), blosum)
        for _ in range(3):
            al.scan(d8)
        ts = []
        for _ in range(5):
            al.scan(d8)
            ts.append(ps.kernel_ms())
        print(f"resident 1/{world} shard: kernels {min(ts):.3f} ms = {len(query) * float(o8[-1]) / min(ts) / 1e6:.0f} GCUPS per GPU", flush=True)
        del d8
    for first, piece in settings:
        os.environ["PSB_SCAN_HOST_FIRST_MB"] = str(first)
        os.environ["PSB_SCAN_HOST_PIECE_MB"] = str(piece)
        for _ in range(3):
            r = step()
        ok = (np.array_equal(r.score, want[0]) and np.array_equal(r.end_query, want[1])
              and np.array_equal(r.end_ref, want[2]))
        ts = []
        for _ in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        print(f"first {first:4d} MB piece {piece:4d} MB: median {cells / ts[len(ts) // 2] / 1e9:7.0f} "
              f"best {cells / ts[0] / 1e9:7.0f} GCUPS  ({ts[len(ts) // 2] * 1e3:.2f} ms, kernels "
              f"{ps.kernel_ms():.2f} ms)  equal={ok}", flush=True)


if __name__ == "__main__":
    main()
