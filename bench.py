#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its config C2:

  Smith-Waterman local with a reused Profile: one 400-aa query vs a synthetic 1M-protein database
  (UniProt-like log-normal lengths), BLOSUM62, open 10 / extend 1, whole-box GCUPS at N B200.

A "step" is one pass of the hot path (psb_scan: packed 16-bit DPX kernel + 32-bit re-runs +
results D2H) over the rank's shard of the database.  The database is sharded by residue count
across ranks with no data-path collective; the total work is fixed as N grows.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Rank 0 prints ONE JSON line.  `--impl reference` times the CPU baseline instead (the
parasail-equivalent striped AVX2 restatement under oracle/, all host threads, bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import psb_data  # noqa: E402

QUERY_SEED, LEN_SEED, RES_SEED = 2001, 2002, 2003
QUERY_LEN = 400
OPEN, GAP = 10, 1
METRIC = "GCUPS (whole box) at 1/2/4/8 B200 vs parasail CPU on host cores"
OPS_PER_CELL = 5          # algorithmic integer lane-ops per Gotoh cell
LANE_OPS_PER_CLK_SM = 64  # ALU-pipe issue rate (B300_MICROARCH.md:85), checked by tools/dpx_bench


def workload_name(n_db):
    return (f"C2: sw_striped_profile_sat, one {QUERY_LEN}-aa query (reused Profile) vs {n_db} synthetic proteins "
            f"(log-normal lengths, seeds {QUERY_SEED}/{LEN_SEED}/{RES_SEED}), BLOSUM62, open {OPEN} ext {GAP}")


def make_inputs(n_db):
    query = psb_data.random_seq(QUERY_SEED, 0, QUERY_LEN)
    cat, off = psb_data.protein_db(LEN_SEED, RES_SEED, n_db, query=query, planted_frac=0.01)
    return query, cat, off


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(query, cat, off, n_sample, threads=0):
    """the striped AVX2 port on the host cores, bounded sample of the same workload"""
    from oracle import oracle as orc
    omat = orc.Matrix.from_table(psb_data.BLOSUM62_ALPHABET, psb_data.blosum62_table())
    n_sample = min(n_sample, len(off) - 1)
    sub_off = off[: n_sample + 1]
    res, secs = orc.striped_sw_scan(query, cat, sub_off, omat, OPEN, GAP, threads=threads)
    cells = float(len(query)) * float(sub_off[-1] - sub_off[0])
    return res, secs, cells


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (parasail is not
    installable here; the striped AVX2 restatement stands in, see oracle/striped_cpu.cpp)."""
    if rank != 0:
        return
    n_sample = min(args.db, 300000)
    query, cat, off = make_inputs(n_sample)
    for _ in range(args.warmup):
        cpu_baseline(query, cat, off, n_sample)
    t0 = time.perf_counter()
    cells = 0.0
    for _ in range(args.steps):
        res, secs, c = cpu_baseline(query, cat, off, n_sample)
        cells += c
    dt = time.perf_counter() - t0
    value = cells / dt / 1e9
    threads = int(res["threads"])
    sample = f"{n_sample} subjects of the C2 database per step ({cells / args.steps:.3g} cells), striped AVX2 8->16->32 bit"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "s8/s16 (sat escalation)", "data": "synthetic",
        "config": {"workload": workload_name(args.db), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_JSON_FD = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print banners there too (NCCL's version line when
    NCCL_DEBUG is set on the box), so everything but the result line is sent to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, line)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--db", type=int, default=1000000, help="number of database proteins (C2: 1M)")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-sample", type=int, default=200000, help="subjects of the CPU-baseline sample (rank 0, N=1)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        # a host-side group for the phase in which rank 0 drives every GPU itself: an NCCL barrier would keep a
        # spinning kernel of the idle ranks' processes on those GPUs, time-sliced against the measured work
        cpu_group = dist.new_group(backend="gloo")
    torch.cuda.set_device(local_rank)
    import parasail_rs_b200 as ps
    from parasail_rs_b200 import _lib
    L = _lib.lib()
    if L.psb_set_device(local_rank) != 0:
        raise SystemExit("psb_set_device: " + ps.last_error())
    stream = torch.cuda.current_stream()
    L.psb_set_stream(stream.cuda_stream)

    # ---- inputs: the full database is a pure function of the seeds; each rank keeps its shard ----
    query, cat, off = make_inputs(args.db)
    lens = np.diff(off)
    total_cells = float(QUERY_LEN) * float(off[-1])
    if world > 1:
        from parasail_rs_b200 import sharding
        shard = ps.shard_plan(off, world)
        my_cat, my_off, mine = sharding.local_shard(cat, off, shard, rank)
    else:
        mine, my_cat, my_off = np.arange(len(lens)), cat, off
    my_cells = float(QUERY_LEN) * float(my_off[-1])
    # pinned host copies: the e2e leg copies from these every step
    pin_cat = torch.empty(len(my_cat), dtype=torch.uint8, pin_memory=True)
    pin_cat.numpy()[:] = my_cat
    pin_off = torch.empty(len(my_off), dtype=torch.int64, pin_memory=True)
    pin_off.numpy()[:] = my_off

    blosum = ps.Matrix.from_name("blosum62")
    profile = ps.Profile.new(query, False, blosum)
    aligner = ps.Aligner.new().local().gap_open(OPEN).gap_extend(GAP).profile(profile).build()
    assert aligner.fn_name == "sw_striped_profile_sat"
    db = ps.Database((pin_cat.numpy(), pin_off.numpy()), blosum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------------------------
    # SURVEY 8d protocol: database resident; query profile upload + kernels + results D2H inside the timed
    # region, so every step builds a fresh Profile (host profile + H2D of query, matrix and packed profile)
    def resident_step():
        p = ps.Profile.new(query, False, blosum)
        a = ps.Aligner.new().local().gap_open(OPEN).gap_extend(GAP).profile(p).build()
        return a.scan(db)
    res = None
    for _ in range(args.warmup):
        res = resident_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = 0.0, 0
    ev0.record(stream)
    for _ in range(args.steps):
        res = resident_step()
        kernel_ms += ps.kernel_ms()
        launches += ps.launches()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    kms = torch.tensor([kernel_ms / args.steps], dtype=torch.float64, device="cuda")
    nl = torch.tensor([launches], dtype=torch.int64, device="cuda")
    nretry = torch.tensor([res.n_retried], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
        dist.all_reduce(nretry, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    value = total_cells * args.steps / (total_ms * 1e-3) / 1e9
    scan_scores = res.score.copy()
    scan_eq, scan_er = res.end_query.copy(), res.end_ref.copy()

    # ---- end to end: host buffers in, host results out, every step ---------------------------------
    # N = 1: one C call with HOST buffers, psb_scan_host (piecewise upload from pinned memory on a copy
    # stream, device-side sort + packing, scan kernels, results D2H, all pipelined).
    # N > 1: ONE process drives the whole box through psb_scan_box: rank 0 hands the full host database to
    # the library, which cuts it into N residue-balanced ranges, runs one worker thread per device and
    # returns ONE caller-order batch; the other ranks idle at the barrier meanwhile.
    e2e_res = None
    if world == 1:
        e2e_cat, e2e_off = pin_cat.numpy(), pin_off.numpy()
    elif rank == 0:
        full_cat = torch.empty(len(cat), dtype=torch.uint8, pin_memory=True)
        full_cat.numpy()[:] = cat
        full_off = torch.empty(len(off), dtype=torch.int64, pin_memory=True)
        full_off.numpy()[:] = off
        e2e_cat, e2e_off = full_cat.numpy(), full_off.numpy()

    def e2e_step():
        p = ps.Profile.new(query, False, blosum)
        a = ps.Aligner.new().local().gap_open(OPEN).gap_extend(GAP).profile(p).build()
        return a.scan_host((e2e_cat, e2e_off)) if world == 1 else a.scan_box((e2e_cat, e2e_off), world)
    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)
    host_barrier()
    if rank == 0:
        for _ in range(6):   # warm-up: contexts of the worker threads, pool growth, staging and pinned result blocks, measured rates of the piece plan
            e2e_res = e2e_step()
    host_barrier()
    e2e_wall = 0.0
    if rank == 0:
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_res = e2e_step()
        e2e_wall = (time.perf_counter() - t0) * 1e3
    host_barrier()
    e2e_ms = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = total_cells * args.e2e_steps / (float(e2e_ms.item()) * 1e-3) / 1e9
    h2d = int(len(cat) + 8 * len(off) + QUERY_LEN + 26 * 512 * 4 * world)
    d2h = int(12 * len(lens) + 4 * world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sw16_scan_kernel): integer-ALU / DPX cell updates ---------
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    peak_s16 = 148 * LANE_OPS_PER_CLK_SM * (sm_mhz * 1e-3) / OPS_PER_CELL * 2  # GCUPS per GPU at the sampled clock
    achieved = my_cells / (float(kms.item()) * 1e-3) / 1e9
    packed_bytes = float(my_off[-1]) * 5 / 8 + 12.0 * len(mine)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "int_alu_dpx", "achieved": achieved, "peak": peak_s16, "unit": "GCUPS", "frac": achieved / peak_s16,
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch on a 200k-subject
        # scan (profiles/r2j_ncu_sw16_C2_keymetrics.csv: 51.76 MB + 0.06 MB for 47.6 MB of packed
        # residues + 2.4 MB of results, which leave L2 later), scaled to this launch's algorithmic bytes
        "traffic": packed_bytes * (51.763968e6 + 0.058112e6) / (47.589941e6 + 2.4e6),
        "traffic_source": "ncu --set full capture profiles/r2j_ncu_sw16_C2 (200k subjects, this build), scaled by algorithmic bytes; not measured in this run",
        "kernel": "sw16_scan_kernel<25> (packed s16x2 DPX, 16-lane groups)", "kernel_ms_per_launch": float(kms.item()),
        "peak_basis": f"148 SM x {LANE_OPS_PER_CLK_SM} lane-ops/clk x {sm_mhz:.0f} MHz (sampled) / {OPS_PER_CELL} ops per cell x 2 cells per s16x2 op",
        "peak_at_max_clock": 148 * LANE_OPS_PER_CLK_SM * 1.965 / OPS_PER_CELL * 2,
        "hbm": {"algorithmic_bytes_per_launch": packed_bytes, "achieved_gbs": packed_bytes / (float(kms.item()) * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- every subject against the cached CPU result (tests/golden/c2_block_hashes.json) -----------------
    verified_all = None
    gpath = os.path.join(ROOT, "tests", "golden", "c2_block_hashes.json" if args.db == 1000000 else f"c2_block_hashes_{args.db}.json")
    if os.path.exists(gpath):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from make_c2_golden import block_hashes
        gold = json.load(open(gpath))
        checks = [("end-to-end call", e2e_res.score, e2e_res.end_query, e2e_res.end_ref)]
        if world == 1:
            checks.append(("resident scan", scan_scores, scan_eq, scan_er))
        for what, sc, eq, er in checks:
            mine_h = block_hashes(sc, eq, er, gold["block"])
            bad = [i for i, (a, b2) in enumerate(zip(mine_h, gold["hashes"])) if a != b2]
            if bad or len(mine_h) != len(gold["hashes"]):
                raise SystemExit(f"{what}: GPU results disagree with the cached CPU result in {len(bad)} blocks of {gold['block']} subjects (first: block {bad[:1]}) -- refusing to report a number")
        verified_all = (f"all {args.db} subjects of {' and '.join(c[0] for c in checks)} equal to the cached CPU result "
                        f"({len(gold['hashes'])} block hashes, {gold['generated_by']})")

    # ---- CPU baseline beside it (rank 0, N = 1 only) + cross-check of the GPU results ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cres, secs, ccells = cpu_baseline(query, cat, off, args.cpu_sample)
        ns = len(cres["score"])
        same = (np.array_equal(cres["score"], scan_scores[:ns]) and np.array_equal(cres["end_query"], scan_eq[:ns])
                and np.array_equal(cres["end_ref"], scan_er[:ns]))
        if not same:
            raise SystemExit("GPU scan disagrees with the CPU baseline on the sample -- refusing to report a number")
        cpu = {"value": ccells / secs / 1e9, "unit": "GCUPS", "cores": int(cres["threads"]), "kind": "port",
               "sample": f"first {ns} subjects of the same database ({ccells:.3g} cells, {secs:.1f} s), "
                         "parasail-equivalent striped AVX2 8->16->32 bit restatement; GPU results on the sample verified equal"}

    out = {
        "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "s16x2 (s32 re-run on overflow)", "data": "synthetic",
        "config": {"workload": workload_name(args.db), "sharding": f"residue-balanced over {world} GPU(s), no collective",
                   "cells_per_step": total_cells, "l2": "packed database (5 bit/residue) exceeds L2; not flushed between steps"
                   if float(off[-1]) * 5 / 8 / world > 126e6 else "shard fits L2; HBM traffic is not the bound (0.002 B/cell)",
                   "subjects_rerun_at_32bit": int(nretry.item()), "verified": verified_all,
                   "timed_region": "per step: Profile::new (host profile + H2D), scan kernels, results D2H; database resident"},
        "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps,
                "what": ("profile create + psb_scan_host (piecewise H2D from pinned host memory overlapped with device packing and the scan kernels) + results D2H"
                         if world == 1 else
                         f"ONE process, one call: profile create + psb_scan_box over {world} GPUs (contiguous residue-balanced ranges, one worker thread per device, "
                         "piecewise H2D from the caller's pinned arrays, device packing, scan kernels, caller-order results D2H into one batch); wall clock on rank 0")},
        "gpu_launches": int(nl.item()), "clocks": clocks, "roofline": roofline,
    }
    if cpu:
        out["cpu_baseline"] = cpu
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
