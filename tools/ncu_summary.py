#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the two small files that are committed under profiles/:
   <out>_keymetrics.csv   one row per key metric (time, pipes, issue, occupancy, DRAM bytes, instruction counts)
   <out>_stalls.txt       warp-state stall reasons per issued instruction
usage: python tools/ncu_summary.py gpurun_out/r2d/ncu_p16_C3.ncu-rep profiles/r2_ncu_pairs16_C3"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    # several launches in one report: the longest one is the kernel of interest (the others are its small companions)
    it = hdr.index("gpu__time_duration.sum")
    def dur(r):
        try:
            return float(r[it].replace(",", "")) * (1000.0 if False else 1.0)
        except (ValueError, IndexError):
            return -1.0
    iu = units[it]
    vals = max(rows[2:], key=lambda r: float(r[hdr.index("sm__cycles_elapsed.max")].replace(",", "")) if len(r) > it and r[it] else -1)
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    with open(out + "_keymetrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for k in KEYS:
            if k in d:
                w.writerow([k, d[k][1], d[k][0]])
    stalls = sorted(((float(v.replace(",", "")), k) for k, (v, u) in d.items()
                     if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
    with open(out + "_stalls.txt", "w") as f:
        f.write(f"{d['Kernel Name'][0]}\nsource: {rep} (ncu --set full --clock-control none --import-source on)\n")
        f.write("warps stalled per issued instruction, by reason:\n")
        for v, k in stalls[:12]:
            f.write(f"  {k.split('stalled_')[1].split('_per')[0]:28s} {v:6.2f}\n")
    print(out, d["Kernel Name"][0], d["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()
