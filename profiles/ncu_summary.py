#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls] [--source N]"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__cycles_elapsed.max",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_lsu.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("==", d.get("Kernel Name", "?")[:80])
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                print(f"  {h:75s} {v:>18s} {u}")
            elif "--stalls" in sys.argv and "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    if float(v.replace(",", "")) >= 0.1:
                        print(f"  {h:75s} {v:>18s}")
                except ValueError:
                    pass


if __name__ == "__main__":
    main()
