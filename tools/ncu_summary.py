#!/usr/bin/env python
"""Text summary of an `ncu -i X.ncu-rep --page raw --csv` export: per captured launch the duration,
DRAM bytes, launch shape, issue / FP64 / DRAM utilisation, shared-memory conflicts and the warp stall
reasons per issued instruction (the numbers profiles/README.md and DESIGN.md quote).
Usage: ncu_summary.py raw.csv "title" "command" > profiles/NAME.txt"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpc__cycles_elapsed.avg.per_second"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    if len(sys.argv) > 2:
        print(sys.argv[2])
    if len(sys.argv) > 3:
        print(sys.argv[3])
    for r in rows[2:]:
        print("---")
        print("Kernel Name:", r[ix["Kernel Name"]])
        for k in KEEP:
            if k in ix:
                print(f"{k}: {r[ix[k]]} {units[ix[k]]}")
        for n in hdr:
            if n.startswith(STALL) and n.endswith("_per_issue_active.ratio"):
                print(f"stall_{n[len(STALL):-len('_per_issue_active.ratio')]}: {float(r[ix[n]]):.6f}")


if __name__ == "__main__":
    main()
