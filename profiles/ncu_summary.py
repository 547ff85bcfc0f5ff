#!/usr/bin/env python
"""Summarise an `ncu -i rep --page raw --csv` dump: one block of key metrics per captured launch."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active", "sm__pipe_fmaheavy",
        "sm__pipe_fmalite", "sm__inst_executed_pipe_alu", "sm__pipe_alu_cycles_active",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput", "lts__t_sector_hit_rate",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", 
        "smsp__average_warps_issue_stalled", "sm__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("=" * 100)
        print(row[hdr.index("Kernel Name")][:120], "grid", row[hdr.index("Grid Size")],
              "block", row[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS) and row[i] not in ("", "0"):
                print(f"  {h:95s} {row[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
