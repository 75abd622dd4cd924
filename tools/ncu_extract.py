#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` (stdin) into the metric,unit,value extract kept
under profiles/ (first profiled launch; the metrics DESIGN.md / bench.py quote).  `--all` prints one
block per profiled launch (helper-kernel captures hold several kernels)."""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct")


def main():
    rows = list(csv.reader(line for line in sys.stdin if line.startswith('"')))
    if len(rows) < 3:
        sys.exit("no raw page on stdin")
    names, units = rows[0], rows[1]
    print("metric,unit,value")
    for first in (rows[2:] if "--all" in sys.argv else rows[2:3]):
        for i, name in enumerate(names):
            if name == "Kernel Name":
                print(f'Kernel Name,,"{first[i]}"')
        for want in KEEP:
            if want in names:
                i = names.index(want)
                print(f"{want},{units[i]},{first[i].replace(',', '')}")


if __name__ == "__main__":
    main()
