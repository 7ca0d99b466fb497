"""Markdown table of the headline ncu metrics of selected launches of an `ncu --page raw --csv` dump.
Usage: python tools/ncu_summary.py raw.csv idx[,idx...] > summary.md"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    sel = [int(x) for x in sys.argv[2].split(",")]
    for i in sel:
        r = rows[2 + i]
        print(f"## {r[hdr.index('Kernel Name')].split('(')[0]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                j = hdr.index(w)
                print(f"| {w} | {r[j]} | {units[j]} |")
        print()


if __name__ == "__main__":
    main()
