#!/usr/bin/env python
"""Key per-launch metrics of an ncu report as a small table (run where ncu is installed).

    python tools/ncu_summary.py <report.ncu-rep> [more reports...]
"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "lim_regs"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"),
        ("sm__inst_executed.avg.per_cycle_active", "ipc"), ("smsp__inst_executed.sum", "warp_inst"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests")]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(rep)
        for r in rows[2:]:
            name = r[col["Kernel Name"]][:60]
            grid, block = r[col["Grid Size"]], r[col["Block Size"]]
            print(f"  {name}  grid {grid} block {block}")
            for k, short in KEYS:
                if k in col:
                    print(f"      {short:10s} {r[col[k]]} {units[col[k]]}")


if __name__ == "__main__":
    main()
