#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.
usage: ncu_summary.py report.ncu-rep "title / command" > profiles/xyz.md"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid (CTAs)"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "SM issue-slot utilisation %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (I-cache) / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (L1TEX/global) / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency) / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch / issue"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (warp inst)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores (warp inst)"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none --import-source on)\n")
    print("| metric | " + " | ".join(f"launch {i}: {r[name_i].split('(')[0]}" for i, r in enumerate(data)) + " |")
    print("|---|" + "---|" * len(data))
    for key, label in KEYS:
        if key in hdr:
            i = hdr.index(key)
            print(f"| {label} (`{key}`) | " + " | ".join(f"{r[i]} {units[i]}" for r in data) + " |")


if __name__ == "__main__":
    main()
