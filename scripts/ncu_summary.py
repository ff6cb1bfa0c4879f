#!/usr/bin/env python
"""Prints the metrics DESIGN.md quotes from an ncu report: `python scripts/ncu_summary.py report.ncu-rep` (runs
`ncu -i report --page raw --csv` and picks duration, DRAM traffic, throughputs, issue utilisation, occupancy and the
stall reasons per launch)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__occupancy_limit_registers", "blocks/SM by registers"),
    ("launch__occupancy_limit_shared_mem", "blocks/SM by shared memory"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no instruction / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle / issue"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== launch {r[ix['ID']]}: {r[ix['Kernel Name']]}  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}")
        for key, label in KEYS:
            if key in ix:
                print(f"   {label:34s} {r[ix[key]]} {units[ix[key]]}")


if __name__ == "__main__":
    main()
