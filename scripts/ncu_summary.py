#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-kernel table kept under profiles/.

usage: python scripts/ncu_summary.py gpurun_out/<tag>_prof.ncu-rep > profiles/<tag>_ncu_full_summary.txt
Means over the captured launches of each kernel (radix passes are split by grid size: P-sized vs R-sized)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for m in METRICS:
        for i, h in enumerate(head):
            if h == m or h.endswith("." + m):
                col[m] = i
                break
    kname = head.index("Kernel Name")
    grid = head.index("Grid Size")
    acc = defaultdict(lambda: defaultdict(list))
    for r in body:
        name = r[kname].split("(")[0].split("::")[-1]
        if name.startswith("radix_pass") or name.startswith("radix_hist"):
            g = int(r[grid].strip("()").split(",")[0])
            name += "[R]" if g > 600 else "[P]"
        for m, i in col.items():
            try:
                acc[name][m].append(float(r[i].replace(",", "")))
            except ValueError:
                pass
    print(f"# {rep}: ncu --set full --clock-control none --import-source on; means over the captured launches")
    print("# units: " + ", ".join(f"{m}={units[i] or '-'}" for m, i in col.items()))
    print()
    for name, ms in acc.items():
        n = max(len(v) for v in ms.values())
        short = lambda m: m.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "").split(".")[0]
        print(f"{name}: n={n} " + " ".join(f"{short(m)}={sum(v) / len(v):.4g}" for m, v in ms.items() if v))


if __name__ == "__main__":
    main()
