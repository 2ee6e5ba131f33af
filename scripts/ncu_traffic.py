#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report: measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
per bench.py stage and per launch of that stage (a stage = the kernels bench.py's gsb_profile_* timers bracket).

usage: python scripts/ncu_traffic.py gpurun_out/<tag>_prof.ncu-rep profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

STAGE_OF = [  # (kernel-name prefix, grid class or None, stage, launches of this kernel per stage launch)
    ("preprocess_kernel", None, "preprocess", 1), ("radix_histogram_kernel", "P", "depth_sort", 1), ("radix_pass_kernel", "P", "depth_sort", 4),
    ("scan_tiles_kernel", None, "emit", 1), ("emit_sorted_kernel", None, "emit", 1), ("init_ranges_kernel", None, "emit", 1),
    ("radix_pass_kernel", "R", "tile_sort", 2), ("render_table_kernel", None, "render", 1), ("render_compact_kernel", None, "render", 1),
    ("to_u8_kernel", None, "to_u8", 1), ("prepare_depth_kernel", None, "prepare_depth", 1), ("mark_bricks_kernel", None, "mark_bricks", 1),
    ("integrate_kernel", None, "integrate", 1),
]


def main():
    rep, out_path = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    kname, grid = head.index("Kernel Name"), head.index("Grid Size")
    rd = next(i for i, h in enumerate(head) if h.endswith("dram__bytes_read.sum"))
    wr = next(i for i, h in enumerate(head) if h.endswith("dram__bytes_write.sum"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = defaultdict(list)
    for r in body:
        name = r[kname].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        g = int(r[grid].strip("()").split(",")[0])
        cls = None
        if name.startswith("radix_"):
            cls = "R" if g > 600 else "P"
        per_kernel[(name, cls)].append(float(r[rd].replace(",", "")) * scale.get(units[rd], 1) + float(r[wr].replace(",", "")) * scale.get(units[wr], 1))
    traffic = defaultdict(float)
    used = {}
    for prefix, cls, stage, mult in STAGE_OF:
        for (name, c), vals in per_kernel.items():
            if name.startswith(prefix) and c == cls:
                traffic[stage] += mult * sum(vals) / len(vals)
                used[f"{name}[{c}]" if c else name] = [round(sum(vals) / len(vals)), len(vals)]
    result = {k: int(v) for k, v in traffic.items()}
    result["_source"] = f"{rep}: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full), summed over the kernels of each stage"
    result["_kernels"] = used
    with open(out_path, "w") as f:
        json.dump(result, f, indent=1)
    print(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
