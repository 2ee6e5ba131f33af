"""Static facts about the shipped kernels, no GPU needed: registers / spills / shared memory per kernel (ptxas -v) and the
SASS instruction mix of each kernel in gs2mesh_b200/libgs2mesh_b200.so (cuobjdump -sass).

    python scripts/static_report.py > profiles/<tag>_static_report.txt

What to look for (B200_PROFILING.md): UBLKCP + SYNCS = 1-D bulk TMA + mbarrier (preprocess), FFMA2 / FADD2 / FMUL2 =
packed f32x2 arithmetic (table blend), MUFU.EX2, MATCH = warp match (radix ranking, brick de-duplication), STL / LDL =
register spills.  No UTCxMMA / UTMALDG is expected: nothing on this path is a contraction or a multi-dimensional tile.
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs2mesh_b200 import build as gsb_build  # noqa: E402

KEY = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMUL", "FADD", "FSETP", "FSEL", "FMNMX", "MUFU.EX2", "MUFU.LG2",
       "MUFU.RCP", "MUFU.RSQ", "MUFU.SQRT", "MATCH", "VOTE", "SHFL", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "LDL",
       "STL", "BAR", "DADD", "DMUL", "DFMA"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    out = [re.sub(r"\((?:int|bool|unsigned int)\)", "", o).replace("<unnamed>::", "").replace("gsb::", "") for o in out]
    out = [re.sub(r"cub::CUB_\w+::", "cub::", o) for o in out]
    return [re.sub(r"^void ", "", re.sub(r"\(.*", "", o))[:100] for o in out]


def ptxas_table():
    with tempfile.TemporaryDirectory() as tmp:
        cmd = ["nvcc"] + gsb_build.NVCC_FLAGS + ["-Xptxas", "-v", "-o", os.path.join(tmp, "lib.so")] + gsb_build._sources()
        err = subprocess.run(cmd, capture_output=True, text=True).stderr
    rows, cur = [], None
    for line in err.split("\n"):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
            continue
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
        if m and cur:
            rows.append((cur, int(m.group(1)), int(m.group(3) or 0), spill))
            cur = None
    names = demangle([r[0] for r in rows])
    print("== ptxas -v (sm_100a): registers / static shared memory / stack, spill stores, spill loads (bytes)")
    for (raw, regs, smem, spill), name in sorted(zip(rows, names), key=lambda t: t[1]):
        print(f"{name:<72s} regs {regs:3d}  smem {smem:6d}  stack {spill[0]:4d}  spill st {spill[1]:4d} ld {spill[2]:4d}")


def sass_mix():
    sass = subprocess.run(["cuobjdump", "-sass", gsb_build.LIB_PATH], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEY:
                if op == k or op.startswith(k + "."):
                    cur[k] += 1
                    break
    names = demangle(list(funcs))
    print("\n== SASS instruction mix per kernel (static counts; cuobjdump -sass)")
    for (raw, c), name in sorted(zip(funcs.items(), names), key=lambda t: t[1]):
        mix = "  ".join(f"{k} {c[k]}" for k in KEY if c[k])
        print(f"{name}\n    total {c['_total']}  {mix}")
    everything = collections.Counter()
    for c in funcs.values():
        everything.update(c)
    print("\n== whole library")
    print("    " + "  ".join(f"{k} {everything[k]}" for k in KEY if everything[k]))
    for absent in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "HMMA", "WGMMA"):
        print(f"    {absent}: {len(re.findall(absent, sass))}")


if __name__ == "__main__":
    ptxas_table()
    sass_mix()
