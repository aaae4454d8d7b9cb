#!/usr/bin/env python3
"""Issue-slot / stall digest per kernel from `ncu -i X.ncu-rep --page raw --csv` of a `--set full` capture."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]


def g(r, k):
    return r[hdr.index(k)] if k in hdr else "NA"


stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
seen = set()
for r in data:
    name = g(r, "Kernel Name").split("(")[0][:44]
    key = (name, g(r, "Grid Size"))
    if key in seen:
        continue
    seen.add(key)
    print("==", name, "grid", g(r, "Grid Size"), "block", g(r, "Block Size"), "| %s us" % g(r, "gpu__time_duration.sum"),
          "| regs", g(r, "launch__registers_per_thread"))
    print("   ipc/SM %s  issue slots busy %s %%  warps resident per sub-partition %s  eligible %s" % (
        g(r, "sm__inst_executed.avg.per_cycle_active"), g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g(r, "smsp__warps_active.avg.per_cycle_active"), g(r, "smsp__warps_eligible.avg.per_cycle_active")))
    st = sorted(((float(g(r, h).replace(",", "")) if g(r, h) not in ("", "NA") else 0.0,
                  h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                 for h in stall), reverse=True)[:6]
    print("   stall cycles per issue:", ", ".join("%s %.2f" % (n, v) for v, n in st))
