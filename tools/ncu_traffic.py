#!/usr/bin/env python3
"""Condenses `ncu -i X.ncu-rep --page raw --csv` of a `--set full` capture into the few per-launch
numbers the roofline discussion uses, and (with --json) writes the per-launch DRAM traffic record
that bench.py reports as roofline.traffic.

    ncu -i gpurun_out/k_bwd_full.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv --json profiles/r01_ncu_traffic.json --workload "..."
"""
import argparse
import csv
import json

ap = argparse.ArgumentParser()
ap.add_argument("raw_csv")
ap.add_argument("--json")
ap.add_argument("--workload", default="")
ap.add_argument("--command", default="")
args = ap.parse_args()

rows = list(csv.reader(open(args.raw_csv)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def val(row, key):
    i = hdr.index(key)
    return float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0)


launches = []
for r in data:
    rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "grid": r[hdr.index("Grid Size")]}
    for k in KEYS:
        if k in hdr:
            rec[k] = val(r, k)
    rec["dram_bytes"] = rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]
    launches.append(rec)
    print("%-44s grid %-14s %8.1f us  dram %7.1f MB (%.1f %% of peak)  sm %.1f %%  regs %d" % (
        rec["kernel"][:44], rec["grid"], rec["gpu__time_duration.sum"], rec["dram_bytes"] / 1e6,
        rec.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
        rec.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0), rec.get("launch__registers_per_thread", 0)))
avg = sum(l["dram_bytes"] for l in launches) / len(launches)
print("launches %d  mean dram bytes per launch %.1f MB  total %.1f MB" % (len(launches), avg / 1e6, avg * len(launches) / 1e6))
if args.json:
    json.dump({"workload": args.workload, "command": args.command, "kernel": launches[0]["kernel"],
               "launches": len(launches), "dram_bytes_per_launch": avg, "per_launch": launches},
              open(args.json, "w"), indent=1)
