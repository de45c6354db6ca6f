#!/usr/bin/env python
"""Markdown table from an `ncu --page raw --csv` export: duration, issue / warp / pipe activity, DRAM traffic, top stalls.
Usage: tools/ncu_summary.py gpurun_out/prof_<tag>_raw.csv [--json out.json]  (the JSON maps kernel -> DRAM bytes/launch)"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
M = {
    "us": "gpu__time_duration.sum", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "dram": "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
units = rows[1]


def num(r, name):
    try:
        return float(r[ix[name]].replace(",", ""))
    except Exception:
        return float("nan")


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


print("| kernel | us | issue % | warps % | tensor % | xu % | fma % | dram % | rd MB | wr MB | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
traffic = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"^void ", "", r[ix["Kernel Name"]]).split("(")[0].replace("nbss::", "")
    us = to_us(num(r, M["us"]), units[ix[M["us"]]])
    rd = to_bytes(num(r, M["rd"]), units[ix[M["rd"]]])
    wr = to_bytes(num(r, M["wr"]), units[ix[M["wr"]]])
    st = sorted(((num(r, s), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:3]
    print(f"| `{name}` | {us:.1f} | {num(r, M['issue']):.1f} | {num(r, M['warps']):.1f} | {num(r, M['tensor']):.1f} | {num(r, M['xu']):.1f} | "
          f"{num(r, M['fma']):.1f} | {num(r, M['dram']):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | " + ", ".join(f"{n} {v:.1f}" for v, n in st) + " |")
    traffic.setdefault(name, []).append(rd + wr)
if "--json" in sys.argv:
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
