#!/usr/bin/env python
"""DRAM bytes (read + written) per C-ABI call from an `ncu --set full --page raw --csv` export of ONE training step of a
one-layer network (tools/profile.sh): kernels are grouped into the calls bench.py times, by name and launch order.
Usage: tools/traffic_from_ncu.py gpurun_out/prof_<tag>_raw.csv <batch> > profiles/<tag>_traffic.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
batch = int(sys.argv[2])
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def b(r, n):
    return float(r[ix[n]].replace(",", "")) * U.get(units[ix[n]], 1)


calls = {}
seen = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"^void ", "", r[ix["Kernel Name"]]).split("(")[0].replace("nbss::", "").split("<")[0]
    k = seen[name] = seen.get(name, 0) + 1
    tot = b(r, "dram__bytes_read.sum") + b(r, "dram__bytes_write.sum")
    if name == "wgrad_kernel":
        key = "ffn_wgrad" if k <= 3 else "mhsa_wgrad"
    elif name in ("mhsa_bwd_core_kernel", "mhsa_bwd_ln_kernel"):
        key = "mhsa_bwd"
    elif name in ("squeeze_fwd_tc_kernel", "unsqueeze_fwd_tc_kernel") or (name == "lg_tc_kernel" and k == 1):
        key = "full_fwd_tc"
    elif name in ("squeeze_bwd_tc_kernel", "unsqueeze_bwd_tc_kernel", "lg_wgrad_kernel") or name == "lg_tc_kernel":
        key = "full_bwd_tc"
    elif name.startswith("fconv_tc_fwd"):
        key, tot = "fconv_tc_fwd", tot / 2  # two launches per layer: average
    elif name.startswith("fconv_tc_bwd"):
        key, tot = "fconv_tc_bwd", tot / 2
    else:
        key = name.replace("_kernel", "").replace("_fft", "")  # stft_fft_kernel -> stft (the C-ABI call name)
    calls[key] = calls.get(key, 0.0) + tot
json.dump({"batch_per_gpu": batch, "source": sys.argv[1], "dram_bytes_per_call": {k: round(v) for k, v in calls.items()}}, sys.stdout, indent=1)
print()
