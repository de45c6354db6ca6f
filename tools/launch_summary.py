"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py file.csv"""
import csv, sys, collections, re, signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
s = sum(a[1] for a in tot.values())
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us  {n:4d} x {us / n:8.1f} us  {100 * us / s:5.1f}%  {k}")
print(f"{s:10.1f} us total, {sum(a[0] for a in tot.values())} launches")
