"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv ...`).
usage: python tools/launch_list.py x.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r is hdr or len(r) <= max(ki, vi) or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[ui], 1e-6)
    name = r[ki].strip()
    name = name[:name.rfind(">") + 1] if ">" in name else re.sub(r"\(.*", "", name)  # drop the parameter list, keep template args
    name = re.sub(r"\(swalbe::\w+\)", "", name)
    tot[name] += v
    cnt[name] += 1
allms = sum(tot.values())
print(f"{sum(cnt.values())} launches, {allms:.3f} ms of kernel time (cold caches, serialised by the profiler)")
for name in sorted(tot, key=tot.get, reverse=True):
    print(f"{100 * tot[name] / allms:6.2f} %  {cnt[name]:5d} x  {tot[name] / cnt[name]:9.4f} ms  {name[:150]}")
