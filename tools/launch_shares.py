"""Developer helper: sum an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.
usage: python tools/launch_shares.py launches.csv passes > shares.csv"""
import csv
import sys
from collections import OrderedDict

passes = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
acc = OrderedDict()
for r in rows:
    name, unit, val = r[4][:56], r[13], float(r[14].replace(",", ""))
    us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}[unit]
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in acc.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches_per_pass", "us_per_pass", "share_pct"])
for name, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    w.writerow([name, round(n / passes, 1), round(us / passes, 1), round(100 * us / tot, 1)])
