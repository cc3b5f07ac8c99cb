"""ncu launch list (gpu__time_duration.sum CSV) -> per-kernel totals and shares.  usage: python tools/launch_shares2.py in.csv out.csv"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(rows):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("cloops::", "")
    name = re.sub(r"<.*", "", name) if name.startswith(("cub::", "at::", "thrust")) else name
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000.0 if row["Metric Unit"] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
tot = sum(v[0] for v in agg.values())
with open(sys.argv[2], "w") as fh:
    fh.write("kernel,launches,total_us,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        fh.write("%s,%d,%.1f,%.4f\n" % (k, v[1], v[0], v[0] / tot))
print("total %.1f ms over %d launches" % (tot / 1e3, sum(v[1] for v in agg.values())))
