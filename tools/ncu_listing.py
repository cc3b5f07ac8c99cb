"""Developer helper: condensed per-instruction listing (executed counts, stall samples) from
`ncu -i X.ncu-rep --page source --csv`.  usage: python tools/ncu_listing.py src.csv [kernel substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    hdr = b["rows"][0]
    iS, iE, iSm, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    body = [r for r in b["rows"][1:] if len(r) > iT]
    tot = sum(int(r[iE]) for r in body)
    tots = sum(int(r[iSm]) for r in body)
    print("==", b["name"][:90], "total warp inst", tot, "samples", tots, "sass lines", len(body))
    cum = 0
    for n, r in enumerate(body):
        e = int(r[iE])
        if e == 0:
            continue
        cum += e
        print("%4d %-64s exec=%8d (%4.1f%%) thr=%5s samp=%5s (%4.1f%%) cum=%5.1f%%" %
              (n, r[iS].strip()[:64], e, 100 * e / tot, r[iT], r[iSm], 100 * int(r[iSm]) / max(tots, 1), 100 * cum / tot))
