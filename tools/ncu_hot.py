#!/usr/bin/env python
"""Hot spots of an ncu report's SASS page: instructions with the most stall samples, with their main stall reason.
usage: ncu -i rep --page source --csv > src.csv; ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for r in body:
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[ix[c]] or 0)
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    s = int(r[ix["# Samples"]])
    why = max(stall_cols, key=lambda c: int(r[ix[c]] or 0))
    print("%5d %5.1f%% %-14s %s" % (i, 100 * s / tot, why[6:], r[ix["Source"]].strip()[:110]))
