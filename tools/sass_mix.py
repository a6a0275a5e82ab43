#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter()
tot = 0; tots = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    op = op.split(".")[0] + ("." + src.split()[0].split(".")[1] if "." in src.split()[0] and op in ("LDG","STG","LDS","STS") else "")
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ex[op] += n; st[op] += s; tot += n; tots += s
print("total warp instructions %d, samples %d" % (tot, tots))
for op, n in ex.most_common(28):
    print("%-14s %12d %5.1f%%   stall samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * st[op] / max(tots, 1)))
