#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list: one line per launch (kernel, grid, block, metrics)."""
import csv, re, sys
from collections import OrderedDict

def load(path):
    rows = OrderedDict()
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr):
            continue
        key = int(r[ix["ID"]])
        d = rows.setdefault(key, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    return rows

def short(name):
    m = re.search(r"fft_lines_kernel<b200fft::Cfg<(\w+), \d+>, (\d), (\d), (\d)>", name)
    if m:
        fl = {"00": "row", "11": "col", "01": "trans"}.get(m.group(2) + m.group(3), "?")
        return "fft_lines<%s,%s%s>" % (m.group(1), fl, "+tw" if m.group(4) == "1" else "")
    return re.sub(r"\(.*", "", name)[:60]

if __name__ == "__main__":
    for p in sys.argv[1:]:
        rows = load(p)
        print("#", p)
        tot = sum(d.get("gpu__time_duration.sum", 0) for d in rows.values())
        for k, d in rows.items():
            t = d.get("gpu__time_duration.sum", 0.0)
            rd_, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            extra = ""
            if rd_ is not None and t > 0:
                extra = "  dram rd %8.1f MB wr %8.1f MB  -> %7.1f GB/s" % (rd_ / 1e6, wr / 1e6, (rd_ + wr) / t)
            print("%4d %-28s grid %-16s block %-12s %10.1f us %5.1f%%%s" % (k, short(d["name"]), d["grid"], d["block"], t / 1e3, 100 * t / tot if tot else 0, extra))
