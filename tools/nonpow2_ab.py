#!/usr/bin/env python
"""A/B of one planner switch over row lengths, in one process: tools/nonpow2_ab.py B200FFT_MAX_PRIME=31 37 61 74 ...
prints, per length, % of the HBM roofline with the switch set and with the default (c64, ~1 GiB batches)."""
import os, sys
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af

PEAK = 6543.7
var, val = sys.argv[1].split("=")
lens = [int(a) for a in sys.argv[2:]]


def run(n):
    batch = (1 << 27) // n
    x = torch.randn(batch * n, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    p = af.Plan("many", [n], af.C2C, batch)
    for _ in range(2): p.exec(x, y, af.FORWARD)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): p.exec(x, y, af.FORWARD)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    d = p.describe().strip().split("\n")[0]
    p.destroy()
    return 100 * 2 * batch * n * 8 / ms / 1e6 / PEAK, d


for n in lens:
    os.environ[var] = val
    a, da = run(n)
    del os.environ[var]
    b, db = run(n)
    print("c64 n=%5d  %s=%s: %5.1f %% (%s)   default: %5.1f %% (%s)" % (n, var, val, a, da[:28], b, db[:60]), flush=True)
