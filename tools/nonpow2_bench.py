#!/usr/bin/env python
"""Non power-of-two row lengths in the reference's test range [33, 1024] (c64, ~1 GiB batches): % of the HBM roofline."""
import sys, math
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af

PEAK = 6543.7
def primes_with_big_factor(n):
    m = n
    for p in (2, 3, 5, 7, 11, 13):
        while m % p == 0: m //= p
    return m != 1
Z = "--c128" in sys.argv      # c128 instead of c64
lens = [int(a) for a in sys.argv[1:] if not a.startswith("-")] or [34, 37, 51, 67, 101, 127, 131, 193, 251, 257, 389, 509, 521, 769, 997, 998, 1009, 1021,
                                          48, 96, 100, 120, 360, 1000]
worst = (1.0, 0)
for n in lens:
    batch = (1 << 27) // n
    if Z: batch //= 2
    x = torch.randn(batch * n, dtype=torch.complex128 if Z else torch.complex64, device="cuda")
    y = torch.empty_like(x)
    p = af.Plan("many", [n], af.Z2Z if Z else af.C2C, batch)
    for _ in range(2): p.exec(x, y, af.FORWARD)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): p.exec(x, y, af.FORWARD)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    frac = 2 * batch * n * (16 if Z else 8) / ms / 1e6 / PEAK
    desc = p.describe().strip().split("\n")[0]
    kind = "bluestein" if "bluestein" in desc else "mixed" if "mixed-radix" in desc else "tiny" if "tiny" in desc else "pow2"
    if n >= 33 and frac < worst[0]: worst = (frac, n)
    print("%s n=%5d b=%8d %-9s %9.1f us  %5.1f %%   %s" % ("c128" if Z else "c64", n, batch, kind, ms * 1e3, 100 * frac, desc[:90]), flush=True)
    p.destroy()
    del x, y
print("worst in [33,1024]: n=%d at %.1f %%" % (worst[1], 100 * worst[0]))
