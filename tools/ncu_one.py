#!/usr/bin/env python
"""Run one plan a few times (for ncu).  usage: ncu_one.py kind type batch dims..."""
import sys, torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af
kind, typ, batch = sys.argv[1], (af.C2C if sys.argv[2] == "f" else af.Z2Z), int(sys.argv[3])
dims = [int(a) for a in sys.argv[4:]]
n = batch
for d in dims: n *= d
dt = torch.complex64 if typ == af.C2C else torch.complex128
x = torch.randn(n, dtype=dt, device="cuda"); y = torch.empty_like(x)
p = af.Plan(kind, dims, typ, batch)
print(p.describe())
for _ in range(3): p.exec(x, y, af.FORWARD)
torch.cuda.synchronize()
