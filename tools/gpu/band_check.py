#!/usr/bin/env python
"""Parity + timing of the band (L2-fused two-phase) kernel: 2D transforms whose column axis takes it."""
import sys, os, math
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af

def check(h, w, mode):
    torch.manual_seed(h + w)
    x = torch.view_as_complex(torch.rand(h, w, 2, device="cuda") * 2 - 1)
    y = af.fft2D(mode, x)
    xr = x.to(torch.complex128)
    ref = torch.fft.fft2(xr) if mode == "Forward" else torch.fft.ifft2(xr) * (1 if mode == "Inverse" else h * w)
    err = float(torch.linalg.vector_norm(y.to(torch.complex128) - ref) / torch.linalg.vector_norm(ref))
    p = af.Plan("2d", [h, w], af.C2C, 1)
    d = p.describe().strip().split("\n")[-1][:90]
    p.destroy()
    print("fft2D %-8s %5dx%-5d rel-L2 %.2e  %s | %s" % (mode, h, w, err, "OK" if err < 1e-5 * math.log2(h * w) else "FAIL", d), flush=True)
    return err

for (h, w) in ((8192, 128), (8192, 256), (8192, 1024)):
    for mode in ("Forward", "Inverse", "Reverse"):
        check(h, w, mode)
torch.cuda.synchronize()
