#!/usr/bin/env python
"""Developer check of the persistent pipelined column kernels (plain and cluster): parity vs numpy, then timing."""
import os, sys, math
import numpy as np
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af
sys.argv = [sys.argv[0], "none"] + sys.argv[1:]
exec(open("tools/cluster_check.py").read().split('if "parity" in sys.argv')[0])

def parity_axis(n, inner, typ, outer=1):
    dt = np.complex64 if typ == af.C2C else np.complex128
    x = (rng.uniform(-1, 1, (outer, n, inner)) + 1j * rng.uniform(-1, 1, (outer, n, inner))).astype(dt)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    refi = np.fft.ifft(x.astype(np.complex128), axis=1)
    p = af.Plan("axis", (outer, n, inner), typ)
    xd = torch.from_numpy(x).cuda(); yd = torch.empty_like(xd)
    p.exec(xd, yd, af.FORWARD); e1 = rel(yd.cpu().numpy(), ref)
    p.exec(xd, yd, af.INVERSE, scale=1.0 / n); e2 = rel(yd.cpu().numpy(), refi)
    bar = (1e-5 if typ == af.C2C else 1e-13) * math.log2(n)
    print("%s n=%d inner=%d outer=%d fwd %.2e inv %.2e bar %.1e | %s" % ("OK " if max(e1, e2) < bar else "BAD", n, inner, outer, e1, e2, bar, p.describe().strip()), flush=True)
    p.destroy()

def parity_1d(n, typ):
    dt = np.complex64 if typ == af.C2C else np.complex128
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(dt)
    ref = np.fft.fft(x.astype(np.complex128))
    p = af.Plan("1d", (n,), typ)
    xd = torch.from_numpy(x).cuda(); yd = torch.empty_like(xd)
    p.exec(xd, yd, af.FORWARD); e1 = rel(yd.cpu().numpy(), ref)
    bar = (1e-5 if typ == af.C2C else 1e-13) * math.log2(n)
    print("%s 1d n=%d fwd %.2e bar %.1e | %s" % ("OK " if e1 < bar else "BAD", n, e1, bar, p.describe().replace("\n", " || ")), flush=True)
    p.destroy()

if "parity" in sys.argv:
    os.environ["B200FFT_PIPE_MIN_TILES"] = "1"; os.environ["B200FFT_PIPE"] = "1"
    for (n, inner, outer) in [(1024, 64, 1), (1024, 512, 3), (2048, 128, 2), (4096, 64, 1), (8192, 64, 1), (8192, 1024, 1), (8192, 256, 2), (16384, 128, 1)]:
        parity_axis(n, inner, af.C2C, outer)
    for (n, inner, outer) in [(512, 64, 1), (512, 256, 3), (1024, 64, 2), (2048, 64, 1), (4096, 128, 1), (8192, 64, 2)]:
        parity_axis(n, inner, af.Z2Z, outer)
    parity_1d(1 << 20, af.C2C)
    parity_1d(1 << 18, af.Z2Z)
    os.environ.pop("B200FFT_PIPE_MIN_TILES"); os.environ["B200FFT_PIPE"] = "0"
if "time" in sys.argv:
    for env in ({"B200FFT_PIPE": "1"}, {"B200FFT_PIPE": "0"}, {"B200FFT_PIPE": "0", "B200FFT_CLUSTER": "0"}):
        tag = " pipe" if env["B200FFT_PIPE"] == "1" else " nopipe" if len(env) == 1 else " nopipe nocluster"
        timing("col 8192 x 8192 c64" + tag, "axis", (1, 8192, 8192), af.C2C, env=env)
        timing("cfg3 2D 8192^2 c64" + tag, "2d", (8192, 8192), af.C2C, env=env)
        timing("col 1024 x 65536 c64" + tag, "axis", (1, 1024, 65536), af.C2C, env=env)
        timing("col 4096 x 8192 c64" + tag, "axis", (1, 4096, 8192), af.C2C, env=env)
        timing("col 2048 x 16384 c64" + tag, "axis", (1, 2048, 16384), af.C2C, env=env)
        timing("col 16384 x 4096 c64" + tag, "axis", (1, 16384, 4096), af.C2C, env=env)
        timing("col 512 x 65536 c128" + tag, "axis", (1, 512, 65536), af.Z2Z, env=env)
        timing("col 4096 x 4096 c128" + tag, "axis", (1, 4096, 4096), af.Z2Z, env=env)
        timing("col 8192 x 4096 c128" + tag, "axis", (1, 8192, 4096), af.Z2Z, env=env)
        timing("cfg5 3D 1024^3 c64" + tag, "3d", (1024, 1024, 1024), af.C2C, iters=3, env=env)
        timing("cfg4 1D 2^28 c64" + tag, "1d", (1 << 28,), af.C2C, iters=3, env=env)
