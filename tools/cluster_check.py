#!/usr/bin/env python
"""Developer check of the cluster (DSMEM) column kernels: parity per variant vs numpy, then timing (run under gpurun)."""
import os, sys, math
import numpy as np
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af
os.environ.setdefault("B200FFT_CLUSTER", "1")   # the cluster kernels are opt-in
os.environ.setdefault("B200FFT_PIPE", "0")

PEAK = 6462.4
rng = np.random.default_rng(3)

def rel(y, ref):
    return float(np.linalg.norm((y - ref).ravel()) / np.linalg.norm(ref.ravel()))

def parity(n, inner, typ, nvar, outer=1):
    dt = np.complex64 if typ == af.C2C else np.complex128
    x = (rng.uniform(-1, 1, (outer, n, inner)) + 1j * rng.uniform(-1, 1, (outer, n, inner))).astype(dt)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    refi = np.fft.ifft(x.astype(np.complex128), axis=1)
    for v in range(nvar):
        os.environ["B200FFT_VARIANTS"] = "k%d%s=%d" % (n, "f" if typ == af.C2C else "d", v)
        p = af.Plan("axis", (outer, n, inner), typ)
        xd = torch.from_numpy(x).cuda(); yd = torch.empty_like(xd)
        p.exec(xd, yd, af.FORWARD); e1 = rel(yd.cpu().numpy(), ref)
        p.exec(xd, yd, af.INVERSE, scale=1.0 / n); e2 = rel(yd.cpu().numpy(), refi)
        bar = (1e-5 if typ == af.C2C else 1e-13) * math.log2(n)
        print("%s n=%d inner=%d outer=%d v%d fwd %.2e inv %.2e bar %.1e | %s" % ("OK " if max(e1, e2) < bar else "BAD", n, inner, outer, v, e1, e2, bar, p.describe().strip()), flush=True)
        p.destroy()
    os.environ["B200FFT_VARIANTS"] = ""

def timing(name, kind, dims, typ, iters=10, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k); os.environ[k] = v
    n_total = 1
    for d in dims: n_total *= d
    dt = torch.complex64 if typ == af.C2C else torch.complex128
    x = torch.randn(n_total, dtype=dt, device="cuda"); y = torch.empty_like(x)
    p = af.Plan(kind, dims, typ)
    for _ in range(3): p.exec(x, y, af.FORWARD)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): p.exec(x, y, af.FORWARD)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    nbytes = n_total * (8 if typ == af.C2C else 16)
    print("%-40s %8.1f us  per-pass-equivalent (1 pass) %6.0f GB/s = %5.1f%%  passes=%d" % (name, ms * 1e3, 2 * nbytes / ms / 1e6, 2 * nbytes / ms / 1e6 / PEAK * 100, p.num_passes), flush=True)
    p.destroy()
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v

if "parity" in sys.argv or len(sys.argv) == 1:
    parity(8192, 64, af.C2C, 7)
    parity(8192, 24, af.C2C, 7, outer=2)
    parity(4096, 48, af.C2C, 2)
    parity(16384, 16, af.C2C, 2)
    parity(4096, 16, af.Z2Z, 1)
    parity(8192, 12, af.Z2Z, 1, outer=2)
if "time" in sys.argv or len(sys.argv) == 1:
    for v in range(7):
        timing("col 8192 x 8192 c64 cluster v%d" % v, "axis", (1, 8192, 8192), af.C2C, env={"B200FFT_VARIANTS": "k8192f=%d" % v})
    timing("col 8192 x 8192 c64 four-step (2 passes)", "axis", (1, 8192, 8192), af.C2C, env={"B200FFT_CLUSTER": "0"})
    timing("rows 8192 x 8192 c64", "many", (8192,), af.C2C) if False else None
    for v in range(5):
        timing("cfg3 2D 8192^2 cluster v%d" % v, "2d", (8192, 8192), af.C2C, env={"B200FFT_VARIANTS": "k8192f=%d" % v})
    timing("cfg3 2D 8192^2 no cluster", "2d", (8192, 8192), af.C2C, env={"B200FFT_CLUSTER": "0"})
    for v in range(2):
        timing("col 4096 x 8192 c64 cluster v%d" % v, "axis", (1, 4096, 8192), af.C2C, env={"B200FFT_VARIANTS": "k4096f=%d" % v})
    timing("col 4096 x 8192 c64 four-step", "axis", (1, 4096, 8192), af.C2C, env={"B200FFT_CLUSTER": "0"})
    for v in range(2):
        timing("col 16384 x 4096 c64 cluster v%d" % v, "axis", (1, 16384, 4096), af.C2C, env={"B200FFT_VARIANTS": "k16384f=%d" % v})
    timing("col 16384 x 4096 c64 four-step", "axis", (1, 16384, 4096), af.C2C, env={"B200FFT_CLUSTER": "0"})
    timing("col 4096 x 4096 c128 cluster", "axis", (1, 4096, 4096), af.Z2Z)
    timing("col 4096 x 4096 c128 four-step", "axis", (1, 4096, 4096), af.Z2Z, env={"B200FFT_CLUSTER": "0"})
    timing("col 8192 x 4096 c128 cluster", "axis", (1, 8192, 4096), af.Z2Z)
    timing("col 8192 x 4096 c128 four-step", "axis", (1, 8192, 4096), af.Z2Z, env={"B200FFT_CLUSTER": "0"})
