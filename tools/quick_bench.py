#!/usr/bin/env python
"""Developer timing of the BASELINE configs (CUDA events, rotating buffers when smaller than L2)."""
import sys, math, time
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af

PEAK = 6462.4  # GB/s measured copy
L2 = 126e6

def bench(name, kind, dims, typ, batch=1, passes_min=1, iters=20):
    esz = 8 if typ == af.C2C else 16
    dt = torch.complex64 if typ == af.C2C else torch.complex128
    n_total = batch
    for d in dims: n_total *= d
    nbytes = n_total * esz
    nbuf = max(1, int(math.ceil(2 * L2 / (2 * nbytes)))) if 2 * nbytes < 4 * L2 else 1
    try:
        plan = af.Plan(kind, dims, typ, batch)
    except Exception as e:
        print(name, "PLAN FAILED", e); return
    xs = [torch.randn(n_total, dtype=dt, device="cuda") for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    for i in range(3):
        plan.exec(xs[i % nbuf], ys[i % nbuf], af.FORWARD)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        plan.exec(xs[i % nbuf], ys[i % nbuf], af.FORWARD)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    npts = 1
    for d in dims: npts *= d
    flops = 5 * npts * math.log2(npts) * batch
    alg = passes_min * 2 * nbytes
    act = plan.num_passes * 2 * nbytes
    print(f"{name:34s} {ms*1e3:10.1f} us  {flops/ms/1e6:9.0f} GFLOP/s  strict {alg/ms/1e6:7.0f} GB/s = {alg/ms/1e6/PEAK*100:5.1f}%  per-pass {act/ms/1e6/PEAK*100:5.1f}% ({plan.num_passes} passes, nbuf={nbuf})", flush=True)
    if "-d" in sys.argv: print(plan.describe())
    plan.destroy()

which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "all"
if which in ("all", "1"): bench("cfg1 c64 n=1024 b=4096", "many", [1024], af.C2C, 4096, 1, 50)
if which in ("all", "2"): bench("cfg2 c128 n=4096 b=65536", "many", [4096], af.Z2Z, 65536, 1, 10)
if which in ("all", "3"): bench("cfg3 c64 8192x8192", "2d", [8192, 8192], af.C2C, 1, 2, 10)
if which in ("all", "4"): bench("cfg4 c64 n=2^28", "1d", [1 << 28], af.C2C, 1, 2, 5)
if which in ("all", "5"): bench("cfg5 c64 1024^3", "3d", [1024, 1024, 1024], af.C2C, 1, 3, 3)
if which in ("all", "sweep"):
    for lg in range(4, 15):
        n = 1 << lg
        bench(f"c64 n={n} b={2**27//n}", "many", [n], af.C2C, 2**27 // n, 1, 10)
    for lg in range(4, 14):
        n = 1 << lg
        bench(f"c128 n={n} b={2**26//n}", "many", [n], af.Z2Z, 2**26 // n, 1, 10)
if which == "var":
    import os
    def withvar(v, f):
        os.environ["B200FFT_VARIANTS"] = v; f(); os.environ["B200FFT_VARIANTS"] = ""
    for v in range(4):
        withvar(f"r4096d={v}", lambda: bench(f"c128 n=4096 b=65536 v{v}", "many", [4096], af.Z2Z, 65536, 1, 10))
    for v in range(2):
        withvar(f"r2048d={v}", lambda: bench(f"c128 n=2048 b=32768 v{v}", "many", [2048], af.Z2Z, 32768, 1, 10))
    for v in range(4):
        withvar(f"r1024d={v}", lambda: bench(f"c128 n=1024 b=65536 v{v}", "many", [1024], af.Z2Z, 65536, 1, 10))
    for v in range(3):
        withvar(f"r8192f={v}", lambda: bench(f"c64 n=8192 b=16384 v{v}", "many", [8192], af.C2C, 16384, 1, 10))
    for v in range(2):
        withvar(f"r16384f={v}", lambda: bench(f"c64 n=16384 b=8192 v{v}", "many", [16384], af.C2C, 8192, 1, 10))
    for v in range(4):
        withvar(f"c1024f={v}", lambda: bench(f"cfg5 1024^3 col v{v}", "3d", [1024, 1024, 1024], af.C2C, 1, 3, 3))
    bench("cfg1 c64 n=1024 b=4096", "many", [1024], af.C2C, 4096, 1, 50)
if which == "ring":
    bench("cfg2 c128 n=4096 b=65536", "many", [4096], af.Z2Z, 65536, 1, 10)
    bench("c128 n=2048 b=32768", "many", [2048], af.Z2Z, 32768, 1, 10)
    bench("c64 n=8192 b=16384", "many", [8192], af.C2C, 16384, 1, 10)
    bench("c64 n=4096 b=32768", "many", [4096], af.C2C, 32768, 1, 10)
