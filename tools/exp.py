#!/usr/bin/env python
"""Developer A/B timing: python tools/exp.py <experiment> -- each line = one (env, shape) combination, CUDA events."""
import math, os, sys
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af

PEAK = 6532.5
L2 = 126e6


def bench(name, kind, dims, typ, batch=1, passes_min=1, iters=10, env=None, desc=False):
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        esz = 8 if typ == af.C2C else 16
        dt = torch.complex64 if typ == af.C2C else torch.complex128
        n_total = batch
        for d in dims:
            n_total *= d
        nbytes = n_total * esz
        nbuf = max(1, int(math.ceil(2 * L2 / (2 * nbytes)))) if 2 * nbytes < 4 * L2 else 1
        plan = af.Plan(kind, dims, typ, batch)
        xs = [torch.randn(n_total, dtype=dt, device="cuda") for _ in range(nbuf)]
        ys = [torch.empty_like(x) for x in xs]
        for i in range(3):
            plan.exec(xs[i % nbuf], ys[i % nbuf], af.FORWARD)
        torch.cuda.synchronize()
        best = 1e30
        tot = 0.0
        reps = 3
        for r in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(iters):
                plan.exec(xs[i % nbuf], ys[i % nbuf], af.FORWARD)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            best = min(best, ms); tot += ms
        ms = tot / reps
        alg = passes_min * 2 * nbytes
        tag = " ".join("%s=%s" % (k.replace("B200FFT_", ""), v) for k, v in env.items())
        print(f"{name:28s} {tag:44s} mean {ms*1e3:9.1f} us best {best*1e3:9.1f} us  strict {alg/ms/1e6/PEAK*100:5.1f}%  ({plan.num_passes} passes)", flush=True)
        if desc:
            print(plan.describe())
        plan.destroy()
        del xs, ys
    except Exception as ex:
        print(name, env, "FAILED", ex, flush=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


which = sys.argv[1]
if which == "cfg3":
    bench("cfg3 8192^2", "2d", [8192, 8192], af.C2C, 1, 2, env={}, desc=True)
    for cv in range(4):
        for pv in range(2):
            bench("cfg3 8192^2", "2d", [8192, 8192], af.C2C, 1, 2, env={"B200FFT_PAIR2D": 1, "B200FFT_VARIANTS": f"c4096f={cv},p8192f={pv}"}, desc=(cv == 0 and pv == 0))
    # the column pass alone, as a strided axis [1][4096][16384] (same bytes as cfg3)
    for cv in range(4):
        bench("col4096 x16384", "axis", [1, 4096, 16384], af.C2C, 1, 1, env={"B200FFT_VARIANTS": f"c4096f={cv}", "B200FFT_MAX_COL_N": 4096})
    bench("4096^2", "2d", [4096, 4096], af.C2C, 1, 2, env={})
    bench("4096^2", "2d", [4096, 4096], af.C2C, 1, 2, env={"B200FFT_PAIR2D": 1})
    bench("4096^2 c128", "2d", [4096, 4096], af.Z2Z, 1, 2, env={})
    bench("4096^2 c128", "2d", [4096, 4096], af.Z2Z, 1, 2, env={"B200FFT_PAIR2D": 1})
if which == "cfg1":
    bench("cfg1 n=1024 b=4096", "many", [1024], af.C2C, 4096, 1, 50, env={"B200FFT_NO_RING": 1})
    for gv in range(2):
        bench("cfg1 n=1024 b=4096", "many", [1024], af.C2C, 4096, 1, 50, env={"B200FFT_VARIANTS": f"g1024f={gv}"}, desc=(gv == 0))
    for gv in range(2):
        bench("n=1024 b=131072", "many", [1024], af.C2C, 131072, 1, 10, env={"B200FFT_VARIANTS": f"g1024f={gv}", "B200FFT_RING_C64_MAX_LINES": 1 << 30})
    bench("n=1024 b=131072", "many", [1024], af.C2C, 131072, 1, 10, env={"B200FFT_NO_RING": 1})
    for b in (1024, 2048, 8192, 16384, 32768):
        bench(f"n=1024 b={b}", "many", [1024], af.C2C, b, 1, 50, env={"B200FFT_NO_RING": 1})
        bench(f"n=1024 b={b}", "many", [1024], af.C2C, b, 1, 50, env={"B200FFT_RING_C64_MAX_LINES": 1 << 30})
if which == "fused":
    bench("cfg3 8192^2", "2d", [8192, 8192], af.C2C, 1, 2, env={})
    for mb in (1, 2, 4):
        for la in (2, 4, 8):
            bench("cfg3 8192^2", "2d", [8192, 8192], af.C2C, 1, 2, env={"B200FFT_FUSED": 1, "B200FFT_BAND_MB": mb, "B200FFT_FUSED_LA": la, "B200FFT_FUSED_SLOTS": la + 4}, desc=(mb == 1 and la == 2))
    bench("cfg4 2^28", "1d", [1 << 28], af.C2C, 1, 2, 5)
    bench("cfg4 2^28", "1d", [1 << 28], af.C2C, 1, 2, 5, env={"B200FFT_FUSED": 1}, desc=True)
    bench("cfg5 1024^3", "3d", [1024, 1024, 1024], af.C2C, 1, 3, 3)
    bench("cfg5 1024^3", "3d", [1024, 1024, 1024], af.C2C, 1, 3, 3, env={"B200FFT_FUSED": 1}, desc=True)
