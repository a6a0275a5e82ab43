#!/usr/bin/env python
"""Developer sweep: parity of the CUDA path vs numpy/scipy over many shapes (run under gpurun)."""
import sys, time, math
import numpy as np
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af

rng = np.random.default_rng(7)
bad = 0

def rnd(shape, dt):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dt)

def rel(y, ref):
    return float(np.linalg.norm((y - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))

def check(name, y, ref, dt, npts):
    global bad
    e = rel(y.astype(np.complex128), ref)
    bar = (1e-5 if dt == np.complex64 else 1e-13) * max(1.0, math.log2(max(npts, 2)))
    tight = (4e-7 if dt == np.complex64 else 8e-16) * max(1.0, math.sqrt(math.log2(max(npts, 2))))
    flag = "OK " if e <= bar else "BAD"
    if e > bar: bad += 1
    if e > bar or e > 4 * tight or "-v" in sys.argv:
        print(f"{flag} {name:50s} rel={e:.3e} bar={bar:.1e}", flush=True)

def run1d(n, batch, dt, mode="Forward"):
    x = rnd((batch, n), dt)
    try:
        y = af.fft(mode, torch.from_numpy(x).cuda()).cpu().numpy()
    except Exception as ex:
        global bad; bad += 1
        print("EXC", n, batch, dt.__name__, mode, ex, flush=True); return
    x128 = x.astype(np.complex128)
    ref = np.fft.fft(x128, axis=-1) if mode == "Forward" else np.fft.ifft(x128, axis=-1) * (n if mode == "Reverse" else 1)
    check(f"fft {mode} n={n} batch={batch} {dt.__name__}", y, ref, dt, n)

t0 = time.time()
for dt in (np.complex64, np.complex128):
    for lg in range(0, 15):
        n = 1 << lg
        if dt == np.complex128 and n > 8192: continue
        for mode in ("Forward", "Reverse", "Inverse"):
            run1d(n, 37, dt, mode)
    print("pow2 rows done", dt.__name__, time.time() - t0, flush=True)
    for n in [1 << 15, 1 << 16, 1 << 17, 1 << 20, 1 << 21, 1 << 22]:
        run1d(n, 3, dt, "Forward"); run1d(n, 1, dt, "Inverse")
    print("four-step done", dt.__name__, time.time() - t0, flush=True)
    for n in [3, 5, 6, 7, 9, 10, 11, 12, 13, 15, 17, 19, 23, 24, 25, 30, 31, 36, 49, 60, 97, 100, 121, 125, 127, 169, 243, 360, 500, 625, 720, 1000, 1001, 1009, 1023, 1024 * 3, 5000, 4099, 65537, 100000]:
        run1d(n, 5, dt, "Forward"); run1d(n, 2, dt, "Inverse")
    print("non-pow2 done", dt.__name__, time.time() - t0, flush=True)
    for (h, w) in [(1, 1), (1, 8), (8, 1), (4, 4), (16, 64), (64, 16), (128, 256), (512, 512), (2048, 64), (64, 4096), (4096, 32), (8192, 16), (32, 16384), (3, 5), (48, 128), (100, 100), (37, 64), (64, 37), (1000, 24), (17, 1024)]:
        x = rnd((h, w), dt)
        for mode in ("Forward", "Inverse"):
            try:
                y = af.fft2D(mode, torch.from_numpy(x).cuda()).cpu().numpy()
            except Exception as ex:
                bad += 1; print("EXC 2d", h, w, ex, flush=True); continue
            x128 = x.astype(np.complex128)
            ref = np.fft.fft2(x128) if mode == "Forward" else np.fft.ifft2(x128)
            check(f"fft2D {mode} {h}x{w} {dt.__name__}", y, ref, dt, h * w)
    print("2D done", dt.__name__, time.time() - t0, flush=True)
    for (d, h, w) in [(1, 1, 1), (2, 2, 2), (8, 16, 32), (32, 8, 16), (64, 64, 64), (16, 32, 64), (3, 5, 7), (16, 32, 64)[::-1], (10, 12, 14), (4, 1024, 8), (1024, 4, 8), (5, 64, 33), (128, 128, 128)]:
        x = rnd((d, h, w), dt)
        for mode in ("Forward", "Inverse"):
            try:
                y = af.fft3D(mode, torch.from_numpy(x).cuda()).cpu().numpy()
            except Exception as ex:
                bad += 1; print("EXC 3d", d, h, w, ex, flush=True); continue
            x128 = x.astype(np.complex128)
            ref = np.fft.fftn(x128) if mode == "Forward" else np.fft.ifftn(x128)
            check(f"fft3D {mode} {d}x{h}x{w} {dt.__name__}", y, ref, dt, d * h * w)
    print("3D done", dt.__name__, time.time() - t0, flush=True)
# rank-3 fft (innermost axis), host path
x = rnd((6, 10, 256), np.complex64)
y = af.run_host("fft", "Forward", x)
check("run_host fft rank3", y, np.fft.fft(x.astype(np.complex128), axis=-1), np.complex64, 256)
print("launches", af.kernel_launches(), "bad", bad, "time", time.time() - t0)
sys.exit(1 if bad else 0)
