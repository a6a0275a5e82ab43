"""Planner fuzz (GPU): seeded random shapes through the public entry points (accfft_fft / fft1D / fft2D / fft3D), each
checked against numpy's float64 FFT (pocketfft) at the north_star tolerance, with three extra checks per case:
the output buffer starts as NaN (an element the plan never stores shows up), the input must be bit-identical afterwards
(out-of-place contract, PTX.hs:92), and a second call must reproduce the first bit for bit (no stale scratch / flags).

The shapes are drawn around the planner's switch points (csrc/plan.cu): line lengths either side of 2048 / 4096 / 8192 /
16384 (lines kernel / ring / four-step), batches that leave a ragged last tile, column axes of 512 ... 32768 points with
inner widths that are not a multiple of the tile width, smooth and prime lengths up to a few hundred thousand.

    python tools/plan_fuzz.py [--cases 300] [--seed 1] [--max-elems 23]
"""
import argparse
import ctypes
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MODES = ["Forward", "Reverse", "Inverse"]
_MODE = {"Forward": 0, "Reverse": 1, "Inverse": 2}


def draw_len(rng, cap):
    """one axis length <= cap: powers of two, their neighbours, smooth composites, primes, plain random"""
    kind = rng.integers(0, 10)
    if kind <= 3:
        n = 1 << int(rng.integers(0, max(1, int(math.log2(cap)) + 1)))
    elif kind == 4:
        n = (1 << int(rng.integers(1, max(2, int(math.log2(cap)) + 1)))) + int(rng.choice([-1, 1]))
    elif kind <= 6:
        n = 1
        while True:
            f = int(rng.choice([2, 2, 3, 3, 5, 7, 11, 13]))
            if n * f > cap:
                break
            n *= f
            if rng.random() < 0.15:
                break
    elif kind == 7:
        n = int(rng.integers(1, max(2, min(cap, 1024)) + 1))     # the reference suite's own range (test/Test/Base.hs:44-45)
    else:
        n = int(rng.integers(1, cap + 1))
    return max(1, min(int(n), cap))


def draw_case(rng, max_log):
    total_cap = 1 << int(rng.integers(10, max_log + 1))
    kind = str(rng.choice(["fft", "fft", "fft", "fft1D", "fft2D", "fft2D", "fft3D"]))
    if kind == "fft1D":
        shape = (draw_len(rng, total_cap),)
    elif kind == "fft":
        n = draw_len(rng, total_cap)
        rest = max(1, total_cap // n)
        r = int(rng.integers(2, 5))
        outer = []
        for _ in range(r - 1):
            b = int(rng.integers(1, rest + 1)) if rng.random() < 0.7 else draw_len(rng, rest)
            outer.append(b)
            rest = max(1, rest // b)
        shape = tuple(outer) + (n,)
    elif kind == "fft2D":
        h = draw_len(rng, min(total_cap, 1 << 16))
        w = draw_len(rng, max(1, total_cap // h))
        shape = (h, w) if rng.random() < 0.5 else (w, h)
    else:
        d = draw_len(rng, min(total_cap, 1 << 12))
        h = draw_len(rng, max(1, min(total_cap // d, 1 << 12)))
        w = draw_len(rng, max(1, total_cap // (d * h)))
        shape = tuple(rng.permutation([d, h, w]).tolist())
    dtype = np.complex64 if rng.random() < 0.6 else np.complex128
    mode = str(rng.choice(MODES))
    return kind, shape, dtype, mode


def targeted_cases():
    """shapes that sit on the planner's switch points and that a uniform draw rarely meets"""
    out = []
    for n in (2048, 4096, 8192, 16384, 32768, 65536, 1 << 17, 1 << 20, 1 << 22):      # rows: lines / ring / four-step
        for b in (1, 3, 37, 149, 297, 300):
            if n * b <= (1 << 23):
                out.append(("fft", (b, n)))
    for h in (512, 1024, 2048, 4096, 8192, 16384, 32768):                                  # column axes and ragged inner widths
        for w in (1, 8, 24, 96, 100, 130, 512):
            if h * w <= (1 << 22):
                out.append(("fft2D", (h, w)))
    for shp in ((1024, 1024, 4), (4, 1024, 1024), (1024, 4, 1024), (512, 2, 2048), (64, 64, 64), (128, 96, 80), (3, 2048, 100),
                (2048, 3, 100), (100, 3, 2048), (16, 4096, 16), (4096, 16, 16), (16, 16, 8192)):
        out.append(("fft3D", shp))
    for n in (4099, 8191, 10007, 65537, 100003, 3 * 4096, 5 * 8192, 30030, 2 * 3 * 5 * 7 * 11 * 13 * 4, 1000000, 999983):
        out.append(("fft1D", (n,)))
        out.append(("fft", (5, n)))
    return out


def reference(kind, mode, x):
    x = x.astype(np.complex128)
    if kind in ("fft", "fft1D"):
        f = np.fft.fft if mode == "Forward" else np.fft.ifft
        y = f(x, axis=-1)
        n = x.shape[-1]
    else:
        f = np.fft.fftn if mode == "Forward" else np.fft.ifftn
        y = f(x)
        n = x.size
    if mode == "Reverse":
        y = y * n          # un-normalised inverse (Mode.hs:15-19)
    return y, n


def run_case(af, torch, kind, shape, dtype, mode, seed):
    from accelerate_fft_b200 import _lib
    rng = np.random.default_rng(seed)
    x = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dtype)
    xin = torch.from_numpy(x).cuda()
    keep = xin.clone()
    outs = []
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    L = af.lib()
    shp = (ctypes.c_int64 * len(shape))(*shape)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        out = torch.full_like(xin, complex(float("nan"), float("nan")))
        if kind == "fft":
            st = L.accfft_fft(_MODE[mode], len(shape), shp, typ, xin.data_ptr(), out.data_ptr(), stream)
        elif kind == "fft1D":
            st = L.accfft_fft1D(_MODE[mode], shape[0], typ, xin.data_ptr(), out.data_ptr(), stream)
        elif kind == "fft2D":
            st = L.accfft_fft2D(_MODE[mode], shape[0], shape[1], typ, xin.data_ptr(), out.data_ptr(), stream)
        else:
            st = L.accfft_fft3D(_MODE[mode], shape[0], shape[1], shape[2], typ, xin.data_ptr(), out.data_ptr(), stream)
        _lib.check(st, kind)
        torch.cuda.synchronize()
        outs.append(out)
    problems = []
    if not torch.equal(xin, keep):
        problems.append("input modified")
    y = outs[0].cpu().numpy()
    if np.isnan(y.real).any() or np.isnan(y.imag).any():
        problems.append("output holds NaN (%d elements never stored?)" % int(np.isnan(y.real).sum()))
    if not torch.equal(outs[0].view(torch.float32 if dtype == np.complex64 else torch.float64),
                       outs[1].view(torch.float32 if dtype == np.complex64 else torch.float64)):
        problems.append("second call differs from the first")
    ref, n = reference(kind, mode, x)
    den = np.linalg.norm(ref.ravel())
    err = float(np.linalg.norm((y.astype(np.complex128) - ref).ravel()) / den) if den > 0 else 0.0
    lg = max(1.0, math.log2(max(2, n)))
    tol = (1e-5 if dtype == np.complex64 else 1e-13) * lg
    if not (err <= tol):
        problems.append("rel-L2 %.3e > %.3e" % (err, tol))
    return err, tol, problems


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-elems", type=int, default=23, help="log2 of the largest array (elements)")
    ap.add_argument("--targeted", action="store_true", help="run the fixed list of switch-point shapes first (both dtypes)")
    ap.add_argument("--budget-s", type=float, default=1e9, help="stop drawing new cases after this many seconds")
    a = ap.parse_args()
    import torch
    import accelerate_fft_b200 as af
    af.lib()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    bad = 0
    worst = 0.0
    done = 0
    cases = []
    if a.targeted:
        for i, (kind, shape) in enumerate(targeted_cases()):
            for dtype in (np.complex64, np.complex128):
                cases.append((kind, shape, dtype, MODES[(i + (dtype == np.complex128)) % 3]))
    cases += [draw_case(rng, a.max_elems) for _ in range(a.cases)]
    for c, (kind, shape, dtype, mode) in enumerate(cases):
        if time.time() - t0 > a.budget_s:
            print("plan_fuzz: time budget reached after %d of %d cases" % (c, len(cases)), flush=True)
            break
        try:
            err, tol, problems = run_case(af, torch, kind, shape, dtype, mode, a.seed * 100003 + c)
        except Exception as e:   # a refused plan is a finding too
            err, tol, problems = float("nan"), 0.0, ["exception: %r" % (e,)]
        done += 1
        worst = max(worst, err / tol if tol > 0 and err == err else 0.0)
        if problems:
            bad += 1
            print("FAIL #%d %s %s %s %s: %s" % (c, kind, shape, np.dtype(dtype).name, mode, "; ".join(problems)), flush=True)
    print("plan_fuzz: %d cases, %d failed, worst err/tol %.3f, %.1f s" % (done, bad, worst, time.time() - t0), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
