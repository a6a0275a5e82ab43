#!/usr/bin/env python
"""Round-2 kernels on small shapes, for compute-sanitizer (memcheck / racecheck): one-launch Bluestein, the band kernel
(strided axis through L2 slots), the grid-stride scatter pass + flag barrier of the slab entry point (one rank), and
back-to-back programmatic dependent launches of the lines / ring kernels."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import accelerate_fft_b200 as af
from accelerate_fft_b200._lib import ALLGATHER_FN
rng = np.random.default_rng(2)

def rc(shape, dt):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dt)

def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)

for dt in (np.complex64, np.complex128):
    for n in (34, 67, 131, 257, 521, 1009):                      # Bluestein on M = 128 ... 2048, ragged last tile
        x = rc((37, n), dt)
        y = af.fft("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
        e = rel(y, np.fft.fft(x.astype(np.complex128), axis=-1))
        print("bluestein n=%d %s rel %.1e" % (n, np.dtype(dt).name, e), flush=True)
        assert e < 1e-4
    xs = [torch.from_numpy(rc((300, 4096 if dt == np.complex128 else 1024), dt)).cuda() for _ in range(3)]
    ys = [af.fft("Forward", x) for x in xs for _ in range(2)]    # back-to-back launches (PDL): ring (c128) / lines (c64)
    torch.cuda.synchronize()
    for x, y in zip([x for x in xs for _ in range(2)], ys):
        assert rel(y.cpu().numpy(), np.fft.fft(x.cpu().numpy().astype(np.complex128), axis=-1)) < 1e-4
    print("pdl back-to-back ok", np.dtype(dt).name, flush=True)
for dt, n in ((np.complex64, 16384), (np.complex64, 8192), (np.complex128, 8192), (np.complex128, 4096)):   # single-stage rings
    x = rc((2 * 148 + 3, n), dt)
    p = af.Plan("many", [n], af.C2C if dt == np.complex64 else af.Z2Z, x.shape[0]); assert "ring: G=1 NS=1" in p.describe(), p.describe(); p.destroy()
    y = af.fft("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
    e = rel(y, np.fft.fft(x.astype(np.complex128), axis=-1))
    print("single-stage ring n=%d %s rel %.1e" % (n, np.dtype(dt).name, e), flush=True)
    assert e < 1e-4
for dt in (np.complex64, np.complex128):                          # single-buffer TMA column kernel: 1024-point columns, 148+ tiles
    tl = 16 if dt == np.complex64 else 8
    x = rc((1024, 150 * tl), dt)
    p = af.Plan("axis", (1, 1024, 150 * tl), af.C2C if dt == np.complex64 else af.Z2Z)
    d = p.describe(); p.destroy()
    xd = torch.from_numpy(x).cuda()
    y = af.fft2D("Forward", xd).cpu().numpy() if False else None
    import os
    os.environ["B200FFT_PIPE"] = "0"; af.lib().accfft_plan_cache_clear()
    p = af.Plan("axis", (1, 1024, 150 * tl), af.C2C if dt == np.complex64 else af.Z2Z)
    assert "ring cols" in p.describe(), p.describe()
    out = torch.empty_like(xd)
    p.exec(xd, out, af.FORWARD); torch.cuda.synchronize(); p.destroy()
    os.environ.pop("B200FFT_PIPE"); af.lib().accfft_plan_cache_clear()
    e = rel(out.cpu().numpy(), np.fft.fft(x.astype(np.complex128), axis=0))
    print("ring cols %s rel %.1e" % (np.dtype(dt).name, e), flush=True)
    assert e < 1e-4
x = rc((8192, 128), np.complex64)                                 # band kernel: 4 bands of 32 columns
p = af.Plan("2d", [8192, 128], af.C2C, 1); assert "band A[" in p.describe(), p.describe(); p.destroy()
y = af.fft2D("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
e = rel(y, np.fft.fft2(x.astype(np.complex128)))
print("band 8192x128 rel %.1e" % e, flush=True)
assert e < 1e-4
lib = af.lib()
cb = ALLGATHER_FN(lambda c, s, r, n: (ctypes.memmove(r, s, n), 0)[1])
for dt, typ in ((np.complex64, af.C2C), (np.complex128, af.Z2Z)):  # slab entry point, one rank, chunked + SM-limited scatter
    d, h, w = 128, 256, 64
    x = rc((d, h, w), dt)
    xd = torch.from_numpy(x).cuda()
    hnd = ctypes.c_void_p()
    assert lib.b200fftPlanSlab3d(ctypes.byref(hnd), d, h, w, typ, 0, 1, 1, cb, None) == 0
    assert lib.b200fftSlabTune(hnd, 0, 2, 2, 6) == 0 and lib.b200fftSlabTune(hnd, 1, 2, 2, 6) == 0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat = torch.empty_like(xd); tr = torch.empty((h, d, w), dtype=xd.dtype, device="cuda")
    assert lib.b200fftExecSlab(hnd, xd.data_ptr(), nat.data_ptr(), af.FORWARD, 1.0, 0, st) == 0
    assert lib.b200fftExecSlab(hnd, xd.data_ptr(), tr.data_ptr(), af.FORWARD, 1.0, 1, st) == 0
    ref = np.fft.fftn(x.astype(np.complex128))
    assert rel(nat.cpu().numpy(), ref) < 1e-4 and rel(tr.cpu().numpy().transpose(1, 0, 2), ref) < 1e-4
    assert lib.b200fftDestroySlab(hnd) == 0
    print("slab entry point ok", np.dtype(dt).name, flush=True)
print("ok")
