#!/usr/bin/env python
"""Small shapes through every kernel family, for compute-sanitizer (memcheck / synccheck / racecheck)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
os.environ["B200FFT_PIPE_MIN_TILES"] = "1"
import accelerate_fft_b200 as af
rng = np.random.default_rng(1)
def run(kind, shape, dt, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    af.lib().accfft_plan_cache_clear()
    x = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dt)
    f = getattr(af, kind)
    y = f("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
    ref = {"fft": lambda a: np.fft.fft(a, axis=-1), "fft2D": np.fft.fft2, "fft3D": np.fft.fftn}[kind](x.astype(np.complex128))
    e = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    print("%-6s %-18s %-10s %s rel %.1e" % (kind, shape, np.dtype(dt).name, env, e), flush=True)
    assert e < 1e-4
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
for dt, n1 in ((np.complex64, 1024), (np.complex128, 512)):
    run("fft", (7, 4096), dt)
    run("fft", (700, 4096), dt)                                   # ring (c128)
    run("fft2D", (n1, 64), dt, B200FFT_PIPE="1")                  # pipe CS=1
    run("fft2D", (n1 * 2, 64), dt, B200FFT_PIPE="1")              # pipe CS=2
    run("fft2D", (n1 * 8, 40), dt, B200FFT_PIPE="1", B200FFT_CLUSTER="1")   # pipe CS=8, 5 tiles
    run("fft2D", (n1 * 16, 16), dt, B200FFT_PIPE="1", B200FFT_CLUSTER="1")  # pipe CS=16
    run("fft2D", (n1 * 8, 24), dt, B200FFT_PIPE="0", B200FFT_CLUSTER="1")   # simple cluster kernel
    run("fft", (3, n1 * n1), dt, B200FFT_PIPE="1")                # four-step with the pipelined col+tw pass
    run("fft3D", (16, 32, 64), dt)
run("fft", (301, 8192), np.complex64, B200FFT_PIPE="1")           # pipe rows
run("fft", (5, 1000), np.complex64); run("fft", (5, 1009), np.complex64)   # mixed radix, Bluestein
for dt in (np.complex64, np.complex128):
    for n in (3, 12, 17, 31):
        run("fft", (300, n), dt); run("fft2D", (n, 70), dt)           # one thread per line, direct sum
    run("fft", (9, 1536), dt); run("fft2D", (360, 50), dt); run("fft", (3, 9000), np.complex64)   # mixed radix rows / strided
for dt in (np.complex64, np.complex128):
    for n in (2, 4, 8, 16, 32):
        run("fft", (1000, n), dt)                                 # rows staged through shared memory (ragged last tile)
    run("fft2D", (4096, 64), dt, B200FFT_CLUSTER_ROWS="0")
    run("fft2D", (4096, 4096), dt, B200FFT_CLUSTER_ROWS="1")      # rows + first radix-8 column stage in a cluster
    run("fft2D", (1024, 2048), dt)                                # default policy: pipelined N=1024 column pass
    x = (rng.uniform(-1, 1, (64, 32, 128)) + 1j * rng.uniform(-1, 1, (64, 32, 128))).astype(dt)
    y = af.fft_centred("Forward", torch.from_numpy(x).cuda()).cpu().numpy()          # rotation folded into the last passes
    assert np.linalg.norm(y - np.fft.fftshift(np.fft.fftn(x.astype(np.complex128)))) / np.linalg.norm(y) < 1e-4
    p = af.Plan("axis", (3, 1024, 96), af.C2C if dt == np.complex64 else af.Z2Z)
    xd = torch.from_numpy((rng.uniform(-1, 1, (3, 1024, 96)) + 0j).astype(dt)).cuda()
    outs = [torch.zeros((3, 128, 96), dtype=xd.dtype, device="cuda") for _ in range(8)]
    p.exec_scatter(xd, [o.data_ptr() for o in outs], 128 * 96, 96, af.FORWARD)        # scatter store
    torch.cuda.synchronize(); p.destroy()
    print("centred + scatter ok", np.dtype(dt).name, flush=True)
print("ok")
