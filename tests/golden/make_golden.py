#!/usr/bin/env python
"""Generates tests/golden/kat.npz -- known-answer vectors for the FFT hot path.

The reference (accelerate-fft) holds no golden vectors of its own and cannot
be run in this image (Haskell), so these fixtures are produced from the mathematical definition
(Mode.hs:21-26: Forward = e^{-2 pi i jk/n}) by an implementation that is independent of both the
product and the oracle under oracle/: scipy.fft (pocketfft) evaluated in long double, rounded to
complex128.  Inputs are seeded U(-1,1) like the reference's generators (test/Test/Base.hs:35-42).

Run:  python tests/golden/make_golden.py      (commits tests/golden/kat.npz)
"""
import os

import numpy as np
import scipy.fft as sf

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261017)

SIZES_1D = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 17, 25, 30, 31, 32, 64, 100, 127, 128, 243, 256, 1000, 1024]
SHAPES_2D = [(1, 1), (2, 3), (4, 4), (5, 8), (16, 16), (12, 10), (32, 8)]
SHAPES_3D = [(1, 1, 1), (2, 2, 2), (2, 3, 4), (4, 4, 4), (8, 4, 2), (3, 5, 7)]


def rnd(shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def exact(x, axes):
    return sf.fftn(x.astype(np.clongdouble), axes=axes).astype(np.complex128)


out = {}
for n in SIZES_1D:
    x = rnd((2, n))
    out[f"in1_{n}"] = x
    out[f"fwd1_{n}"] = exact(x, (-1,))                                   # from the c128 input
    out[f"fwd1_{n}_f32"] = exact(x.astype(np.complex64), (-1,))          # from the input rounded to c64
for s in SHAPES_2D:
    x = rnd(s)
    k = "x".join(map(str, s))
    out[f"in2_{k}"] = x
    out[f"fwd2_{k}"] = exact(x, (0, 1))
    out[f"fwd2_{k}_f32"] = exact(x.astype(np.complex64), (0, 1))
for s in SHAPES_3D:
    x = rnd(s)
    k = "x".join(map(str, s))
    out[f"in3_{k}"] = x
    out[f"fwd3_{k}"] = exact(x, (0, 1, 2))
    out[f"fwd3_{k}_f32"] = exact(x.astype(np.complex64), (0, 1, 2))
np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)
print("wrote", os.path.join(HERE, "kat.npz"), len(out), "arrays")
