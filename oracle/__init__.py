"""CPU oracle for the FFT hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / CPU baseline.  The product
(accelerate_fft_b200) never imports it and has no CPU fallback.

It wraps oracle/liboracle.so (built from adhoc_oracle.c by `make -C oracle`), the plain-C
restatement of the reference's pure-Accelerate path:
  * adhoc_fft / adhoc_fft2d / adhoc_fft3d  <- Adhoc.hs:37-48 (+ FFT.hs:136-138,166-187)
  * fft / fft1D / fft2D / fft3D wrappers that add the `Inverse` scaling of FFT.hs:83,110,141,172
  * exact_dft / exact_bin: the long-double definition of the transform (Mode.hs:21-26).
  * dft / idft (dft_definition.c): the reference's own definition module, DFT.hs:42-59 + DFT/Roots.hs:26-51, in the
    working precision -- a second oracle that shares no algorithm with the first.
PARITY STATUS: unpinned by reference-held vectors (the reference has none); see adhoc_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FORWARD, REVERSE, INVERSE = "Forward", "Reverse", "Inverse"


def sign_of_mode(mode):
    """Mode.hs:21-26 signOfMode."""
    return {FORWARD: -1, REVERSE: +1, INVERSE: +1}[mode]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("adhoc_oracle.c", "adhoc_impl.inc", "dft_definition.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        vp, sz, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        for suf in ("f32", "f64"):
            getattr(lib, "adhoc_fft_" + suf).argtypes = [i, sz, sz, vp, vp, i]
            getattr(lib, "adhoc_fft2d_" + suf).argtypes = [i, sz, sz, vp, vp, i]
            getattr(lib, "adhoc_fft3d_" + suf).argtypes = [i, sz, sz, sz, vp, vp, i]
            for n in ("adhoc_fft_", "adhoc_fft2d_", "adhoc_fft3d_"):
                getattr(lib, n + suf).restype = None
        for suf in ("f32", "f64"):
            getattr(lib, "dft_definition_" + suf).argtypes = [i, sz, sz, vp, vp]
            getattr(lib, "dft_definition_" + suf).restype = None
        lib.exact_dft.argtypes = [i, sz, sz, vp, vp]
        lib.exact_dft.restype = None
        lib.exact_bin.argtypes = [i, sz, sz, sz, vp, vp]
        lib.exact_bin.restype = None
        _LIB = lib
    return _LIB


def _suf(a):
    if a.dtype == np.complex64:
        return "f32"
    if a.dtype == np.complex128:
        return "f64"
    raise TypeError("oracle: only complex64 / complex128 (Type.hs:28-30)")


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def adhoc_fft(sign, a, threads=1):
    """Un-normalised DFT along the innermost axis (Adhoc.hs:37-48)."""
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    n = a.shape[-1] if a.ndim else 1
    batch = a.size // n if n else 0
    if a.size:
        getattr(_lib(), "adhoc_fft_" + _suf(a))(sign, batch, n, _ptr(a), _ptr(out), threads)
    return out


def adhoc_fft2d(sign, a, threads=1):
    a = np.ascontiguousarray(a)
    assert a.ndim == 2
    out = np.empty_like(a)
    if a.size:
        getattr(_lib(), "adhoc_fft2d_" + _suf(a))(sign, a.shape[0], a.shape[1], _ptr(a), _ptr(out), threads)
    return out


def adhoc_fft3d(sign, a, threads=1):
    a = np.ascontiguousarray(a)
    assert a.ndim == 3
    out = np.empty_like(a)
    if a.size:
        getattr(_lib(), "adhoc_fft3d_" + _suf(a))(sign, a.shape[0], a.shape[1], a.shape[2], _ptr(a), _ptr(out), threads)
    return out


def _scaled(mode, out, scale):
    # FFT.hs:82-84 etc.: case mode of Inverse -> A.map (/scale) (go arr)
    if mode == INVERSE:
        real = out.real.dtype.type
        out = (out / real(scale)).astype(out.dtype)
    return out


def fft(mode, a, threads=1):
    """FFT.hs:63-84: innermost-axis transform; Inverse divides by the innermost length."""
    return _scaled(mode, adhoc_fft(sign_of_mode(mode), a, threads), a.shape[-1])


def fft1D(mode, a, threads=1):
    """FFT.hs:92-111."""
    assert a.ndim == 1
    return _scaled(mode, adhoc_fft(sign_of_mode(mode), a, threads), a.shape[0])


def fft2D(mode, a, threads=1):
    """FFT.hs:119-142: scale = size arr."""
    return _scaled(mode, adhoc_fft2d(sign_of_mode(mode), a, threads), a.size)


def fft3D(mode, a, threads=1):
    """FFT.hs:150-173: scale = size arr."""
    return _scaled(mode, adhoc_fft3d(sign_of_mode(mode), a, threads), a.size)


def _dft_definition(inverse, a):
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    n = a.shape[-1]
    if a.size:
        getattr(_lib(), "dft_definition_" + _suf(a))(inverse, a.size // n, n, _ptr(a), _ptr(out))
    return out


def dft(a):
    """DFT.hs:42-45 `dft` (roots: DFT/Roots.hs:26-36) along the innermost axis, in the precision of `a`.  O(n^2)."""
    return _dft_definition(0, a)


def idft(a):
    """DFT.hs:50-59 `idft` (roots: DFT/Roots.hs:41-51; divides by n), in the precision of `a`.  O(n^2)."""
    return _dft_definition(1, a)


def exact_dft(sign, a):
    """Long-double O(n^2) DFT along the innermost axis; returns clongdouble."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    n = a.shape[-1]
    out = np.empty(a.shape, dtype=np.clongdouble)
    if a.size:
        _lib().exact_dft(sign, a.size // n, n, _ptr(a), _ptr(out))
    return out


def exact_bin(sign, a, k, stride=1, n=None):
    """One exact output bin X[k] of the length-n transform over a.flat[::stride]."""
    a = np.ascontiguousarray(a, dtype=np.complex128).reshape(-1)
    if n is None:
        n = a.size // stride
    out = np.empty(1, dtype=np.clongdouble)
    _lib().exact_bin(sign, n, stride, k, _ptr(a), _ptr(out))
    return out[0]


def rel_l2(y, ref):
    """The parity metric of BASELINE.json: ||y-ref||_2 / ||ref||_2 over the whole array."""
    y = np.asarray(y).astype(np.clongdouble)
    ref = np.asarray(ref).astype(np.clongdouble)
    den = np.sqrt(np.sum(np.abs(ref) ** 2))
    num = np.sqrt(np.sum(np.abs(y - ref) ** 2))
    return float(num / den) if den > 0 else float(num)
