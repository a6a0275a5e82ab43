"""accelerate_fft_b200 -- B200-native drop-in for the GPU hot path of
Data.Array.Accelerate.Math.FFT (AccelerateHS/accelerate-fft).

The product is the C-ABI library libb200fft.so (include/b200fft.h, csrc/): hand-written sm_100a
Stockham kernels behind a cuFFT-shaped plan/exec/destroy interface that the reference's PTX
backend binds instead of Hackage `cufft` (haskell/).

This Python package is the host-side mirror of the reference's public API used by the tests and
the bench -- same names, argument meaning and error behaviour as
/root/reference/src/Data/Array/Accelerate/Math/FFT.hs:
    fft(mode, arr)    DFT along the innermost axis of a rank>=1 array   (FFT.hs:63-84)
    fft1D(mode, arr)  vector                                            (FFT.hs:92-111)
    fft2D(mode, arr)  matrix, both axes                                 (FFT.hs:119-142)
    fft3D(mode, arr)  cube, all three axes                              (FFT.hs:150-173)
with mode in {Forward, Reverse, Inverse} (Mode.hs:15-19) over complex64 / complex128 arrays
(Type.hs:28-39).  torch is used only for device memory and streams.
"""
import ctypes

from . import _lib
from ._lib import B200FFTError, C2C, Z2Z, FORWARD, INVERSE, build, lib  # noqa: F401

Forward, Reverse, Inverse = "Forward", "Reverse", "Inverse"
_MODE = {Forward: 0, Reverse: 1, Inverse: 2}


def signOfMode(mode):
    """Mode.hs:21-26."""
    return {Forward: -1, Reverse: 1, Inverse: 1}[mode]


def _torch():
    import torch
    return torch


def _type_of(t):
    torch = _torch()
    if t.dtype == torch.complex64:
        return C2C
    if t.dtype == torch.complex128:
        return Z2Z
    raise TypeError("accelerate_fft_b200: only Complex Float / Complex Double arrays (Type.hs:28-30), got %s" % t.dtype)


def _prep(arr):
    torch = _torch()
    if not isinstance(arr, torch.Tensor):
        raise TypeError("expected a torch tensor holding the Accelerate array")
    if not arr.is_cuda:
        raise RuntimeError("accelerate_fft_b200 runs on the GPU only; got a %s tensor (no CPU fallback)" % arr.device)
    typ = _type_of(arr)
    # lazy conj / neg views share storage with the un-conjugated tensor: materialise them, the C ABI sees raw memory
    arr = arr.resolve_conj().resolve_neg().contiguous()
    out = torch.empty_like(arr)
    stream = torch.cuda.current_stream(arr.device).cuda_stream
    return arr, out, typ, ctypes.c_void_p(stream)


def _shape(arr):
    return (ctypes.c_int64 * arr.dim())(*arr.shape)


def fft(mode, arr):
    """Innermost-axis DFT of an array of rank >= 1 (FFT.hs:63-84)."""
    if arr.dim() < 1:
        raise ValueError("fft needs rank >= 1")
    arr, out, typ, stream = _prep(arr)
    torch = _torch()
    with torch.cuda.device(arr.device):
        _lib.check(lib().accfft_fft(_MODE[mode], arr.dim(), _shape(arr), typ, arr.data_ptr(), out.data_ptr(), stream), "fft")
    return out


def fft1D(mode, arr):
    if arr.dim() != 1:
        raise ValueError("fft1D needs a DIM1 array")
    arr, out, typ, stream = _prep(arr)
    torch = _torch()
    with torch.cuda.device(arr.device):
        _lib.check(lib().accfft_fft1D(_MODE[mode], arr.shape[0], typ, arr.data_ptr(), out.data_ptr(), stream), "fft1D")
    return out


def fft2D(mode, arr):
    if arr.dim() != 2:
        raise ValueError("fft2D needs a DIM2 array")
    arr, out, typ, stream = _prep(arr)
    torch = _torch()
    with torch.cuda.device(arr.device):
        _lib.check(lib().accfft_fft2D(_MODE[mode], arr.shape[0], arr.shape[1], typ, arr.data_ptr(), out.data_ptr(), stream), "fft2D")
    return out


def fft3D(mode, arr):
    if arr.dim() != 3:
        raise ValueError("fft3D needs a DIM3 array")
    arr, out, typ, stream = _prep(arr)
    torch = _torch()
    with torch.cuda.device(arr.device):
        _lib.check(lib().accfft_fft3D(_MODE[mode], arr.shape[0], arr.shape[1], arr.shape[2], typ, arr.data_ptr(),
                                      out.data_ptr(), stream), "fft3D")
    return out


# ---- Data.Array.Accelerate.Math.DFT.Centre (DFT/Centre.hs): the step on either side of the transform ------------

def _centre_like(arr, rank, call, what):
    if arr.dim() != rank:
        raise ValueError("%s needs a DIM%d array" % (what, rank))
    arr, out, typ, stream = _prep(arr)
    torch = _torch()
    with torch.cuda.device(arr.device):
        _lib.check(call(arr, out, typ, stream), what)
    return out


def _centre(arr, rank):
    return _centre_like(arr, rank, lambda a, o, t, s: lib().accfft_centre(rank, _shape(a), t, a.data_ptr(), o.data_ptr(), s),
                        "centre%dD" % rank)


def _shift(arr, rank, inverse):
    return _centre_like(arr, rank, lambda a, o, t, s: lib().accfft_shift(rank, _shape(a), t, inverse, a.data_ptr(), o.data_ptr(), s),
                        ("ishift%dD" if inverse else "shift%dD") % rank)


def centre1D(arr):
    """(-1)^x * arr (Centre.hs:36-43)."""
    return _centre(arr, 1)


def centre2D(arr):
    """(-1)^(y+x) * arr (Centre.hs:47-54)."""
    return _centre(arr, 2)


def centre3D(arr):
    """(-1)^(z+y+x) * arr (Centre.hs:58-65)."""
    return _centre(arr, 3)


def shift1D(arr):
    """out[i] = arr[(i + n/2 + odd n) rem n] (Centre.hs:70-79)."""
    return _shift(arr, 1, 0)


def shift2D(arr):
    return _shift(arr, 2, 0)     # Centre.hs:98-112


def shift3D(arr):
    return _shift(arr, 3, 0)     # Centre.hs:134-151


def ishift1D(arr):
    """out[i] = arr[(i + n/2) rem n] (Centre.hs:84-94)."""
    return _shift(arr, 1, 1)


def ishift2D(arr):
    return _shift(arr, 2, 1)     # Centre.hs:116-130


def ishift3D(arr):
    return _shift(arr, 3, 1)     # Centre.hs:155-164


def fft_centred(mode, arr):
    """shiftND (fftND mode arr) for a DIM1/2/3 array in one call -- zero frequency in the middle; for even extents this is
    fftND mode (centreND arr) (Centre.hs:17-19).  Power-of-two extents: the rotation rides on the stores of each axis' last
    butterfly pass (b200fftExecShifted), no extra pass over the array."""
    if arr.dim() not in (1, 2, 3):
        raise ValueError("fft_centred needs a DIM1, DIM2 or DIM3 array")
    arr, out, typ, stream = _prep(arr)
    torch = _torch()
    with torch.cuda.device(arr.device):
        _lib.check(lib().accfft_fft_centred(arr.dim(), _MODE[mode], _shape(arr), typ, arr.data_ptr(), out.data_ptr(), stream),
                   "fft_centred")
    return out


def run_host(kind, mode, a):
    """Host-buffer entry (numpy in, numpy out): H2D copy + transform + D2H copy through the C ABI.
    kind in {"fft","fft1D","fft2D","fft3D"}."""
    import numpy as np
    a = np.ascontiguousarray(a)
    if a.dtype == np.complex64:
        typ = C2C
    elif a.dtype == np.complex128:
        typ = Z2Z
    else:
        raise TypeError("only complex64 / complex128")
    out = np.empty_like(a)
    k = {"fft": 0, "fft1D": 1, "fft2D": 2, "fft3D": 3}[kind]
    shape = (ctypes.c_int64 * a.ndim)(*a.shape)
    _lib.check(lib().accfft_run_host(k, _MODE[mode], a.ndim, shape, typ, a.ctypes.data, out.ctypes.data), kind)
    return out


def run_host_seq(kind, modes, h_in, h_out=None):
    """Chain of transforms over a HOST array through the C ABI (accfft_run_host_seq): copy in, `modes` applied one
    after the other on the device, copy out; for kind "fft" the rows flow through the device in chunks so both
    copies overlap the kernels.  h_in / h_out: numpy arrays or CPU torch tensors (pinned for asynchronous DMA)."""
    import numpy as np
    torch = None
    if not isinstance(h_in, np.ndarray):
        torch = _torch()
        if h_in.is_cuda:
            raise RuntimeError("run_host_seq takes host arrays")
        h_in = h_in.resolve_conj().resolve_neg().contiguous()
        if h_out is None:
            h_out = torch.empty_like(h_in)
        typ = _type_of(h_in)
        pin, pout, shp = h_in.data_ptr(), h_out.data_ptr(), tuple(h_in.shape)
    else:
        h_in = np.ascontiguousarray(h_in)
        if h_in.dtype == np.complex64:
            typ = C2C
        elif h_in.dtype == np.complex128:
            typ = Z2Z
        else:
            raise TypeError("only complex64 / complex128")
        if h_out is None:
            h_out = np.empty_like(h_in)
        pin, pout, shp = h_in.ctypes.data, h_out.ctypes.data, h_in.shape
    k = {"fft": 0, "fft1D": 1, "fft2D": 2, "fft3D": 3}[kind]
    ms = (ctypes.c_int * len(modes))(*[_MODE[m] for m in modes])
    shape = (ctypes.c_int64 * len(shp))(*shp)
    _lib.check(lib().accfft_run_host_seq(k, ms, len(modes), len(shp), shape, typ, pin, pout), kind)
    return h_out


def set_fused_inverse(on):
    lib().accfft_set_fused_inverse(1 if on else 0)


class Plan:
    """Thin wrapper over the cuFFT-shaped ABI (b200fftPlan* / b200fftExec / b200fftDestroy)."""

    def __init__(self, kind, dims, typ, batch=1):
        self.h = ctypes.c_void_p()
        L = lib()
        if kind == "1d":
            st = L.b200fftPlan1d(ctypes.byref(self.h), dims[0], typ, batch)
        elif kind == "many":
            st = L.b200fftPlanMany1d(ctypes.byref(self.h), dims[0], batch, typ)
        elif kind == "2d":
            st = L.b200fftPlan2d(ctypes.byref(self.h), dims[0], dims[1], typ)
        elif kind == "3d":
            st = L.b200fftPlan3d(ctypes.byref(self.h), dims[0], dims[1], dims[2], typ)
        elif kind == "axis":   # dims = (outer, n, inner)
            st = L.b200fftPlanAxis(ctypes.byref(self.h), dims[0], dims[1], dims[2], typ)
        else:
            raise ValueError(kind)
        _lib.check(st, "plan " + kind)
        self.typ = typ

    def exec(self, src, dst, direction, stream=None, scale=None):
        torch = _torch()
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        if scale is None:
            st = lib().b200fftExec(self.h, src.data_ptr(), dst.data_ptr(), direction, ctypes.c_void_p(stream))
        else:
            st = lib().b200fftExecScaled(self.h, src.data_ptr(), dst.data_ptr(), direction, float(scale), ctypes.c_void_p(stream))
        _lib.check(st, "exec")

    def exec_scatter(self, src, out_ptrs, out_outer_stride, out_n_stride, direction, stream=None, scale=1.0):
        """b200fftExecScatter: the pass's stores go to the buffers `out_ptrs` (raw device addresses, possibly of
        peer GPUs), split by output index."""
        torch = _torch()
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        arr = (ctypes.c_void_p * len(out_ptrs))(*out_ptrs)
        src_ptr = src.data_ptr() if hasattr(src, "data_ptr") else int(src)
        st = lib().b200fftExecScatter(self.h, src_ptr, arr, len(out_ptrs), out_outer_stride, out_n_stride, direction,
                                      float(scale), ctypes.c_void_p(stream))
        _lib.check(st, "exec_scatter")

    @property
    def num_passes(self):
        return lib().b200fftNumPasses(self.h)

    @property
    def scratch_bytes(self):
        return lib().b200fftScratchBytes(self.h)

    def describe(self):
        buf = ctypes.create_string_buffer(8192)
        lib().b200fftDescribe(self.h, buf, 8192)
        return buf.value.decode()

    def destroy(self):
        if self.h:
            lib().b200fftDestroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def kernel_launches():
    return lib().b200fftKernelLaunches()
