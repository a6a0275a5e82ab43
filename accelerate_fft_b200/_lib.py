"""ctypes binding of libb200fft.so (include/b200fft.h).  Fails loudly when the library is
missing: there is no CPU or cuFFT fallback in this package."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fft.so")

C2C = 0x29
Z2Z = 0x69
FORWARD = -1
INVERSE = 1

_lib = None
# b200fftAllgatherFn: int (*)(void* ctx, const void* send, void* recv, size_t bytes)
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

EXPORTS = [
    "b200fftPlan1d", "b200fftPlan2d", "b200fftPlan3d", "b200fftPlanMany1d", "b200fftExec",
    "b200fftExecScaled", "b200fftDestroy", "b200fftErrorString", "b200fftScratchBytes",
    "b200fftNumPasses", "b200fftKernelLaunches", "b200fftDescribe",
    "accfft_fft", "accfft_fft1D", "accfft_fft2D", "accfft_fft3D", "accfft_run_host", "accfft_run_host_seq",
    "accfft_set_fused_inverse", "accfft_plan_cache_size", "accfft_plan_cache_clear",
    "b200fftPlanAxis", "b200fftSlabPack", "b200fftSlabUnpack", "b200fftTrimScratch",
    "b200fftExecShifted", "accfft_centre", "accfft_shift", "accfft_fft_centred",
    "b200fftExecScatterOn", "b200fftPlanAxisView", "b200fftHasExperimental",
    "b200fftPlanSlab3d", "b200fftExecSlab", "b200fftSlabNaturalBuffer", "b200fftSlabTune", "b200fftDestroySlab",
    "b200fftExecScatter", "b200fftPeerAlloc", "b200fftPeerFree", "b200fftPeerExport", "b200fftPeerOpen", "b200fftPeerClose",
]


class B200FFTError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        try:
            msg = lib().b200fftErrorString(status).decode()
        except Exception:  # pragma: no cover
            msg = "status %d" % status
        super().__init__("%s%s" % (what + ": " if what else "", msg))


def build(verbose=False):
    """Compile libb200fft.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libb200fft.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "accelerate_fft_b200: %s is missing.  Build it with `make -C accelerate_fft_b200/csrc` "
            "(or __graft_entry__.build()).  There is no fallback path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i, d = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    pvp = ctypes.POINTER(ctypes.c_void_p)
    L.b200fftPlan1d.argtypes = [pvp, i64, i, i64]
    L.b200fftPlan2d.argtypes = [pvp, i64, i64, i]
    L.b200fftPlan3d.argtypes = [pvp, i64, i64, i64, i]
    L.b200fftPlanMany1d.argtypes = [pvp, i64, i64, i]
    L.b200fftExec.argtypes = [vp, vp, vp, i, vp]
    L.b200fftExecScaled.argtypes = [vp, vp, vp, i, d, vp]
    L.b200fftDestroy.argtypes = [vp]
    L.b200fftErrorString.argtypes = [i]
    L.b200fftErrorString.restype = ctypes.c_char_p
    L.b200fftScratchBytes.argtypes = [vp]
    L.b200fftScratchBytes.restype = ctypes.c_size_t
    L.b200fftNumPasses.argtypes = [vp]
    L.b200fftKernelLaunches.argtypes = []
    L.b200fftKernelLaunches.restype = i64
    L.b200fftDescribe.argtypes = [vp, ctypes.c_char_p, i]
    pi64 = ctypes.POINTER(ctypes.c_int64)
    L.accfft_fft.argtypes = [i, i, pi64, i, vp, vp, vp]
    L.accfft_fft1D.argtypes = [i, i64, i, vp, vp, vp]
    L.accfft_fft2D.argtypes = [i, i64, i64, i, vp, vp, vp]
    L.accfft_fft3D.argtypes = [i, i64, i64, i64, i, vp, vp, vp]
    L.accfft_run_host.argtypes = [i, i, i, pi64, i, vp, vp]
    L.accfft_run_host_seq.argtypes = [i, ctypes.POINTER(ctypes.c_int), i, i, pi64, i, vp, vp]
    L.accfft_set_fused_inverse.argtypes = [i]
    L.accfft_set_fused_inverse.restype = None
    L.accfft_plan_cache_clear.restype = None
    L.b200fftPlanAxis.argtypes = [pvp, i64, i64, i64, i]
    L.b200fftSlabPack.argtypes = [i, vp, vp, i64, i64, i64, i, vp]
    L.b200fftSlabUnpack.argtypes = [i, vp, vp, i64, i64, i64, i, vp]
    L.b200fftExecShifted.argtypes = [vp, vp, vp, i, d, vp]
    L.accfft_centre.argtypes = [i, pi64, i, vp, vp, vp]
    L.accfft_shift.argtypes = [i, pi64, i, i, vp, vp, vp]
    L.accfft_fft_centred.argtypes = [i, i, pi64, i, vp, vp, vp]
    L.b200fftExecScatter.argtypes = [vp, vp, pvp, i, i64, i64, i, d, vp]
    L.b200fftExecScatterOn.argtypes = [vp, vp, pvp, i, i64, i64, i, d, i, vp]
    L.b200fftPlanAxisView.argtypes = [pvp, i64, i64, i64, i64, i64, i]
    L.b200fftPlanSlab3d.argtypes = [pvp, i64, i64, i64, i, i, i, i, ALLGATHER_FN, vp]
    L.b200fftExecSlab.argtypes = [vp, vp, vp, i, d, i, vp]
    L.b200fftSlabNaturalBuffer.argtypes = [vp, pvp]
    L.b200fftSlabTune.argtypes = [vp, i, i, i, i]
    L.b200fftDestroySlab.argtypes = [vp]
    L.b200fftPeerAlloc.argtypes = [pvp, ctypes.c_size_t]
    L.b200fftPeerFree.argtypes = [vp]
    L.b200fftPeerExport.argtypes = [vp, ctypes.c_char_p]
    L.b200fftPeerOpen.argtypes = [ctypes.c_char_p, pvp]
    L.b200fftPeerClose.argtypes = [vp]
    _lib = L
    return L


def check(status, what=""):
    if status != 0:
        raise B200FFTError(status, what)
