"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/b200fft.h declares, and fails loudly (no fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200fft.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:b200fft|accfft_)\w+)\s*\(", src)))


def test_header_declares_the_cufft_replacements():
    syms = declared_symbols()
    for s in ("b200fftPlan1d", "b200fftPlan2d", "b200fftPlan3d", "b200fftPlanMany1d", "b200fftExec", "b200fftDestroy",
              "b200fftErrorString", "accfft_fft", "accfft_fft1D", "accfft_fft2D", "accfft_fft3D"):
        assert s in syms


def test_library_exports_every_declared_symbol(af):
    L = af.lib()
    for s in declared_symbols():
        assert hasattr(L, s), "libb200fft.so does not export %s" % s


def test_library_has_no_cufft_or_cublas_dependency():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "accelerate_fft_b200", "libb200fft.so")], capture_output=True, text=True).stdout
    assert "cufft" not in out.lower() and "cublas" not in out.lower()


def test_error_strings(af):
    L = af.lib()
    assert L.b200fftErrorString(0) == b"B200FFT_SUCCESS"
    assert b"NO_DEVICE" in L.b200fftErrorString(11)
    assert b"UNKNOWN" in L.b200fftErrorString(12345)


def test_argument_validation_without_device(af):
    L = af.lib()
    h = ctypes.c_void_p()
    assert L.b200fftPlanMany1d(ctypes.byref(h), 1024, 4, 0x1234) == 3      # INVALID_TYPE
    assert L.b200fftPlanMany1d(ctypes.byref(h), 0, 4, af.C2C) == 8         # INVALID_SIZE
    assert L.b200fftPlan2d(ctypes.byref(h), -1, 4, af.Z2Z) == 8
    assert L.b200fftPlanMany1d(None, 8, 1, af.C2C) == 4                    # INVALID_VALUE
    assert L.b200fftExec(None, None, None, -1, None) == 1                  # INVALID_PLAN
    assert L.b200fftDestroy(None) == 1


def test_no_cpu_fallback(af):
    """Without a GPU the product must refuse, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(af.B200FFTError) as ei:
        af.run_host("fft", "Forward", np.ones((2, 8), np.complex64))
    assert ei.value.status == 11
    with pytest.raises(af.B200FFTError) as ei:
        af.run_host_seq("fft", ["Forward", "Inverse"], np.ones((2, 8), np.complex64))
    assert ei.value.status == 11
    with pytest.raises(RuntimeError):
        af.fft("Forward", torch.ones(8, dtype=torch.complex64))
    h = ctypes.c_void_p()
    assert af.lib().b200fftPlanMany1d(ctypes.byref(h), 1024, 4, af.C2C) == 11


def test_python_surface_mirrors_reference(af):
    import torch
    assert af.signOfMode(af.Forward) == -1 and af.signOfMode(af.Reverse) == 1 and af.signOfMode(af.Inverse) == 1
    with pytest.raises((TypeError, RuntimeError)):
        af.fft("Forward", torch.ones(8, dtype=torch.float32))
    with pytest.raises(ValueError):
        af.fft2D("Forward", torch.ones(8, dtype=torch.complex64))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "accelerate_fft_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def _build_c_smoke(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "accelerate_fft_b200")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L" + libdir, "-lb200fft", "-lm",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr          # include/b200fft.h is valid, warning-free C99
    return exe


def test_header_is_c_and_the_library_links_from_plain_c(tmp_path):
    """include/b200fft.h compiles as pedantic C99 and a plain-C host (no CUDA headers, no C++, no Python) links the library and
    calls it; without a GPU the call must fail loudly with B200FFT_NO_DEVICE (exit code 3), never fall back."""
    import subprocess
    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "NO_DEVICE" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_plain_c_host_computes_a_correct_transform(tmp_path):
    import subprocess
    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "max abs error" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_slab_entry_point_validates_before_touching_a_device():
    """b200fftPlanSlab3d (the multi-GPU entry point): argument errors are reported as such with or without a GPU; a valid
    request on a box without a device is B200FFT_NO_DEVICE -- there is no host path behind this entry either."""
    import torch
    from accelerate_fft_b200._lib import ALLGATHER_FN, lib
    L = lib()
    calls = []

    def allgather(_ctx, send, recv, nbytes):
        calls.append(nbytes)
        ctypes.memmove(recv, send, nbytes)
        return 0

    cb = ALLGATHER_FN(allgather)
    h = ctypes.c_void_p()
    C2C = 0x29
    assert L.b200fftPlanSlab3d(ctypes.byref(h), 64, 64, 64, C2C, 0, 0, 0, cb, None) == 4        # nranks < 1: INVALID_VALUE
    assert L.b200fftPlanSlab3d(ctypes.byref(h), 64, 64, 64, C2C, 2, 2, 0, cb, None) == 4        # rank out of range
    assert L.b200fftPlanSlab3d(ctypes.byref(h), 64, 64, 64, 0x2a, 0, 1, 0, cb, None) == 3       # not C2C / Z2Z: INVALID_TYPE
    assert L.b200fftPlanSlab3d(ctypes.byref(h), 64, 63, 64, C2C, 0, 2, 0, cb, None) == 8        # H not divisible: INVALID_SIZE
    assert L.b200fftPlanSlab3d(ctypes.byref(h), 48, 32, 16, C2C, 0, 1, 0, cb, None) == 16       # D not a power of two: NOT_SUPPORTED
    assert L.b200fftPlanSlab3d(ctypes.byref(h), 4096, 64, 64, C2C, 0, 1, 0, cb, None) == 16     # D above the scatter pass's reach
    assert not calls                                                                            # nothing was exchanged for a bad request
    if not torch.cuda.is_available():
        assert L.b200fftPlanSlab3d(ctypes.byref(h), 64, 64, 64, C2C, 0, 1, 1, cb, None) == 11   # NO_DEVICE
    assert L.b200fftExecSlab(None, None, None, -1, 1.0, 0, None) == 1                           # INVALID_PLAN
    assert L.b200fftDestroySlab(None) == 1
