import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- shared helpers -----------------------------------------------------------------------

def rand_complex(rng, shape, dtype):
    """re, im ~ U(-1, 1): the reference's generator (test/Test/Base.hs:35-42)."""
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dtype)


def rel_l2(y, ref):
    y = np.asarray(y, dtype=np.complex128).ravel()
    ref = np.asarray(ref, dtype=np.complex128).ravel()
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(y - ref) / den) if den > 0 else float(np.linalg.norm(y - ref))


def bar(dtype, npoints):
    """BASELINE.json north_star tolerance: rel-L2 <= 1e-5*log2(N) (Float) / 1e-13*log2(N) (Double),
    N = points of one transform."""
    lg = max(1.0, math.log2(max(2, npoints)))
    return (1e-5 if np.dtype(dtype) == np.complex64 else 1e-13) * lg


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def af():
    import accelerate_fft_b200 as m
    m.lib()  # fail loudly if the extension is missing
    return m
