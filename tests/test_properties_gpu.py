"""The reference's own property suite (test/Test/FFT.hs:39-236) re-expressed with pytest+hypothesis
against the CUDA path: same seven properties, same generators (test/Test/Base.hs:35-58: components
uniform in [-1,1]; DIM1 n in [1,1024]; DIM2 x in [1,128], y in [1,48]; DIM3 x in [1,64], y in [1,32],
z in [1,16]), all three modes, both element types -- at a far tighter tolerance than `~~~`."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import bar, rel_l2

pytestmark = pytest.mark.gpu

SET = settings(max_examples=50, deadline=None, suppress_health_check=list(HealthCheck))   # CI runs 50 tests (ci.yml:117)

dim1 = st.tuples(st.integers(1, 1024))
dim2 = st.tuples(st.integers(1, 48), st.integers(1, 128))
dim3 = st.tuples(st.integers(1, 16), st.integers(1, 32), st.integers(1, 64))
modes = st.sampled_from(["Forward", "Reverse", "Inverse"])
dtypes = st.sampled_from([np.complex64, np.complex128])
seeds = st.integers(0, 2**31 - 1)


def arr(seed, shape, dtype, k=1):
    rng = np.random.default_rng(seed)
    return [(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dtype) for _ in range(k)]


def T(af, rank_full, mode, x):
    """fft1D / fft2D / fft3D by rank -- the entry points homogeneity/additivity/inverse are tested with."""
    import torch
    f = {1: af.fft1D, 2: af.fft2D, 3: af.fft3D}[rank_full]
    return f(mode, torch.from_numpy(x).cuda()).cpu().numpy()


def F(af, mode, x):
    """`fft` (innermost axis) -- what reverse/conjugate/isometry/unitarity use on DIM2/DIM3 (:66-69,75-78)."""
    import torch
    return af.fft(mode, torch.from_numpy(x).cuda()).cpu().numpy()


def rev(a):
    n = a.shape[-1]
    return np.ascontiguousarray(a[..., (-np.arange(n)) % n])


shapes = st.one_of(dim1, dim2, dim3)


@SET
@given(shapes, modes, dtypes, seeds)
def test_homogeneity(af, shape, mode, dtype, seed):
    (x,) = arr(seed, shape, dtype)
    c = dtype(complex(*np.random.default_rng(seed + 1).uniform(-1, 1, 2)))
    lhs, rhs = T(af, len(shape), mode, (c * x).astype(dtype)), c * T(af, len(shape), mode, x)
    assert rel_l2(lhs, rhs) <= bar(dtype, x.size)


@SET
@given(shapes, modes, dtypes, seeds)
def test_additivity(af, shape, mode, dtype, seed):
    x, y = arr(seed, shape, dtype, 2)
    lhs = T(af, len(shape), mode, (x + y).astype(dtype))
    rhs = T(af, len(shape), mode, x) + T(af, len(shape), mode, y)
    assert rel_l2(lhs, rhs) <= bar(dtype, x.size)


@SET
@given(shapes, dtypes, seeds)
def test_inverse(af, shape, dtype, seed):
    (x,) = arr(seed, shape, dtype)
    z = T(af, len(shape), "Inverse", T(af, len(shape), "Forward", x))
    assert rel_l2(z, x) <= 2 * bar(dtype, x.size)


@SET
@given(shapes, modes, dtypes, seeds)
def test_reverse(af, shape, mode, dtype, seed):
    (x,) = arr(seed, shape, dtype)
    assert rel_l2(rev(F(af, mode, x)), F(af, mode, rev(x))) <= bar(dtype, shape[-1])


@SET
@given(shapes, modes, dtypes, seeds)
def test_conjugate(af, shape, mode, dtype, seed):
    (x,) = arr(seed, shape, dtype)
    assert rel_l2(np.conj(F(af, mode, x)), F(af, mode, np.conj(rev(x)))) <= bar(dtype, shape[-1])


@SET
@given(shapes, st.sampled_from(["Forward", "Reverse"]), dtypes, seeds)
def test_isometry(af, shape, mode, dtype, seed):
    (x,) = arr(seed, shape, dtype)
    n = shape[-1]
    lhs = np.sqrt((np.abs(F(af, mode, x).astype(np.complex128)) ** 2).sum(-1))
    rhs = np.sqrt(n) * np.sqrt((np.abs(x.astype(np.complex128)) ** 2).sum(-1))
    assert np.allclose(lhs, rhs, rtol=50 * bar(dtype, n), atol=1e-30)


@SET
@given(shapes, st.sampled_from(["Forward", "Reverse"]), dtypes, seeds)
def test_unitarity(af, shape, mode, dtype, seed):
    x, y = arr(seed, shape, dtype, 2)
    n = shape[-1]
    fx, fy = F(af, mode, x).astype(np.complex128), F(af, mode, y).astype(np.complex128)
    lhs = (fx * np.conj(fy)).sum(-1)
    rhs = n * (x.astype(np.complex128) * np.conj(y.astype(np.complex128))).sum(-1)
    scale = n * np.sqrt((np.abs(x) ** 2).sum(-1) * (np.abs(y) ** 2).sum(-1)) + 1e-30
    assert np.all(np.abs(lhs - rhs) / scale <= 50 * bar(dtype, n))
