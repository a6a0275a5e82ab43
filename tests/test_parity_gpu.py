"""GPU parity tests: the CUDA path, called through the C ABI (accfft_* / b200fft*), against the
oracle (oracle/: C restatement of Adhoc.hs + exact long-double DFT), the committed known-answer
vectors, and -- at BASELINE.json's full sizes -- size-independent properties.

Tolerance everywhere: relative L2 <= 1e-5*log2(N) for Float, 1e-13*log2(N) for Double
(N = points of one transform), as stated in BASELINE.json's north_star."""
import os

import numpy as np
import pytest

from conftest import bar, rand_complex, rel_l2

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat.npz"))
DTYPES = [np.complex64, np.complex128]
MODES = ["Forward", "Reverse", "Inverse"]


def gpu(af, kind, mode, x):
    import torch
    f = {"fft": af.fft, "fft1D": af.fft1D, "fft2D": af.fft2D, "fft3D": af.fft3D}[kind]
    return f(mode, torch.from_numpy(np.ascontiguousarray(x)).cuda()).cpu().numpy()


def _is_5smooth(n):
    for p in (2, 3, 5):
        while n % p == 0:
            n //= p
    return n == 1


# ---- committed known-answer vectors ------------------------------------------------------------

@pytest.mark.parametrize("dtype", DTYPES)
def test_golden_vectors(af, dtype):
    suf = "_f32" if dtype == np.complex64 else ""
    for k in GOLD.files:
        if not k.startswith("in"):
            continue
        rank = int(k[2])
        x = GOLD[k].astype(dtype)
        ref = GOLD["fwd" + k[2:] + suf]
        if rank == 1:
            y = gpu(af, "fft", "Forward", x)
            n = x.shape[-1]
        else:
            y = gpu(af, "fft%dD" % rank, "Forward", x)
            n = x.size
        assert rel_l2(y, ref) <= bar(dtype, n), k


# ---- against the oracle on seeded inputs ---------------------------------------------------------

@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_fft_pow2_vs_oracle(af, oracle, dtype, mode):
    rng = np.random.default_rng(11)
    for lg in range(0, 15):
        n = 1 << lg
        if dtype == np.complex128 and n > 8192:
            continue
        x = rand_complex(rng, (5, n), dtype)
        y = gpu(af, "fft", mode, x)
        ref = oracle.fft(mode, x, threads=4)
        assert rel_l2(y, ref) <= bar(dtype, n), (n, mode)
        ex = oracle.exact_dft(oracle.sign_of_mode(mode), x[:1]) if n <= 2048 else None
        if ex is not None:
            if mode == "Inverse":
                ex = ex / n
            assert rel_l2(y[:1], ex.astype(np.complex128)) <= bar(dtype, n) / 2, (n, mode)


@pytest.mark.parametrize("n,dtype", [(4096, np.complex128), (16384, np.complex64), (8192, np.complex128), (8192, np.complex64)])
@pytest.mark.parametrize("mode", ["Forward", "Inverse"])
def test_persistent_tma_row_kernel_vs_oracle(af, oracle, n, dtype, mode):
    """Batches large enough (>= 2 tiles per SM) take the persistent TMA-fed ring kernel (ring_kernel.cuh);
    an odd batch exercises uneven tile counts per CTA; with the ring disabled the plan falls back to the
    plain kernel with the same result."""
    import torch
    rng = np.random.default_rng(21)
    batch = 2 * 148 * 3 + 5 if n == 4096 else 2 * 148 + 3      # (the 128 KB lines run with a single-stage ring)
    x = rand_complex(rng, (batch, n), dtype)
    p = af.Plan("many", [n], af.C2C if dtype == np.complex64 else af.Z2Z, batch)
    assert "ring" in p.describe()
    p.destroy()
    y = gpu(af, "fft", mode, x)
    ref = oracle.fft(mode, x, threads=8)
    assert rel_l2(y, ref) <= bar(dtype, n)
    worst = max(rel_l2(y[i], ref[i]) for i in (0, 1, 147, 148, 295, 296, batch - 2, batch - 1))
    assert worst <= bar(dtype, n)
    # the same rows through the plain (non-persistent) kernel must agree bit for bit
    os.environ["B200FFT_NO_RING"] = "1"
    try:
        af.lib().accfft_plan_cache_clear()
        y2 = gpu(af, "fft", mode, x)
    finally:
        os.environ["B200FFT_NO_RING"] = "0"
        af.lib().accfft_plan_cache_clear()
    assert np.array_equal(y, y2)
    if n == 4096:     # the registered alternative (one CTA of two groups alternating over three stages, round 1's default)
        with _env(af, B200FFT_VARIANTS="g4096d=1"):
            p = af.Plan("many", [n], af.Z2Z, batch)
            assert "ring: G=2 NS=3" in p.describe(), p.describe()
            p.destroy()
            assert np.array_equal(y, gpu(af, "fft", mode, x))


@pytest.mark.parametrize("dtype", DTYPES)
def test_fft_every_length_1_to_130_and_reference_range(af, oracle, dtype):
    """cuFFT accepts every length; the reference's PTX suite draws n in [1,1024] (test/Test/Base.hs:44-45)."""
    rng = np.random.default_rng(12)
    sizes = list(range(1, 131)) + [243, 255, 257, 360, 509, 625, 729, 1000, 1009, 1021, 1023]
    for n in sizes:
        x = rand_complex(rng, (3, n), dtype)
        y = gpu(af, "fft", "Forward", x)
        ex = oracle.exact_dft(-1, x).astype(np.complex128)
        assert rel_l2(y, ex) <= bar(dtype, n), n
        if _is_5smooth(n):  # the oracle's chirp branch is only ~1e-4 accurate in Float
            assert rel_l2(y, oracle.fft("Forward", x)) <= bar(dtype, n), n
        z = gpu(af, "fft", "Inverse", y.astype(dtype))
        assert rel_l2(z, x) <= 2 * bar(dtype, n), n


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_prime_radices_17_to_61_one_pass(af, oracle, dtype, mode):
    """Lengths whose largest prime factor is 17 ... 61 (c128: 17, 19, 23) take the one-pass mixed-radix kernel with a prime
    radix of their own instead of Bluestein (generic.cu factor_small): rows against the exact DFT, strided axes through
    fft2D / fft3D against numpy float64; a third of the 992 lengths the reference's suite draws from (test/Test/Base.hs:44-45)."""
    rng = np.random.default_rng(21)
    lens = [34, 38, 46, 58, 62, 17 * 17, 17 * 19, 19 * 23, 23 * 29, 29 * 29, 31 * 31, 3 * 17 * 19, 2 * 17 * 29, 4 * 13 * 19, 1023,
            8 * 31, 32 * 23, 5 * 7 * 29, 37, 41, 43, 47, 53, 59, 61, 2 * 37, 37 * 24, 61 * 16, 53 * 19, 47 * 17, 43 * 23, 41 * 25,
            59 * 13, 3 * 5 * 37, 37 * 41]
    sgn = -1 if mode == "Forward" else 1
    for n in lens:
        p = af.Plan("many", [n], af.C2C if dtype == np.complex64 else af.Z2Z, 5)
        d = p.describe()
        p.destroy()
        # ... when one or two stages do it (with three the measured gain is gone: those lengths stay on Bluestein)
        rad = ({2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 21, 24, 25, 27, 28, 30, 32, 17, 19, 23, 29, 31,
                37, 41, 43, 47, 53, 59, 61} if dtype == np.complex64 else set(range(2, 17)) | {17, 19, 23})
        two_stage = n in rad or any(n % a == 0 and n // a in rad for a in rad)
        assert ("mixed-radix" in d) == two_stage, (n, d)
        x = rand_complex(rng, (5, n), dtype)
        ex = oracle.exact_dft(sgn, x).astype(np.complex128) / (n if mode == "Inverse" else 1)
        assert rel_l2(gpu(af, "fft", mode, x), ex) <= bar(dtype, n), (n, mode)
    for shape in [(34, 57), (323, 40), (96, 17 * 23), (31, 31), (37, 41), (2 * 53, 61 * 4)]:
        x = rand_complex(rng, shape, dtype)
        x128 = x.astype(np.complex128)
        ex = np.fft.fft2(x128) if mode == "Forward" else np.fft.ifft2(x128) * (1 if mode == "Inverse" else x.size)
        assert rel_l2(gpu(af, "fft2D", mode, x), ex) <= bar(dtype, x.size), (shape, mode)
    for shape in [(17, 19, 23), (34, 6, 58), (37, 5, 47)]:
        x = rand_complex(rng, shape, dtype)
        x128 = x.astype(np.complex128)
        ex = np.fft.fftn(x128) if mode == "Forward" else np.fft.ifftn(x128) * (1 if mode == "Inverse" else x.size)
        assert rel_l2(gpu(af, "fft3D", mode, x), ex) <= bar(dtype, x.size), (shape, mode)


@pytest.mark.parametrize("dtype", DTYPES)
def test_large_1d_four_step_vs_library(af, dtype):
    import scipy.fft as sf
    rng = np.random.default_rng(13)
    for n in (1 << 15, 1 << 18, 1 << 21, 3 << 16, 1000003):
        x = rand_complex(rng, (n,), dtype)
        y = gpu(af, "fft1D", "Forward", x)
        ref = sf.fft(x.astype(np.complex128), workers=-1)
        assert rel_l2(y, ref) <= bar(dtype, n), n
        z = gpu(af, "fft1D", "Inverse", y.astype(dtype))
        assert rel_l2(z, x) <= 2 * bar(dtype, n), n


SHAPES_2D = [(1, 1), (1, 16), (16, 1), (8, 8), (48, 128), (128, 48), (37, 64), (64, 37), (100, 100), (31, 17), (512, 256),
             (4096, 16), (16, 4096), (8192, 4), (2, 16384)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_fft2d_vs_oracle(af, oracle, dtype, mode):
    rng = np.random.default_rng(14)
    for shape in SHAPES_2D:
        if dtype == np.complex128 and max(shape) > 8192:
            continue
        x = rand_complex(rng, shape, dtype)
        y = gpu(af, "fft2D", mode, x)
        x128 = x.astype(np.complex128)
        ex = np.fft.fft2(x128) if mode == "Forward" else np.fft.ifft2(x128) * (1 if mode == "Inverse" else x.size)
        assert rel_l2(y, ex) <= bar(dtype, x.size), (shape, mode)
        if all(_is_5smooth(s) for s in shape):
            assert rel_l2(y, oracle.fft2D(mode, x, threads=4)) <= bar(dtype, x.size), (shape, mode)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_fft2d_two_pass_row_pair_plan(af, oracle, dtype, mode):
    """Column lengths of 2 x (longest single-pass column) take the two-pass plan: row pass with the radix-2
    butterfly of rows n2 / n2+H/2 folded into its load, then an in-place H/2-point column pass (plan.cu
    try_pair_2d).  Checked against numpy in double and, on a 4096 x 1024 problem, against the oracle."""
    _needs_experimental(af)
    rng = np.random.default_rng(31)
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    os.environ["B200FFT_PAIR2D"] = "1"     # opt-in plan (measured slower than the default on B200, see plan.cu)
    af.lib().accfft_plan_cache_clear()
    try:
        for shape in [(4096, 1024), (4096, 2048), (8192, 1024)]:
            if dtype == np.complex128 and shape[0] > 4096:
                continue      # c128 columns of 4096 have no single-pass kernel: that shape keeps the four-step plan
            p = af.Plan("2d", list(shape), typ, 1)
            d = p.describe()
            p.destroy()
            assert "pair-rows" in d and "pair-cols" in d and len(d.strip().split("\n")) == 2, d
            x = rand_complex(rng, shape, dtype)
            y = gpu(af, "fft2D", mode, x)
            x128 = x.astype(np.complex128)
            ex = np.fft.fft2(x128) if mode == "Forward" else np.fft.ifft2(x128) * (1 if mode == "Inverse" else x.size)
            assert rel_l2(y, ex) <= bar(dtype, x.size), (shape, mode)
            if shape == (4096, 1024):
                assert rel_l2(y, oracle.fft2D(mode, x, threads=8)) <= bar(dtype, x.size), (shape, mode)
    finally:
        os.environ["B200FFT_PAIR2D"] = "0"
        af.lib().accfft_plan_cache_clear()


def _needs_experimental(af):
    """The opt-in kernels that measured slower than the default plans are only in `make B200FFT_EXPERIMENTAL=1` builds."""
    if not af.lib().b200fftHasExperimental():
        pytest.skip("libb200fft.so built without B200FFT_EXPERIMENTAL=1 (cluster / row-pair / lock-step fused kernels left out)")


class _env:
    """Set planner switches for the duration of a test and drop the cached plans on both sides."""

    def __init__(self, af, **kv):
        self.af, self.kv, self.old = af, kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v
        self.af.lib().accfft_plan_cache_clear()

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        self.af.lib().accfft_plan_cache_clear()


def _np_fft2(mode, x):
    x128 = x.astype(np.complex128)
    return np.fft.fft2(x128) if mode == "Forward" else np.fft.ifft2(x128) * (1 if mode == "Inverse" else x.size)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", [(8192, 128), (8192, 512), (8192, 96)])
def test_band_kernel_l2_fused_column_axis(af, oracle, mode, shape):
    """Strided axes of 8192 points (c64): both four-step phases in ONE persistent TMA-fed launch with the
    intermediate in L2 scratch slots (band_kernel.cuh) -- the default plan for cfg3's column axis.  Against the oracle's
    fft2D; the same transform planned without the band pass (B200FFT_BAND=0) must agree to rounding; a width that is not
    a multiple of the band (96 columns) and a buffer that is only 8-byte aligned take the fallback plan."""
    import torch
    h, w = shape
    rng = np.random.default_rng(h + w)
    x = rand_complex(rng, (h, w), np.complex64)
    p = af.Plan("2d", [h, w], af.C2C, 1)
    desc = p.describe()
    p.destroy()
    assert ("band A[" in desc) == (w % 32 == 0), desc
    y = gpu(af, "fft2D", mode, x)
    assert rel_l2(y, oracle.fft2D(mode, x, threads=8)) <= bar(np.complex64, x.size)
    with _env(af, B200FFT_BAND="0"):
        y0 = gpu(af, "fft2D", mode, x)
    assert rel_l2(y, y0) <= 4e-7
    if w % 32 == 0:
        # 8-byte aligned (not 16) input and output: TMA cannot serve them, the band-free fallback plan does
        buf = torch.empty(h * w + 1, dtype=torch.complex64, device="cuda")
        xin = buf[1:].view(h, w)
        xin.copy_(torch.from_numpy(x))
        obuf = torch.empty(h * w + 1, dtype=torch.complex64, device="cuda")
        out = obuf[1:].view(h, w)
        pl = af.Plan("2d", [h, w], af.C2C, 1)
        pl.exec(xin, out, af.FORWARD)
        pl.destroy()
        assert rel_l2(out.cpu().numpy(), oracle.fft2D("Forward", x, threads=8)) <= bar(np.complex64, x.size)


@pytest.mark.parametrize("mode", ["Forward", "Inverse"])
def test_band_two_pass_large_1d(af, mode):
    """Contiguous 1D transforms of 2^27 points (c64) in TWO HBM round trips: the 8192-point strided axis with the outer
    four-step twiddle (band kernel, MODE_STRIDED + OUTER) and the 16384-point rows with the transposed store (MODE_ROWS),
    each with its intermediate in L2 slots -- the opt-in two-round-trip plan of cfg4 (2^28 = 2^14 x 2^14).  Full-array
    compare against a double-precision library transform, and against the band-free three-pass plan of the same library."""
    import scipy.fft as sf
    n = 1 << 27
    rng = np.random.default_rng(27)
    x = rand_complex(rng, (n,), np.complex64)
    with _env(af, B200FFT_BAND_1D="1"):     # opt-in: measured 3 % slower than the three lines passes on cfg4
        p = af.Plan("1d", [n], af.C2C, 1)
        desc = p.describe()
        p.destroy()
        assert "+outer tw" in desc and "4step-rows: band" in desc, desc
        y = gpu(af, "fft1D", mode, x)
    x128 = x.astype(np.complex128)
    ref = sf.fft(x128, workers=-1) if mode == "Forward" else sf.ifft(x128, workers=-1)
    del x128
    assert rel_l2(y, ref) <= bar(np.complex64, n)
    del ref
    with _env(af, B200FFT_BAND="0"):
        y0 = gpu(af, "fft1D", mode, x)
    assert rel_l2(y, y0) <= 1e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_cluster_column_kernel(af, oracle, dtype, mode):
    """Column axes of 4096..16384 points in ONE pass by a thread-block cluster exchanging through distributed
    shared memory (cluster_kernel.cuh; opt-in, B200FFT_CLUSTER=1).  Includes a ragged last column tile."""
    _needs_experimental(af)
    rng = np.random.default_rng(41)
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    shapes = [(4096, 64), (8192, 24), (8192, 100), (16384, 16)] if dtype == np.complex64 else [(4096, 16), (8192, 12), (8192, 37)]
    with _env(af, B200FFT_CLUSTER="1", B200FFT_PIPE="0"):
        for shape in shapes:
            p = af.Plan("2d", list(shape), typ, 1)
            d = p.describe()
            p.destroy()
            assert "cluster cols" in d and len(d.strip().split("\n")) == 2, d
            x = rand_complex(rng, shape, dtype)
            y = gpu(af, "fft2D", mode, x)
            assert rel_l2(y, _np_fft2(mode, x)) <= bar(dtype, x.size), (shape, mode)
            if shape[0] == 4096:
                assert rel_l2(y, oracle.fft2D(mode, x, threads=8)) <= bar(dtype, x.size), (shape, mode)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_cluster_rows_plus_first_column_stage(af, oracle, dtype, mode):
    """2D with H = 8*M: rows + the radix-8 first stage of the column axis in one cluster pass (exchange across the 8 rows
    through distributed shared memory), then ONE M-point column pass (cluster_kernel.cuh fft_cluster_rows_kernel;
    opt-in, B200FFT_CLUSTER_ROWS=1 -- measured slower than the three-pass plan, profiles/r01_pipe_and_cluster.txt)."""
    _needs_experimental(af)
    rng = np.random.default_rng(45)
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    shapes = [(4096, 4096), (8192, 4096)] if mode != "Reverse" else [(4096, 4096)]
    with _env(af, B200FFT_CLUSTER_ROWS="1"):
        for shape in shapes:
            p = af.Plan("2d", list(shape), typ, 1)
            d = p.describe()
            p.destroy()
            assert "rows+radix8" in d and len(d.strip().split("\n")) == 2, d
            x = rand_complex(rng, shape, dtype)
            y = gpu(af, "fft2D", mode, x)
            assert rel_l2(y, _np_fft2(mode, x)) <= bar(dtype, x.size), (shape, mode)
    x = rand_complex(rng, (4096, 4096), dtype)
    with _env(af, B200FFT_CLUSTER_ROWS="1"):
        y = gpu(af, "fft2D", "Forward", x)
    assert rel_l2(y, oracle.fft2D("Forward", x, threads=8)) <= bar(dtype, x.size)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", ["Forward", "Inverse"])
def test_pipelined_column_kernels(af, oracle, dtype, mode):
    """The persistent software-pipelined column kernel (pipe_kernel.cuh): cp.async landing FIFO, two thread groups and,
    for clusters, the st.async + mbarrier exchange.  Forced on for every legal shape (B200FFT_PIPE=1); one tile per
    CTA up to several tiles per CTA with an odd count; buffers that are not 16-byte aligned take the fallback."""
    import torch
    rng = np.random.default_rng(43)
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    n1 = 1024 if dtype == np.complex64 else 512          # the CTA share of the instantiated kernels
    with _env(af, B200FFT_CLUSTER="1", B200FFT_PIPE="1", B200FFT_PIPE_MIN_TILES="1"):
        cases = [(1, 64), (1, 8 * 148 * 3 + 8), (2, 128), (4, 64), (8, 64), (8, 8 * 37), (16, 128)]
        if not af.lib().b200fftHasExperimental():
            cases = [c for c in cases if c[0] <= 2]       # clusters of 4..16 are opt-in kernels of the experimental build
        for cs, inner in cases + ([(0, 256)] if dtype == np.complex64 else []):
            shape = (n1 * cs, inner) if cs else (512, inner)       # cs == 0: the c64 512-point single-CTA kernel
            p = af.Plan("2d", list(shape), typ, 1)
            d = p.describe()
            p.destroy()
            if cs == 0:
                want = "pipe: N=512x1"
            elif dtype == np.complex128 and cs == 2:
                want = "pipe: N=1024x1"                            # c128 N=1024: one CTA with 4 columns is the default
            else:
                want = "pipe: N=%dx%d" % (n1, cs)
            assert want in d, d
            x = rand_complex(rng, shape, dtype)
            y = gpu(af, "fft2D", mode, x)
            assert rel_l2(y, _np_fft2(mode, x)) <= bar(dtype, x.size), (shape, mode)
            if cs in (1, 8) and inner == 64:
                assert rel_l2(y, oracle.fft2D(mode, x, threads=8)) <= bar(dtype, x.size), (shape, mode)
        # the four-step first pass (column transform + twiddle at the store) through the pipelined kernel
        n = n1 * n1
        x = rand_complex(rng, (2, n), dtype)
        p = af.Plan("many", [n], typ, 2)
        assert "col+tw" in p.describe() and "pipe:" in p.describe(), p.describe()
        p.destroy()
        y = gpu(af, "fft", mode, x)
        x128 = x.astype(np.complex128)
        assert rel_l2(y, np.fft.fft(x128) if mode == "Forward" else np.fft.ifft(x128)) <= bar(dtype, n)
        # a buffer that is only element-aligned must take the lock-step kernel and still be right
        shape = (n1 * 8, 64)
        x = rand_complex(rng, shape, dtype)
        buf = torch.empty(x.size + 1, dtype=torch.from_numpy(x).dtype, device="cuda")
        xin = buf[1:].view(shape)
        xin.copy_(torch.from_numpy(x))
        if xin.data_ptr() % 16:
            y = af.fft2D(mode, xin).cpu().numpy()
            assert rel_l2(y, _np_fft2(mode, x)) <= bar(dtype, x.size)


@pytest.mark.parametrize("dtype", DTYPES)
def test_pipelined_kernels_default_policy_and_rows(af, dtype):
    """Without any switch the planner takes the pipelined column kernel where it was measured to win (plain / CS=2,
    tile span <= 256 MB) and nowhere else; the row variant (B200FFT_PIPE=1) agrees with the lock-step row kernel."""
    import torch
    rng = np.random.default_rng(47)
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    n1 = 1024 if dtype == np.complex64 else 512
    p = af.Plan("axis", [16, 1024, 512], typ, 1)
    assert "pipe: N=1024x1" in p.describe(), p.describe()          # both types: N=1024 in one CTA
    p.destroy()
    p = af.Plan("axis", [16, n1, 512], typ, 1)
    assert ("pipe:" in p.describe()) == (dtype == np.complex64), p.describe()      # c128 N=512 stays on the lock-step kernel
    x = rand_complex(rng, (16, n1, 512), dtype)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    p.exec(xd, yd, af.FORWARD)
    p.destroy()
    assert rel_l2(yd.cpu().numpy(), np.fft.fft(x.astype(np.complex128), axis=1)) <= bar(dtype, n1)
    for shape in [(1, n1, 1 << 16), (1, n1 * 8, 1024)]:      # span 512 MB / a cluster of 8: not taken by default
        p = af.Plan("axis", list(shape), typ, 1)
        assert "pipe:" not in p.describe(), p.describe()
        p.destroy()
    if not af.lib().b200fftHasExperimental():
        return                                               # the row variant is an opt-in kernel of the experimental build
    n = 8192 if dtype == np.complex64 else 4096
    x = rand_complex(rng, (2 * 148 + 5, n), dtype)
    y0 = af.fft("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
    with _env(af, B200FFT_PIPE="1"):
        p = af.Plan("many", [n], typ, x.shape[0])
        assert "pipe rows" in p.describe(), p.describe()
        p.destroy()
        y1 = af.fft("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel_l2(y1, np.fft.fft(x.astype(np.complex128), axis=1)) <= bar(dtype, n)
    assert rel_l2(y1, y0) <= bar(dtype, n) / 10


# ---- DFT/Centre.hs: centre / shift / ishift and the fused shift(fft(x)) -------------------------------------------

def _ref_shift(x, inverse):
    """Centre.hs:70-164 restated: backpermute with roll i = (i + shift) rem n per axis."""
    out = x
    for ax, n in enumerate(x.shape):
        sh = n // 2 + (0 if (inverse or n % 2 == 0) else 1)
        out = np.take(out, (np.arange(n) + sh) % n, axis=ax)
    return out


def _ref_centre(x):
    """Centre.hs:36-66 restated: (-1)^(sum of indices) * x."""
    idx = np.indices(x.shape).sum(axis=0)
    return np.where(idx % 2 == 0, x, -x)


@pytest.mark.parametrize("dtype", DTYPES)
def test_centre_shift_ishift_exact(af, dtype):
    import torch
    rng = np.random.default_rng(51)
    for shape in [(1,), (2,), (7,), (8,), (1000,), (5, 6), (8, 3), (64, 64), (3, 4, 5), (8, 8, 8), (7, 1, 2)]:
        x = rand_complex(rng, shape, dtype)
        xd = torch.from_numpy(x).cuda()
        r = len(shape)
        c = getattr(af, "centre%dD" % r)(xd).cpu().numpy()
        s = getattr(af, "shift%dD" % r)(xd).cpu().numpy()
        i = getattr(af, "ishift%dD" % r)(xd).cpu().numpy()
        assert np.array_equal(c, _ref_centre(x)), shape
        assert np.array_equal(s, _ref_shift(x, False)) and np.array_equal(s, np.fft.fftshift(x)), shape
        assert np.array_equal(i, _ref_shift(x, True)) and np.array_equal(i, np.fft.ifftshift(x)), shape
        back = getattr(af, "ishift%dD" % r)(torch.from_numpy(s).cuda()).cpu().numpy()
        assert np.array_equal(back, x), shape                       # Centre.hs:84-86


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_fft_centred_fused_and_fallback(af, oracle, dtype, mode):
    """shift(fft(x)) in one call: for power-of-two extents the rotation is folded into the last pass of every axis (same
    number of kernel launches as the plain transform); other extents take transform + stand-alone shift.  For even
    extents it equals fft(centre(x)) (Centre.hs:17-19)."""
    import torch
    rng = np.random.default_rng(53)
    shapes = [(2,), (64,), (4096,), (1 << 16,), (8, 16), (256, 64), (2048, 32), (4096, 64), (4, 8, 16), (64, 32, 128), (16, 2048, 8),
              (1, 64), (64, 1), (6,), (1000,), (12, 10), (8, 7), (3, 4, 6)]
    for shape in shapes:
        x = rand_complex(rng, shape, dtype)
        xd = torch.from_numpy(x).cuda()
        kind = "fft%dD" % len(shape)
        pow2 = all(n & (n - 1) == 0 for n in shape)
        plain = getattr(af, kind)(mode, xd)
        l0 = af.kernel_launches()
        getattr(af, kind)(mode, xd)
        l1 = af.kernel_launches()
        y = af.fft_centred(mode, xd)
        l2 = af.kernel_launches()
        if pow2 and not (mode == "Inverse"):
            assert l2 - l1 == l1 - l0, (shape, l1 - l0, l2 - l1)      # fused: no extra kernel
        y = y.cpu().numpy()
        ref = _ref_shift(plain.cpu().numpy(), False)
        npts = x.size if len(shape) > 1 else shape[0]
        assert rel_l2(y, ref) <= bar(dtype, npts) / 10, (shape, mode)
        if pow2 and mode != "Inverse":
            assert np.array_equal(y, ref), (shape, mode)              # same butterflies, permuted stores
        if all(n % 2 == 0 for n in shape):
            z = getattr(af, kind)(mode, getattr(af, "centre%dD" % len(shape))(xd)).cpu().numpy()
            assert rel_l2(z, ref) <= bar(dtype, npts), (shape, mode)
    x = rand_complex(rng, (64, 128), dtype)
    y = af.fft_centred("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel_l2(y, _ref_shift(oracle.fft2D("Forward", x), False)) <= bar(dtype, x.size)
    # plans built by the whole-transform builders do not carry the rotation: the call must fall back, not mis-shift
    with _env(af, B200FFT_CLUSTER_ROWS="1", B200FFT_PAIR2D="1"):
        x = rand_complex(rng, (4096, 4096), dtype)
        y = af.fft_centred("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
        assert rel_l2(y, np.fft.fftshift(np.fft.fft2(x.astype(np.complex128)))) <= bar(dtype, x.size)


@pytest.mark.parametrize("dtype", DTYPES)
def test_exec_scatter_single_gpu(af, dtype):
    """b200fftExecScatter on one device: the stores of a strided-axis pass split by output index over several buffers
    (the slab transform's exchange; with peer-mapped buffers in tests/test_multigpu.py).  Both the lock-step kernel
    (several targets) and the pipelined one (single target, transposing store) against the plain exec."""
    import torch
    rng = np.random.default_rng(61)
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    for (outer, n, inner, npeers) in [(6, 64, 40, 4), (3, 1024, 96, 8), (2, 1024, 96, 1), (1, 2048, 64, 2)]:
        x = rand_complex(rng, (outer, n, inner), dtype)
        xd = torch.from_numpy(x).cuda()
        ref = torch.empty_like(xd)
        with _env(af, B200FFT_PIPE_MIN_TILES="1"):
            p = af.Plan("axis", [outer, n, inner], typ, 1)
            p.exec(xd, ref, af.FORWARD)
            nl = n // npeers
            outs = [torch.zeros((outer, nl, inner), dtype=xd.dtype, device="cuda") for _ in range(npeers)]
            p.exec_scatter(xd, [o.data_ptr() for o in outs], nl * inner, inner, af.FORWARD)
            torch.cuda.synchronize()
            for r in range(npeers):
                # (the plain exec of N=2048 c64 takes the cluster-of-2 pipelined kernel: same values, different rounding)
                assert rel_l2(outs[r].cpu().numpy(), ref[:, r * nl:(r + 1) * nl, :].cpu().numpy()) <= bar(dtype, n) / 20, (outer, n, inner, npeers, r)
            if npeers == 1:   # transposing store: (o, k, i) -> [k][o][i]
                tr = torch.zeros((n, outer, inner), dtype=xd.dtype, device="cuda")
                p.exec_scatter(xd, [tr.data_ptr()], inner, outer * inner, af.FORWARD)
                torch.cuda.synchronize()
                assert rel_l2(tr.cpu().numpy(), ref.transpose(0, 1).contiguous().cpu().numpy()) <= bar(dtype, n) / 20
            p.destroy()
    assert rel_l2(ref.cpu().numpy(), np.fft.fft(x.astype(np.complex128), axis=1)) <= bar(dtype, 2048)


SHAPES_3D = [(1, 1, 1), (2, 2, 2), (16, 32, 64), (64, 32, 16), (10, 12, 14), (3, 5, 7), (5, 64, 33), (4, 1024, 8), (1024, 4, 8),
             (64, 64, 64), (16, 16, 4096)]


@pytest.mark.parametrize("dtype", DTYPES)
def test_slab_entry_point_single_rank(af, oracle, dtype):
    """The multi-GPU entry point of the C ABI (b200fftPlanSlab3d / b200fftExecSlab, csrc/slab.cu) with ONE rank: the exchange
    degenerates to scattering into this GPU's own receive buffer, but the whole pipeline runs -- window plans, the chunked
    y pass as a grid-stride loop on a few CTAs, side streams, flag barriers, both output layouts, the Inverse scale -- so it
    is exercised on a one-GPU box too (two ranks over NVLink: tests/test_multigpu.py).  Against the oracle's fft3D."""
    import ctypes
    import torch
    from accelerate_fft_b200._lib import ALLGATHER_FN
    lib = af.lib()
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    rng = np.random.default_rng(77)

    def allgather(_ctx, send, recv, nbytes):          # one rank: the gathered blob is my own
        ctypes.memmove(recv, send, nbytes)
        return 0

    cb = ALLGATHER_FN(allgather)
    for (d, h, w) in [(16, 32, 64), (64, 128, 96), (256, 64, 512)]:
        x = rand_complex(rng, (d, h, w), dtype)
        ref = oracle.fft3D("Forward", x, threads=8)
        xd = torch.from_numpy(x).cuda()
        hnd = ctypes.c_void_p()
        assert lib.b200fftPlanSlab3d(ctypes.byref(hnd), d, h, w, typ, 0, 1, 1, cb, None) == 0
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for (cp, ck, yc) in [(None, None, None), (1, 1, 0), (2, 2, 5), (4, 3, 40)]:
            if cp is not None:
                assert lib.b200fftSlabTune(hnd, 0, cp, ck, yc) == 0 and lib.b200fftSlabTune(hnd, 1, cp, ck, yc) == 0
            nat = torch.empty_like(xd)
            tr = torch.empty((h, d, w), dtype=xd.dtype, device="cuda")
            assert lib.b200fftExecSlab(hnd, xd.data_ptr(), nat.data_ptr(), af.FORWARD, 1.0, 0, st) == 0
            assert lib.b200fftExecSlab(hnd, xd.data_ptr(), tr.data_ptr(), af.FORWARD, 1.0, 1, st) == 0
            assert rel_l2(nat.cpu().numpy(), ref) <= bar(dtype, x.size), ((d, h, w), cp, ck, yc)
            assert rel_l2(tr.cpu().numpy().transpose(1, 0, 2), ref) <= bar(dtype, x.size), ((d, h, w), cp, ck, yc)   # [H][D][W]
            back = torch.empty_like(xd)
            assert lib.b200fftExecSlab(hnd, nat.data_ptr(), back.data_ptr(), af.INVERSE, 1.0 / x.size, 0, st) == 0
            assert rel_l2(back.cpu().numpy(), x) <= 2 * bar(dtype, x.size)
        buf = ctypes.c_void_p()
        assert lib.b200fftSlabNaturalBuffer(hnd, ctypes.byref(buf)) == 0 and buf.value
        assert lib.b200fftExecSlab(hnd, xd.data_ptr(), xd.data_ptr(), af.FORWARD, 1.0, 0, st) != 0     # in == out is refused
        assert lib.b200fftDestroySlab(hnd) == 0
    hnd = ctypes.c_void_p()
    assert lib.b200fftPlanSlab3d(ctypes.byref(hnd), 48, 32, 16, typ, 0, 1, 1, cb, None) == 16            # D not a power of two: NOT_SUPPORTED


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", ["Forward", "Inverse"])
def test_single_buffer_tma_column_kernel(af, oracle, dtype, mode):
    """1024-point columns whose 128 KB tile fills an SM: the persistent single-buffer TMA-fed column kernel (ringcol_kernel.cuh,
    cfg5's z axis) -- taken when the pipelined column kernel is not (here: switched off), from 148 tiles on; against the oracle's
    fft2D, bit-identical to the lock-step kernel, and the fallback for a buffer TMA cannot address."""
    import torch
    tl = 16 if dtype == np.complex64 else 8
    typ = af.C2C if dtype == np.complex64 else af.Z2Z
    w = 150 * tl
    rng = np.random.default_rng(91)
    x = rand_complex(rng, (1024, w), dtype)
    with _env(af, B200FFT_PIPE="0"):
        p = af.Plan("2d", [1024, w], typ, 1)
        assert "ring cols" in p.describe(), p.describe()
        p.destroy()
        y = gpu(af, "fft2D", mode, x)
        with _env(af, B200FFT_PIPE="0", B200FFT_RINGCOL="0"):
            y0 = gpu(af, "fft2D", mode, x)
        # the column pass alone: an input TMA cannot address (element-aligned only) takes the lock-step kernel, same result
        pa = af.Plan("axis", (1, 1024, w), typ)
        assert "ring cols" in pa.describe(), pa.describe()
        xd = torch.from_numpy(x).cuda()
        buf = torch.empty(x.size + 1, dtype=xd.dtype, device="cuda")
        xin = buf[1:].view(1024, w)
        xin.copy_(xd)
        oa, ob = torch.empty_like(xd), torch.empty_like(xd)
        pa.exec(xd, oa, af.FORWARD)
        if xin.data_ptr() % 16:
            pa.exec(xin, ob, af.FORWARD)
            assert torch.equal(oa, ob)
        pa.destroy()
        assert rel_l2(oa.cpu().numpy(), np.fft.fft(x.astype(np.complex128), axis=0)) <= bar(dtype, 1024)
    assert rel_l2(y, oracle.fft2D(mode, x, threads=8)) <= bar(dtype, x.size)
    assert np.array_equal(y, y0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", MODES)
def test_fft3d_vs_oracle(af, oracle, dtype, mode):
    rng = np.random.default_rng(15)
    for shape in SHAPES_3D:
        x = rand_complex(rng, shape, dtype)
        y = gpu(af, "fft3D", mode, x)
        x128 = x.astype(np.complex128)
        ex = np.fft.fftn(x128) if mode == "Forward" else np.fft.ifftn(x128) * (1 if mode == "Inverse" else x.size)
        assert rel_l2(y, ex) <= bar(dtype, x.size), (shape, mode)
        if all(_is_5smooth(s) for s in shape) and x.size <= 1 << 18:
            assert rel_l2(y, oracle.fft3D(mode, x, threads=4)) <= bar(dtype, x.size), (shape, mode)


@pytest.mark.parametrize("dtype", DTYPES)
def test_fft_innermost_axis_of_rank_2_3_4(af, oracle, dtype):
    """FFT.hs:63-84 `fft` transforms only the innermost axis; rank 2/3 use the batched plans (PTX.hs:59-60)."""
    rng = np.random.default_rng(16)
    for shape in [(7, 64), (3, 5, 128), (4, 6, 100), (2, 3, 4, 32)]:
        x = rand_complex(rng, shape, dtype)
        for mode in MODES:
            assert rel_l2(gpu(af, "fft", mode, x), oracle.fft(mode, x)) <= bar(dtype, shape[-1]), (shape, mode)


def test_edge_cases(af):
    import torch
    # empty arrays: nothing to do, shape preserved
    e = torch.empty((0, 16), dtype=torch.complex64, device="cuda")
    assert af.fft("Forward", e).shape == (0, 16)
    # length-1 transforms are the identity (Adhoc.hs:45)
    x = torch.randn(9, 1, dtype=torch.complex128, device="cuda")
    assert torch.equal(af.fft("Forward", x), x)
    assert torch.equal(af.fft("Inverse", x), x)
    # the input is never written (out-of-place contract, PTX.hs:92)
    a = torch.randn(64, 4096, dtype=torch.complex64, device="cuda")
    b = a.clone()
    af.fft2D("Inverse", a)
    assert torch.equal(a, b)
    # non-contiguous input is accepted (made dense, like Accelerate's arrays always are)
    t = torch.randn(32, 48, dtype=torch.complex64, device="cuda").t()
    y = af.fft("Forward", t)
    assert rel_l2(y.cpu().numpy(), np.fft.fft(t.cpu().numpy().astype(np.complex128), axis=-1)) < 1e-5
    # lazy conj / neg views (x.conj(), x.mH) share storage with the plain tensor: they must be materialised, not ignored
    c = torch.randn(8, 64, dtype=torch.complex64, device="cuda")
    ref = np.fft.fft(np.conj(c.cpu().numpy()).astype(np.complex128), axis=-1)
    assert c.conj().is_conj()
    assert rel_l2(af.fft("Forward", c.conj()).cpu().numpy(), ref) < 1e-5
    assert rel_l2(af.fft2D("Forward", c.conj()).cpu().numpy(), np.fft.fft2(np.conj(c.cpu().numpy()).astype(np.complex128))) < 1e-5
    hc = torch.randn(4, 32, dtype=torch.complex128).conj()
    assert rel_l2(af.run_host_seq("fft", ["Forward"], hc).numpy(), np.fft.fft(hc.resolve_conj().numpy(), axis=-1)) < 1e-12


def test_host_buffer_entry_and_fused_inverse(af, oracle):
    rng = np.random.default_rng(17)
    x = rand_complex(rng, (6, 10, 256), np.complex64)
    assert rel_l2(af.run_host("fft", "Forward", x), oracle.fft("Forward", x)) <= bar(np.complex64, 256)
    x2 = rand_complex(rng, (96, 80), np.complex128)
    ref = oracle.fft2D("Inverse", x2)
    assert rel_l2(af.run_host("fft2D", "Inverse", x2), ref) <= bar(np.complex128, x2.size)
    af.set_fused_inverse(True)
    try:
        assert rel_l2(af.run_host("fft2D", "Inverse", x2), ref) <= bar(np.complex128, x2.size)
        x3 = rand_complex(rng, (1 << 16,), np.complex64)
        assert rel_l2(af.run_host("fft1D", "Inverse", x3), np.fft.ifft(x3.astype(np.complex128))) <= bar(np.complex64, 1 << 16)
    finally:
        af.set_fused_inverse(False)


def test_host_chain_pipelined_in_row_chunks(af, oracle):
    """accfft_run_host_seq: a chain of transforms over a HOST array; for `fft` the rows flow through the device in
    chunks on three streams (>= 32 MiB arrays).  A chunk of rows runs the same kernel as the whole array, so the
    result must equal the device-resident path bit for bit; a sample of rows is checked against the oracle."""
    import torch
    rng = np.random.default_rng(23)
    for dtype, shape in ((np.complex64, (8200, 1024)), (np.complex128, (3, 1030, 1024))):   # 64 MiB / 48 MiB, ragged last chunk
        x = rand_complex(rng, shape, dtype)
        hx = torch.from_numpy(x).pin_memory()
        y = af.run_host_seq("fft", ["Forward"], hx).numpy()
        yd = af.fft("Forward", torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(y, yd)
        rows = x.reshape(-1, 1024)[::97]
        assert rel_l2(y.reshape(-1, 1024)[::97], oracle.fft("Forward", rows)) <= bar(dtype, 1024)
        z = af.run_host_seq("fft", ["Forward", "Inverse"], x)           # pageable numpy buffers work too
        assert rel_l2(z, x) <= 2 * bar(dtype, 1024)
        z2 = af.fft("Inverse", af.fft("Forward", torch.from_numpy(x).cuda())).cpu().numpy()
        assert np.array_equal(z, z2)
    # whole-array kinds and small arrays take the unchunked route
    x2 = rand_complex(rng, (64, 48), np.complex64)
    assert rel_l2(af.run_host_seq("fft2D", ["Forward", "Reverse"], x2), x2.astype(np.complex128) * x2.size) <= 2 * bar(np.complex64, x2.size)
    with pytest.raises(af.B200FFTError):
        af.run_host_seq("fft2D", ["Forward"], np.ones((2, 3, 4), np.complex64))


def test_plan_cache_and_streams(af):
    """Plans are cached per (context, shape, type) (PTX/Plans.hs:66-86) and exec takes the stream as an argument."""
    import torch
    af.lib().accfft_plan_cache_clear()
    x = torch.randn(16, 512, dtype=torch.complex64, device="cuda")
    af.fft("Forward", x)
    n1 = af.lib().accfft_plan_cache_size()
    af.fft("Inverse", x)          # same shape, other direction: same plan
    assert af.lib().accfft_plan_cache_size() == n1
    af.fft("Forward", x.to(torch.complex128))
    assert af.lib().accfft_plan_cache_size() == n1 + 1
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ref = torch.fft.fft(x.to(torch.complex128), dim=-1)  # test-side reference only
    outs = []
    for s in (s1, s2, s1, s2):
        with torch.cuda.stream(s):
            outs.append(af.fft("Forward", x))
    torch.cuda.synchronize()
    for o in outs:
        assert rel_l2(o.cpu().numpy(), ref.cpu().numpy()) < 1e-5


# ---- BASELINE.json configurations at full size ---------------------------------------------------

def test_cfg1_full_vs_oracle(af, oracle):
    """cfg 1: c64 [4096,1024] Forward -- full-array compare with both CPU paths."""
    import scipy.fft as sf
    rng = np.random.default_rng(1001)
    x = rand_complex(rng, (4096, 1024), np.complex64)
    y = gpu(af, "fft", "Forward", x)
    assert rel_l2(y, oracle.fft("Forward", x, threads=8)) <= bar(np.complex64, 1024)              # pure-Accelerate path
    assert rel_l2(y, sf.fft(x.astype(np.complex128), workers=-1)) <= bar(np.complex64, 1024) / 4  # FFTW stand-in, exact-ish


def test_cfg2_full_size(af, oracle):
    """cfg 2: c128 [65536,4096] Forward + Inverse.  Sampled rows against the oracle and the exact
    long-double transform; whole-array round trip, Parseval and linearity on the device."""
    import scipy.fft as sf
    import torch
    n, batch = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(1002)
    x = (torch.rand(batch, n, 2, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
    x = torch.view_as_complex(x)
    y = af.fft("Forward", x)
    rows = torch.tensor(sorted({0, 1, 255, 4097, 32768, 65535} | set(np.random.default_rng(2).integers(0, batch, 58).tolist())), device="cuda")
    xs, ys = x[rows].cpu().numpy(), y[rows].cpu().numpy()
    assert rel_l2(ys, oracle.fft("Forward", xs, threads=8)) <= bar(np.complex128, n)
    ex = sf.fft(xs.astype(np.clongdouble)).astype(np.complex128)
    assert rel_l2(ys, ex) <= bar(np.complex128, n) / 2
    # Parseval per row: ||F x||^2 = n ||x||^2  (isometry property, test/Test/FFT.hs:201-217)
    e_in = (x.real ** 2 + x.imag ** 2).sum(-1)
    e_out = (y.real ** 2 + y.imag ** 2).sum(-1)
    assert float(((e_out - n * e_in).abs() / (n * e_in)).max()) < 1e-12
    z = af.fft("Inverse", y)                      # includes the 1/n scale (FFT.hs:83)
    err = torch.linalg.vector_norm(z - x) / torch.linalg.vector_norm(x)
    assert float(err) <= 2 * bar(np.complex128, n)
    del z, e_in, e_out
    # linearity with a complex scalar on the whole array
    c = complex(0.3, -0.7)
    y2 = af.fft("Forward", x * c)
    err = torch.linalg.vector_norm(y2 - y * c) / torch.linalg.vector_norm(y)
    assert float(err) <= bar(np.complex128, n)


def test_cfg3_full_vs_library(af, oracle):
    """cfg 3: c64 8192x8192 fft2D Forward -- full-array compare against a double-precision library transform,
    the oracle on an embedded 256x256 problem, and the round trip."""
    import scipy.fft as sf
    import torch
    rng = np.random.default_rng(1003)
    x = rand_complex(rng, (8192, 8192), np.complex64)
    xd = torch.from_numpy(x).cuda()
    yd = af.fft2D("Forward", xd)
    y = yd.cpu().numpy()
    ref = sf.fft2(x.astype(np.complex128), workers=-1)
    assert rel_l2(y, ref) <= bar(np.complex64, x.size)
    del ref
    zd = af.fft2D("Inverse", yd)
    assert float(torch.linalg.vector_norm(zd - xd) / torch.linalg.vector_norm(xd)) <= 2 * bar(np.complex64, x.size)
    xs = x[:256, :256].copy()
    assert rel_l2(gpu(af, "fft2D", "Forward", xs), oracle.fft2D("Forward", xs, threads=8)) <= bar(np.complex64, xs.size)


def _fold(x, axis, m):
    """sum_{j} x[n + j*m] along `axis` (length m result): the DFT of the folded sequence equals the
    DFT of x sampled at every (len/m)-th bin."""
    import torch
    shp = list(x.shape)
    L = shp[axis]
    new = shp[:axis] + [L // m, m] + shp[axis + 1:]
    return x.reshape(new).to(torch.complex128).sum(dim=axis)


def test_cfg4_full_size(af, oracle):
    """cfg 4: c64 n=2^28 fft1D.  Closed-form tones, the folding (bin-decimation) identity against the
    oracle on random data, round trip; plus a full compare at 2^24 against a library transform."""
    import scipy.fft as sf
    import torch
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(1004)
    x = torch.view_as_complex(torch.rand(n, 2, dtype=torch.float32, device="cuda", generator=g) * 2 - 1)
    y = af.fft1D("Forward", x)
    # folding: X[k * 2^14] for k < 2^14 == DFT_{2^14}( sum_j x[i + j*2^14] )
    m = 1 << 14
    folded = _fold(x, 0, m).cpu().numpy()
    ref = oracle.exact_dft(-1, folded[None, :])[0].astype(np.complex128) if m <= 2048 else sf.fft(folded.astype(np.clongdouble)).astype(np.complex128)
    got = y[:: n // m].cpu().numpy()
    assert rel_l2(got, ref) <= bar(np.complex64, n)
    assert rel_l2(oracle.fft("Forward", folded.astype(np.complex128)), ref) < 1e-12   # the oracle agrees on the folded problem
    # offset folding catches errors in the other residue classes: X[k*2^14 + r] via modulated fold
    r = 12345
    ph = torch.exp(-2j * torch.pi * r * torch.arange(n, device="cuda", dtype=torch.float64) / n)
    folded_r = _fold(x.to(torch.complex128) * ph, 0, m).cpu().numpy()
    got_r = y[r:: n // m].cpu().numpy()
    assert rel_l2(got_r, sf.fft(folded_r.astype(np.clongdouble)).astype(np.complex128)) <= bar(np.complex64, n)
    del ph
    z = af.fft1D("Inverse", y)
    assert float(torch.linalg.vector_norm(z - x) / torch.linalg.vector_norm(x)) <= 2 * bar(np.complex64, n)
    del z, y
    # tones: x[j] = e^{2 pi i f j / n} -> n * delta_f
    for f in (1, 98765, n - 3):
        j = torch.arange(n, device="cuda", dtype=torch.float64)
        ang = 2 * torch.pi * ((j * f) % n) / n
        t = torch.complex(torch.cos(ang), torch.sin(ang)).to(torch.complex64)
        del j, ang
        yt = af.fft1D("Forward", t)
        peak = yt[f].item()
        assert abs(peak - n) / n < 1e-5
        yt[f] = 0
        assert float(torch.linalg.vector_norm(yt)) / n < 1e-4   # everything else is rounding noise of the float input
        del t, yt
    # full compare at 2^24
    rng = np.random.default_rng(1004)
    xs = rand_complex(rng, (1 << 24,), np.complex64)
    assert rel_l2(gpu(af, "fft1D", "Forward", xs), sf.fft(xs.astype(np.complex128), workers=-1)) <= bar(np.complex64, 1 << 24)


def test_cfg5_full_size(af, oracle):
    """cfg 5: c64 1024^3 fft3D (single GPU).  Folding identity per axis against the oracle's fft3D,
    plane-wave known answers, round trip."""
    import torch
    d = 1024
    g = torch.Generator(device="cuda").manual_seed(1005)
    x = torch.view_as_complex(torch.rand(d, d, d, 2, dtype=torch.float32, device="cuda", generator=g) * 2 - 1)
    y = af.fft3D("Forward", x)
    m = 16
    f = _fold(_fold(_fold(x, 2, m), 1, m), 0, m).cpu().numpy()   # 16^3, complex128
    ref = oracle.fft3D("Forward", f)
    got = y[:: d // m, :: d // m, :: d // m].cpu().numpy()
    assert rel_l2(got, ref) <= bar(np.complex64, d ** 3)
    z = af.fft3D("Inverse", y)
    assert float(torch.linalg.vector_norm(z - x) / torch.linalg.vector_norm(x)) <= 2 * bar(np.complex64, d ** 3)
    del z, y, x
    # plane wave e^{2 pi i (a z + b y + c x)/d} -> d^3 at (a,b,c)
    a, b, c = 3, 1000, 517
    i = torch.arange(d, device="cuda", dtype=torch.float64)
    wz = torch.exp(2j * torch.pi * ((a * i) % d) / d)
    wy = torch.exp(2j * torch.pi * ((b * i) % d) / d)
    wx = torch.exp(2j * torch.pi * ((c * i) % d) / d)
    t = (wz[:, None, None] * wy[None, :, None] * wx[None, None, :]).to(torch.complex64)
    yt = af.fft3D("Forward", t)
    del t
    peak = yt[a, b, c].item()
    assert abs(peak - d ** 3) / d ** 3 < 1e-5
    yt[a, b, c] = 0
    assert float(torch.linalg.vector_norm(yt)) / d ** 3 < 1e-4
