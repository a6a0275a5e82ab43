"""Seeded planner fuzz on the GPU (tools/plan_fuzz.py): shapes drawn around the planner's switch points through the public
entry points, against numpy's float64 FFT at the north_star tolerance; the output starts as NaN (an element the plan never
stores shows up), the input must survive bit for bit (PTX.hs:92) and a second call must reproduce the first exactly.
The full run (484 cases, profiles/r02_plan_fuzz.txt) is `python tools/plan_fuzz.py --targeted --cases 250`."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_random_shapes_every_entry_point(af):
    import torch
    import plan_fuzz as pf
    rng = np.random.default_rng(7)
    failures = []
    for c in range(90):
        kind, shape, dtype, mode = pf.draw_case(rng, 21)
        err, tol, problems = pf.run_case(af, torch, kind, shape, dtype, mode, 700000 + c)
        if problems:
            failures.append((kind, shape, np.dtype(dtype).name, mode, problems))
    assert not failures, failures


def test_switch_point_shapes(af):
    import torch
    import plan_fuzz as pf
    failures = []
    for i, (kind, shape) in enumerate(pf.targeted_cases()):
        if i % 3 != 0:          # a third of the list per run keeps the suite short; the tool runs all of it
            continue
        dtype = np.complex64 if i % 2 == 0 else np.complex128
        mode = pf.MODES[(i // 3) % 3]
        err, tol, problems = pf.run_case(af, torch, kind, shape, dtype, mode, 800000 + i)
        if problems:
            failures.append((kind, shape, np.dtype(dtype).name, mode, problems))
    assert not failures, failures


@pytest.mark.timeout(180, method="thread")     # a wedged stream must end the process, not hold the box
def test_concurrent_host_threads_share_plans(af):
    """SURVEY.md 8(b) threading: GHC runs callers on different OS threads; plan creation is serialised per cache
    (PTX/Plans.hs:71) but exec is not (:86), so exec must be re-entrant on a shared plan -- here six host threads, each on
    its own stream, race on creating and then executing the same plans (lines, four-step with scratch, the cooperative
    band kernel, Bluestein with a workspace, 3D); every result must equal the single-threaded one bit for bit."""
    import threading

    import torch
    af.lib().accfft_plan_cache_clear()
    g = torch.Generator(device="cuda").manual_seed(5)

    def rnd(shape, dt):
        return torch.view_as_complex(torch.rand(tuple(shape) + (2,), generator=g, device="cuda",
                                                dtype=torch.float32 if dt == torch.complex64 else torch.float64) * 2 - 1)
    work = [(af.fft, "Forward", rnd((16, 512), torch.complex64)),
            (af.fft, "Inverse", rnd((3, 4096), torch.complex128)),
            (af.fft2D, "Forward", rnd((8192, 128), torch.complex64)),
            (af.fft1D, "Forward", rnd((1 << 20,), torch.complex64)),
            (af.fft, "Reverse", rnd((7, 1009), torch.complex64)),
            (af.fft, "Forward", rnd((5, 100003), torch.complex128)),
            (af.fft3D, "Inverse", rnd((64, 64, 64), torch.complex64)),
            (af.fft2D, "Forward", rnd((96, 100), torch.complex128))]
    nthreads, rounds = 6, 3
    results = [[None] * len(work) for _ in range(nthreads)]
    errors = []
    gate = threading.Barrier(nthreads)

    def body(t):
        try:
            s = torch.cuda.Stream()
            gate.wait()
            with torch.cuda.stream(s):
                for r in range(rounds):
                    for i in range(len(work)):
                        j = (i + t) % len(work)       # threads meet on different plans at different times
                        f, mode, x = work[j]
                        results[t][j] = f(mode, x)
            s.synchronize()
        except Exception as e:   # noqa: BLE001
            errors.append((t, repr(e)))

    torch.cuda.synchronize()
    ths = [threading.Thread(target=body, args=(t,)) for t in range(nthreads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    assert not errors, errors
    assert af.lib().accfft_plan_cache_size() <= len(work) + 2      # racing creators end up sharing one plan per key
    for j, (f, mode, x) in enumerate(work):
        ref = f(mode, x)
        torch.cuda.synchronize()
        for t in range(nthreads):
            assert torch.equal(torch.view_as_real(results[t][j]), torch.view_as_real(ref)), (t, j, mode, tuple(x.shape))
