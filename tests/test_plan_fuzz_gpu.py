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
