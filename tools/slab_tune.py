#!/usr/bin/env python
"""Developer sweep of the C-ABI slab transform's pipelining knobs (plane chunks, column chunks, CTAs of the scatter pass).
torchrun --nproc-per-node N tools/slab_tune.py   -- 1024^3 c64 Forward, ms per transform (max over ranks)."""
import os, sys, itertools
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import accelerate_fft_b200 as af
from accelerate_fft_b200.slab import SlabPlan

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
d = h = w = 1024
x = torch.view_as_complex(torch.rand(d // world, h, w, 2, dtype=torch.float32, device="cuda") * 2 - 1)
sp = SlabPlan(d, h, w, torch.complex64, None)
outs = {True: torch.empty((h // world, d, w), dtype=torch.complex64, device="cuda"),
        False: sp.natural_buffer()}     # the library-owned buffer: no final copy

def timeit(tr, steps=10):
    for _ in range(3):
        sp(af.Forward, x, transposed_out=tr, out=outs[tr])
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sp(af.Forward, x, transposed_out=tr, out=outs[tr])
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps

t0 = (timeit(True), timeit(False))
if rank == 0:
    print("library defaults              transposed-out %.3f ms   natural %.3f ms" % t0, flush=True)
grid = [(1, 1, 0), (1, 1, 96), (1, 1, 120)]
grid += [(cp, ck, y) for ck in (2, 4) for cp in (1, 2) for y in (0, 80, 96, 112, 128)]
if len(sys.argv) > 1:
    grid = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (cp, ck, y) in grid:
    sp.tune(True, cp, ck, y)
    sp.tune(False, cp, ck, y)
    t_tr, t_nat = timeit(True), timeit(False)
    if rank == 0:
        print("planes=%d cols=%d y_ctas=%3d   transposed-out %.3f ms   natural %.3f ms" % (cp, ck, y, t_tr, t_nat), flush=True)
sp.close()
dist.barrier()
dist.destroy_process_group()
