"""Slab-decomposed 3D FFT across the GPUs of one NVLink box (BASELINE.json config 5).

New capability -- the reference is single-device (its plan cache is merely per-context,
PTX/Plans.hs:68-73).  Parity is defined against the same single-array fft3D (FFT.hs:150-173).

Layout: rank g owns z-planes [g*D/P, (g+1)*D/P) of a dense (D, H, W) array.
  1. local x- and y-axis passes on the slab                       (b200fftPlanAxis x2, hand-written kernels)
  2. pack into P peer blocks [P][dl][H/P][W]                       (b200fftSlabPack)
  3. all-to-all over NVLink: afterwards rank g owns ky-rows [g*H/P,(g+1)*H/P) for ALL z, laid out
     [D][H/P][W] with no unpack needed                             (NCCL; chunked so the exchange of chunk c
                                                                    overlaps the x/y passes of chunk c+1)
  4. z-axis pass on [D][H/P*W]                                     (b200fftPlanAxis)
  5. `transposed_out=False` only: second all-to-all + b200fftSlabUnpack back to z-slabs, so the result has
     exactly fft3D's layout.

The decomposition logic is written against a small `ops` interface so the CPU tests can run it with
world_size 2 over gloo with numpy standing in for the local passes (tests/test_slab_cpu.py).
"""
import ctypes

import numpy as np


class SlabGeometry:
    def __init__(self, d, h, w, world):
        if d % world or h % world:
            raise ValueError("slab decomposition needs D and H divisible by the number of ranks (D=%d H=%d P=%d)" % (d, h, world))
        self.d, self.h, self.w, self.world = d, h, w, world
        self.dl, self.hl = d // world, h // world

    def chunk_bounds(self, chunks):
        chunks = max(1, min(chunks, self.dl))
        while self.dl % chunks:
            chunks -= 1
        c = self.dl // chunks
        return [(i * c, (i + 1) * c) for i in range(chunks)]


def slab_fft3d(geom, ops, x_local, sign, transposed_out=False, chunks=1):
    """The distributed algorithm, generic over `ops` (see GpuOps below / NumpyOps in the tests).
    x_local: (dl, H, W).  Returns (D, hl, W) [ky-slab, transposed_out] or (dl, H, W) [z-slab]."""
    g = geom
    recv = ops.empty((g.d, g.hl, g.w))                 # B[z][kyl][kx]
    works = []
    for (z0, z1) in g.chunk_bounds(chunks):
        a = ops.fft_xy(x_local[z0:z1], sign)           # (dlc, H, W): x then y axis
        send = ops.pack(a, g.world)                    # (P, dlc, hl, W)
        outs = [recv[s * g.dl + z0: s * g.dl + z1] for s in range(g.world)]
        ins = [send[r] for r in range(g.world)]
        works.append((ops.all_to_all(outs, ins), send))  # keep `send` alive until the exchange is done
    for wk, _ in works:
        ops.wait(wk)
    c = ops.fft_z(recv, sign)                          # (D, hl, W): z axis on all z
    if transposed_out:
        return c
    back = ops.empty((g.world, g.dl, g.hl, g.w))
    outs = [back[s] for s in range(g.world)]
    ins = [c[r * g.dl:(r + 1) * g.dl] for r in range(g.world)]
    ops.wait(ops.all_to_all(outs, ins))
    return ops.unpack(back, g.world)                   # (dl, H, W)


class GpuOps:
    """Local passes through the C ABI; exchange through torch.distributed (NCCL)."""

    def __init__(self, geom, dtype, group=None):
        import torch
        import torch.distributed as dist
        from . import Plan, C2C, Z2Z, lib
        self.torch, self.dist, self.group, self.lib = torch, dist, group, lib()
        self.dtype = dtype
        self.typ = C2C if dtype == torch.complex64 else Z2Z
        self.g = geom
        self._plans = {}
        self.Plan = Plan

    def _plan(self, outer, n, inner):
        key = (outer, n, inner)
        if key not in self._plans:
            self._plans[key] = self.Plan("axis", key, self.typ)
        return self._plans[key]

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.dtype, device="cuda")

    def fft_xy(self, a, sign):
        dlc, h, w = a.shape
        t = self.torch.empty_like(a)
        self._plan(dlc * h, w, 1).exec(a, t, sign)
        o = self.torch.empty_like(a)
        self._plan(dlc, h, w).exec(t, o, sign)
        return o

    def fft_z(self, b, sign):
        d, hl, w = b.shape
        o = self.torch.empty_like(b)
        self._plan(1, d, hl * w).exec(b, o, sign)
        return o

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def pack(self, a, world):
        from ._lib import check
        dlc, h, w = a.shape
        out = self.torch.empty((world, dlc, h // world, w), dtype=a.dtype, device="cuda")
        check(self.lib.b200fftSlabPack(self.typ, a.data_ptr(), out.data_ptr(), dlc, h, w, world, self._stream()), "slab pack")
        return out

    def unpack(self, back, world):
        from ._lib import check
        _, dl, hl, w = back.shape
        out = self.torch.empty((dl, hl * world, w), dtype=back.dtype, device="cuda")
        check(self.lib.b200fftSlabUnpack(self.typ, back.data_ptr(), out.data_ptr(), dl, hl * world, w, world, self._stream()), "slab unpack")
        return out

    def all_to_all(self, outs, ins):
        return self.dist.all_to_all(outs, ins, group=self.group, async_op=True)

    def wait(self, work):
        work.wait()


class SlabFFT3D:
    """fft3D of a (D,H,W) array whose z-slabs live on the ranks of `group` (one process per GPU)."""

    def __init__(self, d, h, w, dtype, group=None, chunks=4):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        self.geom = SlabGeometry(d, h, w, world)
        self.ops = GpuOps(self.geom, dtype, group)
        self.chunks = chunks

    def __call__(self, mode, x_local, transposed_out=False):
        from . import FORWARD, INVERSE, Inverse, Forward
        sign = FORWARD if mode == Forward else INVERSE
        y = slab_fft3d(self.geom, self.ops, x_local, sign, transposed_out, self.chunks)
        if mode == Inverse:   # FFT.hs:155,172: scale by the whole size
            y = y / float(self.geom.d * self.geom.h * self.geom.w)
        return y


def bench_slab(args, af, dist, rank, local, world, desc, measured_peak, ClockSampler):
    """bench.py --config cfg5 at N>1 GPUs: 1024^3 c64, z-slabs, strong scaling."""
    import json
    import math
    import torch
    d = h = w = 1024
    geom = SlabGeometry(d, h, w, world)
    torch.manual_seed(1005 + rank)
    x = torch.view_as_complex(torch.rand(geom.dl, h, w, 2, dtype=torch.float32, device="cuda") * 2 - 1)
    fft = SlabFFT3D(d, h, w, torch.complex64, None, chunks=args.chunks if hasattr(args, "chunks") else 4)
    results = {}
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = af.kernel_launches()
    for name, tr in (("transposed_out", True), ("natural_out", False)):
        for _ in range(max(3, args.warmup)):
            y = fft(af.Forward, x, transposed_out=tr)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = fft(af.Forward, x, transposed_out=tr)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[name] = float(t.item()) / args.steps
        del y
    launches = af.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    flops = 5.0 * d * h * w * math.log2(d * h * w)
    ms = results["natural_out"]
    peak, peak_src = measured_peak()
    slab_bytes = geom.dl * h * w * 8
    hbm_alg = 3 * 2 * slab_bytes                       # three axis passes over the local slab
    nvl_out = slab_bytes * (world - 1) / world          # bytes each GPU sends per exchange
    if rank == 0:
        line = {
            "metric": "fft_gflops_5nlog2n", "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + ", z-slab decomposed, NCCL all-to-all", "per_gpu_shape": [geom.dl, h, w],
                       "output_layout": "natural z-slabs (two exchanges); transposed-out (one exchange) reported beside it",
                       "chunks": fft.chunks, "l2": "slab larger than L2"},
            "transposed_out": {"ms_per_step": results["transposed_out"], "value": flops / (results["transposed_out"] * 1e-3) / 1e9,
                               "unit": "GFLOP/s"},
            "roofline": {"bound": "hbm", "achieved": hbm_alg / (results["transposed_out"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": hbm_alg / (results["transposed_out"] * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "note": "per-GPU algorithmic HBM bytes (3 passes x 2 x slab) over the whole transposed-out step, exchange included",
                         "nvlink_out_bytes_per_gpu_per_exchange": nvl_out,
                         "nvlink_gbs_if_exchange_were_the_whole_step": nvl_out / (results["transposed_out"] * 1e-3) / 1e9},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0
