#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native FFT hot path (contract in the task brief).

Default workload = BASELINE.json configs[1]: batched 1D complex Double FFT, n=4096, batch=65536,
one step = Forward then Inverse (normalised) over the whole batch.  The rows are independent units: at N GPUs
every rank transforms a full batch of its own with no data-path collective (weak scaling; --scaling strong splits
BASELINE's 65536 rows over the ranks instead).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg1..cfg5] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = whole-job GFLOP/s (5 N log2 N per transform) with inputs
resident in HBM; `e2e` = the same through host buffers (pinned H2D of the step's input, D2H of its
result inside the timed region); `roofline` = algorithmic bytes / measured launch time of the
dominant kernel against the measured HBM copy bandwidth; `cpu_baseline` = the oracle port of the
reference's pure-Accelerate path (oracle/, C) on the box's host cores, bounded sample.
--impl reference times that CPU path alone (the reference's GPU/CPU bindings need GHC+cuFFT/FFTW,
which this image does not have).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (kind, dims, dtype, batch, min passes, description)
    "cfg1": ("fft", (1024,), "c64", 4096, 1, "1D c64 n=1024 batch=4096 Forward"),
    "cfg2": ("fft", (4096,), "c128", 65536, 1, "batched 1D c128 n=4096 batch=65536 Forward+Inverse"),
    "cfg3": ("fft2D", (8192, 8192), "c64", 1, 2, "2D c64 8192x8192 Forward"),
    "cfg4": ("fft1D", (1 << 28,), "c64", 1, 2, "1D c64 n=2^28 Forward (four-step)"),
    "cfg5": ("fft3D", (1024, 1024, 1024), "c64", 1, 3, "3D c64 1024^3 Forward"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(cfg):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(cfg)
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                c, m, pw = float(r[0]), float(r[1]), float(r[2])
            except Exception:
                continue
            mx = m
            if c >= 0.5 * m:   # under load (idle parks at ~120 MHz)
                sm.append(c)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            sm = [float(r[0]) for r in rows if r and r[0].strip().replace(".", "").isdigit()]
        out["sm_mhz"] = statistics.median(sm) if sm else None
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        out["samples"] = len(rows)
        return out


def dominant_kernel(plan_desc):
    if any("band A[" in d for d in plan_desc):
        return "fft_band_kernel (+ fft_ring_rows_kernel rows)"
    names = []
    if any("| ring:" in d for d in plan_desc):
        names.append("fft_ring_rows_kernel")
    if any("| pipe" in d for d in plan_desc):
        names.append("fft_pipe_cols_kernel")
    if any("| ring cols" in d for d in plan_desc):
        names.append("fft_ringcol_kernel")
    if any("| ring trans" in d for d in plan_desc):
        names.append("fft_ringtrans_kernel")
    if any("|" not in d for d in plan_desc) or not names:
        names.insert(0, "fft_lines_kernel")
    return " + ".join(names)


def flops_of(cfg, batch):
    kind, dims, dt, _, _, _ = CONFIGS[cfg]
    npts = 1
    for d in dims:
        npts *= d
    per = 5.0 * npts * math.log2(npts) * batch
    return per * (2 if cfg == "cfg2" else 1)      # cfg2's step is Forward + Inverse


def workload_geometry(cfg, world, scaling_mode):
    """How one config is laid over `world` ranks: (replicas, rows per rank, scaling label, per-rank shape, bytes per
    buffer, rotating buffers).  Pure arithmetic -- both arms describe the workload with it."""
    kind, dims, dtp, batch, _, _ = CONFIGS[cfg]
    esz = 8 if dtp == "c64" else 16
    if cfg in ("cfg3", "cfg4", "cfg5") and world > 1:
        # BASELINE.json names these single-GPU (cfg5's multi-GPU form is the slab transform, reported under "slab")
        replicas, my_batch, scaling = world, 1, "weak"
    else:
        if batch % world:
            raise SystemExit("batch %d not divisible by %d ranks" % (batch, world))
        if scaling_mode == "strong":      # BASELINE's rows split over the ranks ("2 @ P": [65536/P, 4096] per GPU)
            replicas, my_batch, scaling = 1, batch // world, "strong"
        else:                             # every rank transforms a full BASELINE batch of its own
            replicas, my_batch, scaling = world, batch, "weak"
    shape = ((my_batch,) + dims) if kind == "fft" else dims
    n_local = 1
    for d in shape:
        n_local *= d
    nbytes = n_local * esz
    # several distinct buffer pairs when one fits in L2 (cfg1): rotate so every step streams from HBM
    nbuf = 1 if nbytes * 2 > 4 * 126e6 else int(math.ceil(4 * 126e6 / (2 * nbytes)))
    return replicas, my_batch, scaling, shape, nbytes, nbuf


def config_object(cfg, world, scaling_mode):
    """The `config` key of the JSON line: identical in the b200 arm and the --impl reference arm (same workload, same
    sharding); what the reference arm actually samples of it is said in its cpu_baseline.sample."""
    kind, dims, dtp, batch, _, desc = CONFIGS[cfg]
    _, my_batch, scaling, shape, nbytes, nbuf = workload_geometry(cfg, world, scaling_mode)
    return {"workload": desc, "per_gpu_shape": list(shape),
            "sharding": ("the batch rows are the units: every rank owns %d of the %d rows, no data-path collective"
                         % (my_batch, my_batch * world if scaling == "weak" else batch)) if kind == "fft" else "independent replicas",
            "l2": "inputs larger than L2 (%.0f MB per buffer x %d rotating buffers)" % (nbytes / 1e6, nbuf),
            "inverse_scale": "fused into the last pass", "step": "Forward+Inverse" if cfg == "cfg2" else "Forward"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------------
def cpu_step(cfg, sample_batch, threads, np, oracle, x):
    kind, dims, dt, _, _, _ = CONFIGS[cfg]
    if cfg == "cfg2":
        y = oracle.fft("Forward", x, threads=threads)
        return oracle.fft("Inverse", y, threads=threads)
    return getattr(oracle, kind)("Forward", x, threads=threads)


def cpu_sample(cfg, np):
    """A bounded sample of the workload (same shape per transform, fewer transforms)."""
    kind, dims, dt, batch, _, _ = CONFIGS[cfg]
    dtype = np.complex64 if dt == "c64" else np.complex128
    rng = np.random.default_rng(1000 + int(cfg[3]))
    if cfg == "cfg1":
        shape, sb, what = (4096, 1024), 4096, "the whole workload (4096 rows)"
    elif cfg == "cfg2":
        shape, sb, what = (4096, 4096), 4096, "4096 of the 65536 rows (1/16 of the batch), Forward+Inverse"
    elif cfg == "cfg3":
        shape, sb, what = (2048, 2048), 1.0 / 16, "a 2048x2048 fft2D (1/16 of the points; GFLOP/s by its own 5NlogN)"
    elif cfg == "cfg4":
        shape, sb, what = (1 << 24,), 1.0 / 16, "a 2^24-point fft1D (1/16 of the points; GFLOP/s by its own 5NlogN)"
    else:
        shape, sb, what = (256, 256, 256), 1.0 / 64, "a 256^3 fft3D (1/64 of the points; GFLOP/s by its own 5NlogN)"
    x = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dtype)
    if cfg in ("cfg1", "cfg2"):
        fl = flops_of(cfg, sb)
    else:
        npts = x.size
        fl = 5.0 * npts * math.log2(npts)
    return x, fl, what


def run_cpu_baseline(cfg, steps=1, warmup=0):
    import numpy as np
    import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    x, fl, what = cpu_sample(cfg, np)
    for _ in range(warmup):
        cpu_step(cfg, None, threads, np, oracle, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(cfg, None, threads, np, oracle, x)
    dt = (time.perf_counter() - t0) / steps
    out = {"value": fl / dt / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
           "sample": what + "; oracle/ C port of Adhoc.hs (pure-Accelerate path), OpenMP over rows", "ms_per_step": dt * 1e3}
    # FFTW-class stand-in for the reference's llvm-cpu path (FFTW itself is not in the image)
    try:
        import scipy.fft as sf
        kind = CONFIGS[cfg][0]
        f = {"fft": lambda a: sf.fft(a, axis=-1, workers=threads), "fft1D": lambda a: sf.fft(a, workers=threads),
             "fft2D": lambda a: sf.fft2(a, workers=threads), "fft3D": lambda a: sf.fftn(a, workers=threads)}[kind]
        f(x)
        t0 = time.perf_counter()
        y = f(x)
        if cfg == "cfg2":
            sf.ifft(y, axis=-1, workers=threads)
        d2 = time.perf_counter() - t0
        out["fftw_standin"] = {"impl": "scipy.fft (pocketfft)", "value": fl / d2 / 1e9, "unit": "GFLOP/s", "cores": threads}
    except Exception:
        pass
    return out, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = args.config
    base, dt = run_cpu_baseline(cfg, steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
    kind, dims, dtp, batch, _, desc = CONFIGS[cfg]
    line = {
        "impl": "reference", "metric": "fft_gflops_5nlog2n", "value": base["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": workload_geometry(cfg, max(1, args.gpus), args.scaling)[2],
        "vs_baseline": None, "dtype": "f64" if dtp == "c128" else "f32", "data": "synthetic",
        "config": config_object(cfg, max(1, args.gpus), args.scaling),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's own bindings need GHC + FFTW/cuFFT (absent here): this arm is the oracle port of its pure-Accelerate CPU path",
    }
    if "fftw_standin" in base:
        line["fftw_standin"] = base["fftw_standin"]
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def device_run(cfg, args, af, torch, dist, world, rank, scaling_mode, keep_alive=False):
    """Device-resident timing of one BASELINE config through the drop-in boundary (b200fftExec / b200fftExecScaled on a
    cached plan with caller-owned buffers -- what PTX.hs:77-106 does after allocateRemote).  Returns a dict with the
    whole-job value, the per-transform launch time measured inside the timed region and the roofline object; with
    keep_alive the buffers / plan / step function come back too (the headline's e2e + clock legs re-use them)."""
    kind, dims, dtp, batch, min_passes, desc = CONFIGS[cfg]
    dt = torch.complex64 if dtp == "c64" else torch.complex128
    esz = 8 if dtp == "c64" else 16
    replicas, my_batch, scaling, shape, nbytes, nbuf = workload_geometry(cfg, world, scaling_mode)
    torch.manual_seed(1000 + int(cfg[3]) + 7919 * rank)
    xs = [torch.view_as_complex(torch.rand(shape + (2,), dtype=torch.float32 if dtp == "c64" else torch.float64, device="cuda") * 2 - 1)
          for _ in range(nbuf)]
    plan_kind = {"fft": "many", "fft1D": "1d", "fft2D": "2d", "fft3D": "3d"}[kind]
    plan = af.Plan(plan_kind, list(dims), af.C2C if dtp == "c64" else af.Z2Z, my_batch if kind == "fft" else 1)
    ys = [torch.empty_like(x) for x in xs]
    zs = [torch.empty_like(x) for x in xs] if cfg == "cfg2" else None
    inv_scale = 1.0 / dims[-1]       # 1/n folded into the last butterfly pass (same result as FFT.hs:83's extra map)

    def step(i):
        x, y = xs[i % nbuf], ys[i % nbuf]
        plan.exec(x, y, af.FORWARD)
        if cfg == "cfg2":
            plan.exec(y, zs[i % nbuf], af.INVERSE, scale=inv_scale)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps = args.steps
    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    npass = plan.num_passes
    plan_desc = plan.describe().strip().split("\n")
    per_step = 2 if cfg == "cfg2" else 1
    inner_events = nbytes >= (256 << 20)     # short transforms: an event pair per transform would be timed, not the kernel

    # Short transforms (cfg1: ~15 us) are captured once into a CUDA graph -- the K steps of the timed region, back to back
    # on one stream over the rotating buffers -- so the timed region holds the library's kernels and nothing of the Python /
    # ctypes dispatch (~2 us per call); single-pass plans only (no stream-ordered scratch inside a capture).
    graph = None
    if not inner_events and plan.scratch_bytes == 0 and not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for i in range(steps):
                        step(i)
            torch.cuda.current_stream().wait_stream(side)
            g.replay()
            torch.cuda.synchronize()
            graph = g
        except Exception as ex:   # capture refused: time the plain loop
            print("bench.py: CUDA graph capture failed for %s (%r); timing the direct loop" % (cfg, ex), file=sys.stderr)
            torch.cuda.synchronize()
            graph = None

    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(per_step + 1)] for _ in range(steps)] if inner_events else None
    l0 = af.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(steps):
            x, y = xs[i % nbuf], ys[i % nbuf]
            if inner_events:
                evs[i][0].record()
            plan.exec(x, y, af.FORWARD)
            if inner_events:
                evs[i][1].record()
            if cfg == "cfg2":
                plan.exec(y, zs[i % nbuf], af.INVERSE, scale=inv_scale)
                if inner_events:
                    evs[i][2].record()
    ev1.record()
    barrier()
    launches = (af.kernel_launches() - l0) if graph is None else steps * per_step * npass
    ms_total = ev0.elapsed_time(ev1)
    repeats = 1
    if graph is not None and getattr(args, "repeat", 1) > 1:
        # a graph replay of K ~15 us transforms is a ~150 us region: one sample of it is clock-ramp noise (0.67-0.73 run to run).
        # The side configs take the median of several identical regions.
        regions = [ms_total]
        for _ in range(args.repeat - 1):
            barrier()
            ev0.record()
            graph.replay()
            ev1.record()
            barrier()
            regions.append(ev0.elapsed_time(ev1))
        ms_total = statistics.median(regions)
        repeats = len(regions)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    # strong: the BASELINE batch is the whole job (replicas = 1); weak / replicas: every rank a full unit of its own
    total_flops = flops_of(cfg, batch if kind == "fft" else 1) * replicas
    value = total_flops / (ms_step * 1e-3) / 1e9
    if inner_events:
        samples = [e[j].elapsed_time(e[j + 1]) for e in evs for j in range(per_step)]
        exec_ms = statistics.mean(samples)
        how = "CUDA events around every transform inside the timed region, mean over %d transforms" % len(samples)
    else:   # launches back to back on one stream: the step time is the launch time
        samples = [ms_step / per_step] * (steps * per_step)
        exec_ms = ms_step / per_step
        how = ("CUDA events around the %d back-to-back transforms of the timed region / %d" % (len(samples), len(samples))) + \
              ("; the region is one CUDA graph replay (no host dispatch between the launches)" if graph is not None else "") + \
              ("; median of %d such regions" % repeats if repeats > 1 else "")
    peak, peak_src = measured_peak()
    alg_bytes = min_passes * 2 * nbytes              # min passes x 2 x N_total x sizeof(complex)
    achieved = alg_bytes / (exec_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(cfg), "peak_source": peak_src,
                "kernel": "%s (%d launch(es) per transform direction; algorithmic passes %d)" % (dominant_kernel(plan_desc), npass, min_passes),
                "algorithmic_bytes_per_exec": alg_bytes, "exec_ms": exec_ms, "exec_samples": len(samples), "how": how,
                "per_pass_frac": (npass * 2 * nbytes) / (exec_ms * 1e-3) / 1e9 / peak, "plan": plan_desc}
    # a device copy of the SAME size through the same rotation: what "HBM speed" means for this working set
    # (a 32 MB copy reaches ~71 % of the large-copy peak: ramp-up and drain of a ~15 us kernel)
    try:
        for i in range(3):
            ys[i % nbuf].copy_(xs[i % nbuf])
        torch.cuda.synchronize()
        ncopy = max(5, steps)
        cgraph = None
        if graph is not None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                cgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cgraph, stream=side):
                    for i in range(ncopy):
                        ys[i % nbuf].copy_(xs[i % nbuf])
            torch.cuda.current_stream().wait_stream(side)
            cgraph.replay()
            torch.cuda.synchronize()
        ev0.record()
        if cgraph is not None:
            cgraph.replay()
        else:
            for i in range(ncopy):
                ys[i % nbuf].copy_(xs[i % nbuf])
        ev1.record()
        torch.cuda.synchronize()
        copy_ms = ev0.elapsed_time(ev1) / ncopy
        roofline["same_size_copy_gbs"] = 2 * nbytes / (copy_ms * 1e-3) / 1e9
        roofline["frac_of_same_size_copy"] = achieved / roofline["same_size_copy_gbs"]
    except Exception:
        pass
    res = {"cfg": cfg, "value": value, "ms_step": ms_step, "roofline": roofline, "launches": int(launches), "scaling": scaling,
           "shape": list(shape), "nbytes": nbytes, "nbuf": nbuf, "my_batch": my_batch, "batch": batch, "replicas": replicas,
           "total_flops": total_flops, "graph": graph is not None}
    if keep_alive:
        res.update({"step": step, "xs": xs, "plan": plan, "barrier": barrier})
    else:
        del xs, ys, zs, graph
        plan.destroy()
        torch.cuda.empty_cache()
        af.lib().b200fftTrimScratch()
    return res


def slab_object(args, af, torch, dist, world, rank, local):
    """cfg5 (1024^3 c64 fft3D) z-slab decomposed over the `world` GPUs, strong scaling: parity first (folded-bin identity
    against the oracle's fft3D on the 16^3 fold of the same input), then peer-scatter / NCCL x transposed-out /
    natural-out timings, NVLink rate and the efficiency against the single-GPU transform timed in the same run (rank 0)."""
    import numpy as np
    from accelerate_fft_b200 import slab as S
    d = h = w = 1024
    geom = S.SlabGeometry(d, h, w, world)
    torch.manual_seed(1005 + 7919 * rank)
    x = torch.view_as_complex(torch.rand(geom.dl, h, w, 2, dtype=torch.float32, device="cuda") * 2 - 1)
    m = 16
    out = {"workload": "fft3D c64 1024^3 Forward, z-slabs [%d,1024,1024] per GPU (strong scaling)" % geom.dl, "n_gpus": world}

    def fold_local(a, axis):
        shp = list(a.shape)
        L = shp[axis]
        return a.reshape(shp[:axis] + [L // m, m] + shp[axis + 1:]).to(torch.complex128).sum(dim=axis)

    # fold of the GLOBAL input: local fold over x and y, then over the local z (dl is a multiple of 16), summed over ranks
    f = fold_local(fold_local(fold_local(x, 2), 1), 0)
    fr = torch.view_as_real(f).contiguous()
    dist.all_reduce(fr)
    folded = torch.view_as_complex(fr).cpu().numpy()
    ref = None
    if rank == 0:
        try:
            import oracle          # the checker (test infrastructure), never on the timed path
            oracle.build()
            ref = oracle.fft3D("Forward", folded)
            out["parity_oracle"] = "oracle.fft3D (C port of FFT.hs:166-187 + Adhoc.hs) on the 16^3 fold of the input"
        except Exception as ex:
            ref = np.fft.fftn(folded)
            out["parity_oracle"] = "numpy fftn on the 16^3 fold (oracle unavailable: %r)" % (ex,)

    def gather_bins(y, natural, ky_major=False):
        """Y[64 kz, 64 ky, 64 kx] from the distributed result, summed over ranks (each bin lives on one rank)."""
        st = d // m
        bins = torch.zeros(m, m, m, dtype=torch.complex128, device="cuda")
        if ky_major:     # the C ABI's transposed-out layout [hl][D][W]
            y = y.transpose(0, 1)
        if natural:      # y: [dl][H][W], rank owns z in [rank*dl, (rank+1)*dl)
            for kz in range(m):
                z = kz * st
                if rank * geom.dl <= z < (rank + 1) * geom.dl:
                    bins[kz] = y[z - rank * geom.dl, ::st, ::st].to(torch.complex128)
        else:            # y: [D][hl][W], rank owns ky in [rank*hl, (rank+1)*hl)
            for ky in range(m):
                yy = ky * st
                if rank * geom.hl <= yy < (rank + 1) * geom.hl:
                    bins[:, ky] = y[::st, yy - rank * geom.hl, ::st].to(torch.complex128)
        br = torch.view_as_real(bins).contiguous()
        dist.all_reduce(br)
        return torch.view_as_complex(br).cpu().numpy()

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    def time_it(fn):
        for _ in range(max(3, args.warmup)):
            y = fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del y
        return float(t.item()) / args.steps

    ms, par = {}, {}
    l0 = af.kernel_launches()
    try:
        # the C-ABI entry point (b200fftPlanSlab3d / b200fftExecSlab): pipeline, peer mappings and barriers in the library
        sp = S.SlabPlan(d, h, w, torch.complex64, None)
        outs = {True: torch.empty((geom.hl, d, w), dtype=torch.complex64, device="cuda"), False: sp.natural_buffer()}
        for name, tr in (("cabi_transposed_out", True), ("cabi_natural_out", False)):
            y = sp(af.Forward, x, transposed_out=tr, out=outs[tr])
            got = gather_bins(y, natural=not tr, ky_major=tr)
            if rank == 0:
                par[name] = rel(got, ref)
            del y
            ms[name] = time_it(lambda: sp(af.Forward, x, transposed_out=tr, out=outs[tr]))
        del outs
        sp.close()
        del sp
    except Exception as ex:   # peer mappings unavailable (no P2P between the GPUs): the NCCL path stands
        out["cabi_error"] = repr(ex)[:300]
    torch.cuda.empty_cache()
    try:
        nf = S.SlabFFT3D(d, h, w, torch.complex64, None, chunks=4)
        for name, tr in (("nccl_transposed_out", True), ("nccl_natural_out", False)):
            y = nf(af.Forward, x, transposed_out=tr)
            got = gather_bins(y, natural=not tr)
            if rank == 0:
                par[name] = rel(got, ref)
            del y
            ms[name] = time_it(lambda: nf(af.Forward, x, transposed_out=tr))
        del nf
    except Exception as ex:
        out["nccl_error"] = repr(ex)[:300]
    out["gpu_launches"] = int(af.kernel_launches() - l0)
    del x
    torch.cuda.empty_cache()
    # the single-GPU transform of the same problem, timed here on rank 0 alone (the others wait at the barrier)
    single = None
    if rank == 0:
        xs = torch.view_as_complex(torch.rand(d, h, w, 2, dtype=torch.float32, device="cuda") * 2 - 1)
        ys = torch.empty_like(xs)
        p1 = af.Plan("3d", [d, h, w], af.C2C, 1)
        for _ in range(3):
            p1.exec(xs, ys, af.FORWARD)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            p1.exec(xs, ys, af.FORWARD)
        e1.record()
        torch.cuda.synchronize()
        single = e0.elapsed_time(e1) / args.steps
        p1.destroy()
        del xs, ys
        torch.cuda.empty_cache()
    dist.barrier()
    if rank != 0:
        return None
    flops = 5.0 * d * h * w * math.log2(d * h * w)
    slab_bytes = geom.dl * h * w * 8
    nvl_out = slab_bytes * (world - 1) / world
    bar = 1e-5 * math.log2(d * h * w)
    out.update({
        "ms_per_step": ms, "parity_rel_l2": max(par.values()) if par else None, "parity_rel_l2_by_variant": par, "parity_bar": bar,
        "parity_ok": bool(par) and max(par.values()) <= bar,
        "parity_how": "Y[64kz,64ky,64kx] of the distributed result == fft3D of the input folded to 16^3 (the bin-decimation identity of tests/test_parity_gpu.py::test_cfg5_full_size), checked before timing for every variant",
        "single_gpu_ms": single,
        "efficiency": {k: single / (world * v) for k, v in ms.items()} if single else None,
        "gflops": {k: flops / (v * 1e-3) / 1e9 for k, v in ms.items()},
        "nvlink_out_bytes_per_gpu_per_exchange": nvl_out,
        "nvlink_out_gbs_per_gpu": {k: nvl_out * (1 if "transposed" in k else 2) / (v * 1e-3) / 1e9 for k, v in ms.items()},
        "nvlink_note": "bytes each GPU sends (one exchange transposed-out, two natural) / the WHOLE step time: a lower bound on the link rate during the exchange; measured peer copy 770 GB/s per direction (B200_PROFILING.md)",
        "hbm_frac_per_gpu": {k: 3 * 2 * slab_bytes / (v * 1e-3) / 1e9 / measured_peak()[0] for k, v in ms.items()},
        "variants": "cabi_* = b200fftExecSlab (C ABI, csrc/slab.cu): the y (and z) pass stores scattered straight into the owning rank's memory over NVLink peer mappings, z of a column chunk beside y of the next, flag barriers in peer memory, no NCCL on the data path; transposed-out = [H/P][D][W]; natural-out lands in the library-owned buffer (b200fftSlabNaturalBuffer; a caller-owned output adds one device copy of the slab).  nccl_* = pack + NCCL all-to-all (+ unpack), 4 chunks overlapped",
    })
    return out


def main_gpu(args):
    import torch
    import accelerate_fft_b200 as af
    af.lib()   # fail loudly if the extension is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world and rank == 0 and world > 1:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)

    # clocks are sampled from before the warm-up until after the launch-time measurement (the timed
    # region itself can be shorter than one nvidia-smi period)
    sampler = ClockSampler(local) if rank == 0 else None
    cfg = args.config
    kind, dims, dtp, batch, min_passes, desc = CONFIGS[cfg]
    dt = torch.complex64 if dtp == "c64" else torch.complex128
    if cfg == "cfg5" and world > 1:
        from accelerate_fft_b200 import slab
        return slab.bench_slab(args, af, dist, rank, local, world, desc, measured_peak, ClockSampler)
    head = device_run(cfg, args, af, torch, dist, world, rank, args.scaling, keep_alive=True)
    step, xs, barrier = head["step"], head["xs"], head["barrier"]
    shape, nbytes, nbuf = tuple(head["shape"]), head["nbytes"], head["nbuf"]
    scaling, my_batch, total_flops = head["scaling"], head["my_batch"], head["total_flops"]

    # keep the GPU loaded a little longer for the clock sampler (the timed region can be shorter than one
    # nvidia-smi period); nothing measured here is reported
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:
        for i in range(4):
            step(i)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    # ---- end to end through host buffers (pinned H2D of the input, D2H of the result) ----------
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(shape, dtype=dt).pin_memory()
        hx.copy_(xs[0].cpu())
        hy = torch.empty(shape, dtype=dt).pin_memory()
        ksteps = max(1, min(args.steps, 3 if nbytes > 1e9 else args.steps))
        modes = ["Forward", "Inverse"] if cfg == "cfg2" else ["Forward"]

        def e2e_step():
            # the repo's public host-buffer call (accfft_run_host_seq): H2D of the step's input, the transforms,
            # D2H of the result -- chunked over rows for `fft` so the copies overlap the kernels; synchronous
            af.run_host_seq(kind, modes, hx, hy)
        e2e_step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(ksteps):
            e2e_step()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / ksteps
        e2e = {"value": total_flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_ms, "steps": ksteps,
               "api": "accfft_run_host_seq(kind=%s, modes=%s) on pinned host buffers: H2D + transforms + D2H inside the C ABI call%s"
                      % (kind, "+".join(modes), ", rows pipelined in chunks over 3 streams" if kind == "fft" else "")}
        del hx, hy
    # release the headline's buffers before the other configs
    head["plan"].destroy()
    for k in ("step", "xs", "plan", "barrier"):
        head.pop(k, None)
    del step, xs
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, device-resident, same method (N=1) -------------------------
    configs = None
    total_launches = head["launches"]
    if world == 1 and not args.no_configs:
        configs = {}
        for c in sorted(CONFIGS):
            if c == cfg:
                r = head
            else:
                try:
                    # the side configs are diagnostics beside the headline: at least 10 transforms each, whatever --steps says
                    # (three transforms of a 15 us kernel measure the launch ramp, not the kernel)
                    cargs = argparse.Namespace(**vars(args))
                    cargs.steps = max(args.steps, 10)
                    cargs.repeat = 7
                    r = device_run(c, cargs, af, torch, None, 1, 0, "strong")
                    total_launches += r["launches"]
                except Exception as ex:
                    configs[c] = {"error": repr(ex)[:300]}
                    continue
            rf = r["roofline"]
            configs[c] = {"workload": CONFIGS[c][5], "gflops": r["value"], "ms_per_step": r["ms_step"], "exec_ms": rf["exec_ms"],
                          "frac": rf["frac"], "per_pass_frac": rf["per_pass_frac"], "achieved_gbs": rf["achieved"],
                          "algorithmic_bytes_per_exec": rf["algorithmic_bytes_per_exec"], "traffic": rf["traffic"],
                          "passes": len(rf["plan"]), "min_passes": CONFIGS[c][4], "plan": rf["plan"], "how": rf["how"],
                          "same_size_copy_gbs": rf.get("same_size_copy_gbs"), "frac_of_same_size_copy": rf.get("frac_of_same_size_copy"),
                          "cuda_graph": r["graph"]}

    # ---- the slab-decomposed 3D transform (the one config with an exchange step), N > 1 ---------
    slab = None
    if world > 1 and not args.no_slab:
        try:
            slab = slab_object(args, af, torch, dist, world, rank, local)
        except Exception as ex:
            slab = {"error": repr(ex)[:500]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = run_cpu_baseline(cfg)

    if rank == 0:
        roofline = head["roofline"]
        alg_bytes = roofline["algorithmic_bytes_per_exec"]
        ms_step = head["ms_step"]
        line = {
            "metric": "fft_gflops_5nlog2n", "value": head["value"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64" if dtp == "c128" else "f32", "data": "synthetic",
            "config": config_object(cfg, world, args.scaling),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(total_launches), "clocks": clocks,
            "hbm_gbs_whole_step": (2 if cfg == "cfg2" else 1) * alg_bytes * world / (ms_step * 1e-3) / 1e9,
        }
        if configs is not None:
            line["configs"] = configs
        if slab is not None:
            line["slab"] = slab
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version on the first
    communicator, torchrun children inherit the descriptor): everything but the final line goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="batched-1D configs at N>1 GPUs: strong = BASELINE's batch split over the ranks (default: [65536/P,4096] per GPU), weak = a full batch per GPU")
    ap.add_argument("--no-configs", action="store_true", help="N=1: skip the `configs` object (the other four BASELINE configs)")
    ap.add_argument("--no-slab", action="store_true", help="N>1: skip the `slab` object (cfg5 1024^3 z-slab decomposed)")
    ap.add_argument("--no-graph", action="store_true", help="time short transforms through the plain Python loop instead of one CUDA graph replay")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
