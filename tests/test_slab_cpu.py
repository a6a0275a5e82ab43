"""Host-side logic of the multi-GPU paths, exercised on CPU with world_size 2 over gloo:
  * the slab decomposition / pack / all-to-all / unpack index logic of accelerate_fft_b200.slab
    (numpy stands in for the local GPU passes -- this tests the plumbing, not the kernels);
  * the batch sharding used by bench.py for the batched-1D configs."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class NumpyOps:
    """Same interface as slab.GpuOps; local passes by numpy, exchange by gloo send/recv."""

    def __init__(self, dist, rank, world):
        self.dist, self.rank, self.world = dist, rank, world

    def empty(self, shape):
        return np.empty(shape, dtype=np.complex128)

    def fft_xy(self, a, sign):
        f = np.fft.fft if sign < 0 else (lambda v, axis: np.fft.ifft(v, axis=axis) * v.shape[axis])
        return f(f(a, axis=2), axis=1)

    def fft_z(self, b, sign):
        f = np.fft.fft if sign < 0 else (lambda v, axis: np.fft.ifft(v, axis=axis) * v.shape[axis])
        return f(b, axis=0)

    def pack(self, a, world):        # numpy mirror of b200fftSlabPack: [dl][h][w] -> [P][dl][h/P][w]
        dl, h, w = a.shape
        return np.ascontiguousarray(a.reshape(dl, world, h // world, w).transpose(1, 0, 2, 3))

    def unpack(self, back, world):   # mirror of b200fftSlabUnpack
        _, dl, hl, w = back.shape
        return np.ascontiguousarray(back.transpose(1, 0, 2, 3).reshape(dl, hl * world, w))

    def all_to_all(self, outs, ins):
        import torch
        reqs, tmp = [], []
        for r in range(self.world):
            if r == self.rank:
                outs[r][...] = ins[r]
                continue
            t_in = torch.from_numpy(np.ascontiguousarray(ins[r]).view(np.float64))
            t_out = torch.empty(outs[r].size * 2, dtype=torch.float64)
            reqs.append(self.dist.isend(t_in.reshape(-1), dst=r))
            reqs.append(self.dist.irecv(t_out, src=r))
            tmp.append((r, t_out))
        return (reqs, tmp, outs)

    def wait(self, work):
        reqs, tmp, outs = work
        for q in reqs:
            q.wait()
        for r, t in tmp:
            outs[r][...] = t.numpy().view(np.complex128).reshape(outs[r].shape)


def _worker(rank, world, port, shape, chunks, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from accelerate_fft_b200.slab import SlabGeometry, slab_fft3d
        d, h, w = shape
        rng = np.random.default_rng(99)
        full = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)
        geom = SlabGeometry(d, h, w, world)
        ops = NumpyOps(dist, rank, world)
        mine = full[rank * geom.dl:(rank + 1) * geom.dl]
        ref = np.fft.fftn(full)
        nat = slab_fft3d(geom, ops, mine, -1, transposed_out=False, chunks=chunks)
        tr = slab_fft3d(geom, ops, mine, -1, transposed_out=True, chunks=chunks)
        e1 = np.abs(nat - ref[rank * geom.dl:(rank + 1) * geom.dl]).max()
        e2 = np.abs(tr - ref[:, rank * geom.hl:(rank + 1) * geom.hl, :]).max()
        inv = slab_fft3d(geom, ops, nat, +1, transposed_out=False, chunks=chunks) / full.size
        e3 = np.abs(inv - mine).max()
        # batch sharding of the batched-1D configs: contiguous split, no collective
        batch = 64
        rows = np.arange(batch)[rank * batch // world:(rank + 1) * batch // world]
        q.put((rank, float(e1), float(e2), float(e3), rows.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,chunks", [((8, 6, 5), 1), ((8, 4, 16), 2), ((4, 8, 3), 4)])
def test_slab_decomposition_world2_gloo(shape, chunks):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = []
    for rank, e1, e2, e3, rr in sorted(res):
        assert e1 < 1e-10 and e2 < 1e-10 and e3 < 1e-12, (rank, e1, e2, e3)
        rows += rr
    assert rows == list(range(64))          # the shards tile the batch exactly once, in order


def test_slab_geometry_validation():
    from accelerate_fft_b200.slab import SlabGeometry
    g = SlabGeometry(1024, 1024, 1024, 8)
    assert (g.dl, g.hl) == (128, 128)
    assert g.chunk_bounds(4) == [(0, 32), (32, 64), (64, 96), (96, 128)]
    assert g.chunk_bounds(3) == [(0, 64), (64, 128)]     # falls back to a divisor
    with pytest.raises(ValueError):
        SlabGeometry(10, 8, 8, 4)


def _emulate_exec_scatter(y, outs, offset, outer_stride, n_stride):
    """The index contract of b200fftExecScatter (include/b200fft.h) in numpy: `y` is the transformed [outer][n][inner]
    array; output index k of the axis goes to outs[k // (n/npeers)] (flat buffers) at
    offset + o*outer_stride + (k % (n/npeers))*n_stride + i."""
    outer, n, inner = y.shape
    nl = n // len(outs)
    for o in range(outer):
        for k in range(n):
            base = offset + o * outer_stride + (k % nl) * n_stride
            outs[k // nl][base:base + inner] = y[o, k]


@pytest.mark.parametrize("world", [2, 4])
def test_peer_scatter_index_arithmetic(world):
    """slab.scatter_targets + the ExecScatter contract reproduce fft3D: every rank's y pass scatters its ky rows
    into the owners' [D][hl][W] buffers, the z pass scatters back into z-slabs (PeerSlabFFT3D without the GPUs)."""
    from accelerate_fft_b200.slab import SlabGeometry, scatter_targets
    d, h, w = 8, 12, 5
    rng = np.random.default_rng(3)
    full = rng.uniform(-1, 1, (d, h, w)) + 1j * rng.uniform(-1, 1, (d, h, w))
    ref = np.fft.fftn(full)
    g = SlabGeometry(d, h, w, world)
    recv = [np.zeros(d * g.hl * w, dtype=np.complex128) for _ in range(world)]
    back = [np.zeros(g.dl * h * w, dtype=np.complex128) for _ in range(world)]
    for r in range(world):
        mine = full[r * g.dl:(r + 1) * g.dl]
        a = np.fft.fft(np.fft.fft(mine, axis=2), axis=1)                  # x then y on the slab: [dl][H][W]
        ty, _ = scatter_targets(g, r)
        _emulate_exec_scatter(a, recv, ty["offset"], ty["outer_stride"], ty["n_stride"])
    for r in range(world):
        b = recv[r].reshape(d, g.hl, w)
        assert np.allclose(np.fft.fft(b, axis=0), ref[:, r * g.hl:(r + 1) * g.hl], atol=1e-9)     # transposed-out result
        c = np.fft.fft(b, axis=0).reshape(1, d, g.hl * w)                 # z pass viewed as [1][D][hl*W]
        _, tz = scatter_targets(g, r)
        _emulate_exec_scatter(c, back, tz["offset"], tz["outer_stride"], tz["n_stride"])
    for r in range(world):
        assert np.allclose(back[r].reshape(g.dl, h, w), ref[r * g.dl:(r + 1) * g.dl], atol=1e-9)
