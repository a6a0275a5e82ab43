"""Multi-GPU parity (needs >= 2 GPUs on the box; `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import accelerate_fft_b200 as af
from accelerate_fft_b200.slab import SlabFFT3D
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
d, h, w = 64, 32, 48
rng = np.random.default_rng(5)
full = (rng.uniform(-1, 1, (d, h, w)) + 1j * rng.uniform(-1, 1, (d, h, w))).astype(np.complex64)
ref = np.fft.fftn(full.astype(np.complex128))
dl, hl = d // world, h // world
mine = torch.from_numpy(full[rank * dl:(rank + 1) * dl]).cuda()
for chunks in (1, 4):
    f = SlabFFT3D(d, h, w, torch.complex64, None, chunks=chunks)
    nat = f(af.Forward, mine).cpu().numpy()
    tr = f(af.Forward, mine, transposed_out=True).cpu().numpy()
    e1 = np.linalg.norm(nat - ref[rank * dl:(rank + 1) * dl]) / np.linalg.norm(ref[rank * dl:(rank + 1) * dl])
    e2 = np.linalg.norm(tr - ref[:, rank * hl:(rank + 1) * hl]) / np.linalg.norm(ref[:, rank * hl:(rank + 1) * hl])
    back = f(af.Inverse, torch.from_numpy(nat).cuda()).cpu().numpy()
    e3 = np.linalg.norm(back - full[rank * dl:(rank + 1) * dl]) / np.linalg.norm(full[rank * dl:(rank + 1) * dl])
    assert e1 < 1e-5 * 17 and e2 < 1e-5 * 17 and e3 < 2e-5 * 17, (rank, chunks, e1, e2, e3)
# the exchange folded into the kernels' stores over peer memory (b200fftExecScatter + CUDA IPC): no NCCL data path
from accelerate_fft_b200.slab import PeerSlabFFT3D
for (d2, h2, w2) in ((64, 32, 48), (256, 1024, 64)):
    full2 = (rng.uniform(-1, 1, (d2, h2, w2)) + 1j * rng.uniform(-1, 1, (d2, h2, w2))).astype(np.complex64)
    ref2 = np.fft.fftn(full2.astype(np.complex128))
    dl2, hl2 = d2 // world, h2 // world
    mine2 = torch.from_numpy(full2[rank * dl2:(rank + 1) * dl2]).cuda()
    pf = PeerSlabFFT3D(d2, h2, w2, torch.complex64, None, chunks=(1 if d2 == 64 else 4))
    for rep in range(2):
        tr = pf(af.Forward, mine2, transposed_out=True).cpu().numpy()
        nat = pf(af.Forward, mine2).cpu().numpy()
        e1 = np.linalg.norm(nat - ref2[rank * dl2:(rank + 1) * dl2]) / np.linalg.norm(ref2[rank * dl2:(rank + 1) * dl2])
        e2 = np.linalg.norm(tr - ref2[:, rank * hl2:(rank + 1) * hl2]) / np.linalg.norm(ref2[:, rank * hl2:(rank + 1) * hl2])
        back = pf(af.Inverse, torch.from_numpy(nat).cuda()).cpu().numpy()
        e3 = np.linalg.norm(back - full2[rank * dl2:(rank + 1) * dl2]) / np.linalg.norm(full2[rank * dl2:(rank + 1) * dl2])
        assert e1 < 1e-5 * 24 and e2 < 1e-5 * 24 and e3 < 2e-5 * 24, ("peer", rank, (d2, h2, w2), rep, e1, e2, e3)
    pf.close()
# batched 1D sharded by rows: every rank transforms its contiguous share, results tile the full answer
x = (rng.uniform(-1, 1, (64, 4096)) + 1j * rng.uniform(-1, 1, (64, 4096)))
lo, hi = rank * 64 // world, (rank + 1) * 64 // world
y = af.fft(af.Forward, torch.from_numpy(x[lo:hi]).cuda()).cpu().numpy()
assert np.linalg.norm(y - np.fft.fft(x[lo:hi])) / np.linalg.norm(y) < 1e-13 * 12
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_slab_fft3d_and_batch_sharding_2gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
