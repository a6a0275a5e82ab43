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
# the same transform behind the C ABI (b200fftPlanSlab3d / b200fftExecSlab, csrc/slab.cu): pipeline, peer mappings and
# barriers inside the library; default and chunked / SM-limited pipelines, both layouts, caller-owned outputs
from accelerate_fft_b200.slab import SlabPlan
for (d3, h3, w3) in ((64, 32, 48), (256, 1024, 64), (128, 256, 512)):
    full3 = (rng.uniform(-1, 1, (d3, h3, w3)) + 1j * rng.uniform(-1, 1, (d3, h3, w3))).astype(np.complex64)
    ref3 = np.fft.fftn(full3.astype(np.complex128))
    dl3, hl3 = d3 // world, h3 // world
    mine3 = torch.from_numpy(full3[rank * dl3:(rank + 1) * dl3]).cuda()
    sp = SlabPlan(d3, h3, w3, torch.complex64, None)
    for (cp, ck, yc) in ((None, None, None), (1, 1, 0), (2, 2, 24), (4, 4, 48), (1, 3, 7)):
        if cp is not None:     # (the first round runs the library's defaults)
            sp.tune(True, cp, ck, yc)
            sp.tune(False, cp, ck, yc)
        for rep in range(2):
            tr = sp(af.Forward, mine3, transposed_out=True).cpu().numpy().transpose(1, 0, 2)     # [hl][D][W] -> [D][hl][W]
            nat_d = sp(af.Forward, mine3) if rep == 0 else sp(af.Forward, mine3, out=sp.natural_buffer()).clone()
            nat = nat_d.cpu().numpy()
            e1 = np.linalg.norm(nat - ref3[rank * dl3:(rank + 1) * dl3]) / np.linalg.norm(ref3[rank * dl3:(rank + 1) * dl3])
            e2 = np.linalg.norm(tr - ref3[:, rank * hl3:(rank + 1) * hl3]) / np.linalg.norm(ref3[:, rank * hl3:(rank + 1) * hl3])
            back = sp(af.Inverse, nat_d).cpu().numpy()
            e3 = np.linalg.norm(back - full3[rank * dl3:(rank + 1) * dl3]) / np.linalg.norm(full3[rank * dl3:(rank + 1) * dl3])
            assert e1 < 1e-5 * 26 and e2 < 1e-5 * 26 and e3 < 2e-5 * 26, ("cabi", rank, (d3, h3, w3), (cp, ck, yc), rep, e1, e2, e3)
    sp.close()
zd = (rng.uniform(-1, 1, (64, 32, 16)) + 1j * rng.uniform(-1, 1, (64, 32, 16)))
spd = SlabPlan(64, 32, 16, torch.complex128, None)
got = spd(af.Forward, torch.from_numpy(zd[rank * 64 // world:(rank + 1) * 64 // world]).cuda()).cpu().numpy()
refd = np.fft.fftn(zd)[rank * 64 // world:(rank + 1) * 64 // world]
assert np.linalg.norm(got - refd) / np.linalg.norm(refd) < 1e-13 * 15
spd.close()
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
