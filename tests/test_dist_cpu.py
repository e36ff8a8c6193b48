"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch sharding, max-over-ranks timing reduce,
the three all-to-all exchanges of the distributed six-step) with a numpy stand-in for the per-rank
transform.  The CUDA kernels are not involved here; the GPU run covers them."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))


def _numpy_local_fft(z, n, batch, log2_total, first_col):
    x = z[0].double().numpy() + 1j * z[1].double().numpy()
    y = np.fft.fft(x.reshape(batch, n), axis=1) / n
    if log2_total:
        k = np.arange(n)[None, :]
        col = (first_col + np.arange(batch))[:, None]
        y = y * np.exp(-2j * np.pi * ((k * col) % (1 << log2_total)) / (1 << log2_total))
    return torch.stack([torch.from_numpy(y.real.copy()), torch.from_numpy(y.imag.copy())])


def _worker(rank, world, port, n1, n2, q):
    from tfft import dist as tdist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = n1 * n2
        rng = np.random.default_rng(7)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        lo, hi = tdist.shard_range(n1, rank, world)
        slab = x.reshape(n1, n2)[lo:hi]
        plan = tdist.SixStepPlan(n1, n2, rank, world, _numpy_local_fft)
        o_re, o_im = plan.forward(torch.from_numpy(slab.real.copy()), torch.from_numpy(slab.imag.copy()))
        want = (np.fft.fft(x) / n)[rank * n // world:(rank + 1) * n // world]
        err = np.abs(o_re.numpy() + 1j * o_im.numpy() - want).max()
        slowest = tdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
        q.put((rank, float(err), slowest, plan.nvlink_bytes_per_rank()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n1,n2", [(16, 32), (64, 64)])
def test_six_step_exchange_world2(n1, n2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n1) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n1, n2, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, err, slowest, nbytes in res:
        assert err < 1e-12
        assert slowest == 2.0                                  # max over ranks
        assert nbytes == 3 * 1 * 4 * n1 * n2 // 4              # A * (G-1) * 4N / G^2


def test_shard_range_tiles_the_batch():
    from tfft import dist as tdist
    for total, world in [(4096, 8), (4096, 3), (5, 4), (1, 2)]:
        ranges = [tdist.shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1
