"""GPU: ONE 1-D transform sharded over several ranks (BASELINE config C4) through the C ABI tfft_mg_*.

* `test_ranks_on_one_gpu`: `world` ranks live in this process on ONE GPU, each with its own plan and buffers; the peers'
  buffers are plain device pointers.  The ranks of a phase run one after the other on ONE stream (tfft_mg_exec_phase,
  no flag barrier: kernels that spin on each other must never share a GPU), so everything else that runs on a
  multi-GPU box runs here too -- the tile transposes that store into the owner's buffer, the peer tables, the twiddled
  and plain local transforms, buffer reuse across execs.  world = 1 runs tfft_mg_exec itself, barrier kernel included.
* `test_two_gpus_*`: two processes, one per GPU (torch.distributed.run, NCCL for the plumbing and for the NCCL
  all-to-all version SixStepPlan); skipped on a single-GPU box.
Checked against the fp64 FFT of the fp16-quantised input; tolerance = 1.25 x the error level measured on B200 for the
two-pass sizes (profiles/r01_sweep_c3.json: 4.4e-4 at 2^20, 4.7e-4 at 2^24), below the reference's own level."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import tfft

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1.25 * 4.7e-4


@pytest.mark.parametrize("lg,world,staged", [(16, 1, None), (20, 1, None), (18, 2, None), (20, 2, None), (20, 4, None),
                                             (22, 8, None), (21, 4, None), (24, 2, None), (26, 8, None),
                                             (20, 2, "1"), (22, 4, "0"), (20, 1, "1")])
def test_ranks_on_one_gpu(lg, world, staged, monkeypatch):
    """staged: None = the plan's own choice of exchange layout (direct row-major peer stores for 1-2 ranks, source-rank-major
    staging + local unpack above), "0" / "1" = the other one forced (developer variable TFFT_MG_STAGED)."""
    if staged is not None:
        monkeypatch.setenv("TFFT_MG_STAGED", staged)
    n = 1 << lg
    m = n // world
    rng = np.random.default_rng(lg * 16 + world)
    re, im = rng.standard_normal(n).astype(np.float16), rng.standard_normal(n).astype(np.float16)
    want = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64)) / n
    plans = [tfft.MgPlan(n, r, world) for r in range(world)]
    tfft.MgPlan.connect_local(plans)
    ins = [(torch.from_numpy(re[r * m:(r + 1) * m]).cuda(), torch.from_numpy(im[r * m:(r + 1) * m]).cuda())
           for r in range(world)]
    outs = [(torch.empty(m, dtype=torch.float16, device="cuda"), torch.empty(m, dtype=torch.float16, device="cuda"))
            for _ in range(world)]
    for rep in range(2):                # the second exec reuses every buffer
        if world == 1:
            plans[0].set_timeout_ms(5000)
            plans[0].exec(ins[0][0], ins[0][1], outs[0][0], outs[0][1])
        else:
            for phase in range(4):
                for r, p in enumerate(plans):
                    p.exec_phase(phase, ins[r][0], ins[r][1], outs[r][0], outs[r][1])
    torch.cuda.synchronize()
    for p in plans:
        p.status()                      # raises on a barrier timeout
    got = np.concatenate([o[0].cpu().numpy().astype(np.float64) + 1j * o[1].cpu().numpy().astype(np.float64)
                          for o in outs])
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err <= (TOL if lg <= 24 else 1.25 * 6.5e-4), err
    # the inputs are never written
    for r in range(world):
        assert bool(torch.equal(ins[r][0].cpu(), torch.from_numpy(re[r * m:(r + 1) * m])))
    z_re, z_im = plans[-1].result()      # zero-copy view of the plan-owned result planes
    assert bool(torch.equal(z_re, outs[-1][0])) and bool(torch.equal(z_im, outs[-1][1]))
    info = plans[0].info
    assert info["exchanges"] == 3 and info["exchange_bytes_per_rank"] == 3 * (world - 1) * 4 * n // (world * world)
    for p in plans:
        p.close()


def test_mg_argument_checks():
    with pytest.raises(tfft.TfftError):
        tfft.MgPlan(1 << 14, 0, 2)          # too short to split into 64 x 64 tiles per rank
    with pytest.raises(tfft.TfftError):
        tfft.MgPlan(1 << 20, 2, 2)          # rank out of range
    with pytest.raises(tfft.TfftError):
        tfft.MgPlan(1 << 20, 0, 3)          # world not a power of two
    p = tfft.MgPlan(1 << 20, 0, 2)
    x = torch.zeros(1 << 19, dtype=torch.float16, device="cuda")
    with pytest.raises(tfft.TfftError):
        p.exec(x, x)                        # not connected yet
    p.close()


def _torchrun(nproc, script, *args, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29617", script, *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("lg", [20, 26])
def test_two_gpus_sixstep_nccl_and_peer_store(lg):
    out = _torchrun(2, os.path.join(ROOT, "tests", "_mg_worker.py"), lg)
    assert out["world"] == 2
    assert out["rel_l2_sixstep_nccl"] <= 1.25 * 6.5e-4, out   # three-pass level covers 2^26
    assert out["rel_l2_mg_peer"] <= 1.25 * 6.5e-4, out
    assert out["zero_copy_view_matches"]
