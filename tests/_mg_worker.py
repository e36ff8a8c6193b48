"""Worker of tests/test_gpu_dist.py::test_two_gpus_*: launched with torch.distributed.run, one rank per GPU.
Runs ONE transform of length 2^lg both ways -- SixStepPlan (three NCCL all-to-alls) and MgPlan (peer stores over
NVLink, C ABI tfft_mg_*) -- and prints on rank 0 a JSON line with the worst rank's rel-L2 error of each against the
fp64 FFT of the fp16-quantised input (numpy pocketfft: the definition the oracle is pinned to, tests/test_oracle.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import numpy as np
import torch
import torch.distributed as dist
import tfft
from tfft import dist as tdist

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << lg
rng = np.random.default_rng(900 + lg)
re = rng.standard_normal(n).astype(np.float16)
im = rng.standard_normal(n).astype(np.float16)
want = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64)) / n
m = n // world
sl = slice(rank * m, (rank + 1) * m)
d_re, d_im = torch.from_numpy(re[sl]).cuda(), torch.from_numpy(im[sl]).cuda()


def rel(o_re, o_im):
    got = o_re.cpu().numpy().astype(np.float64) + 1j * o_im.cpu().numpy().astype(np.float64)
    return float(np.linalg.norm(got - want[sl]) / np.linalg.norm(want[sl]))


n1 = 1 << ((lg + 1) // 2)
six = tdist.SixStepPlan(n1, n // n1, rank, world, tdist.tfft_local_fft())
o_re, o_im = six.forward(d_re, d_im)
torch.cuda.synchronize()
e_six = rel(o_re, o_im)
mg = tdist.make_mg_plan(n)
mg.set_timeout_ms(20000)
out_re, out_im = torch.empty_like(d_re), torch.empty_like(d_im)
for _ in range(3):                      # repeated execs: the epoch counters and buffer reuse across execs
    mg.exec(d_re, d_im, out_re, out_im)
torch.cuda.synchronize()
mg.status()
e_mg = rel(out_re, out_im)
z_re, z_im = mg.result()
same = bool(torch.equal(z_re, out_re)) and bool(torch.equal(z_im, out_im))
t = torch.tensor([e_six, e_mg, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"lg": lg, "world": world, "rel_l2_sixstep_nccl": t[0].item(), "rel_l2_mg_peer": t[1].item(),
                      "zero_copy_view_matches": t[2].item() == 0.0}), flush=True)
dist.barrier()
mg.close()
dist.destroy_process_group()
