"""BASELINE configs[3]: ONE 1-D fp16 C2C FFT of length N = 2^28 (default) sharded over the GPUs of a box:
distributed six-step, three NCCL all-to-all exchanges over NVLink (tfft.dist.SixStepPlan).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
             --master-port P tools/bench_sixstep.py [--log2n 28] [--steps 10]
Prints one JSON line on rank 0 (device-timed, max over ranks) and checks every rank's slice of the
spectrum against a complex64 torch FFT of the same (replicated, seeded) input."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch, torch.distributed as dist
import tfft
from tfft import dist as tdist

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lg = args.log2n
n1 = 1 << (lg // 2); n2 = 1 << (lg - lg // 2); n = n1 * n2
g = torch.Generator(device="cuda"); g.manual_seed(2024)
x_re = torch.randn(n, generator=g, device="cuda").to(torch.float16)
x_im = torch.randn(n, generator=g, device="cuda").to(torch.float16)
lo, hi = tdist.shard_range(n1, rank, world)
s_re, s_im = x_re.view(n1, n2)[lo:hi].contiguous(), x_im.view(n1, n2)[lo:hi].contiguous()
plan = tdist.SixStepPlan(n1, n2, rank, world, tdist.tfft_local_fft())
o_re, o_im = plan.forward(s_re, s_im)
torch.cuda.synchronize()
want = torch.fft.fft(torch.complex(x_re.float(), x_im.float())) / n
sl = slice(rank * n // world, (rank + 1) * n // world)
got = torch.complex(o_re.float(), o_im.float())
err = float(torch.linalg.vector_norm(got - want[sl]) / torch.linalg.vector_norm(want[sl]))
del want
for _ in range(args.warmup):
    plan.forward(s_re, s_im)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    plan.forward(s_re, s_im)
e1.record()
dist.barrier(); torch.cuda.synchronize()
ms = tdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, torch.device("cuda"))
worst = tdist.max_over_ranks(err, torch.device("cuda"))
if rank == 0:
    nv = plan.nvlink_bytes_per_rank()
    print(json.dumps({"metric": "single 1-D fp16 C2C FFT, distributed six-step", "log2n": lg, "n_gpus": world,
                      "ms_per_transform": round(ms, 4), "gflops": round(5.0 * n * lg / (ms * 1e-3) / 1e9, 1),
                      "rel_l2_vs_complex64_fft_worst_rank": worst, "all_to_alls": plan.all_to_alls,
                      "nvlink_bytes_sent_per_rank": nv,
                      "nvlink_gbs_per_rank": round(nv / (ms * 1e-3) / 1e9, 1),
                      "hbm_algorithmic_bytes_per_rank": 2 * 8 * n // world}), flush=True)
dist.destroy_process_group()
