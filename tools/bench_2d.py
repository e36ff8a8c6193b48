"""BASELINE config C5: batched 2-D transforms, 8192 x 8192 planar fp16 images, 2 images per GPU
(16 images over 8 GPUs, batch-sharded, no collective).  CUDA-event time of exec with the images resident in HBM,
HBM GB/s = 8*ny*nx*batch*2 passes / t against the measured roofline, GFLOP/s = 5*ny*nx*log2(ny*nx)*batch / t,
per-pass times, rel-L2 vs an fp32 torch fft2, and cuFFT fp16 (torch.fft.fft2 on complex32) timed in the same run.
Usage: python tools/bench_2d.py [ny nx batch]   -> one JSON line (also gpurun_out/bench_2d.json).
Under torchrun (--nproc-per-node G) every rank transforms its own `batch` images (weak scaling, no collective); the time
is the max over ranks and the throughput figures are whole-job aggregates."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import numpy as np, torch
import tfft

ny, nx, b = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (8192, 8192, 2)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = ny * nx
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

def timed(fn, warm=3, iters=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
x = torch.randn(b * 2 * n, generator=g, device="cuda").to(torch.float16)
y = torch.empty_like(x)
plan = tfft.NativePlan(n, b, 0, shape2d=(ny, nx))
if world > 1:
    dist.barrier()
ms = timed(lambda: plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n))
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
lg = int(np.log2(n))
row = {"workload": f"C5: {b} x 2-D {ny}x{nx} fp16 planar per GPU", "n_gpus": world, "ms": round(ms, 4),
       "passes": plan.info["passes"], "gflops": round(5.0 * n * lg * b * world / (ms * 1e-3) / 1e9, 1),
       "hbm_gbs": round(8.0 * n * b * 2 * world / (ms * 1e-3) / 1e9, 1), "env_ybits": os.environ.get("TFFT_2D_YBITS")}
row["roofline_frac"] = round(row["hbm_gbs"] / PEAK / world, 4)
xs = torch.complex(x[:n].float(), x[n:2 * n].float()).view(ny, nx)
want = torch.fft.fft2(xs) / n
got = torch.complex(y[:n].float(), y[n:2 * n].float()).view(ny, nx)
row["rel_l2_vs_fp32_fft2"] = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
del xs, got
try:
    xc = torch.view_as_complex(torch.randn(b, ny, nx, 2, device="cuda", dtype=torch.float16).contiguous())
    row["cufft_fp16_ms"] = round(timed(lambda: torch.fft.fft2(xc), warm=2, iters=5), 4)
    xc1 = torch.complex(x[:n].view(ny, nx), x[n:2 * n].view(ny, nx))
    yc = torch.fft.fft2(xc1).to(torch.complex64) / n
    row["cufft_fp16_rel_l2"] = float(torch.linalg.vector_norm(yc - want) / torch.linalg.vector_norm(want))
except Exception as e:  # noqa
    row["cufft_fp16_error"] = repr(e)[:160]
if world > 1:
    dist.destroy_process_group()
if rank != 0:
    sys.exit(0)
print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_2d.json"), "a") as f:
    f.write(json.dumps(row) + "\n")
