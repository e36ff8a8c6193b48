"""Developer tool: batched three-pass plans (one launch per pass for the whole batch) -- every transform of the batch
against an fp32 FFT, and the time per exec.  usage: three_check.py lg batch"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch
import tfft
lg, b = int(sys.argv[1]), int(sys.argv[2])
n = 1 << lg
g = torch.Generator(device="cuda"); g.manual_seed(lg)
x0 = torch.randn(2 * n * b, generator=g, device="cuda").to(torch.float16)
plan = tfft.NativePlan(n, b, 0)
x = x0.clone(); y = torch.zeros_like(x)
plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
worst = 0.0
for i in range(b):
    xi = x0.view(b, 2, n)[i]
    ref = torch.fft.fft(torch.complex(xi[0].float(), xi[1].float())) / n
    yi = y.view(b, 2, n)[i]
    got = torch.complex(yi[0].float(), yi[1].float())
    worst = max(worst, float(torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref)))
for _ in range(2):
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
e1.record(); torch.cuda.synchronize()
print(f"lg {lg} batch {b} passes {plan.info['passes']}: {e0.elapsed_time(e1) / reps:.4f} ms/exec, worst rel-L2 over the batch {worst:.2e}", flush=True)
