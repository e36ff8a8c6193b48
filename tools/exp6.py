"""Developer: C2 kernel time against the batch (quantisation of 4096 units over 148 SMs, fixed cost per launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch, tfft
n = 16384
for b in (2072, 3996, 4096, 4144, 8288, 16384, 16576):
    x = torch.randn(2 * n * b, device="cuda").to(torch.float16); y = torch.empty_like(x)
    plan = tfft.NativePlan(n, b)
    for _ in range(5): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 50
    print(f"batch {b:6d} units/SM {b/148:7.2f}  {us:8.2f} us  {us/b*4096:8.2f} us per 4096  frac {8.0*n*b/us/1e3/6525.2:.4f}")
    del x, y, plan
