"""Developer tool: run ONE shape a few times (short command line for ncu, see /opt/skills/guides/B200_PROFILING.md).
    python tools/prof_case.py c2 | c1 | n15 | n24 | n22 | c5 | n10   [reps]
Prints the CUDA-event time per exec."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch
import tfft

case = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shape2d = None
if case == "c2":
    n, b = 16384, 4096
elif case == "c1":
    n, b = 4096, 1
elif case == "c5":
    n, b, shape2d = 8192 * 8192, 2, (8192, 8192)
else:
    lg = int(case[1:])
    n = 1 << lg
    b = max(1, (1 << 28) // n)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn(2 * n * b, generator=g, device="cuda").to(torch.float16)
y = torch.empty_like(x)
plan = tfft.NativePlan(n, b, 0, shape2d=shape2d)
for _ in range(3):
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
e1.record(); torch.cuda.synchronize()
print(case, n, b, "passes", plan.info["passes"], "ms/exec", e0.elapsed_time(e1) / reps)
