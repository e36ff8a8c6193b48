"""Developer tool: time the four-step split N = 2^lg1 * 2^lg2 for every legal lg1 (TFFT_FOURSTEP_LG1) at the C3 sizes.
One process per (lg, lg1) because the knob is read at plan creation.  Usage: python tools/tune_fourstep.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, os.path.join(%r, "tensor-fft_b200"))
import torch, tfft
lg = int(sys.argv[1]); n = 1 << lg; b = (1 << 28) // n
x = torch.randn(2 * (1 << 28), device="cuda").to(torch.float16); y = torch.empty_like(x)
plan = tfft.NativePlan(n, b)
for _ in range(3): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lg": lg, "lg1": os.environ.get("TFFT_FOURSTEP_LG1"), "ms": e0.elapsed_time(e1) / 10}))
''' % ROOT
for lg in range(16, 25):
    for lg1 in range(8, 13):
        if not (8 <= lg - lg1 <= 12):
            continue
        env = dict(os.environ, TFFT_DEVELOPER="1", TFFT_FOURSTEP_LG1=str(lg1))
        r = subprocess.run([sys.executable, "-c", CHILD, str(lg)], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
