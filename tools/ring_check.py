"""Developer tool: landing-ring kernel (tuner key ring=1, the default) against the single-unit kernel (ring=0) on the
shapes that use it -- bit-identical output (same stages, same arithmetic) and the time per exec of both."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch
import tfft

cases = [int(a) for a in sys.argv[1:]] or [15, 22, 23, 24]
reps = 10
tmp = tempfile.mkdtemp()
for lg in cases:
    n = 1 << lg
    b = max(1, (1 << 28) // n)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    x0 = torch.randn(2 * n * b, generator=g, device="cuda").to(torch.float16)
    outs, times = [], []
    for ring in (1, 0):
        f = os.path.join(tmp, f"t{lg}_{ring}.dat")
        with open(f, "w") as fh:
            fh.write(f"256 256 8 8 256\n{n} 256 8 8 256 ring={ring}\n")
        plan = tfft.NativePlan(n, b, 0, tuner_file=f)
        x = x0.clone(); y = torch.zeros_like(x)
        plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
        torch.cuda.synchronize()
        outs.append(y.clone())
        for _ in range(2):
            plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / reps)
        plan.close()
    same = bool(torch.equal(outs[0], outs[1]))
    ref = torch.fft.fft(torch.complex(x0.view(b, 2, n)[:1, 0].float(), x0.view(b, 2, n)[:1, 1].float())) / n
    got = torch.complex(outs[0].view(b, 2, n)[:1, 0].float(), outs[0].view(b, 2, n)[:1, 1].float())
    err = float(torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref))
    print(f"lg {lg} batch {b}: ring {times[0]:.4f} ms  single-unit {times[1]:.4f} ms  bit-identical {same}  rel-L2 vs fp32 fft {err:.2e}", flush=True)
