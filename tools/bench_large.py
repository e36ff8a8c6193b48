"""Single large 1-D transforms N = 2^25 .. 2^29 on one GPU (the sizes the reference's FFTBenchSinlge.cu goes up to):
three HBM passes (N = N1*Na*Nb).  CUDA-event time, HBM GB/s = 8*N*3 / t, rel-L2 of a 2^20-point slice of the
spectrum against a complex64 torch FFT.  Usage: python tools/bench_large.py [lo=25] [hi=29]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch, tfft
lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (25, 29)
for lg in range(lo, hi + 1):
    n = 1 << lg
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    x = torch.randn(2 * n, generator=g, device="cuda").to(torch.float16)
    keep = x.clone(); y = torch.empty_like(x)
    plan = tfft.NativePlan(n, 1)
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n); torch.cuda.synchronize()
    want = torch.fft.fft(torch.complex(keep[:n].float(), keep[n:].float()))[: 1 << 20] / n
    got = torch.complex(y[: 1 << 20].float(), y[n:n + (1 << 20)].float())
    rel = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
    del want, got
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(5):
        x.copy_(keep)                      # multi-pass sizes use the input planes as scratch
        e0.record(); plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[len(ms) // 2]
    print(json.dumps({"log2n": lg, "passes": plan.info["passes"], "ms": round(t, 4), "gflops": round(5.0 * n * lg / (t * 1e-3) / 1e9, 1),
                      "hbm_gbs": round(8.0 * n * plan.info["passes"] / (t * 1e-3) / 1e9, 1), "rel_l2_slice": rel,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("TFFT_")}}), flush=True)
    del x, y, keep
