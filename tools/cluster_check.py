"""Developer check of the CTA-pair (cluster) units: accuracy vs a complex64 cuFFT and time per exec, for the sizes that use
them (N = 65536 single pass; four-step sizes whose column pass is 4096 long; 2-D with 4096-point columns)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch
import tfft

def timed(fn, warm=3, iters=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

cases = [("1d", 16, 3), ("1d", 16, 4096), ("1d", 21, 128), ("1d", 22, 64), ("1d", 24, 16), ("2d", (8192, 8192), 2), ("2d", (4096, 1024), 8)]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[1]) in sys.argv[1:]]
for kind, shp, b in cases:
    if kind == "1d":
        n = 1 << shp
        plan = tfft.NativePlan(n, b)
    else:
        n = shp[0] * shp[1]
        plan = tfft.NativePlan(n, b, 0, shape2d=shp)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x = torch.randn(2 * n * b, generator=g, device="cuda").to(torch.float16)
    keep = x.clone()
    y = torch.zeros_like(x)
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    xv, yv = keep.view(b, 2, -1), y.view(b, 2, -1)
    nb = min(b, 4)
    xs = torch.complex(xv[:nb, 0].float(), xv[:nb, 1].float())
    if kind == "1d":
        want = torch.fft.fft(xs, dim=1) / n
    else:
        want = (torch.fft.fft2(xs.view(nb, *shp)) / n).reshape(nb, -1)
    got = torch.complex(yv[:nb, 0].float(), yv[:nb, 1].float())
    rel = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
    last = torch.complex(yv[b - 1, 0].float(), yv[b - 1, 1].float())
    xl = torch.complex(xv[b - 1, 0].float(), xv[b - 1, 1].float())
    wl = (torch.fft.fft(xl) / n) if kind == "1d" else (torch.fft.fft2(xl.view(*shp)) / n).reshape(-1)
    rel_last = float(torch.linalg.vector_norm(last - wl) / torch.linalg.vector_norm(wl))
    x.copy_(keep)
    ms = timed(lambda: plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n))
    print(json.dumps({"case": [kind, shp, b], "passes": plan.info["passes"], "rel_l2": rel, "rel_l2_last": rel_last,
                      "ms": round(ms, 4), "hbm_gbs": round(8.0 * n * b * plan.info["passes"] / ms / 1e6, 1),
                      "no_cluster": bool(os.environ.get("TFFT_NO_CLUSTER"))}), flush=True)
