"""BASELINE configs[2]: size sweep N = 2^8 .. 2^24, batch = 2^28 / N (1 GiB in, 1 GiB out).
Per size: CUDA-event time of exec (data resident), GFLOP/s (5 N log2 N), HBM GB/s = 8*N*batch*passes / t,
fraction of the measured HBM roofline, rel-L2 vs the fp64 oracle on a subset, and the same for two
comparison points in the same run: cuFFT fp16 (torch.fft on complex32) and the unmodified reference
kernels (oracle/_ref, single-transform path on a capped batch).  Writes gpurun_out/sweep.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import tfft, oracle as O

TOTAL = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
FAST = len(sys.argv) > 2 and sys.argv[2] == "fast"      # developer mode: our timings and accuracy only
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

def timed(fn, warm=5, iters=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

rows = []
g = torch.Generator(device="cuda"); g.manual_seed(1234)
x = torch.randn(2 * TOTAL, generator=g, device="cuda").to(torch.float16)
y = torch.empty_like(x)
for lg in range(8, 25):
    n = 1 << lg; b = TOTAL // n
    plan = tfft.NativePlan(n, b)
    passes = plan.info["passes"]
    if passes > 1:
        x.copy_(torch.randn(2 * TOTAL, generator=g, device="cuda").to(torch.float16))   # multi-pass sizes consume the input
    ref_in = x.view(b, 2, n)[:8].clone()
    ms = timed(lambda: plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n))
    # accuracy on the first 8 transforms (fresh input copy)
    xin = ref_in.clone().view(-1); yo = torch.empty_like(xin)
    p8 = tfft.NativePlan(n, 8); p8.exec(xin, xin[n:], yo, yo[n:], 2 * n, 2 * n); torch.cuda.synchronize()
    got = yo.view(8, 2, n).cpu().numpy().astype(np.float64)
    src = ref_in.cpu().numpy().astype(np.float64)
    nb = 8 if lg <= 20 else 1
    w_re, w_im = O.fft_f64(src[:nb, 0], src[:nb, 1])
    err = O.error_stats(got[:nb, 0], got[:nb, 1], w_re, w_im)["rel_l2"]
    row = {"log2n": lg, "batch": b, "passes": passes, "ms": round(ms, 4),
           "gflops": round(5.0 * n * lg * b / (ms * 1e-3) / 1e9, 1),
           "hbm_gbs": round(8.0 * n * b * passes / (ms * 1e-3) / 1e9, 1),
           "hbm_gbs_p1": round(8.0 * n * b / (ms * 1e-3) / 1e9, 1), "rel_l2": err}
    row["roofline_frac"] = round(row["hbm_gbs"] / PEAK, 4)
    # cuFFT fp16 (interleaved complex32, unscaled) through torch.fft
    try:
        if FAST: raise RuntimeError("skipped (fast mode)")
        xc = torch.complex(ref_in[:, 0], ref_in[:, 1])   # complex32
        big = torch.view_as_complex(torch.randn(b, n, 2, device="cuda", dtype=torch.float16).contiguous()) if lg < 25 else None
        ms_c = timed(lambda: torch.fft.fft(big, dim=1), warm=3, iters=10)
        yc = (torch.fft.fft(xc, dim=1).to(torch.complex64) / n).cpu().numpy()
        errc = O.error_stats(yc[:nb].real, yc[:nb].imag, w_re, w_im)["rel_l2"]
        row.update({"cufft_fp16_ms": round(ms_c, 4), "cufft_fp16_rel_l2": errc})
        del big
    except Exception as e:  # noqa
        row["cufft_fp16_error"] = repr(e)[:120]
    # the reference's own kernels (single-transform path, capped batch), kernel-only time
    try:
        if O.ref_lib() is not None and not FAST:
            cap = int(min(b, max(1, (1 << 22) // n), 64))
            k_ms, _ = O.ref_bench_gpu(n, cap, 5, 2, mode=0, use_batch_api=False)
            r_re, r_im = O.ref_fft_gpu(ref_in[:nb, 0].cpu().numpy(), ref_in[:nb, 1].cpu().numpy())
            row.update({"ref_ms_per_transform": round(float(np.mean(k_ms)) / cap, 5), "ref_batch_cap": cap,
                        "ref_rel_l2": O.error_stats(r_re.astype(np.float64), r_im.astype(np.float64), w_re, w_im)["rel_l2"],
                        "ours_ms_per_transform": round(ms / b, 6)})
    except Exception as e:  # noqa
        row["ref_error"] = repr(e)[:120]
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if not FAST: json.dump({"total_elements": TOTAL, "hbm_peak_gbs": PEAK, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
