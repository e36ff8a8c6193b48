"""Tuner: measures the knobs of the B200 kernels per transform length at the C3 working set
(2^28 elements per exec) and writes a tuner file in the reference's line format
`N mode base_warps r16_warps r2_block` (src/base/Plan.h:211-236, src/testing/FileWriter.h:250-269) followed by the
`key=value` knobs tfft_plan_create_from_file understands.  Every candidate is timed in its own process THROUGH a
one-line tuner file, so the file path itself is what is exercised.
Usage: python tools/tune.py [out_file=gpurun_out/TunerResults.dat] [lo=8] [hi=24]"""
import itertools, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "TunerResults.dat")
LO, HI = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 24)
CHILD = r'''
import sys, os, json
sys.path.insert(0, os.path.join(%r, "tensor-fft_b200"))
import torch, tfft
lg = int(sys.argv[1]); n = 1 << lg; b = (1 << 28) // n
x = torch.randn(2 * (1 << 28), device="cuda").to(torch.float16); y = torch.empty_like(x)
plan = tfft.NativePlan(n, b, tuner_file=sys.argv[2])
for _ in range(3): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"ms": e0.elapsed_time(e1) / 10}))
''' % ROOT


def line_for(n, knobs):
    return f"{n} 256 8 8 256 " + " ".join(f"{k}={v}" for k, v in sorted(knobs.items()))


def measure(lg, knobs):
    with tempfile.NamedTemporaryFile("w", suffix=".dat", delete=False) as f:
        f.write(line_for(1 << lg, knobs) + "\n")
        path = f.name
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, str(lg), path], capture_output=True, text=True, timeout=300)
        return json.loads(r.stdout.strip().splitlines()[-1])["ms"]
    except Exception:
        return float("inf")
    finally:
        os.unlink(path)


def candidates(lg):
    if lg <= 15:
        c = [{}, {"prefetch": 1}, {"prefetch": 0}]
        if lg >= 11:
            c += [{"tma": 0}]
        if lg in (13, 14):
            c += [{"two_slot": 0}, {"pipe": 0}]
        return c, []
    first = [{"lg1": v} for v in range(8, 13) if 8 <= lg - v <= 12]
    second = [dict(tma_col=a, prefetch=b) for a, b in itertools.product((0, 1), (0, 1))]
    return first, second


os.makedirs(os.path.dirname(OUT), exist_ok=True)
lines = []
for lg in range(LO, HI + 1):
    first, second = candidates(lg)
    res = [(measure(lg, k), k) for k in first]
    best_ms, best = min(res, key=lambda t: t[0])
    for extra in second:
        k = dict(best, **extra)
        ms = measure(lg, k)
        res.append((ms, k))
        if ms < best_ms:
            best_ms, best = ms, k
    print(json.dumps({"log2n": lg, "best_ms": round(best_ms, 4), "best": best,
                      "all": [(round(m, 4), k) for m, k in sorted(res, key=lambda t: t[0])]}), flush=True)
    lines.append(line_for(1 << lg, best))
with open(OUT, "w") as f:
    f.write("\n".join(lines) + "\n")
print("wrote", OUT)
