"""Developer tool: phase timeline of the fused kernel (needs libtfft.so built with EXTRA=-DTFFT_TRACE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import numpy as np, torch, tfft
n, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 4096)
L = tfft.lib()
x = torch.randn(b * 2 * n, device="cuda").to(torch.float16); y = torch.empty_like(x)
plan = tfft.NativePlan(n, b)
for _ in range(3): plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
trace = torch.zeros(2048 * 4 * 32, dtype=torch.int64, device="cuda")
L.tfft_debug_set_trace.argtypes = [ctypes.c_void_p]
L.tfft_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
torch.cuda.synchronize()
L.tfft_debug_set_trace(None)
t = trace.cpu().numpy().reshape(2048, 4, 32)
base = t[0, 0, 0]
names = ["(issue)", "land wait", "stage0", "stage1", "stage2", "sync", "store", "sync2"]
for row in (0, 1, 2, 3, 294, 295):
    for u in range(4):
        r = t[row, u]
        if r[8] == 0: continue
        d = np.diff(r[:9])
        print(f"cta {row//2} slot {row%2} unit#{u} start+{int(r[0]-base):7d}: " + "  ".join(f"{nm}={int(v)}" for nm, v in zip(names, d)) + f"  total={int(r[8]-r[0])}")
for st in range(3):
    a = t[:296, 1:3, 9 + 2 * st]; bdone = t[:296, 1:3, 10 + 2 * st]; end = t[:296, 1:3, 3 + st]; beg = t[:296, 1:3, 2 + st]
    print(f"stage{st}: pre-sync {int((a - beg).mean())}  mma-first-half-wait {int((bdone - a).mean())}  epilogue(+2nd half wait) {int((end - bdone).mean())}")
for st in range(3):
    a = t[:296, 1:3, 9 + 2 * st]
    print(f"stage{st} mma thread: half0 ready +{int((t[:296,1:3,16+4*st]-a).mean())}  half1 ready +{int((t[:296,1:3,17+4*st]-a).mean())}  all issued +{int((t[:296,1:3,18+4*st]-a).mean())}  first-half done +{int((t[:296,1:3,10+2*st]-a).mean())}")
tt = t[:296, 1:3, :9]
print("mean (units 1-2):", {nm: int(v) for nm, v in zip(names, np.diff(tt, axis=2).mean(axis=(0, 1)))}, "total", int((tt[:, :, 8] - tt[:, :, 0]).mean()))
