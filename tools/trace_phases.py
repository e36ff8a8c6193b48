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
names = ["load issue", "load wait", "stage0", "stage1", "stage2", "sync", "store", "sync2"]
for cta in (0, 1, 147, 148, 295):
    for u in range(4):
        r = t[cta, u]
        if r[0] == 0: continue
        d = np.diff(r[:9])
        print(f"cta {cta:3d} unit#{u} start+{r[0]-t[cta,0,0]:7d}: " + "  ".join(f"{nm}={int(v)}" for nm, v in zip(names, d)))
for st in range(3):
    a = t[:296, 1:3, 9 + 2 * st]; bdone = t[:296, 1:3, 10 + 2 * st]; end = t[:296, 1:3, 3 + st]; beg = t[:296, 1:3, 2 + st]
    print(f"stage{st}: pre-sync {int((a - beg).mean())}  mma(issue+wait+sync) {int((bdone - a).mean())}  epilogue {int((end - bdone).mean())}")
ep = t[:296, 1:3, :]
print("stage0 epilogue detail (warp 0): mma_done->wait0 %d | " % int((ep[:, :, 16] - ep[:, :, 10]).mean()) + " | ".join(
    "item%d proc %d, wait-next %d" % (i, int((ep[:, :, 17 + 2 * i] - ep[:, :, 16 + 2 * i]).mean()),
                                      int((ep[:, :, 18 + 2 * i] - ep[:, :, 17 + 2 * i]).mean()) if i < 3 else 0) for i in range(4)))
tt = t[:, 1:3, :9]; tt = tt[tt[:, :, 0] > 0]
print("mean over CTAs (units 1-2):", {nm: int(v) for nm, v in zip(names, np.diff(tt, axis=1).mean(axis=0))}, "total", int((tt[:, 8] - tt[:, 0]).mean()))
