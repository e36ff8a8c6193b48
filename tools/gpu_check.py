"""Developer GPU smoke: product kernel vs fp64 oracle (and the real reference kernels) over sizes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import tfft, oracle as O

def run(n, batch, with_ref=True, stride_mult=2):
    re, im = O.gauss_fixture(n, batch, seed=1234 + n)
    stride = stride_mult * n
    buf = np.zeros((batch, stride), dtype=np.float16)
    buf[:, :n] = re
    if stride_mult == 2: buf[:, n:] = im
    d_in = torch.from_numpy(buf).cuda().reshape(-1)
    d_im = d_in[n:] if stride_mult == 2 else torch.from_numpy(np.ascontiguousarray(im)).cuda().reshape(-1)
    d_out = torch.zeros(batch * 2 * n, dtype=torch.float16, device="cuda")
    plan = tfft.NativePlan(n, batch)
    plan.exec(d_in, d_im, d_out, d_out[n:], stride if stride_mult == 2 else stride, 2 * n)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().reshape(batch, 2, n).astype(np.float64)
    nb = min(batch, 64)
    w_re, w_im = O.fft_f64(re[:nb].astype(np.float64), im[:nb].astype(np.float64))
    st = O.error_stats(out[:nb, 0], out[:nb, 1], w_re, w_im)
    res = {"n": n, "batch": batch, "rel_l2": st["rel_l2"], "max": st["max"], "info": plan.info}
    if with_ref and O.ref_lib() is not None and n * min(batch, 8) <= (1 << 22):
        rb = min(batch, 8)
        r_re, r_im = O.ref_fft_gpu(re[:rb], im[:rb])
        sr = O.error_stats(r_re.astype(np.float64), r_im.astype(np.float64), w_re[:rb], w_im[:rb])
        res["ref_rel_l2"] = sr["rel_l2"]
        sd = O.error_stats(out[:rb, 0], out[:rb, 1], r_re.astype(np.float64), r_im.astype(np.float64))
        res["vs_ref_rel_l2"] = sd["rel_l2"]
    return res

if __name__ == "__main__":
    cases = [(4096, 1), (16384, 8), (256, 64), (512, 5), (1024, 33), (2048, 16), (8192, 4), (32768, 2),
             (65536, 2), (1 << 20, 1)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    out = []
    for n, b in cases:
        try:
            r = run(n, b)
        except Exception as e:  # noqa
            r = {"n": n, "batch": b, "error": repr(e)}
        print(json.dumps({k: v for k, v in r.items() if k != "info"}), flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)
