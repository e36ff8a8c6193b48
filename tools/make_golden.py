"""Generates tests/golden/*.npz ON A B200 (run through gpurun): inputs (fp16) plus the outputs of the
REAL reference kernels (oracle/_ref/libtfft_ref.so = unmodified /root/reference/src/base behind
oracle/ref_driver.cu) for the default plan (Mode_256) and Mode_4096.  The committed fixtures pin the
CPU oracle and the CUDA path to the reference's own fp16 tensor-core output.
Usage (from the repo root):  gpurun -- 'python tools/make_golden.py'  then copy gpurun_out/golden/*.npz
to tests/golden/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as O

out_dir = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out_dir, exist_ok=True)
cases = []
# BASELINE configs[0]: N=4096, batch 1, the reference's sine fixture (AccuracyTest.cu:18-28)
re, im = O.sine_fixture(4096, cutoff=256, seed_re=42, seed_im=42 * 42)
cases.append(("sine4096", re.astype(np.float16)[None], im.astype(np.float16)[None]))
for n, b, seed in [(256, 4, 11), (512, 2, 12), (1024, 2, 13), (2048, 2, 14), (4096, 2, 15), (8192, 2, 16),
                   (16384, 2, 17), (32768, 1, 18), (65536, 1, 19)]:
    r, i = O.gauss_fixture(n, b, seed=seed)
    cases.append((f"gauss{n}", r, i))
for name, r16, i16 in cases:
    n = r16.shape[1]
    ref_re, ref_im = O.ref_fft_gpu(r16, i16, mode=0)
    d = {"in_re": r16, "in_im": i16, "ref256_re": ref_re, "ref256_im": ref_im}
    if n >= 4096:
        m_re, m_im = O.ref_fft_gpu(r16, i16, mode=1)
        d["ref4096_re"], d["ref4096_im"] = m_re, m_im
    bre, bim = O.ref_fft_gpu(r16, i16, mode=0, use_batch_api=True)
    assert np.array_equal(bre.view(np.uint16), ref_re.view(np.uint16)), "batch overload differs from single"
    w_re, w_im = O.fft_f64(r16.astype(np.float64), i16.astype(np.float64))
    st = O.error_stats(ref_re.astype(np.float64), ref_im.astype(np.float64), w_re, w_im)
    print(name, "reference rel-L2 vs fp64:", st)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
print("golden written to", out_dir)
