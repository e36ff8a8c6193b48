"""The reference's own example programs compile UNMODIFIED against the compat shim
(tensor-fft_b200/compat/base replaces src/base).  The examples #include "../base/ComputeFFT.h" relative
to their own directory, so a scratch tree of symlinks is built: <tmp>/src/testing -> the reference's
files (not copied), <tmp>/src/base -> the shim.  Needs /root/reference, i.e. runs in the build
container only; the binaries are also placed under oracle/_ref/ so that the GPU run can execute them."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
SHIM = os.path.join(ROOT, "tensor-fft_b200", "compat", "base")
OUT = os.path.join(ROOT, "oracle", "_ref")


def _tree(tmp_path):
    src = tmp_path / "src"
    (src / "base").mkdir(parents=True)
    for f in os.listdir(SHIM):
        os.symlink(os.path.join(SHIM, f), src / "base" / f)
    os.symlink(os.path.join(ROOT, "include", "tfft.h"), src / "base" / "tfft.h")
    shutil.copytree(os.path.join(REF, "testing"), src / "testing", symlinks=False,
                    copy_function=lambda s, d: os.symlink(s, d))
    return src


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("example", ["ExampleSingleFFT.cu", "ExampleBatchFFT.cu", "benchmarks/AccuracyTest.cu",
                                     "benchmarks/FFTBenchSinlge.cu"])
def test_reference_example_compiles_against_shim(tmp_path, example):
    """the four acceptance programs -- the two examples, the accuracy sweep
    (benchmarks/AccuracyTest.cu:17-86 through unitTesting/FFTTest.cu:24-88) and the single-transform benchmark that
    builds its plans from a tuner file (benchmarks/FFTBenchSinlge.cu:10-44 -> Bench.h:153-228)."""
    src = _tree(tmp_path)
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "shim_" + os.path.basename(example).replace(".cu", ""))
    cmd = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
           str(src / "testing" / example), "-L" + os.path.join(ROOT, "tensor-fft_b200", "tfft"), "-ltfft",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../tensor-fft_b200/tfft", "-lcufft"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(exe)
