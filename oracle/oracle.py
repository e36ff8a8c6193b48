"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the CPU oracle (oracle/tfft_oracle.cpp)
and of the unmodified-reference driver (oracle/ref_driver.cu -> oracle/_ref/libtfft_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (tensor-fft_b200/tfft) never does.

Reference citations (relative to /root/reference): transform definition and 1/N scale
src/base/ComputeFFT.h:1-16, src/testing/AccuracyCalculator.h:70-84; fixture
src/testing/TestingDataCreation.h:15-27,89-117; statistics
src/testing/AccuracyCalculator.h:86-148; reference seeds/cutoff
src/testing/benchmarks/AccuracyTest.cu:18-28.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_u16p = ctypes.POINTER(ctypes.c_uint16)


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists) with oracle/Makefile."""
    so = os.path.join(_HERE, "_build", "libtfft_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
            os.path.join(_HERE, "tfft_oracle.cpp")):
        subprocess.run(["make", "-C", _HERE, "_build/libtfft_oracle.so"], check=True, capture_output=True)
    ref_so = os.path.join(_HERE, "_ref", "libtfft_ref.so")
    if os.path.isdir("/root/reference/src/base") and (force or not os.path.exists(ref_so) or os.path.getmtime(
            ref_so) < os.path.getmtime(os.path.join(_HERE, "ref_driver.cu"))):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        _LIB = ctypes.CDLL(os.path.join(_HERE, "_build", "libtfft_oracle.so"))
        L = _LIB
        L.oracle_random_weights.argtypes = [ctypes.c_int, ctypes.c_int, _fp]
        L.oracle_sine_fixture.argtypes = [ctypes.c_int64, ctypes.c_int, _fp, _fp, _dp, _dp]
        for name in ("oracle_dft_f64", "oracle_fft_f64"):
            f = getattr(L, name)
            f.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
            f.restype = ctypes.c_int
        L.oracle_fft2_f64.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.oracle_fft2_f64.restype = ctypes.c_int
        L.oracle_ref_algorithm.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_int64, ctypes.c_int]
        L.oracle_ref_algorithm.restype = ctypes.c_int
        L.oracle_ref_input_index.argtypes = [ctypes.c_int64, ctypes.c_int64]
        L.oracle_ref_input_index.restype = ctypes.c_int64
        L.oracle_error_stats.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_int64, _dp]
        L.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _d(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


# ------------------------------------------------------------------ fixtures
def random_weights(count: int, seed: int) -> np.ndarray:
    """GetRandomWeights (TestingDataCreation.h:15-27): libstdc++ default_random_engine."""
    out = np.empty(count, dtype=np.float32)
    lib().oracle_random_weights(count, seed, out.ctypes.data_as(_fp))
    return out


def sine_fixture(n: int, cutoff: int = 256, seed_re: int = 42, seed_im: int = 42 * 42):
    """Reference accuracy fixture (AccuracyTest.cu:18-28): returns (re, im) float64, un-quantised."""
    w_re, w_im = random_weights(cutoff, seed_re), random_weights(cutoff, seed_im)
    re, im = np.empty(n), np.empty(n)
    lib().oracle_sine_fixture(n, cutoff, w_re.ctypes.data_as(_fp), w_im.ctypes.data_as(_fp), _d(re), _d(im))
    return re, im


def gauss_fixture(n: int, batch: int, seed: int = 1234):
    """iid N(0,1) re/im (SURVEY 8d C2/C3 fixture), already rounded to fp16. Shapes (batch, n)."""
    rng = np.random.default_rng(seed)
    re = rng.standard_normal((batch, n)).astype(np.float16)
    im = rng.standard_normal((batch, n)).astype(np.float16)
    return re, im


# ------------------------------------------------------------------ transforms
def _run(fn, re, im, n, batch, nthreads):
    re = np.ascontiguousarray(re, dtype=np.float64).reshape(batch, n)
    im = np.ascontiguousarray(im, dtype=np.float64).reshape(batch, n)
    o_re, o_im = np.empty_like(re), np.empty_like(im)
    rc = fn(_d(re), _d(im), _d(o_re), _d(o_im), n, batch, nthreads)
    if rc != 0:
        raise ValueError(f"oracle rejected n={n} batch={batch}")
    return o_re, o_im


def dft_f64(re, im, nthreads: int = 0):
    """Naive O(N^2) fp64 DFT / N on the last axis. Inputs (batch, n) or (n,)."""
    re = np.atleast_2d(re)
    im = np.atleast_2d(im)
    return _run(lib().oracle_dft_f64, re, im, re.shape[-1], re.shape[0], nthreads)


def fft_f64(re, im, nthreads: int = 0):
    """fp64 FFT / N on the last axis (power-of-two length)."""
    re = np.atleast_2d(re)
    im = np.atleast_2d(im)
    return _run(lib().oracle_fft_f64, re, im, re.shape[-1], re.shape[0], nthreads)


def fft2_f64(re, im, nthreads: int = 0):
    """2-D fp64 FFT / (ny*nx) on the last two axes. Inputs (batch, ny, nx)."""
    re = np.ascontiguousarray(re, dtype=np.float64)
    im = np.ascontiguousarray(im, dtype=np.float64)
    b, ny, nx = re.shape
    o_re, o_im = np.empty_like(re), np.empty_like(im)
    rc = lib().oracle_fft2_f64(_d(re), _d(im), _d(o_re), _d(o_im), ny, nx, b, nthreads)
    if rc != 0:
        raise ValueError("oracle rejected 2-D shape")
    return o_re, o_im


def ref_algorithm(re, im, emulate_fp16: bool):
    """The reference's staged algorithm (Mode_256 plan) for ONE transform, on the CPU."""
    re = np.ascontiguousarray(re, dtype=np.float64).ravel()
    im = np.ascontiguousarray(im, dtype=np.float64).ravel()
    o_re, o_im = np.empty_like(re), np.empty_like(im)
    rc = lib().oracle_ref_algorithm(_d(re), _d(im), _d(o_re), _d(o_im), re.size, 1 if emulate_fp16 else 0)
    if rc != 0:
        raise ValueError("reference algorithm needs a power of two >= 256")
    return o_re, o_im


def ref_input_index(o: int, n: int) -> int:
    return int(lib().oracle_ref_input_index(o, n))


# ------------------------------------------------------------------ metrics
def error_stats(a_re, a_im, b_re, b_im) -> dict:
    """rel-L2 of a against b plus the reference's (max, avg, sigma) triple."""
    arrs = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (a_re, a_im, b_re, b_im)]
    st = np.empty(4)
    lib().oracle_error_stats(*[_d(x) for x in arrs], arrs[0].size, _d(st))
    return {"rel_l2": float(st[0]), "max": float(st[1]), "avg": float(st[2]), "sigma": float(st[3])}


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ------------------------------------------------------------------ the real reference (GPU)
def ref_lib():
    """oracle/_ref/libtfft_ref.so: the unmodified reference behind oracle/ref_driver.cu."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libtfft_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        _REF = ctypes.CDLL(path)
        _REF.ref_fft.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u16p, _u16p]
        _REF.ref_fft.restype = ctypes.c_int
        _REF.ref_bench.argtypes = [ctypes.c_int] * 6 + [_u16p, _u16p, _dp, _dp]
        _REF.ref_bench.restype = ctypes.c_int
        _REF.ref_plan_info.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        _REF.ref_plan_info.restype = ctypes.c_int
    return _REF


def ref_fft_gpu(re16: np.ndarray, im16: np.ndarray, mode: int = 0, use_batch_api: bool = False):
    """Run the real reference kernels. re16/im16: float16 (batch, n). Returns float16 (batch, n) x2."""
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libtfft_ref.so missing")
    re16 = np.atleast_2d(np.asarray(re16, dtype=np.float16))
    im16 = np.atleast_2d(np.asarray(im16, dtype=np.float16))
    b, n = re16.shape
    packed = np.ascontiguousarray(np.stack([re16, im16], axis=1)).view(np.uint16)  # (b, 2, n)
    out = np.empty_like(packed)
    rc = L.ref_fft(n, b, mode, 1 if use_batch_api else 0, packed.ctypes.data_as(_u16p), out.ctypes.data_as(_u16p))
    if rc != 0:
        raise RuntimeError(f"reference driver failed rc={rc}")
    o = out.view(np.float16)
    return o[:, 0, :].copy(), o[:, 1, :].copy()


def ref_bench_gpu(n: int, batch: int, steps: int, warmup: int, mode: int = 0, use_batch_api: bool = True,
                  seed: int = 1234):
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libtfft_ref.so missing")
    re, im = gauss_fixture(n, batch, seed)
    packed = np.ascontiguousarray(np.stack([re, im], axis=1)).view(np.uint16)
    out = np.empty_like(packed)
    k_ms, e_ms = np.zeros(steps), np.zeros(steps)
    rc = L.ref_bench(n, batch, mode, 1 if use_batch_api else 0, steps, warmup, packed.ctypes.data_as(_u16p),
                     out.ctypes.data_as(_u16p), _d(k_ms), _d(e_ms))
    if rc != 0:
        raise RuntimeError(f"reference bench failed rc={rc}")
    return k_ms, e_ms
