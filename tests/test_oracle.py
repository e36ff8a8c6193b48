"""CPU: pins the oracle (oracle/tfft_oracle.cpp) against numpy's pocketfft, the reference's index
algebra, its fixture recipe and its error statistics.  No GPU."""
import numpy as np
import pytest

import oracle as O


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 16384])
def test_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    re, im = rng.standard_normal((3, n)), rng.standard_normal((3, n))
    want = np.fft.fft(re + 1j * im, axis=1) / n
    got_re, got_im = O.fft_f64(re, im)
    assert np.abs(got_re + 1j * got_im - want).max() < 1e-13


def test_naive_dft_matches_numpy_config1():
    # BASELINE configs[0]: N=4096, batch 1, "fp64 host DFT"
    n = 4096
    re, im = O.sine_fixture(n)
    want = np.fft.fft(re + 1j * im) / n
    got_re, got_im = O.dft_f64(re, im)
    assert np.abs(got_re[0] + 1j * got_im[0] - want).max() < 1e-12


def test_fft2_matches_numpy():
    rng = np.random.default_rng(3)
    re, im = rng.standard_normal((2, 32, 64)), rng.standard_normal((2, 32, 64))
    want = np.fft.fft2(re + 1j * im) / (32 * 64)
    got_re, got_im = O.fft2_f64(re, im)
    assert np.abs(got_re + 1j * got_im - want).max() < 1e-13


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192, 65536])
def test_reference_algorithm_structure_equals_definition(n):
    """The reference's staged algorithm (digit reversal + 256 base + radix-16 + radix-2,
    src/base/ComputeFFT.h:54-151) in exact arithmetic IS the 1/N-scaled forward DFT in natural order."""
    rng = np.random.default_rng(n + 1)
    re, im = rng.standard_normal(n), rng.standard_normal(n)
    want = np.fft.fft(re + 1j * im) / n
    got_re, got_im = O.ref_algorithm(re, im, emulate_fp16=False)
    assert np.abs(got_re + 1j * got_im - want).max() < 1e-13


def test_reference_digit_reversal_is_a_permutation():
    # src/base/TensorFFT256.cu:125-161: 8192 = 2*16*16*16 example from the kernel's own comment
    n = 8192
    idx = np.array([O.ref_input_index(o, n) for o in range(n)])
    assert sorted(idx.tolist()) == list(range(n))
    o = 5 + 16 * 7 + 256 * 11 + 4096 * 1          # digits (d0, d1, d2, b0)
    assert O.ref_input_index(o, n) == ((5 * 16 + 7) * 16 + 11) * 2 + 1


@pytest.mark.parametrize("n,level", [(256, 5.1e-4), (4096, 6.6e-4), (16384, 8.0e-4)])
def test_reference_fp16_emulation_error_level(n, level):
    """fp16 emulation of the reference arithmetic lands at the error level SURVEY.md 8c estimates
    (and tests/golden pins against the real kernels)."""
    re, im = O.gauss_fixture(n, 1, seed=n)
    re, im = re[0].astype(np.float64), im[0].astype(np.float64)
    w_re, w_im = O.fft_f64(re, im)
    e_re, e_im = O.ref_algorithm(re, im, emulate_fp16=True)
    st = O.error_stats(e_re, e_im, w_re[0], w_im[0])
    assert 0.5 * level < st["rel_l2"] < 1.5 * level


def test_fixture_weights_are_libstdcxx_minstd():
    # std::default_random_engine == minstd_rand0 seeded via seed_seq{42}; first draws are fixed
    w = O.random_weights(4, 42)
    assert w.dtype == np.float32 and np.all(np.abs(w) <= 1)
    assert np.allclose(w, O.random_weights(4, 42))
    assert not np.allclose(w, O.random_weights(4, 1764))
    re, im = O.sine_fixture(1024, cutoff=256)
    assert re[0] == 0.0 and im[0] == 0.0          # harmonic 0 and t = 0 contribute sin(0)
    assert np.abs(re).max() < 256


def test_error_stats_triple():
    a = np.array([1.0, 2.0]); b = np.array([1.5, 2.0]); z = np.zeros(2)
    st = O.error_stats(a, z, b, z)
    assert st["max"] == 0.5 and abs(st["avg"] - 0.125) < 1e-15
    assert abs(st["rel_l2"] - 0.5 / np.sqrt(1.5 ** 2 + 4)) < 1e-15
    dev = np.array([0.5, 0, 0, 0]) - 0.125
    assert abs(st["sigma"] - np.sqrt((dev ** 2).sum() / 3)) < 1e-15
