"""GPU (-m gpu), round 2: regression guards at the measured error level, the sizes / modes / knobs that round 1 left
untested on hardware (VERDICT r1 "What's weak" 1-2), and the two remaining acceptance programs of the reference.
Everything goes through the C ABI; the checkers are the fp64 oracle (oracle/), a complex64 cuFFT for sizes the host
oracle would take minutes on, and size-independent properties.  Nothing here reads /root/reference."""
import os
import subprocess

import numpy as np
import pytest
import torch

import oracle as O
import tfft

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# rel-L2 vs fp64 measured on B200 for N(0,1) inputs (profiles/r01_sweep_c3.json); a kernel change that loses accuracy
# must trip these guards (1.25 x measured), not hide below the reference's 5e-4 ... 9e-4
MEASURED = {8: 3.23e-4, 9: 3.28e-4, 10: 3.28e-4, 11: 3.30e-4, 12: 3.34e-4, 13: 4.02e-4, 14: 4.00e-4, 15: 4.05e-4,
            16: 4.79e-4, 17: 4.70e-4, 18: 4.66e-4, 19: 4.65e-4, 20: 4.68e-4, 21: 4.67e-4, 22: 4.70e-4, 23: 4.70e-4,
            24: 6.12e-4}   # 2^24: three passes of 256 since round 2 (one more fp16 rounding than the four-step plan's 4.73e-4)
GUARD = 1.25


def _planar(re, im):
    return torch.from_numpy(np.ascontiguousarray(np.stack([re, im], axis=1)).reshape(-1)).cuda()


def _run(n, b, x, flags=0, tuner_file=None, inplace=False):
    y = x if inplace else torch.full_like(x, float("nan"))
    plan = tfft.NativePlan(n, b, flags, tuner_file=tuner_file)
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("lg", sorted(MEASURED))
def test_regression_guard_at_measured_error_level(lg):
    n = 1 << lg
    b = max(1, min(8, (1 << 22) // n))
    re, im = O.gauss_fixture(n, b, seed=4000 + lg)
    y = _run(n, b, _planar(re, im)).view(b, 2, n).cpu().numpy().astype(np.float64)
    w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
    err = O.error_stats(y[:, 0], y[:, 1], w_re, w_im)["rel_l2"]
    assert err <= GUARD * MEASURED[lg], (lg, err, MEASURED[lg])


@pytest.mark.parametrize("lg", [27, 28, 29])
def test_three_pass_sizes(lg):
    """N = 2^27 ... 2^29 (the reference's FFTBenchSinlge.cu goes up to 2^29): against a complex64 cuFFT of the same
    input (the host oracle needs minutes here), Parseval, and an impulse (pure phase ramp, checked exactly)."""
    n = 1 << lg
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    x = torch.randn(2 * n, generator=g, device="cuda").to(torch.float16)
    xs = torch.complex(x[:n].float(), x[n:].float())
    want = torch.fft.fft(xs) / n
    ex = float((xs.abs() ** 2).sum().double())
    del xs
    y = torch.empty_like(x)
    plan = tfft.NativePlan(n, 1)
    assert plan.info["passes"] == 3
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    got = torch.complex(y[:n].float(), y[n:].float())
    rel = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
    assert rel <= GUARD * 6.5e-4, rel          # measured 5.7e-4 ... 6.5e-4
    ey = float((got.abs() ** 2).sum().double()) * n
    assert abs(ex - ey) / ex < 2e-3
    del want, got
    pos, amp = (1 << (lg - 3)) + 12345, 16384.0
    x.zero_()
    x[pos] = amp
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    k = torch.arange(0, n, 4099, device="cuda", dtype=torch.int64)      # a sample of the outputs
    ang = -2 * np.pi * ((k * pos) % n).double() / n
    err = torch.sqrt(((y[k].double() - torch.cos(ang) * amp / n) ** 2 + (y[n + k].double() - torch.sin(ang) * amp / n) ** 2).mean())
    assert float(err) / (amp / n) < 1e-3


def test_three_pass_preserve_input():
    """TFFT_PRESERVE_INPUT above 2^24 (used to index past a table): pass A writes into a plan-owned scratch."""
    n, b = 1 << 25, 2
    g = torch.Generator(device="cuda"); g.manual_seed(25)
    x = torch.randn(b * 2 * n, generator=g, device="cuda").to(torch.float16)
    keep = x.clone()
    y = _run(n, b, x, flags=tfft.TFFT_PRESERVE_INPUT)
    assert bool(torch.equal(x, keep))
    y0 = _run(n, b, x)                                  # default: same arithmetic, input consumed
    assert bool(torch.equal(y, y0)) and not bool(torch.equal(x, keep))
    xs = torch.complex(keep[:n].float(), keep[n:2 * n].float())
    want = torch.fft.fft(xs) / n
    got = torch.complex(y[:n].float(), y[n:2 * n].float())
    assert float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want)) <= GUARD * 6.5e-4


@pytest.mark.parametrize("lg,b", [(24, 3), (25, 2), (26, 2)])
def test_three_pass_batched_equals_looped(monkeypatch, lg, b):
    """Three-pass plans run the whole batch in three launches (the transform index is an outer level of the unit index);
    the developer knob TFFT_THREEPASS_LOOP runs one transform at a time as in round 1: bit-identical, and every transform
    of the batch is checked against a complex64 FFT."""
    n = 1 << lg
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    x0 = torch.randn(b * 2 * n, generator=g, device="cuda").to(torch.float16)
    plan = tfft.NativePlan(n, b)
    assert plan.info["passes"] == 3
    y = torch.full_like(x0, float("nan"))
    x = x0.clone()
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    monkeypatch.setenv("TFFT_THREEPASS_LOOP", "1")
    plan2 = tfft.NativePlan(n, b)
    y2 = torch.full_like(x0, float("nan"))
    x = x0.clone()
    plan2.exec(x, x[n:], y2, y2[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    assert bool(torch.equal(y, y2))
    for i in range(b):
        xi, yi = x0.view(b, 2, n)[i], y.view(b, 2, n)[i]
        want = torch.fft.fft(torch.complex(xi[0].float(), xi[1].float())) / n
        got = torch.complex(yi[0].float(), yi[1].float())
        assert float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want)) <= GUARD * 6.5e-4, i


def test_three_pass_batched_with_a_ragged_transform_stride():
    """The batched three-pass plan loads its last pass through a 5-D tensor map whose transform coordinate folds the user's
    batch (stride = a whole number of matrix rows).  A stride that is not (2n + 8) takes the cp.async twin of that pass:
    same stages, same result up to the operand-layout dependent split of a twiddle (one fp16 rounding)."""
    n, b = 1 << 24, 2
    g = torch.Generator(device="cuda"); g.manual_seed(77)
    x0 = torch.randn(b, 2, n, generator=g, device="cuda").to(torch.float16)
    plan = tfft.NativePlan(n, b)
    x = x0.clone().view(-1)
    y = torch.full_like(x, float("nan"))
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    stride = 2 * n + 8
    xs = torch.zeros(b * stride, dtype=torch.float16, device="cuda")
    for i in range(b):
        xs[i * stride:i * stride + 2 * n] = x0[i].reshape(-1)
    ys = torch.full_like(xs, float("nan"))
    plan.exec(xs, xs[n:], ys, ys[n:], stride, stride)
    torch.cuda.synchronize()
    for i in range(b):
        got = ys[i * stride:i * stride + 2 * n].float()
        want = y[i * 2 * n:(i + 1) * 2 * n].float()
        d = got - want
        assert float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(want)) < 3e-4, i
        ref = torch.fft.fft(torch.complex(x0[i, 0].float(), x0[i, 1].float())) / n
        gc = torch.complex(got[:n], got[n:])
        assert float(torch.linalg.vector_norm(gc - ref) / torch.linalg.vector_norm(ref)) <= GUARD * 6.5e-4, i


def test_three_pass_from_2_24_developer_knob(monkeypatch):
    monkeypatch.setenv("TFFT_THREEPASS_LG", "24")
    n = 1 << 24
    re, im = O.gauss_fixture(n, 1, seed=24)
    plan = tfft.NativePlan(n, 1)
    assert plan.info["passes"] == 3
    x = _planar(re, im)
    y = torch.empty_like(x)
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
    o = y.cpu().numpy().astype(np.float64)
    assert O.error_stats(o[:n], o[n:], w_re[0], w_im[0])["rel_l2"] <= GUARD * 6.5e-4


@pytest.mark.parametrize("lg,b", [(8, 64), (10, 7), (12, 5), (13, 3), (14, 4), (15, 2)])
def test_in_place_single_pass(lg, b):
    """out == in is legal for n <= 32768 (include/tfft.h): bit-identical to the out-of-place result."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=700 + lg)
    want = _run(n, b, _planar(re, im))
    got = _run(n, b, _planar(re, im), inplace=True)
    assert bool(torch.equal(got, want))


@pytest.mark.parametrize("lg,b,knobs", [
    (14, 64, "two_slot=0"), (14, 64, "pipe=0"), (14, 64, "tma=0"), (14, 64, "two_slot=0 pipe=0 prefetch=1"),
    (13, 9, "two_slot=0"), (13, 9, "tma=0 pipe=0"), (12, 33, "tma=0"), (12, 33, "prefetch=0"), (10, 100, "tma=0"),
    (15, 3, "tma=0"), (15, 3, "prefetch=0"), (20, 2, "tma_col=0"), (20, 2, "tma=0"), (22, 1, "tma_col=0 prefetch=1"),
    # landing-ring kernel: on by default for the 4096-point column pass (ring=0 switches it off), opt-in (ring=2) for
    # N = 32768 and for the 4096-point row pass of 2^23
    (22, 3, "ring=0"), (15, 5, "ring=2"), (15, 300, "ring=2"), (23, 2, "ring=2"),
    # column tiles of 16 columns instead of the 64- / 32-column tiles of the 256- / 512-point column passes
    (16, 40, "tma_col=2"), (24, 2, "tma_col=2"), (16, 40, "tma_col=0"), (18, 10, "tma_col=2")])
def test_tuner_knobs_leave_the_result_unchanged(tmp_path, lg, b, knobs):
    """A tuner-file plan (tfft_plan_create_from_file, the reference's CreatePlan(N, file) overload, Plan.h:197-255) with
    non-default kernel knobs: the load path / pipelining / prefetch choices do not change the stages or the DFT matrices;
    only the split of an inter-stage twiddle into its per-thread and per-tile factor follows the operand layout, so
    single results may differ by one fp16 rounding (rel-L2 < 3e-4, no element off by more than 2^-9 of the largest)."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=800 + lg)
    want = _run(n, b, _planar(re, im))
    f = tmp_path / "TunerResults.dat"
    f.write_text(f"256 256 8 8 256\n{n} 256 8 8 256 {knobs}\n")
    got = _run(n, b, _planar(re, im), tuner_file=str(f))
    d = got.float() - want.float()
    assert float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(want.float())) < 3e-4, knobs
    assert float(d.abs().max()) <= float(want.float().abs().max()) * 2.0 ** -9, knobs


@pytest.mark.parametrize("lg,lg1", [(16, 8), (20, 9), (20, 11), (22, 10), (22, 11), (24, 12)])
def test_four_step_split_knob(tmp_path, lg, lg1):
    """lg1 = log2 of the column-pass length: another factorisation, same transform (within the guard)."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, 1, seed=900 + lg + lg1)
    f = tmp_path / "TunerResults.dat"
    f.write_text(f"{n} 256 8 8 256 lg1={lg1}\n")
    y = _run(n, 1, _planar(re, im), tuner_file=str(f)).cpu().numpy().astype(np.float64)
    w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
    assert O.error_stats(y[:n], y[n:], w_re[0], w_im[0])["rel_l2"] <= GUARD * 4.8e-4


@pytest.mark.parametrize("lg,b,lgt,first", [(8, 32, 20, 100), (12, 8, 24, 4000), (14, 4, 28, 16380), (15, 2, 30, 7)])
def test_exec_twiddled_against_oracle(lg, b, lgt, first):
    """tfft_exec_twiddled: output k of transform t times exp(-2 pi i k (first + t) / 2^lgt)."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=1000 + lg)
    x = _planar(re, im)
    y = torch.empty_like(x)
    plan = tfft.NativePlan(n, b)
    plan.exec_twiddled(x, x[n:], y, y[n:], 2 * n, 2 * n, lgt, first)
    torch.cuda.synchronize()
    w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
    k = np.arange(n)[None, :]
    col = (first + np.arange(b))[:, None]
    tw = np.exp(-2j * np.pi * ((k * col) % (1 << lgt)) / (1 << lgt))
    want = (w_re + 1j * w_im) * tw
    o = y.view(b, 2, n).cpu().numpy().astype(np.float64)
    assert O.error_stats(o[:, 0], o[:, 1], want.real, want.imag)["rel_l2"] <= GUARD * 4.8e-4


@pytest.mark.parametrize("case", ["four_step", "two_d", "ragged_tail", "single_chunk"])
def test_exec_host_pipeline_variants(case):
    """tfft_exec_host (chunk ring: upload / transform / download on three streams) equals tfft_exec on device buffers."""
    if case == "two_d":
        ny, nx, b = 512, 1024, 40          # 2 MiB per image: 4 images per chunk, 10 chunks through 4 slots
        n = ny * nx
        plan_args = dict(shape2d=(ny, nx))
    else:
        n, b = {"four_step": (1 << 16, 100), "ragged_tail": (1 << 14, 300), "single_chunk": (1 << 12, 5)}[case]
        plan_args = {}
    rng = np.random.default_rng(len(case))
    host = rng.standard_normal(2 * n * b).astype(np.float16)
    x = torch.from_numpy(host).cuda()
    y = torch.empty_like(x)
    plan = tfft.NativePlan(n, b, 0, **plan_args)
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    out = np.empty_like(host)
    for _ in range(2):                     # twice: slot reuse across calls
        out[:] = 0
        plan.exec_host(host, out)
    assert np.array_equal(out.view(np.uint16), y.cpu().numpy().view(np.uint16))


def _shim(exe):
    path = os.path.join(ROOT, "oracle", "_ref", exe)
    if not os.path.exists(path):
        pytest.skip("shim binaries not built (tests/test_compat_shim.py needs /root/reference at build time)")
    return path


def test_reference_accuracy_sweep_runs_against_the_shim(tmp_path):
    """benchmarks/AccuracyTest.cu, compiled unmodified against compat/base: N = 2^8 ... 2^28 on the reference's own
    sine fixture against a double-precision cuFFT; Accuracy_Test.dat (N avg sigma max per line, FileWriter.h:207-225)
    must stay inside the reference's unit-test thresholds (unitTesting/UnitTest.cu:14-16)."""
    r = subprocess.run([_shim("shim_AccuracyTest")], capture_output=True, text=True, timeout=1500, cwd=tmp_path)
    assert r.returncode == 1, (r.returncode, r.stdout[-800:], r.stderr[-800:])       # `return true;` from main
    rows = [l.split() for l in (tmp_path / "Accuracy_Test.dat").read_text().splitlines() if l.strip()]
    assert [int(x[0]) for x in rows] == [1 << lg for lg in range(8, 29)]
    for n, avg, sigma, mx in rows:
        assert float(avg) <= 1e-3 and float(sigma) <= 1e-2 and float(mx) <= 0.5, (n, avg, sigma, mx)
    with open(os.path.join(ROOT, "gpurun_out", "Accuracy_Test_shim.dat"), "w") as f:
        f.write((tmp_path / "Accuracy_Test.dat").read_text())


def test_reference_single_benchmark_runs_from_a_tuner_file(tmp_path):
    """benchmarks/FFTBenchSinlge.cu: N = 2^12 ... 2^29, every plan created from TunerResults.dat in the working directory
    (Bench.h:153-228 -> CreatePlan(N, file)); here the file tools/tune.py wrote on B200 (profiles/r01_TunerResults.dat, with
    key=value knobs) extended by default lines for the three-pass sizes."""
    lines = open(os.path.join(ROOT, "profiles", "r01_TunerResults.dat")).read().rstrip("\n").split("\n")
    have = {int(l.split()[0]) for l in lines if l.strip()}
    for lg in range(12, 30):
        if (1 << lg) not in have:
            lines.append(f"{1 << lg} 256 8 8 256")
    (tmp_path / "TunerResults.dat").write_text("\n".join(lines) + "\n")
    r = subprocess.run([_shim("shim_FFTBenchSinlge")], capture_output=True, text=True, timeout=1500, cwd=tmp_path)
    assert r.returncode == 1, (r.returncode, r.stdout[-800:], r.stderr[-800:])
    assert r.stdout.count("Benchmarking fft_length") == 18 and "rror" not in r.stdout
    outs = [p for p in os.listdir(tmp_path) if p != "TunerResults.dat"]
    assert outs, "no benchmark file written"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for p in outs:
        with open(os.path.join(ROOT, "gpurun_out", "shim_" + p), "w") as f:
            f.write((tmp_path / p).read_text())


@pytest.mark.parametrize("lg,b", [(16, 5), (16, 300), (22, 2), (24, 1)])
def test_cluster_units_through_the_tuner_file(tmp_path, lg, b):
    """Tuner key cluster=1: units of 2^16 elements shared by a CTA pair (thread-block cluster of 2, stage-1 outputs
    exchanged through distributed shared memory).  N = 65536 then runs in ONE HBM pass (passes == 1, input preserved);
    for 2^22 / 2^24 the 4096-point column pass uses 16-column units.  Same guard as the default plans."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=1600 + lg)
    f = tmp_path / "TunerResults.dat"
    f.write_text(f"{n} 256 8 8 256 cluster=1\n")
    x = _planar(re, im)
    keep = x.clone()
    plan = tfft.NativePlan(n, b, 0, tuner_file=str(f))
    assert plan.info["passes"] == (1 if lg == 16 else 2)
    y = torch.full_like(x, float("nan"))
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    if lg == 16:
        assert bool(torch.equal(x, keep))                 # a single pass never writes its input
    nb = min(b, 4)
    o = y.view(b, 2, n)[:nb].cpu().numpy().astype(np.float64)
    w_re, w_im = O.fft_f64(re[:nb].astype(np.float64), im[:nb].astype(np.float64))
    err = O.error_stats(o[:, 0], o[:, 1], w_re, w_im)["rel_l2"]
    assert err <= GUARD * (4.05e-4 if lg == 16 else MEASURED[lg]), err
    # last transform too (ragged CTA-pair waves)
    ol = y.view(b, 2, n)[b - 1].cpu().numpy().astype(np.float64)
    wl_re, wl_im = O.fft_f64(re[b - 1:].astype(np.float64), im[b - 1:].astype(np.float64))
    assert O.error_stats(ol[0], ol[1], wl_re[0], wl_im[0])["rel_l2"] <= GUARD * (4.05e-4 if lg == 16 else MEASURED[lg])


@pytest.mark.parametrize("lg,b,segs,lgt", [(11, 64, 8, 0), (12, 32, 4, 24), (13, 24, 16, 0), (14, 16, 8, 28), (14, 5, 2, 0),
                                           (15, 4, 8, 0), (10, 16, 4, 0)])
def test_exec_segmented_gathers_in_the_tma_load(lg, b, segs, lgt):
    """tfft_exec_segmented: piece q of transform t at in + q*segment_stride + t*in_stride (the source-rank-major staging planes
    of the multi-GPU transform).  Must equal tfft_exec_twiddled on the gathered input bit for bit; shapes whose row tiles
    are not 64-row atoms (n <= 1024) answer TFFT_E_UNSUPPORTED."""
    n = 1 << lg
    seg = n // segs
    re, im = O.gauss_fixture(n, b, seed=1700 + lg)
    # staging layout [q][t][r]
    st_re = torch.from_numpy(np.ascontiguousarray(re.reshape(b, segs, seg).transpose(1, 0, 2))).cuda().reshape(-1)
    st_im = torch.from_numpy(np.ascontiguousarray(im.reshape(b, segs, seg).transpose(1, 0, 2))).cuda().reshape(-1)
    g_re, g_im = torch.from_numpy(re).cuda().reshape(-1), torch.from_numpy(im).cuda().reshape(-1)
    plan = tfft.NativePlan(n, b)
    want_re, want_im = torch.empty_like(g_re), torch.empty_like(g_im)
    if lgt:
        plan.exec_twiddled(g_re, g_im, want_re, want_im, n, n, lgt, 3)
    else:
        plan.exec(g_re, g_im, want_re, want_im, n, n)
    out_re, out_im = torch.full_like(g_re, float("nan")), torch.full_like(g_im, float("nan"))
    if lg <= 10:
        with pytest.raises(tfft.TfftError):
            plan.exec_segmented(st_re, st_im, out_re, out_im, seg, n, segs, b * seg, lgt, 3 if lgt else 0)
        return
    plan.exec_segmented(st_re, st_im, out_re, out_im, seg, n, segs, b * seg, lgt, 3 if lgt else 0)
    torch.cuda.synchronize()
    assert bool(torch.equal(out_re, want_re)) and bool(torch.equal(out_im, want_im))


@pytest.mark.parametrize("lg,b", [(12, 1), (14, 64), (20, 2)])
def test_exec_can_be_captured_into_a_cuda_graph(lg, b):
    """tfft_plan_prepare does every lazy device initialisation up front, so a later exec only enqueues kernels and can be
    recorded by stream capture (launch-bound loops of small transforms replay as one graph launch)."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=1900 + lg)
    x = _planar(re, im)
    src = x.clone()
    plan = tfft.NativePlan(n, b, tfft.TFFT_PRESERVE_INPUT if lg > 15 else 0)
    plan.prepare()
    want = torch.empty_like(x)
    plan.exec(x, x[n:], want, want[n:], 2 * n, 2 * n)          # also fills the plan's launch cache for these buffers
    torch.cuda.synchronize()
    y = torch.zeros_like(x)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)            # warm-up on the capture stream
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(4):
                plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    y.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert bool(torch.equal(y, want)) and bool(torch.equal(x, src))
    g.replay()
    torch.cuda.synchronize()
    assert bool(torch.equal(y, want))


def _close_to_rounding(got, want):
    d = got.float() - want.float()
    assert float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(want.float())) < 3e-4
    assert float(d.abs().max()) <= float(want.float().abs().max()) * 2.0 ** -9


@pytest.mark.parametrize("lg,b,flags", [(25, 2, 0), (26, 1, tfft.TFFT_INVERSE)])
def test_interleaved_three_pass(lg, b, flags):
    """TFFT_INTERLEAVED above 2^24 (round 1: unsupported): pass A reads half2 pairs, pass C writes them, planar scratch in
    between; same stages as the planar plan, input preserved."""
    n = 1 << lg
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    planar = torch.randn(b * 2 * n, generator=g, device="cuda").to(torch.float16)
    inter = planar.view(b, 2, n).permute(0, 2, 1).contiguous().view(-1)         # (b, n, 2)
    keep = inter.clone()
    want = torch.empty_like(planar)
    tfft.NativePlan(n, b, flags | tfft.TFFT_PRESERVE_INPUT).exec(planar, planar[n:], want, want[n:], 2 * n, 2 * n)
    out = torch.full_like(inter, float("nan"))
    plan = tfft.NativePlan(n, b, flags | tfft.TFFT_INTERLEAVED)
    assert plan.info["passes"] == 3
    plan.exec(inter, inter, out, out, n, n)
    torch.cuda.synchronize()
    assert bool(torch.equal(inter, keep))
    got = out.view(b, n, 2)
    _close_to_rounding(got[:, :, 0], want.view(b, 2, n)[:, 0])
    _close_to_rounding(got[:, :, 1], want.view(b, 2, n)[:, 1])


@pytest.mark.parametrize("ny,nx,b,flags", [(512, 1024, 3, 0), (8192, 4096, 1, 0), (2048, 2048, 2, tfft.TFFT_INVERSE)])
def test_interleaved_two_d(ny, nx, b, flags):
    """TFFT_INTERLEAVED 2-D plans (round 1: unsupported): half2 images in and out, the intermediate lives in the output."""
    n = ny * nx
    g = torch.Generator(device="cuda"); g.manual_seed(ny + nx)
    planar = torch.randn(b * 2 * n, generator=g, device="cuda").to(torch.float16)
    inter = planar.view(b, 2, n).permute(0, 2, 1).contiguous().view(-1)
    keep = inter.clone()
    want = torch.empty_like(planar)
    tfft.NativePlan(n, b, flags, shape2d=(ny, nx)).exec(planar, planar[n:], want, want[n:], 2 * n, 2 * n)
    out = torch.full_like(inter, float("nan"))
    plan = tfft.NativePlan(n, b, flags | tfft.TFFT_INTERLEAVED, shape2d=(ny, nx))
    plan.exec(inter, inter, out, out, n, n)
    torch.cuda.synchronize()
    assert bool(torch.equal(inter, keep))
    got = out.view(b, n, 2)
    _close_to_rounding(got[:, :, 0], want.view(b, 2, n)[:, 0])
    _close_to_rounding(got[:, :, 1], want.view(b, 2, n)[:, 1])
