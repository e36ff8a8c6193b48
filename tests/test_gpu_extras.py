"""GPU (-m gpu): the steps either side of the hot path through the C ABI:
inverse / unscaled transforms (the cuFFT conventions the reference compares itself with,
src/testing/unitTesting/CuFFTTest.h:25-57) and the device-side fixture + deviation statistics
(src/testing/TestingDataCreation.h:89-117, src/testing/AccuracyCalculator.h:86-148)."""
import numpy as np
import pytest
import torch

import oracle as O
import tfft

pytestmark = pytest.mark.gpu


def _exec(n, b, flags, x):
    y = torch.empty_like(x)
    plan = tfft.NativePlan(n, b, flags)
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("lg,b", [(8, 9), (12, 3), (14, 2), (15, 1), (18, 1)])
def test_inverse_flag_is_the_conjugate_transform(lg, b):
    """F^-1(x) == conj(F(conj(x))) up to fp32 summation order (the planes are exchanged instead of negated), and
    within the reference's fp16 error level of the fp64 oracle."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=300 + lg)
    host = np.ascontiguousarray(np.stack([re, im], axis=1)).reshape(-1)
    x = torch.from_numpy(host).cuda()
    inv = _exec(n, b, tfft.TFFT_INVERSE, x.clone()).view(b, 2, n)
    xc = x.clone().view(b, 2, n)
    xc[:, 1] = -xc[:, 1]
    fc = _exec(n, b, 0, xc.view(-1)).view(b, 2, n)
    d = torch.stack([inv[:, 0].float() - fc[:, 0].float(), inv[:, 1].float() + fc[:, 1].float()])
    assert float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(fc.float())) < 2e-4
    w_re, w_im = O.fft_f64(re.astype(np.float64), -im.astype(np.float64))        # fp64 oracle of the conjugate
    st = O.error_stats(inv[:, 0].cpu().numpy().astype(np.float64), inv[:, 1].cpu().numpy().astype(np.float64), w_re, -w_im)
    assert st["rel_l2"] <= 9.0e-4, st


@pytest.mark.parametrize("lg,b", [(8, 9), (11, 4), (14, 2), (16, 1)])
def test_unscaled_flag_matches_cufft_convention(lg, b):
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=400 + lg)
    host = np.ascontiguousarray(np.stack([re, im], axis=1)).reshape(-1)
    x = torch.from_numpy(host).cuda()
    y = _exec(n, b, tfft.TFFT_UNSCALED, x.clone()).view(b, 2, n).cpu().numpy().astype(np.float64)
    w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
    st = O.error_stats(y[:, 0], y[:, 1], w_re * n, w_im * n)                       # oracle is 1/N scaled
    assert st["rel_l2"] <= 9.0e-4, st
    # round trip: unscaled forward, then scaled inverse gives x back (to fp16 accuracy of two transforms)
    z = _exec(n, b, tfft.TFFT_INVERSE, torch.from_numpy(np.ascontiguousarray(y.astype(np.float16)).reshape(-1)).cuda())
    z = z.view(b, 2, n).cpu().numpy().astype(np.float64)
    assert O.error_stats(z[:, 0], z[:, 1], re.astype(np.float64), im.astype(np.float64))["rel_l2"] <= 2e-3


def test_device_sine_fixture_matches_the_oracle_fixture():
    """Same recipe as the reference's fixture kernel; the host oracle evaluates it with libm's sinf, the device
    with CUDA's: results may differ by one fp16 ulp on a small fraction of samples."""
    n, cutoff = 4096, 256
    w_re, w_im = O.random_weights(cutoff, 42), O.random_weights(cutoff, 42 * 42)
    re = torch.empty(n, dtype=torch.float16, device="cuda")
    im = torch.empty(n, dtype=torch.float16, device="cuda")
    tfft.fixture_sine(re, im, n, w_re, w_im)
    h_re, h_im = O.sine_fixture(n, cutoff=cutoff, seed_re=42, seed_im=42 * 42)
    for got, want in ((re, h_re), (im, h_im)):
        g = got.cpu().numpy().astype(np.float64)
        w = np.asarray(want).astype(np.float16).astype(np.float64)
        ulp = np.maximum(np.abs(w), 2.0 ** -14) * 2.0 ** -10
        assert np.all(np.abs(g - w) <= ulp)
        assert np.mean(g != w) < 0.02


def test_device_error_stats_match_the_oracle_statistics():
    n, b = 8192, 3
    re, im = O.gauss_fixture(n, b, seed=77)
    host = np.ascontiguousarray(np.stack([re, im], axis=1)).reshape(-1)
    y = _exec(n, b, 0, torch.from_numpy(host).cuda()).view(b, 2, n)
    w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
    a_re, a_im = y[:, 0].contiguous(), y[:, 1].contiguous()
    got = tfft.error_stats(a_re, a_im, torch.from_numpy(w_re).cuda(), torch.from_numpy(w_im).cuda())
    want = O.error_stats(a_re.cpu().numpy().astype(np.float64), a_im.cpu().numpy().astype(np.float64), w_re, w_im)
    for k in ("max", "avg", "sigma", "rel_l2"):
        assert abs(got[k] - want[k]) <= 1e-9 * max(1.0, abs(want[k])) + 1e-6 * abs(want[k]), (k, got, want)


def test_exchange_pack_unpack_kernels_match_torch():
    """tfft_transpose_blocks / tfft_copy_runs (the pack / unpack of the six-step's all-to-all) against the torch
    permute they replace: bit-exact data movement."""
    from tfft import dist as tdist
    for world, rows_local, cols in ((2, 128, 512), (4, 64, 1024), (8, 256, 512)):
        z = torch.randn(2, rows_local, cols, device="cuda").to(torch.float16)
        cl = cols // world
        want_send = z.reshape(2, rows_local, world, cl).permute(2, 0, 3, 1).contiguous()
        got_send = tdist._gpu_pack(z, world)
        assert bool(torch.equal(got_send, want_send))
        want_out = want_send.permute(1, 2, 0, 3).reshape(2, cl, world * rows_local).contiguous()
        got_out = tdist._gpu_unpack(got_send, world)
        torch.cuda.synchronize()
        assert bool(torch.equal(got_out, want_out))


@pytest.mark.parametrize("lg,b,flags", [(8, 13, 0), (10, 5, 0), (12, 3, 0), (13, 2, 0), (14, 2, tfft.TFFT_INVERSE),
                                        (15, 1, 0), (16, 2, 0), (20, 1, tfft.TFFT_INVERSE), (22, 1, 0)])
def test_interleaved_layout_equals_planar_bit_for_bit(lg, b, flags):
    """TFFT_INTERLEAVED (cuFFT's half2 layout, CuFFTTest.h:25-57): same arithmetic as the planar transform, so the
    results agree bit for bit (the planar run uses the cp.async load path too: TFFT_NO_TMA is not needed because the
    tensor-core stages do not depend on how the operand was loaded)."""
    n = 1 << lg
    re, im = O.gauss_fixture(n, b, seed=500 + lg)
    planar = torch.from_numpy(np.ascontiguousarray(np.stack([re, im], axis=1)).reshape(-1)).cuda()
    want = _exec(n, b, flags, planar.clone()).view(b, 2, n)
    inter = torch.from_numpy(np.ascontiguousarray(np.stack([re, im], axis=2)).reshape(-1)).cuda()   # (b, n, 2)
    keep = inter.clone()
    out = torch.full_like(inter, float("nan"))
    plan = tfft.NativePlan(n, b, flags | tfft.TFFT_INTERLEAVED)
    plan.exec(inter, inter, out, out, n, n)
    torch.cuda.synchronize()
    got = out.view(b, n, 2)
    # same stages and DFT matrices; the inter-stage twiddles are products of a per-thread and a per-tile factor whose
    # split follows the operand layout (cp.async vs TMA tiles), so single results may differ by an fp16 rounding
    for g_, w_ in ((got[:, :, 0], want[:, 0]), (got[:, :, 1], want[:, 1])):
        d = (g_.float() - w_.float())
        # measured 1.3e-4 at 2^16 (the three-term recurrence amplifies a seed difference of one fp32 ulp to a few 1e-6,
        # which flips a few per cent of the fp16 roundings); the transform's own error is 4.8e-4
        assert float(torch.linalg.vector_norm(d) / torch.linalg.vector_norm(w_.float())) < 3e-4
        assert float(d.abs().max()) <= float(w_.float().abs().max()) * 2.0 ** -9
    assert bool(torch.equal(inter, keep))                     # interleaved plans never overwrite their input
    # host path with interleaved buffers
    if lg <= 14:
        h_out = np.empty(2 * n * b, dtype=np.float16)
        plan.exec_host(keep.cpu().numpy(), h_out)
        assert np.array_equal(h_out.view(np.uint16), out.cpu().numpy().view(np.uint16))


def test_interleaved_with_twiddled_and_segmented_exec_fails_loudly():
    """Round 2 supports TFFT_INTERLEAVED for every size and for 2-D (tests/test_gpu_round2.py); what stays planar-only are
    the building blocks of the multi-GPU transform."""
    plan = tfft.NativePlan(4096, 4, tfft.TFFT_INTERLEAVED)
    x = torch.zeros(2 * 4096 * 4, dtype=torch.float16, device="cuda")
    with pytest.raises(tfft.TfftError):
        plan.exec_twiddled(x, x, x, x, 4096, 4096, 20, 0)
    with pytest.raises(tfft.TfftError):
        plan.exec_segmented(x, x, x, x, 512, 4096, 8, 2048)


def test_dependent_launches_back_to_back_are_ordered():
    """Every kernel is launched with programmatic stream serialization (the next kernel's prologue may start while the
    previous one drains) and waits with griddepcontrol.wait before touching data.  A chain forward -> inverse ->
    forward ... on ONE stream without host synchronisation must behave like serialized launches: after 2k launches
    the data is back (up to fp16 rounding), and the result equals the same chain run with a sync after every launch."""
    n, b = 16384, 1024
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    x0 = (torch.randn(b * 2 * n, generator=g, device="cuda") * 0.5).to(torch.float16)
    fwd = tfft.NativePlan(n, b, tfft.TFFT_UNSCALED)
    inv = tfft.NativePlan(n, b, tfft.TFFT_INVERSE)

    def chain(sync):
        a, c = x0.clone(), torch.empty_like(x0)
        for _ in range(4):
            fwd.exec(a, a[n:], c, c[n:], 2 * n, 2 * n)
            if sync:
                torch.cuda.synchronize()
            inv.exec(c, c[n:], a, a[n:], 2 * n, 2 * n)
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return a

    a_async, a_sync = chain(False), chain(True)
    assert bool(torch.equal(a_async, a_sync))
    rel = float(torch.linalg.vector_norm(a_async.float() - x0.float()) / torch.linalg.vector_norm(x0.float()))
    assert rel < 5e-3, rel


def test_two_plans_on_two_streams_concurrently():
    """Plans are immutable after creation: execs of different plans on different streams may overlap on the device
    (CTAs of both kernels share SMs, shared memory and tensor memory) and must give the serial results."""
    n1, b1, n2, b2 = 4096, 2048, 1 << 18, 8
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    x1 = torch.randn(b1 * 2 * n1, generator=g, device="cuda").to(torch.float16)
    x2 = torch.randn(b2 * 2 * n2, generator=g, device="cuda").to(torch.float16)
    p1, p2 = tfft.NativePlan(n1, b1), tfft.NativePlan(n2, b2, tfft.TFFT_PRESERVE_INPUT)
    w1, w2 = torch.empty_like(x1), torch.empty_like(x2)
    p1.exec(x1, x1[n1:], w1, w1[n1:], 2 * n1, 2 * n1)
    p2.exec(x2, x2[n2:], w2, w2[n2:], 2 * n2, 2 * n2)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    y1, y2 = torch.empty_like(x1), torch.empty_like(x2)
    for _ in range(5):
        with torch.cuda.stream(s1):
            p1.exec(x1, x1[n1:], y1, y1[n1:], 2 * n1, 2 * n1)
        with torch.cuda.stream(s2):
            p2.exec(x2, x2[n2:], y2, y2[n2:], 2 * n2, 2 * n2)
    torch.cuda.synchronize()
    assert bool(torch.equal(y1, w1)) and bool(torch.equal(y2, w2))


def test_four_step_odd_batch_stride_takes_the_cp_async_twin():
    """The four-step row pass loads TMA row tiles when the batch stride is a whole number of rows; any other stride
    (multiple of 8 elements) must still work, through the cp.async twin of that pass, with identical bits."""
    n, b = 1 << 16, 3
    re, im = O.gauss_fixture(n, b, seed=91)
    want = None
    for stride in (2 * n, 2 * n + 8, 2 * n + 264):
        buf = np.zeros((b, stride), dtype=np.float16)
        buf[:, :n], buf[:, n:2 * n] = re, im
        x = torch.from_numpy(buf).cuda().reshape(-1)
        y = torch.empty(b * 2 * n, dtype=torch.float16, device="cuda")
        plan = tfft.NativePlan(n, b)
        plan.exec(x, x[n:], y, y[n:], stride, 2 * n)
        torch.cuda.synchronize()
        if want is None:
            want = y.clone()
            w_re, w_im = O.fft_f64(re.astype(np.float64), im.astype(np.float64))
            got = y.view(b, 2, n).cpu().numpy().astype(np.float64)
            assert O.error_stats(got[:, 0], got[:, 1], w_re, w_im)["rel_l2"] <= 9.0e-4
        else:
            assert bool(torch.equal(y, want)), stride
