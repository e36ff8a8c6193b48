"""GPU (-m gpu): 2-D transform (BASELINE config C5) through the C ABI
(tfft_plan_create_2d + tfft_exec) against the fp64 oracle (oracle.fft2_f64) and, at the full
8192 x 8192 size, a plain fp32 torch FFT of the same op plus size-independent properties.
The reference has no 2-D path; each 1-D transform keeps the reference's conventions (forward,
1/N scaled, planar fp16), so the tolerance is the reference's error level for a 1-D transform of
the same total length (tests/test_gpu_parity.py REF_LEVEL, 9.0e-4 above 2^15)."""
import numpy as np
import pytest
import torch

import oracle as O
import tfft

pytestmark = pytest.mark.gpu
TOL = 9.0e-4


def run_2d(re16, im16, img_stride=None, flags=0):
    """re16/im16 (batch, ny, nx) float16 -> (batch, ny, nx) float64 x2; planar images [RE_b | IM_b]."""
    b, ny, nx = re16.shape
    n = ny * nx
    stride = img_stride or 2 * n
    buf = np.zeros((b, stride), dtype=np.float16)
    buf[:, :n], buf[:, n:2 * n] = re16.reshape(b, n), im16.reshape(b, n)
    d_in = torch.from_numpy(buf).cuda().reshape(-1)
    d_out = torch.full((b * 2 * n,), float("nan"), dtype=torch.float16, device="cuda")
    plan = tfft.NativePlan(n, b, flags, shape2d=(ny, nx))
    plan.exec(d_in, d_in[n:], d_out, d_out[n:], stride, 2 * n)
    torch.cuda.synchronize()
    assert np.array_equal(d_in.cpu().numpy().view(np.uint16), buf.reshape(-1).view(np.uint16))   # input preserved
    o = d_out.cpu().numpy().reshape(b, 2, ny, nx)
    return o[:, 0].astype(np.float64), o[:, 1].astype(np.float64), plan


# (ny, nx, batch): plain row pass (ny <= 4096) and one Kronecker shape (4 rows per unit) at sizes the fp64 host
# oracle finishes in seconds
SHAPES = [(256, 256, 3), (512, 4096, 2), (4096, 256, 1), (2048, 2048, 1), (8192, 2048, 1)]


@pytest.mark.parametrize("ny,nx,batch", SHAPES)
def test_2d_vs_fp64_oracle(ny, nx, batch):
    rng = np.random.default_rng(ny + nx)
    re = rng.standard_normal((batch, ny, nx)).astype(np.float16)
    im = rng.standard_normal((batch, ny, nx)).astype(np.float16)
    g_re, g_im, plan = run_2d(re, im)
    assert plan.info["passes"] == 2                       # touches HBM twice: rows, columns
    w_re, w_im = O.fft2_f64(re.astype(np.float64), im.astype(np.float64))
    st = O.error_stats(g_re.reshape(batch, -1), g_im.reshape(batch, -1), w_re.reshape(batch, -1), w_im.reshape(batch, -1))
    assert st["rel_l2"] <= TOL, st
    assert np.isfinite(g_re).all() and np.isfinite(g_im).all()


# bigger shapes (Kronecker units of 2 and 4 rows, 8K/16K/32K-element units, 32768-point rows): the checker is a
# plain torch fp64 fft2 of the same op on the GPU
BIG = [(1024, 32768, 1), (8192, 4096, 2), (16384, 4096, 1), (16384, 2048, 2), (8192, 16384, 1), (16384, 8192, 1)]


@pytest.mark.parametrize("ny,nx,batch", BIG)
def test_2d_big_shapes_vs_torch_fp64(ny, nx, batch):
    n = ny * nx
    g = torch.Generator(device="cuda"); g.manual_seed(ny * 3 + nx)
    x = torch.randn(batch * 2 * n, generator=g, device="cuda").to(torch.float16)
    keep = x.clone()
    y = torch.empty_like(x)
    plan = tfft.NativePlan(n, batch, 0, shape2d=(ny, nx))
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    assert bool(torch.equal(x, keep))                     # input preserved
    xv, yv = x.view(batch, 2, ny, nx), y.view(batch, 2, ny, nx)
    for i in range(batch):
        want = torch.fft.fft2(torch.complex(xv[i, 0].double(), xv[i, 1].double())) / n
        got = torch.complex(yv[i, 0].double(), yv[i, 1].double())
        rel = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
        assert rel <= TOL, (i, rel)
        del want, got


def test_2d_padded_image_stride_is_bit_identical():
    """Images that are not contiguous (stride > 2*ny*nx) take the non-TMA load path: same bits."""
    ny, nx, b = 8192, 4096, 2
    rng = np.random.default_rng(77)
    re = rng.standard_normal((b, ny, nx)).astype(np.float16)
    im = rng.standard_normal((b, ny, nx)).astype(np.float16)
    a_re, a_im, _ = run_2d(re, im)
    b_re, b_im, _ = run_2d(re, im, img_stride=2 * ny * nx + 4096)
    assert np.array_equal(a_re, b_re) and np.array_equal(a_im, b_im)


def test_2d_unsupported_and_invalid_shapes_fail_loudly():
    for ny, nx in ((8192, 256), (100, 256), (256, 65536), (128, 4096)):
        with pytest.raises(tfft.TfftError):
            tfft.NativePlan(ny * nx, 1, 0, shape2d=(ny, nx))


@pytest.mark.parametrize("ybits", [None, "2"])
def test_config5_full_size_8192x8192(ybits, monkeypatch):
    """ybits: rows per unit of the row pass = 2 (default for ny = 8192: 16K-element units) or 4 (32K-element units).
    C5: 8192 x 8192 images (two of the 16; the batch is sharded over GPUs) against an fp32 torch fft2,
    Parseval, an impulse (a pure 2-D plane wave) and bit-exact determinism across images."""
    if ybits:
        monkeypatch.setenv("TFFT_2D_YBITS", ybits)
    ny = nx = 8192
    n, b = ny * nx, 2
    g = torch.Generator(device="cuda"); g.manual_seed(55)
    x = torch.randn(2 * n, generator=g, device="cuda").to(torch.float16)
    x = torch.cat([x, x])                                  # both images identical
    y = torch.empty_like(x)
    plan = tfft.NativePlan(n, b, 0, shape2d=(ny, nx))
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    yv = y.view(b, 2, ny, nx)
    assert bool(torch.equal(yv[0], yv[1]))
    xs = torch.complex(x[:n].float(), x[n:2 * n].float()).view(ny, nx)
    want = torch.fft.fft2(xs) / n
    got = torch.complex(yv[0, 0].float(), yv[0, 1].float())
    rel = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
    assert rel <= TOL, rel
    ex = float((xs.abs() ** 2).sum().double())
    ey = float((got.abs() ** 2).sum().double()) * n
    assert abs(ex - ey) / ex < 2e-3
    del want, got, xs
    # impulse at (py, px) -> exp(-2 pi i (ky py / ny + kx px / nx)) * A / n
    py, px, amp = 4099, 77, 16384.0
    x.zero_()
    x.view(b, 2, ny, nx)[0, 0, py, px] = amp
    plan.exec(x, x[n:], y, y[n:], 2 * n, 2 * n)
    torch.cuda.synchronize()
    ky = torch.arange(ny, device="cuda", dtype=torch.float64)[:, None]
    kx = torch.arange(nx, device="cuda", dtype=torch.float64)[None, :]
    ang = -2 * np.pi * (((ky * py) % ny) / ny + ((kx * px) % nx) / nx)
    w_re, w_im = torch.cos(ang) * amp / n, torch.sin(ang) * amp / n
    err = torch.sqrt(((yv[0, 0].double() - w_re) ** 2 + (yv[0, 1].double() - w_im) ** 2).sum())
    assert float(err) / (np.sqrt(n) * amp / n) < TOL
    assert float(yv[1].abs().max()) == 0.0
