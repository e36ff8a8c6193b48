"""CPU: the kernel's index algebra (tensor-fft_b200/csrc/unit_plan.h) interpreted on the host
(tests/sim/plan_sim.cpp) equals the DFT for every supported shape, is free of shared-memory bank
conflicts, and predicts an fp16 error level below the reference's."""
import ctypes
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def sim():
    L = ctypes.CDLL(os.path.join(HERE, "sim", "libplansim.so"))
    L.plansim_run.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int,
                                                   dp, dp, dp, dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    return L


def _rows(sim, lg, ups, n_units=2, h=0, tstride=None, flags=0):
    n, U = 1 << lg, 1 << ups
    tstride = tstride or n
    tot = n_units * U * tstride
    rng = np.random.default_rng(lg * 10 + ups)
    re, im = rng.standard_normal(tot), rng.standard_normal(tot)
    ore, oim = np.zeros(tot), np.zeros(tot)
    st = (ctypes.c_int64 * 9)(tstride, 1, tstride, 1, 0, U * tstride, 0, U * tstride, 1 << 30)
    conf = (ctypes.c_int * 4)()
    rc = sim.plansim_run(lg, ups, flags, 0, st, 0, n_units, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                         ore.ctypes.data_as(dp), oim.ctypes.data_as(dp), h, conf)
    x = (re + 1j * im).reshape(n_units * U, tstride)[:, :n]
    if h:
        x = x.real.astype(np.float16).astype(np.float64) + 1j * x.imag.astype(np.float16).astype(np.float64)
    want = np.fft.fft(x, axis=1) / n
    got = (ore + 1j * oim).reshape(n_units * U, tstride)[:, :n]
    return rc, list(conf)[:3], np.linalg.norm(got - want) / np.linalg.norm(want)


SHAPES = [(lg, ups) for lg in range(8, 16) for ups in range(0, 8) if 13 <= lg + ups <= 15]


@pytest.mark.parametrize("lg,ups", SHAPES)
def test_row_pass_is_the_dft_and_conflict_free(sim, lg, ups):
    rc, conf, err = _rows(sim, lg, ups)
    assert rc == 0 and conf == [0, 0, 0]
    assert err < 1e-13


@pytest.mark.parametrize("lg,ups", [(8, 6), (8, 5), (9, 5), (10, 4), (10, 3), (11, 2), (12, 2), (13, 1), (14, 0), (15, 0)])
def test_row_pass_with_tma_tiles(sim, lg, ups):
    """Stage-1 operand filled by a SWIZZLE_128B TMA tile (natural row order); for N <= 1024 an atom of 64 rows is
    64/M consecutive transforms.  flags: 2 = TMA, 4 = stage-2 pipelining."""
    rc, conf, err = _rows(sim, lg, ups, flags=6 if lg >= 13 else 2)
    assert rc == 0 and conf == [0, 0, 0]
    assert err < 1e-13


def test_reference_batch_layout_stride(sim):
    rc, conf, err = _rows(sim, 12, 2, tstride=8192)      # [RE_b|IM_b]: stride 2N
    assert rc == 0 and err < 1e-13


@pytest.mark.parametrize("lg,ups,ref_level", [(8, 5, 5.1e-4), (12, 2, 6.6e-4), (14, 0, 8.0e-4)])
def test_predicted_fp16_error_below_reference(sim, lg, ups, ref_level):
    rc, _, err = _rows(sim, lg, ups, h=1)
    assert rc == 0 and err < ref_level


@pytest.mark.parametrize("lg1,lg2,u1,u2,tma", [(8, 8, 5, 5, 0), (8, 8, 6, 6, 2), (8, 8, 6, 6, 2 | 32), (10, 10, 4, 4, 0), (11, 11, 3, 3, 0),
                                               (12, 10, 3, 4, 0), (9, 8, 5, 6, 2), (9, 8, 5, 6, 2 | 32), (12, 9, 3, 5, 2), (11, 9, 3, 5, 2),
                                               (10, 8, 4, 6, 2)])
def test_four_step_passes(sim, lg1, lg2, u1, u2, tma):
    """N = N1*N2: column pass (+ exp(-2 pi i k1 n2/N)) then row pass with transposed store."""
    N1, N2 = 1 << lg1, 1 << lg2
    N = N1 * N2
    rng = np.random.default_rng(5)
    re, im = rng.standard_normal(N), rng.standard_normal(N)
    t_re, t_im, o_re, o_im = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
    U1, U2 = 1 << u1, 1 << u2
    conf = (ctypes.c_int * 4)()
    st = (ctypes.c_int64 * 9)(0, N2, 0, N2, 0, U1, 0, U1, 1 << 30)
    rc1 = sim.plansim_run(lg1, u1, 1 | tma, 1, st, lg1 + lg2, N2 // U1, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                          t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp), 0, conf)
    c1 = list(conf)[:3]
    st = (ctypes.c_int64 * 9)(N2, 1, 0, N1, 0, U2 * N2, 0, U2, 1 << 30)
    # pass 2 (contiguous rows in, transposed out): cp.async chunks or, with tma, row tiles (SWIZZLE_32B / 128B atoms)
    rc2 = sim.plansim_run(lg2, u2, tma & 2, 1, st, 0, N1 // U2, t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp),
                          o_re.ctypes.data_as(dp), o_im.ctypes.data_as(dp), 0, conf)
    want = np.fft.fft(re + 1j * im) / N
    assert rc1 == 0 and rc2 == 0 and c1 == [0, 0, 0] and list(conf)[:3] == [0, 0, 0]   # tma: column tiles loaded by TMA
    assert np.linalg.norm(o_re + 1j * o_im - want) / np.linalg.norm(want) < 1e-13


def test_ring_units(sim):
    """Landing-ring units (flag 16; fft_unit_kernel_ring): N = 32768 in one unit, and the two 4096-point passes of
    2^24 = 4096 x 4096 (8 columns in place with dense staging; 8 rows per unit from row tiles, stored transposed)."""
    rc, conf, err = _rows(sim, 15, 0, flags=2 | 4 | 16)
    assert rc == 0 and conf == [0, 0, 0] and err < 1e-13
    lg1 = lg2 = 12
    N1, N2 = 1 << lg1, 1 << lg2
    N = N1 * N2
    rng = np.random.default_rng(7)
    # only a few units of each pass are simulated (a full 2^24 pass is slow on the CPU): 2 column units, 2 row units
    re, im = rng.standard_normal(N), rng.standard_normal(N)
    t_re, t_im = re.copy(), im.copy()
    conf = (ctypes.c_int * 4)()
    st = (ctypes.c_int64 * 9)(0, N2, 0, N2, 0, 8, 0, 8, 1 << 30)
    rc1 = sim.plansim_run(lg1, 3, 1 | 2 | 16, 1, st, lg1 + lg2, 2, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                          t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp), 0, conf)
    assert rc1 == 0 and list(conf)[:3] == [0, 0, 0]
    x = (re + 1j * im).reshape(N1, N2)[:, :16]
    k1, n2 = np.arange(N1)[:, None], np.arange(16)[None, :]
    want = np.fft.fft(x, axis=0) / N1 * np.exp(-2j * np.pi * k1 * n2 / N)
    got = (t_re + 1j * t_im).reshape(N1, N2)[:, :16]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13
    o_re, o_im = np.zeros(N), np.zeros(N)
    st = (ctypes.c_int64 * 9)(N2, 1, 0, N1, 0, 8 * N2, 0, 8, 1 << 30)
    rc2 = sim.plansim_run(lg2, 3, 2 | 16, 1, st, 0, 2, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                          o_re.ctypes.data_as(dp), o_im.ctypes.data_as(dp), 0, conf)
    assert rc2 == 0 and list(conf)[:3] == [0, 0, 0]
    want = (np.fft.fft((re + 1j * im).reshape(N1, N2)[:16], axis=1) / N2).T      # X[k1 + N1*k2], k1 < 16
    got = (o_re + 1j * o_im).reshape(N2, N1)[:, :16]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13


@pytest.mark.parametrize("lgy,lgx,yb,flags", [(9, 12, 1, 6), (10, 11, 2, 0)])
def test_two_d_passes(sim, lgy, lgx, yb, flags):
    """2-D ny x nx: row pass on Kronecker units (U = 2^yb rows y_lo + u*ny/U, last tensor stage = F_x (x) F_y, row
    twiddle exp(-2 pi i k_y y_lo/ny)) then in-place column pass of length ny/U; equals fft2/(ny*nx)."""
    sim.plansim_run_ex.argtypes = list(sim.plansim_run.argtypes) + [ctypes.POINTER(ctypes.c_int64)]
    ny, nx = 1 << lgy, 1 << lgx
    rng = np.random.default_rng(11)
    re, im = rng.standard_normal(ny * nx), rng.standard_normal(ny * nx)
    t_re, t_im = np.zeros(ny * nx), np.zeros(ny * nx)
    conf = (ctypes.c_int * 4)()
    if yb:
        U, ups = 1 << yb, yb
        st = (ctypes.c_int64 * 9)((ny >> yb) * nx, 1, nx, 1, 0, nx, 0, U * nx, ny // U)
    else:
        ups = 14 - lgx
        U = 1 << ups
        st = (ctypes.c_int64 * 9)(nx, 1, nx, 1, 0, U * nx, 0, U * nx, ny // U)
    ext = (ctypes.c_int64 * 5)(yb, lgy if yb else 0, 1, 0, 0)
    rc1 = sim.plansim_run_ex(lgx, ups, flags, 0, st, 0, ny // U, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                             t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp), 0, conf, ext)
    c1 = list(conf)[:3]
    lg2 = lgy - yb
    u2 = max(3, 14 - lg2)
    U2 = 1 << u2
    st = (ctypes.c_int64 * 9)(0, nx << yb, 0, nx << yb, 0, U2, 0, U2, 1 << 30)
    o_re, o_im = t_re.copy(), t_im.copy()
    rc2 = sim.plansim_run(lg2, u2, 1, 1, st, 0, (nx << yb) // U2, t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp),
                          o_re.ctypes.data_as(dp), o_im.ctypes.data_as(dp), 0, conf)
    want = np.fft.fft2((re + 1j * im).reshape(ny, nx)) / (ny * nx)
    got = (o_re + 1j * o_im).reshape(ny, nx)
    assert rc1 == 0 and rc2 == 0 and c1 == [0, 0, 0] and list(conf)[:3] == [0, 0, 0]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13


@pytest.mark.parametrize("lgy,lgx,yb,flags", [(13, 13, 1, 6), (13, 13, 2, 6), (13, 13, 1, 0), (14, 12, 2, 2), (13, 11, 2, 0)])
def test_kronecker_units_of_large_images(sim, lgy, lgx, yb, flags):
    """Two units of the row pass of a big image (C5 is 8192 x 8192) against the direct formula
    Z[k_y][k_x] = exp(-2 pi i k_y y_lo/ny) / (U nx) * sum_u sum_x in[y_lo + u ny/U][x] exp(-2 pi i (x k_x/nx + u k_y/U))."""
    sim.plansim_run_ex.argtypes = list(sim.plansim_run.argtypes) + [ctypes.POINTER(ctypes.c_int64)]
    ny, nx, U = 1 << lgy, 1 << lgx, 1 << yb
    rng = np.random.default_rng(3)
    y_los = [5, 6]   # units 5 and 6 of the image
    rows = {(yl, u): rng.standard_normal(nx) + 1j * rng.standard_normal(nx) for yl in y_los for u in range(U)}
    span = (ny // U) * nx
    # sparse image: only the rows the two units touch are materialised, at offset (y - 5) * nx in a compact buffer
    comp = {}
    size_in = (U - 1) * span + (y_los[-1] - y_los[0] + 1) * nx
    re, im = np.zeros(size_in), np.zeros(size_in)
    for (yl, u), v in rows.items():
        o = u * span + (yl - y_los[0]) * nx
        re[o:o + nx], im[o:o + nx] = v.real, v.imag
    o_re, o_im = np.zeros(2 * U * nx), np.zeros(2 * U * nx)
    conf = (ctypes.c_int * 4)()
    st = (ctypes.c_int64 * 9)(span, 1, nx, 1, 0, nx, 0, U * nx, ny // U)
    ext = (ctypes.c_int64 * 5)(yb, lgy, 1, 0, 0)
    # the simulator numbers units from 0: shift so that unit index == y_lo by passing pointers moved back by 5 units
    shift_in, shift_out = y_los[0] * nx, y_los[0] * U * nx
    full_re, full_im = np.zeros(shift_in + size_in), np.zeros(shift_in + size_in)
    full_re[shift_in:], full_im[shift_in:] = re, im
    full_ore, full_oim = np.zeros(shift_out + 2 * U * nx), np.zeros(shift_out + 2 * U * nx)
    rc = sim.plansim_run_ex(lgx, yb, flags, 0, st, 0, y_los[-1] + 1, full_re.ctypes.data_as(dp), full_im.ctypes.data_as(dp),
                            full_ore.ctypes.data_as(dp), full_oim.ctypes.data_as(dp), 0, conf, ext)
    assert rc == 0 and list(conf)[:3] == [0, 0, 0]
    got = (full_ore + 1j * full_oim)[shift_out:].reshape(2, U, nx)
    for i, yl in enumerate(y_los):
        x = np.stack([rows[(yl, u)] for u in range(U)])
        want = np.fft.fft2(x) / (U * nx) * np.exp(-2j * np.pi * np.arange(U) * yl / ny)[:, None]
        assert np.linalg.norm(got[i] - want) / np.linalg.norm(want) < 1e-13


@pytest.mark.parametrize("lgy,lgx,yb,flags", [(11, 12, 1, 6)])
def test_two_d_passes_with_tiled_intermediate(sim, lgy, lgx, yb, flags):
    """Same 2-D transform, but the row pass writes the intermediate in the column units' operand order
    [k_y][x/8][y_lo][x%8] (one contiguous chunk per column unit) and the column pass reads it with a 16-byte row
    stride and writes the natural layout."""
    sim.plansim_run_ex.argtypes = list(sim.plansim_run.argtypes) + [ctypes.POINTER(ctypes.c_int64)]
    ny, nx, U = 1 << lgy, 1 << lgx, 1 << yb
    n = ny * nx
    rng = np.random.default_rng(12)
    re, im = rng.standard_normal(n), rng.standard_normal(n)
    t_re, t_im, o_re, o_im = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    conf = (ctypes.c_int * 4)()
    rows2 = ny >> yb                                            # column-pass length
    st = (ctypes.c_int64 * 9)(rows2 * nx, 1, rows2 * nx, 1, 0, nx, 0, 8, ny // U)   # out: k_y stride, unit (y_lo) stride 8
    ext = (ctypes.c_int64 * 5)(yb, lgy, 1, 3, rows2 * 8)
    rc1 = sim.plansim_run_ex(lgx, yb, flags, 0, st, 0, ny // U, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                             t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp), 0, conf, ext)
    c1 = list(conf)[:3]
    lg2, u2 = lgy - yb, 3
    # column units: 8 columns, input row stride 8 (contiguous chunk of rows2*8 elements per unit), natural output
    st = (ctypes.c_int64 * 9)(0, 8, 0, nx << yb, 0, rows2 * 8, 0, 8, 1 << 30)
    rc2 = sim.plansim_run(lg2, u2, 1, 1, st, 0, (nx << yb) // 8, t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp),
                          o_re.ctypes.data_as(dp), o_im.ctypes.data_as(dp), 0, conf)
    want = np.fft.fft2((re + 1j * im).reshape(ny, nx)) / n
    got = (o_re + 1j * o_im).reshape(ny, nx)
    assert rc1 == 0 and rc2 == 0 and c1 == [0, 0, 0] and list(conf)[:3] == [0, 0, 0]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13


@pytest.mark.parametrize("flags", [8, 10])
def test_cluster_unit_single_pass_65536(sim, flags):
    """N = 65536 in ONE pass: a unit of 2^16 elements shared by a CTA pair (flag 8).  CTA r loads the rows with m-half r,
    its stage-1 epilogue stores the outputs with top k_1 bit d into CTA d, stages 2-3 and the store are local.
    flags: 8 = 16-byte copies, 10 = + TMA row tiles."""
    rc, conf, err = _rows(sim, 16, 0, n_units=2, flags=flags)
    assert rc == 0 and conf == [0, 0, 0]
    assert err < 1e-13


def test_cluster_unit_predicted_fp16_error(sim):
    rc, _, err = _rows(sim, 16, 0, n_units=1, h=1, flags=10)
    assert rc == 0 and err < 6e-4


@pytest.mark.parametrize("lg1,lg2,tma", [(12, 8, 2), (12, 8, 0), (12, 10, 2)])
def test_four_step_with_cluster_column_units(sim, lg1, lg2, tma):
    """Four-step N = N1*N2 whose column pass (length N1 = 4096) runs on cluster units of 16 columns (32-byte pieces):
    column pass with the fused twiddle on CTA pairs, then the ordinary row pass with transposed store."""
    N1, N2 = 1 << lg1, 1 << lg2
    N = N1 * N2
    rng = np.random.default_rng(6)
    re, im = rng.standard_normal(N), rng.standard_normal(N)
    t_re, t_im, o_re, o_im = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
    u1, U1 = 4, 16
    conf = (ctypes.c_int * 4)()
    st = (ctypes.c_int64 * 9)(0, N2, 0, N2, 0, U1, 0, U1, 1 << 30)
    rc1 = sim.plansim_run(lg1, u1, 1 | tma | 8, 1, st, lg1 + lg2, N2 // U1, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                          t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp), 0, conf)
    c1 = list(conf)[:3]
    u2 = 14 - lg2
    U2 = 1 << u2
    st = (ctypes.c_int64 * 9)(N2, 1, 0, N1, 0, U2 * N2, 0, U2, 1 << 30)
    rc2 = sim.plansim_run(lg2, u2, 0, 1, st, 0, N1 // U2, t_re.ctypes.data_as(dp), t_im.ctypes.data_as(dp),
                          o_re.ctypes.data_as(dp), o_im.ctypes.data_as(dp), 0, conf)
    want = np.fft.fft(re + 1j * im) / N
    assert rc1 == 0 and rc2 == 0 and c1 == [0, 0, 0]
    assert np.linalg.norm(o_re + 1j * o_im - want) / np.linalg.norm(want) < 1e-13


@pytest.mark.parametrize("lg,ups,bound", [(8, 6, 0), (8, 5, 0), (9, 5, 0), (10, 4, 96), (11, 3, 96), (12, 2, 128), (12, 1, 64)])
def test_twiddle_table_lookups_are_modelled(sim, lg, ups, bound):
    """The only shared-memory accesses besides the 16-byte operand stores and staging loads are the LDS.64 lookups of the
    two-level twiddle table in the epilogue of 2-stage plans (3-stage plans keep their seeds in registers).  The simulator
    counts their EXCESS wavefronts per unit (conflicts[3]): none for a radix-16 first stage; for radix 32 / 64 the second
    lookup (x * 16 g) folds lanes onto the same banks -- bounded, and measured on hardware as the only instructions with
    `L1 Wavefronts Shared Excessive` > 0 (profiles/r02_ncu_c2.txt discussion in DESIGN.md 3)."""
    n, U = 1 << lg, 1 << ups
    tot = U * n
    rng = np.random.default_rng(1)
    re, im = rng.standard_normal(tot), rng.standard_normal(tot)
    ore, oim = np.zeros(tot), np.zeros(tot)
    st = (ctypes.c_int64 * 9)(n, 1, n, 1, 0, U * n, 0, U * n, 1 << 30)
    conf = (ctypes.c_int * 4)()
    rc = sim.plansim_run(lg, ups, 2, 0, st, 0, 1, re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                         ore.ctypes.data_as(dp), oim.ctypes.data_as(dp), 0, conf)
    assert rc == 0 and list(conf)[:3] == [0, 0, 0]
    assert 0 <= conf[3] <= bound, conf[3]
    if bound == 0:
        assert conf[3] == 0
