"""CPU: the C-ABI library loads and exports every symbol include/tfft.h declares; plan logic and the
reference-interface mirror behave like the reference's host code; compute calls fail loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import tfft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "tfft.h")).read()
    declared = set(re.findall(r"\b(tfft_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(tfft.EXPORTS)
    L = tfft.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.tfft_version() >= 100


def test_plan_info_matches_reference_plan_table():
    # r16 = log2N/4 - 1, r2 = log2N % 4 (src/base/Plan.h:99-100)
    for lg, r16, r2 in [(8, 1, 0), (9, 1, 1), (12, 2, 0), (14, 2, 2), (15, 2, 3), (16, 3, 0), (20, 4, 0), (24, 5, 0), (25, 5, 1), (28, 6, 0)]:
        p = tfft.NativePlan(1 << lg, 2)
        assert (p.info["amount_of_r16_steps"], p.info["amount_of_r2_steps"]) == (r16, r2)
        assert p.info["results_in_results"] == 1
        assert p.info["passes"] == (1 if lg <= 15 else 2 if lg <= 23 else 3)   # three passes of 256 from 2^24 on
        assert p.info["algorithmic_bytes"] == 8 * (1 << lg) * 2 * p.info["passes"]
        assert p.info["smem_bytes"] <= 227 * 1024 and p.info["tmem_columns"] <= 512
        p.close()


def test_config2_plan_shape():
    p = tfft.NativePlan(16384, 4096)
    # 16384 = 16 * 32 * 32: three tensor-core stages, the non-power-of-16 factor 4 is folded into them
    assert p.info["tail_radix"] == 2 and p.info["r16_stages"] == 3 and p.info["grid"] == 4096
    assert p.info["algorithmic_bytes"] == 536870912


@pytest.mark.parametrize("n", [0, 100, 128, 255, 3 << 10, 1 << 31])
def test_invalid_sizes_are_rejected(n):
    with pytest.raises(tfft.TfftError):
        tfft.NativePlan(n, 1)


def test_create_plan_mirror_matches_reference_rules(capsys):
    assert tfft.create_plan(1000) is None                       # Plan.h:85-88
    assert "power of 2" in capsys.readouterr().out
    assert tfft.create_plan(128) is None                        # Plan.h:92-96
    assert tfft.create_plan(2048, tfft.MODE_4096) is None       # Plan.h:102-106
    p = tfft.create_plan(16384)
    assert (p.amount_of_r16_steps_, p.amount_of_r2_steps_) == (2, 2)
    assert p.base_fft_gridsize_ == 8 and p.base_fft_shared_mem_in_bytes_ == 16384   # Appendix B row 14
    assert p.results_in_results_ is True
    p = tfft.create_plan(4096, tfft.MODE_4096, 16, 16, 512)     # ExampleSingleFFT.cu call
    assert p.base_fft_warps_per_block_ == 16 and p.base_fft_gridsize_ == 1
    assert tfft.create_plan(512, tfft.MODE_256, 8, 8, 256).base_fft_warps_per_block_ == 2   # clamped, Plan.h:119-127
    assert tfft.create_plan(1 << 14, tfft.MODE_256, 8, 8, 8192) is None             # Plan.h:178-190


def test_create_plan_from_tuner_file(tmp_path):
    f = tmp_path / "TunerResults.dat"
    f.write_text("4096 4096 16 16 512\n16384 256 8 4 256\n")     # FileWriter.h:250-269 line format
    p = tfft.create_plan_from_file(16384, str(f))
    assert p.base_fft_mode_ == tfft.MODE_256 and p.r16_warps_per_block_ == 4
    assert tfft.create_plan_from_file(8192, str(f)) is None
    assert tfft.create_plan_from_file(8192, str(tmp_path / "missing.dat")) is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_exec_fails_loudly_without_gpu():
    p = tfft.NativePlan(4096, 1)
    x = torch.zeros(8192, dtype=torch.float16)
    with pytest.raises(tfft.TfftError):
        p.exec(x, x[4096:], x, x[4096:], 8192, 8192)
    host = np.zeros(8192, dtype=np.float16)
    with pytest.raises(tfft.TfftError):
        p.exec_host(host, host.copy())


def test_misaligned_arguments_rejected():
    L = tfft.lib()
    p = tfft.NativePlan(4096, 2)
    rc = L.tfft_exec(p._h, 16, 32, 48, 64, 4100, 8192, None)     # stride not a multiple of 8
    assert rc == -2
    rc = L.tfft_exec(p._h, 2, 32, 48, 64, 8192, 8192, None)      # misaligned pointer
    assert rc == -2
    assert b"invalid argument" in L.tfft_error_string(-2)


def test_tuner_file_plan_creation(tmp_path):
    """tfft_plan_create_from_file (CreatePlan(N, tuner_file), Plan.h:197-255): the reference's 5-column lines are
    accepted, the key=value knobs of the new kernels are applied, a missing length is TFFT_E_NOT_IN_FILE (-6)."""
    import ctypes
    f = tmp_path / "TunerResults.dat"
    f.write_text("4096 256 8 8 256\n1048576 4096 16 16 512 lg1=8 tma_col=0 prefetch=1\n16384 256 8 8 256 two_slot=0 pipe=0\n")
    L = tfft.lib()
    for n, passes in ((4096, 1), (1048576, 2), (16384, 1)):
        p = tfft.NativePlan(n, 4, tuner_file=str(f))
        assert p.info["passes"] == passes and p.info["n"] == n
    h = ctypes.c_void_p()
    assert L.tfft_plan_create_from_file(ctypes.byref(h), 8192, 1, 0, str(f).encode()) == -6
    assert L.tfft_plan_create_from_file(ctypes.byref(h), 8192, 1, 0, b"/nonexistent/file") == -2
    assert b"tuner file" in L.tfft_error_string(-6)


def test_multi_gpu_and_segmented_entry_points_fail_loudly_without_gpu():
    """The round-2 entry points (tfft_mg_*, tfft_exec_segmented, tfft_plan_prepare) validate their arguments on the host and
    answer with an error -- never a CPU result -- when there is no device."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    L = tfft.lib()
    h = ctypes.c_void_p()
    assert L.tfft_mg_plan_create(ctypes.byref(h), 1 << 20, 0, 3, 0) == -2          # world not a power of two
    assert L.tfft_mg_plan_create(ctypes.byref(h), 1 << 20, 2, 2, 0) == -2          # rank out of range
    assert L.tfft_mg_plan_create(ctypes.byref(h), 1 << 12, 0, 2, 0) == -1          # too short to shard
    assert L.tfft_mg_plan_create(ctypes.byref(h), 1 << 20, 0, 2, 0) == -3          # TFFT_E_NO_DEVICE
    assert not h.value
    p = tfft.NativePlan(16384, 4)
    with pytest.raises(tfft.TfftError):
        p.prepare()                                                                 # needs the device
    buf = (ctypes.c_uint16 * 64)()
    a = ctypes.addressof(buf)
    a += (-a) % 16
    rc = L.tfft_exec_segmented(p._h, a, a, a, a, 2048, 16384, 8, 8192, 0, 0, None)
    assert rc == -3                                                                 # arguments fine, no device
    assert L.tfft_exec_segmented(p._h, a, a, a, a, 2048, 16384, 0, 8192, 0, 0, None) == -2   # segments < 1
    assert L.tfft_exec_segmented(p._h, a + 2, a, a, a, 2048, 16384, 8, 8192, 0, 0, None) == -2   # misaligned
    assert b"barrier" in L.tfft_error_string(-7)
    p.close()


def test_cluster_tuner_key_builds_a_single_pass_plan(tmp_path):
    f = tmp_path / "TunerResults.dat"
    f.write_text("65536 256 8 8 256 cluster=1\n4194304 256 8 8 256 cluster=1 lg1=12\n")
    p = tfft.NativePlan(65536, 7, tuner_file=str(f))
    assert p.info["passes"] == 1 and p.info["algorithmic_bytes"] == 8 * 65536 * 7      # one HBM pass on CTA-pair units
    assert p.info["smem_bytes"] <= 227 * 1024 and p.info["tmem_columns"] == 512
    q = tfft.NativePlan(65536, 7)
    assert q.info["passes"] == 2                                                        # default: four-step
    r = tfft.NativePlan(1 << 22, 1, tuner_file=str(f))
    assert r.info["passes"] == 2 and r.info["transforms_per_cta"] == 16                 # 16 columns per CTA pair
    for x in (p, q, r):
        x.close()


def test_round2_second_session_tuner_keys(tmp_path):
    """ring (landing-ring kernel) and tma_col=2 (no 64-column tiles) are accepted; a four-step split or cluster=1 keeps 2^24
    on the two-pass plan, the default is three passes; every plan stays inside the 227 KiB / 512 columns of an SM."""
    f = tmp_path / "TunerResults.dat"
    f.write_text("4194304 256 8 8 256 ring=0\n32768 256 8 8 256 ring=2\n65536 256 8 8 256 tma_col=2\n16777216 256 8 8 256 lg1=12\n")
    for n, passes in ((1 << 22, 2), (32768, 1), (65536, 2), (1 << 24, 2)):
        p = tfft.NativePlan(n, 3, tuner_file=str(f))
        assert p.info["passes"] == passes, n
        assert p.info["smem_bytes"] <= 227 * 1024 and p.info["tmem_columns"] <= 512
        p.close()
    for n, passes in ((1 << 22, 2), (1 << 24, 3), (1 << 26, 3)):
        p = tfft.NativePlan(n, 3)
        assert p.info["passes"] == passes and p.info["smem_bytes"] <= 227 * 1024
        p.close()
