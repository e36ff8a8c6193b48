"""tfft -- Python host side of the B200-native fp16 FFT (product code).

A thin ctypes binding of the C ABI in include/tfft.h plus a mirror of the reference's host
interface (same names and call sequence as /root/reference/src/base):

    plan = create_plan(n)                       # CreatePlan            Plan.h:77-194
    plan_works_on_device(plan, 0)               # PlanWorksOnDevice     Plan.h:257-296
    h = DataHandler(n)  / DataBatchHandler(n, b)  #                     DataHandler.h:22-166
    h.copy_data_host_to_device(host)            # CopyDataHostToDevice  DataHandler.h:45-53
    err = compute_fft(plan, h)                  # ComputeFFT            ComputeFFT.h:54-293
    h.copy_results_device_to_host(out, plan.results_in_results_)

PyTorch is used only for device memory and streams.  There is no CPU fallback: importing works
anywhere (so plan logic can be tested on CPU) but every compute call raises if the CUDA
library or a GPU is missing.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("TFFT_LIB", os.path.join(_HERE, "libtfft.so"))   # TFFT_LIB: developer override
_lib = None

MODE_256 = 0      # BaseFFTMode::Mode_256   (Plan.h:14)
MODE_4096 = 1     # BaseFFTMode::Mode_4096

TFFT_PRESERVE_INPUT = 1
TFFT_INVERSE = 2      # exp(+2 pi i n k / N)
TFFT_UNSCALED = 4     # no 1/N (cuFFT convention)
TFFT_INTERLEAVED = 8  # half2 (re, im) arrays instead of planes


class TfftError(RuntimeError):
    pass


class _PlanInfo(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int64), ("batch", ctypes.c_int64), ("r16_stages", ctypes.c_int32),
        ("tail_radix", ctypes.c_int32), ("passes", ctypes.c_int32), ("results_in_results", ctypes.c_int32),
        ("amount_of_r16_steps", ctypes.c_int32), ("amount_of_r2_steps", ctypes.c_int32),
        ("transforms_per_cta", ctypes.c_int32), ("smem_bytes", ctypes.c_int32), ("tmem_columns", ctypes.c_int32),
        ("grid", ctypes.c_int64), ("workspace_bytes", ctypes.c_int64), ("algorithmic_bytes", ctypes.c_int64),
    ]


EXPORTS = ("tfft_plan_create", "tfft_plan_create_2d", "tfft_plan_info", "tfft_plan_destroy", "tfft_exec",
           "tfft_exec_twiddled", "tfft_exec_host", "tfft_error_string", "tfft_version", "tfft_fixture_sine",
           "tfft_error_stats", "tfft_transpose_blocks", "tfft_copy_runs", "tfft_plan_create_from_file",
           "tfft_mg_plan_create", "tfft_mg_plan_handle", "tfft_mg_plan_connect", "tfft_mg_plan_info", "tfft_mg_exec",
           "tfft_mg_status", "tfft_mg_set_timeout_ms", "tfft_mg_plan_destroy", "tfft_plan_prepare", "tfft_mg_exec_phase", "tfft_exec_segmented")

TFFT_MG_HANDLE_BYTES = 128


class _MgInfo(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("n1", ctypes.c_int64), ("n2", ctypes.c_int64), ("local_elems", ctypes.c_int64),
                ("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("exchanges", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("exchange_bytes_per_rank", ctypes.c_int64),
                ("device_bytes", ctypes.c_int64), ("result_re", ctypes.c_void_p), ("result_im", ctypes.c_void_p)]


def lib() -> ctypes.CDLL:
    """Load tensor-fft_b200/tfft/libtfft.so (built by __graft_entry__.build()). Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise TfftError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = ctypes.CDLL(_LIB_PATH)
        vp, i64, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32
        L.tfft_plan_create.argtypes = [ctypes.POINTER(vp), i64, i64, u32]
        L.tfft_plan_create_2d.argtypes = [ctypes.POINTER(vp), i64, i64, i64, u32]
        L.tfft_plan_create_from_file.argtypes = [ctypes.POINTER(vp), i64, i64, u32, ctypes.c_char_p]
        L.tfft_plan_info.argtypes = [vp, ctypes.POINTER(_PlanInfo)]
        L.tfft_plan_destroy.argtypes = [vp]
        L.tfft_plan_prepare.argtypes = [vp]
        L.tfft_exec.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp]
        L.tfft_exec_twiddled.argtypes = [vp, vp, vp, vp, vp, i64, i64, ctypes.c_int32, i64, vp]
        L.tfft_exec_host.argtypes = [vp, vp, vp]
        L.tfft_exec_segmented.argtypes = [vp, vp, vp, vp, vp, i64, i64, ctypes.c_int32, i64, ctypes.c_int32, i64, vp]
        fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
        L.tfft_fixture_sine.argtypes = [vp, vp, i64, i64, i64, fp, fp, ctypes.c_int32, vp]
        L.tfft_error_stats.argtypes = [vp, vp, vp, vp, i64, dp, vp]
        L.tfft_transpose_blocks.argtypes = [vp, vp] + [i64] * 10 + [vp]
        L.tfft_copy_runs.argtypes = [vp, vp] + [i64] * 10 + [vp]
        L.tfft_mg_plan_create.argtypes = [ctypes.POINTER(vp), i64, ctypes.c_int32, ctypes.c_int32, u32]
        L.tfft_mg_plan_handle.argtypes = [vp, vp]
        L.tfft_mg_plan_connect.argtypes = [vp, vp]
        L.tfft_mg_plan_info.argtypes = [vp, ctypes.POINTER(_MgInfo)]
        L.tfft_mg_exec.argtypes = [vp, vp, vp, vp, vp, vp]
        L.tfft_mg_status.argtypes = [vp]
        L.tfft_mg_exec_phase.argtypes = [vp, ctypes.c_int32, vp, vp, vp, vp, vp]
        L.tfft_mg_set_timeout_ms.argtypes = [vp, i64]
        L.tfft_mg_plan_destroy.argtypes = [vp]
        L.tfft_error_string.argtypes = [ctypes.c_int]
        L.tfft_error_string.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise TfftError(f"tfft error {rc}: {lib().tfft_error_string(rc).decode()}")


def fixture_sine(re, im, n: int, weights_re, weights_im, stride: Optional[int] = None) -> None:
    """tfft_fixture_sine: fill fp16 CUDA tensors with the reference's sine superposition
    (TestingDataCreation.h:89-117). weights_*: float32 numpy arrays (batch, cutoff)."""
    import numpy as np
    import torch
    wr = np.ascontiguousarray(np.atleast_2d(weights_re), dtype=np.float32)
    wi = np.ascontiguousarray(np.atleast_2d(weights_im), dtype=np.float32)
    for t in (re, im):
        if not (t.is_cuda and t.dtype == torch.float16):
            raise TfftError("fixture_sine needs CUDA float16 tensors (no CPU fallback)")
    fp = ctypes.POINTER(ctypes.c_float)
    _check(lib().tfft_fixture_sine(re.data_ptr(), im.data_ptr(), n, wr.shape[0], stride or n, wr.ctypes.data_as(fp),
                                   wi.ctypes.data_as(fp), wr.shape[1],
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def error_stats(a_re, a_im, b_re, b_im) -> dict:
    """tfft_error_stats: fp16 CUDA result planes vs float64 CUDA reference planes ->
    {max, avg, sigma, rel_l2} (AccuracyCalculator.h:86-148 on the device)."""
    import torch
    if not (a_re.is_cuda and a_re.dtype == torch.float16 and b_re.is_cuda and b_re.dtype == torch.float64):
        raise TfftError("error_stats needs CUDA float16 results and CUDA float64 reference values")
    out = (ctypes.c_double * 4)()
    _check(lib().tfft_error_stats(a_re.data_ptr(), a_im.data_ptr(), b_re.data_ptr(), b_im.data_ptr(), a_re.numel(),
                                  out, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return {"max": out[0], "avg": out[1], "sigma": out[2], "rel_l2": out[3]}


class NativePlan:
    """Owner of a tfft_plan_t."""

    def __init__(self, n: int, batch: int = 1, flags: int = 0, shape2d: Optional[tuple] = None,
                 tuner_file: Optional[str] = None):
        self._h = ctypes.c_void_p()
        if tuner_file is not None:
            _check(lib().tfft_plan_create_from_file(ctypes.byref(self._h), n, batch, flags, tuner_file.encode()))
        elif shape2d is None:
            _check(lib().tfft_plan_create(ctypes.byref(self._h), n, batch, flags))
        else:
            _check(lib().tfft_plan_create_2d(ctypes.byref(self._h), shape2d[0], shape2d[1], batch, flags))
        info = _PlanInfo()
        _check(lib().tfft_plan_info(self._h, ctypes.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in _PlanInfo._fields_}
        self.n, self.batch = n, batch

    def exec(self, in_re, in_im, out_re, out_im, in_stride: int, out_stride: int, stream=None) -> None:
        """Launch on torch CUDA tensors (fp16). Strides in elements between consecutive transforms."""
        import torch
        for t in (in_re, in_im, out_re, out_im):
            if not (t.is_cuda and t.dtype == torch.float16):
                raise TfftError("tfft exec needs CUDA float16 tensors (no CPU fallback)")
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _check(lib().tfft_exec(self._h, in_re.data_ptr(), in_im.data_ptr(), out_re.data_ptr(), out_im.data_ptr(),
                               in_stride, out_stride, ctypes.c_void_p(s)))

    def exec_twiddled(self, in_re, in_im, out_re, out_im, in_stride: int, out_stride: int, log2_total: int,
                      first_col: int, stream=None) -> None:
        """exec + output k of transform b times exp(-2 pi i k (first_col + b) / 2^log2_total)."""
        import torch
        for t in (in_re, in_im, out_re, out_im):
            if not (t.is_cuda and t.dtype == torch.float16):
                raise TfftError("tfft exec needs CUDA float16 tensors (no CPU fallback)")
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _check(lib().tfft_exec_twiddled(self._h, in_re.data_ptr(), in_im.data_ptr(), out_re.data_ptr(),
                                        out_im.data_ptr(), in_stride, out_stride, log2_total, first_col,
                                        ctypes.c_void_p(s)))

    def exec_segmented(self, in_re, in_im, out_re, out_im, in_stride: int, out_stride: int, segments: int,
                       segment_stride: int, log2_total: int = 0, first_col: int = 0, stream=None) -> None:
        """tfft_exec_segmented: piece q of transform b at in + q*segment_stride + b*in_stride (gathered by the TMA load)."""
        import torch
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _check(lib().tfft_exec_segmented(self._h, in_re.data_ptr(), in_im.data_ptr(), out_re.data_ptr(), out_im.data_ptr(),
                                         in_stride, out_stride, segments, segment_stride, log2_total, first_col,
                                         ctypes.c_void_p(s)))

    def prepare(self) -> None:
        """tfft_plan_prepare: all lazy one-time device initialisation now (before graph capture / spin-waiting peers)."""
        _check(lib().tfft_plan_prepare(self._h))

    def exec_host(self, host_in, host_out) -> None:
        """numpy float16 arrays of 2*n*batch values laid out [RE|IM] per transform."""
        _check(lib().tfft_exec_host(self._h, host_in.ctypes.data_as(ctypes.c_void_p),
                                    host_out.ctypes.data_as(ctypes.c_void_p)))

    def close(self) -> None:
        if self._h:
            lib().tfft_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevPlane:
    """Zero-copy view of a plan-owned device plane for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f2", "data": (ptr, False), "version": 3}


class MgPlan:
    """Owner of a tfft_mg_plan_t: ONE transform of length n sharded over `world` GPUs (include/tfft.h, tfft_mg_*).

    The exchange of the IPC handles is the caller's plumbing: `connect_with(gather)` takes a function
    bytes -> list[bytes] (an all-gather over the ranks, e.g. torch.distributed.all_gather_object);
    `connect_local(plans)` links ranks that live in one process (tests: several ranks on one GPU)."""

    def __init__(self, n: int, rank: int, world: int):
        self._h = ctypes.c_void_p()
        _check(lib().tfft_mg_plan_create(ctypes.byref(self._h), n, rank, world, 0))
        info = _MgInfo()
        _check(lib().tfft_mg_plan_info(self._h, ctypes.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in _MgInfo._fields_}
        self.n, self.rank, self.world = n, rank, world

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(TFFT_MG_HANDLE_BYTES)
        _check(lib().tfft_mg_plan_handle(self._h, buf))
        return buf.raw

    def connect(self, handles) -> None:
        blob = b"".join(handles)
        if len(blob) != self.world * TFFT_MG_HANDLE_BYTES:
            raise TfftError("connect needs one handle per rank")
        _check(lib().tfft_mg_plan_connect(self._h, ctypes.c_char_p(blob)))

    def connect_with(self, gather) -> None:
        self.connect(gather(self.handle()))

    @staticmethod
    def connect_local(plans) -> None:
        hs = [p.handle() for p in plans]
        for p in plans:
            p.connect(hs)

    def exec(self, in_re, in_im, out_re=None, out_im=None, stream=None) -> None:
        import torch
        for t in (in_re, in_im) + ((out_re, out_im) if out_re is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float16 and t.numel() >= self.info["local_elems"]):
                raise TfftError("tfft mg exec needs CUDA float16 tensors of n/world elements (no CPU fallback)")
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _check(lib().tfft_mg_exec(self._h, in_re.data_ptr(), in_im.data_ptr(),
                                  out_re.data_ptr() if out_re is not None else None,
                                  out_im.data_ptr() if out_im is not None else None, ctypes.c_void_p(s)))

    def exec_phase(self, phase: int, in_re=None, in_im=None, out_re=None, out_im=None, stream=None) -> None:
        """tfft_mg_exec_phase: one phase without its flag barrier (the caller orders the ranks)."""
        import torch
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        _check(lib().tfft_mg_exec_phase(self._h, phase, ptr(in_re), ptr(in_im), ptr(out_re), ptr(out_im),
                                        ctypes.c_void_p(s)))

    def result(self):
        """Zero-copy torch views of the plan-owned result planes (valid until the next exec)."""
        import torch
        m = self.info["local_elems"]
        return (torch.as_tensor(_DevPlane(self.info["result_re"], m), device="cuda"),
                torch.as_tensor(_DevPlane(self.info["result_im"], m), device="cuda"))

    def status(self) -> None:
        _check(lib().tfft_mg_status(self._h))

    def set_timeout_ms(self, ms: int) -> None:
        _check(lib().tfft_mg_set_timeout_ms(self._h, ms))

    def close(self) -> None:
        if self._h:
            lib().tfft_mg_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------
# Mirror of the reference host interface (src/base/Plan.h, DataHandler.h, ComputeFFT.h)
# --------------------------------------------------------------------------------------------
@dataclass
class Plan:
    """Plan<Integer> (Plan.h:18-39). Launch-shape fields are informational for the new kernels."""
    fft_length_: int
    amount_of_r16_steps_: int
    amount_of_r2_steps_: int
    base_fft_mode_: int
    results_in_results_: bool
    base_fft_warps_per_block_: int
    base_fft_blocksize_: int
    base_fft_gridsize_: int
    base_fft_shared_mem_in_bytes_: int
    r16_warps_per_block_: int
    r16_blocksize_: int
    r16_gridsize_: int
    r16_shared_mem_in_bytes_: int
    r2_blocksize_: int


def is_power_of_2(x: int) -> bool:
    """IsPowerOf2 (Plan.h:41-47)."""
    return x != 0 and (x & (x - 1)) == 0


def exact_log2(x: int) -> int:
    """ExactLog2 (Plan.h:50-67), without the reference's 32-bit truncation."""
    return x.bit_length() - 1


def create_plan(fft_length: int, mode: int = MODE_256, base_fft_warps_per_block: int = 8,
                r16_warps_per_block: int = 8, r2_blocksize: int = 256) -> Optional[Plan]:
    """CreatePlan (Plan.h:77-194): same validation and error behaviour (message on stdout, None).

    The tuning arguments are validated like the reference validates them, so callers that pass
    reference tuner values keep working, but they do not steer the B200 kernels.
    """
    if not is_power_of_2(fft_length):
        print("Error! Input size has to be a power of 2!")
        return None
    lg = exact_log2(fft_length)
    if lg < 8:
        print("Error! Input size has to be larger than 256 i.e. 16^2")
        return None
    if mode == MODE_4096 and fft_length < 4096:
        print("Error! Baselayer fft length cant be longer that fft_length.")
        return None
    r16, r2 = lg // 4 - 1, lg % 4
    total_warps = fft_length // 256
    if total_warps < base_fft_warps_per_block:
        base_w = total_warps
    else:
        if total_warps % base_fft_warps_per_block != 0:
            print("Error! Total amount of warps (fft_length/256) has to be evenly devisable by "
                  "base_fft_warps_per_block.")
            return None
        base_w = 16 if mode == MODE_4096 else base_fft_warps_per_block
    if total_warps < r16_warps_per_block:
        r16_w = total_warps
    else:
        if total_warps % r16_warps_per_block != 0:
            print("Error! Total amount of warps (fft_length/256) has to be evenly devisable by "
                  "amount_of_r16_warps_per_block.")
            return None
        r16_w = r16_warps_per_block
    if (fft_length >> r2) % r2_blocksize != 0:
        print("Error! smallest_r2_subfft_length i.e. pow(2,(log2_of_fft_lenght / 4)) has to be evenly "
              "devisable by r2_blocksize.")
        return None
    return Plan(fft_length_=fft_length, amount_of_r16_steps_=r16, amount_of_r2_steps_=r2, base_fft_mode_=mode,
                results_in_results_=True,   # the fused kernels always write the results planes
                base_fft_warps_per_block_=base_w, base_fft_blocksize_=base_w * 32,
                base_fft_gridsize_=total_warps // base_w, base_fft_shared_mem_in_bytes_=base_w * 1024 * 2,
                r16_warps_per_block_=r16_w, r16_blocksize_=r16_w * 32, r16_gridsize_=total_warps // r16_w,
                r16_shared_mem_in_bytes_=r16_w * 768 * 2, r2_blocksize_=r2_blocksize)


def create_plan_from_file(fft_length: int, tuner_results_file: str) -> Optional[Plan]:
    """CreatePlan(N, tuner_file) (Plan.h:197-255): line format `N mode base_warps r16_warps r2_block`."""
    try:
        with open(tuner_results_file) as f:
            for line in f:
                parts = line.split()
                if parts and int(float(parts[0])) == fft_length:
                    mode = MODE_256 if int(parts[1]) == 256 else MODE_4096
                    global _tuner_file
                    _tuner_file = tuner_results_file   # compute_fft builds its native plans from this file
                    return create_plan(fft_length, mode, int(parts[2]), int(parts[3]), int(parts[4]))
    except OSError:
        print("Error! Failed to open tuner file.")
        return None
    print("Error! Tuner file didnt contain requested fft length.")
    return None


def get_max_no_optin_shared_mem(device_id: int = 0) -> int:
    """GetMaxNoOptInSharedMem (Plan.h:298-303)."""
    import torch
    return int(torch.cuda.get_device_properties(device_id).shared_memory_per_block)


def plan_works_on_device(plan: Plan, device_id: int = 0) -> bool:
    """PlanWorksOnDevice (Plan.h:257-296), for the B200 kernels: needs compute capability 10.x."""
    import torch
    if not torch.cuda.is_available():
        print("Error! No CUDA device.")
        return False
    major, _ = torch.cuda.get_device_capability(device_id)
    if major != 10:
        print("Error! Compute capability 10.x (sm_100a) is required.")
        return False
    return True


class DataHandler:
    """DataHandler<Integer> (DataHandler.h:22-82): one allocation of 4*N halves
    [in_RE | in_IM | out_RE | out_IM]."""

    def __init__(self, fft_length: int, device: str = "cuda"):
        import torch
        self.fft_length_ = fft_length
        self.amount_of_ffts_ = 1
        self.dptr_data_ = torch.empty(4 * fft_length, dtype=torch.float16, device=device)
        n = fft_length
        self.dptr_input_RE_ = self.dptr_data_[0:n]
        self.dptr_input_IM_ = self.dptr_data_[n:2 * n]
        self.dptr_results_RE_ = self.dptr_data_[2 * n:3 * n]
        self.dptr_results_IM_ = self.dptr_data_[3 * n:4 * n]

    def peak_at_last_error(self) -> Optional[str]:
        return None

    def copy_data_host_to_device(self, data) -> Optional[str]:
        import torch
        src = torch.from_numpy(data).view(torch.float16).reshape(-1)
        self.dptr_data_[:2 * self.fft_length_ * self.amount_of_ffts_].copy_(src)
        return None

    def copy_results_device_to_host(self, data, results_in_results: bool) -> Optional[str]:
        import torch
        k = 2 * self.fft_length_ * self.amount_of_ffts_
        src = self.dptr_data_[k:2 * k] if results_in_results else self.dptr_data_[:k]
        torch.from_numpy(data).view(torch.float16).reshape(-1).copy_(src)
        return None


class DataBatchHandler(DataHandler):
    """DataBatchHandler<Integer> (DataHandler.h:86-166): inputs [RE_0|IM_0|RE_1|IM_1|...] then the
    results in the same order."""

    def __init__(self, fft_length: int, amount_of_ffts: int, device: str = "cuda"):
        import torch
        self.fft_length_ = fft_length
        self.amount_of_ffts_ = amount_of_ffts
        n, b = fft_length, amount_of_ffts
        self.dptr_data_ = torch.empty(4 * n * b, dtype=torch.float16, device=device)
        self.dptr_input_RE_ = [self.dptr_data_[2 * i * n:(2 * i + 1) * n] for i in range(b)]
        self.dptr_input_IM_ = [self.dptr_data_[(2 * i + 1) * n:(2 * i + 2) * n] for i in range(b)]
        off = 2 * n * b
        self.dptr_results_RE_ = [self.dptr_data_[off + 2 * i * n:off + (2 * i + 1) * n] for i in range(b)]
        self.dptr_results_IM_ = [self.dptr_data_[off + (2 * i + 1) * n:off + (2 * i + 2) * n] for i in range(b)]


_plan_cache: dict = {}
_tuner_file: Optional[str] = None   # set by create_plan_from_file (the key=value knobs of the same lines steer the kernels)


def compute_fft(fft_plan: Plan, data: DataHandler, max_no_optin_shared_mem: int = 32768) -> Optional[str]:
    """ComputeFFT (ComputeFFT.h:54-151 single, :162-293 batch). Returns None on success, else the
    error string -- the reference's std::optional<std::string> convention. Asynchronous for the
    single handler, synchronised for the batch handler, like the reference."""
    import torch
    n, b = fft_plan.fft_length_, data.amount_of_ffts_
    if n != data.fft_length_:
        return "plan / data handler length mismatch"
    try:
        key = (n, b, data.dptr_data_.device.index, _tuner_file)
        native = _plan_cache.get(key)
        if native is None:
            try:
                native = NativePlan(n, b, tuner_file=_tuner_file)
            except TfftError:
                if _tuner_file is None:
                    raise
                native = NativePlan(n, b)           # length not in the file: default knobs
            _plan_cache[key] = native
        buf = data.dptr_data_
        half = 2 * n * b
        native.exec(buf[0:], buf[n:], buf[half:], buf[half + n:], 2 * n, 2 * n)
        if b > 1:
            torch.cuda.synchronize()
    except TfftError as e:
        return str(e)
    return None
