"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two cases (the reference itself has no working multi-GPU code, its
ComputeFFTMultiGPU in src/base/ComputeFFT.h:295-557 is commented-out replica code):

* batched transforms shard by batch -- contiguous ranges, no collective (`shard_range`);
* ONE huge 1-D transform of length N = N1*N2 is slab-distributed and computed as a six-step
  with all-to-all exchanges (`SixStepPlan`):
      x[i1*N2 + i2], rank r owns rows i1 in block r
      A2A #1  -> rank r owns columns i2 in block r, all i1          (transpose)
      FFT over i1 (length N1, batch N2/G) fused with * exp(-2 pi i k1 i2 / N)   [tfft_exec_twiddled]
      A2A #2  -> rank r owns k1 in block r, all i2
      FFT over i2 (length N2, batch N1/G)
      A2A #3  -> natural order X[k1 + N1*k2], rank r owns the r-th contiguous N/G slice
  The local transform is injected (`local_fft`) so that the exchange logic can be tested on CPU ranks
  (gloo) with a stand-in; on GPUs it is the tfft C ABI.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch range [lo, hi) of `rank`; ranges tile [0, total) and differ by at most 1."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: torch.device) -> float:
    """Max of a host scalar over all ranks (timings are reported as the slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _gpu_pack(z: torch.Tensor, world: int) -> torch.Tensor:
    """send[p][plane][c][r] = z[plane][r][p*cols_local + c] with the library's 64x64-tile transpose kernel."""
    import ctypes
    from . import lib, _check
    _, rows_local, cols = z.shape
    cl = cols // world
    send = torch.empty((world, 2, cl, rows_local), dtype=z.dtype, device=z.device)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # matrices indexed (b0 = plane, b1 = peer): source block z[plane][:, p*cl:(p+1)*cl], destination send[p][plane]
    _check(lib().tfft_transpose_blocks(z.data_ptr(), send.data_ptr(), rows_local, cl, cols, rows_local, 2, world,
                                       rows_local * cols, cl, cl * rows_local, 2 * cl * rows_local, s))
    return send


def _gpu_unpack(recv: torch.Tensor, world: int) -> torch.Tensor:
    """out[plane][c][p*rows_local + r] = recv[p][plane][c][r]: runs of rows_local elements."""
    import ctypes
    from . import lib, _check
    _, _, cl, rows_local = recv.shape
    out = torch.empty((2, cl, world * rows_local), dtype=recv.dtype, device=recv.device)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # levels (i0 = c, i1 = plane, i2 = peer)
    _check(lib().tfft_copy_runs(recv.data_ptr(), out.data_ptr(), rows_local, cl, 2, world,
                                rows_local, cl * rows_local, 2 * cl * rows_local,
                                world * rows_local, cl * world * rows_local, rows_local, s))
    return out


def _a2a_transpose(z: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """z: (2, rows_local, cols) on every rank -- the real and the imaginary plane, rows block-distributed.
    Returns (2, cols_local, rows) with the columns block-distributed: the distributed transpose of both planes
    in ONE all_to_all_single."""
    _, rows_local, cols = z.shape
    cols_local = cols // world
    # block for peer p = my rows x p's columns of both planes, sent transposed so the receiver only concatenates
    fast = (z.is_cuda and z.dtype == torch.float16 and z.is_contiguous() and rows_local % 64 == 0
            and cols_local % 64 == 0 and world * 2 <= 65535)
    if fast:
        send = _gpu_pack(z, world)
    else:   # CPU ranks (gloo tests) and odd shapes
        send = z.reshape(2, rows_local, world, cols_local).permute(2, 0, 3, 1).contiguous()   # (peer, plane, cols_local, rows_local)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[p] = (plane, my cols_local, p's rows_local)  ->  (plane, cols_local, world * rows_local)
    if fast:
        return _gpu_unpack(recv, world)
    return recv.permute(1, 2, 0, 3).reshape(2, cols_local, world * rows_local).contiguous()


class SixStepPlan:
    """Distributed 1-D FFT of length n1*n2 over `world` ranks, natural order in and out, scale 1/N.

    local_fft(z, n, batch, log2_total, first_col) -> z': `batch` contiguous transforms of length n on the
    stacked planes z = (2, batch, n) (z[0] real, z[1] imaginary); when log2_total > 0 output k of transform b
    is also multiplied by exp(-2 pi i k (first_col + b) / 2^log2_total).
    """

    def __init__(self, n1: int, n2: int, rank: int, world: int,
                 local_fft: Callable[..., Tuple[torch.Tensor, torch.Tensor]], group=None):
        if n1 % world or n2 % world:
            raise ValueError("both factors must be divisible by the number of ranks")
        self.n1, self.n2, self.rank, self.world, self.local_fft, self.group = n1, n2, rank, world, local_fft, group
        self.log2_total = (n1 * n2).bit_length() - 1
        self.all_to_alls = 3

    def forward(self, re: torch.Tensor, im: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """re/im: this rank's slab, n1/world rows of n2 (1-D of length N/world or 2-D). Returns the rank's
        contiguous N/world slice of the spectrum as 1-D tensors."""
        n1, n2, w = self.n1, self.n2, self.world
        z = torch.stack([re.reshape(n1 // w, n2), im.reshape(n1 // w, n2)])             # (2, n1/w, n2)
        # A2A #1: columns of the N1 x N2 matrix become local rows
        z = _a2a_transpose(z, w, self.group)                                             # (2, n2/w, n1)
        first_col = self.rank * (n2 // w)
        z = self.local_fft(z, n1, n2 // w, self.log2_total, first_col)                   # Y[i2][k1] * w^(k1 i2)
        # A2A #2
        z = _a2a_transpose(z, w, self.group)                                             # (2, n1/w, n2)
        z = self.local_fft(z, n2, n1 // w, 0, 0)                                         # Z[k1][k2]
        # A2A #3: natural order k = k1 + n1*k2  ->  rank r owns k2 in block r
        z = _a2a_transpose(z, w, self.group)                                             # (2, n2/w, n1)
        return z[0].reshape(-1), z[1].reshape(-1)

    def nvlink_bytes_per_rank(self) -> int:
        """fp16 planar bytes this rank sends per transform: A * (G-1)/G * 4N/G."""
        n = self.n1 * self.n2
        return self.all_to_alls * (self.world - 1) * 4 * n // (self.world * self.world)


def tfft_local_fft():
    """local_fft for SixStepPlan backed by the C ABI (CUDA tensors, fp16). Plans are cached per shape."""
    from . import NativePlan
    cache = {}

    def run(z, n, batch, log2_total, first_col):
        plan = cache.get((n, batch))
        if plan is None:
            plan = cache[(n, batch)] = NativePlan(n, batch)
        o = torch.empty_like(z)
        if log2_total:
            plan.exec_twiddled(z[0], z[1], o[0], o[1], n, n, log2_total, first_col)
        else:
            plan.exec(z[0], z[1], o[0], o[1], n, n)
        return o

    return run


def make_mg_plan(n: int, group=None):
    """tfft.MgPlan for this rank of `group` (default: the world): the IPC handles of the plan-owned exchange buffers are
    all-gathered over torch.distributed -- plumbing only; the data path of MgPlan.exec is the library's own kernels
    storing into peer memory over NVLink (include/tfft.h, tfft_mg_*)."""
    from . import MgPlan
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    plan = MgPlan(n, rank, world)

    def gather(h):
        out = [None] * world
        dist.all_gather_object(out, h, group=group)
        return out

    plan.connect_with(gather)
    return plan
