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


def _a2a_transpose(x: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """x: (rows_local, cols) on every rank, rows block-distributed.  Returns (cols_local, rows) with the
    columns block-distributed: the distributed transpose, one all_to_all_single."""
    rows_local, cols = x.shape
    cols_local = cols // world
    # block for peer p = my rows x p's columns, sent transposed so the receiver only concatenates
    send = x.reshape(rows_local, world, cols_local).permute(1, 2, 0).contiguous()   # (peer, cols_local, rows_local)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    # recv[p] = (my cols_local, p's rows_local)  ->  (cols_local, world * rows_local)
    return recv.permute(1, 0, 2).reshape(cols_local, world * rows_local).contiguous()


class SixStepPlan:
    """Distributed 1-D FFT of length n1*n2 over `world` ranks, natural order in and out, scale 1/N.

    local_fft(re, im, n, batch, log2_total, first_col) -> (re, im): `batch` contiguous transforms of
    length n on 2-D tensors (batch, n); when log2_total > 0 output k of transform b is also multiplied
    by exp(-2 pi i k (first_col + b) / 2^log2_total).
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
        re, im = re.reshape(n1 // w, n2), im.reshape(n1 // w, n2)
        # A2A #1: columns of the N1 x N2 matrix become local rows
        re, im = _a2a_transpose(re, w, self.group), _a2a_transpose(im, w, self.group)     # (n2/w, n1)
        first_col = self.rank * (n2 // w)
        re, im = self.local_fft(re, im, n1, n2 // w, self.log2_total, first_col)           # Y[i2][k1] * w^(k1 i2)
        # A2A #2
        re, im = _a2a_transpose(re, w, self.group), _a2a_transpose(im, w, self.group)     # (n1/w, n2)
        re, im = self.local_fft(re, im, n2, n1 // w, 0, 0)                                 # Z[k1][k2]
        # A2A #3: natural order k = k1 + n1*k2  ->  rank r owns k2 in block r
        re, im = _a2a_transpose(re, w, self.group), _a2a_transpose(im, w, self.group)     # (n2/w, n1)
        return re.reshape(-1), im.reshape(-1)

    def nvlink_bytes_per_rank(self) -> int:
        """fp16 planar bytes this rank sends per transform: A * (G-1)/G * 4N/G."""
        n = self.n1 * self.n2
        return self.all_to_alls * (self.world - 1) * 4 * n // (self.world * self.world)


def tfft_local_fft():
    """local_fft for SixStepPlan backed by the C ABI (CUDA tensors, fp16). Plans are cached per shape."""
    from . import NativePlan
    cache = {}

    def run(re, im, n, batch, log2_total, first_col):
        plan = cache.get((n, batch))
        if plan is None:
            plan = cache[(n, batch)] = NativePlan(n, batch)
        o_re, o_im = torch.empty_like(re), torch.empty_like(im)
        if log2_total:
            plan.exec_twiddled(re, im, o_re, o_im, n, n, log2_total, first_col)
        else:
            plan.exec(re, im, o_re, o_im, n, n)
        return o_re, o_im

    return run
