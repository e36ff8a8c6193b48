// Multi-GPU 1-D transform (BASELINE config C4): exchange kernels that write straight into the peers'
// memory over NVLink (peer-mapped buffers, one process per GPU) and the flag barrier between the phases.  The reference
// has no working multi-GPU path (its ComputeFFTMultiGPU, src/base/ComputeFFT.h:295-557, is commented-out replica code).
//
//   mg_transpose_send : the distributed transpose of one exchange as ONE kernel: 64 x 32 fp16 tiles of the local slab are
//                       transposed in registers and stored, as whole 128-byte lines, into the buffer of the rank that
//                       owns those columns -- pack, all-to-all and unpack of the NCCL version in one pass.
//   mg_barrier        : every rank raises its flag on all peers (system-scope release after the data kernel of the same
//                       stream has completed) and waits until all peers have raised theirs (acquire); bounded spin.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tfft {

constexpr int kMgMaxRanks = 16;

struct MgPeers {
  __half* re[kMgMaxRanks];   // destination planes of every rank (peer-mapped; [rank] is the local one)
  __half* im[kMgMaxRanks];
};

// Local slab: `rows_local` rows of `cols` elements (both planes), rank `rank` of `world`.  Peer p owns columns
// [p*cl, (p+1)*cl), cl = cols / world, and receives them transposed into its staging buffer, SOURCE-RANK MAJOR:
// stage_p[rank][c][r] = src[r][p*cl + c] -- one contiguous block of cl * rows_local elements per source rank.  (Writing
// the row-major matrix dst_p[c][rank*rows_local + r] directly scatters 128-byte..4-KiB pieces over every page of the
// peer's buffer; measured at 8 GPUs that costs 0.13 ms per exchange over the 0.19 ms the same bytes take when the
// destination is contiguous.)  mg_unpack then builds the row-major matrix locally.  With staged == 0 the kernel writes
// the row-major matrix directly (no unpack pass): cheaper for 2 ranks, where the scattered stores cost nothing.
// One warp moves a tile of 64 source rows x 32 source columns entirely in registers: lane (rg = lane / 4, ch = lane % 4)
// loads the 8 x 8 block of rows 8*rg .. 8*rg+7, columns 8*ch .. 8*ch+7 (eight 16-byte loads; a warp instruction covers
// 8 rows x 64 contiguous bytes), transposes it with byte permutes and stores eight 16-byte pieces; a warp store
// instruction writes 4 destination rows x 128 contiguous bytes -- whole lines, which is what matters on NVLink.
// Grid: (rows_local / 64, ceil(cols / 32 / 8), 2 planes), 256 threads.  Column groups are visited peer-interleaved
// (consecutive warps target different peers, starting at rank + 1) so that all NVLink ports carry traffic at any moment.
__global__ void __launch_bounds__(256)
mg_transpose_send(const __half* __restrict__ src_re, const __half* __restrict__ src_im, const MgPeers peers,
                  int rows_local, int cols, int rank, int world, int64_t src_row_stride, int staged) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // blockIdx.x runs over the 64-row blocks (fastest), blockIdx.y over groups of 8 column groups: CTAs that run together
  // write neighbouring 128-byte segments of the same destination rows
  const int lin = blockIdx.y * 8 + warp;             // column group (32 columns) in peer-interleaved order
  if (lin >= cols / 32) return;
  const int cl = cols / world;
  const int peer = (lin + rank + 1) % world;
  const int cg = lin / world;                        // column group inside the peer's block
  const int rg = lane >> 2, ch = lane & 3;
  const int plane = blockIdx.z;
  const __half* s = (plane ? src_im : src_re) + (static_cast<int64_t>(blockIdx.x) * 64 + 8 * rg) * src_row_stride +
                    static_cast<int64_t>(peer) * cl + cg * 32 + 8 * ch;
  uint4 a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = __ldcs(reinterpret_cast<const uint4*>(s + j * src_row_stride));
  // staged: stage_p[rank][c][r] (contiguous block per source rank); direct: the row-major matrix dst_p[c][rank*rows_local + r]
  const int64_t d_row = staged ? rows_local : static_cast<int64_t>(world) * rows_local;
  __half* d = (plane ? peers.im[peer] : peers.re[peer]) +
              (staged ? static_cast<int64_t>(rank) * cl * rows_local : static_cast<int64_t>(rank) * rows_local) +
              (static_cast<int64_t>(cg) * 32 + 8 * ch) * d_row + blockIdx.x * 64 + 8 * rg;
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    const uint32_t sel = (cc & 1) ? 0x7632u : 0x5410u;
    uint32_t w[4];
#pragma unroll
    for (int i2 = 0; i2 < 4; ++i2) {
      const uint32_t lo = reinterpret_cast<const uint32_t*>(&a[2 * i2])[cc >> 1];
      const uint32_t hi = reinterpret_cast<const uint32_t*>(&a[2 * i2 + 1])[cc >> 1];
      w[i2] = __byte_perm(lo, hi, sel);
    }
    *reinterpret_cast<uint4*>(d + cc * d_row) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// dst[c][q*rows_local + r] = stage[q][c][r] for q < world, c < cl: runs of rows_local elements (both planes: blockIdx.z).
// One warp per run.  Grid: (ceil(cl / 8), world, 2), 256 threads.
__global__ void __launch_bounds__(256)
mg_unpack(const __half* __restrict__ st_re, const __half* __restrict__ st_im, __half* __restrict__ dst_re,
          __half* __restrict__ dst_im, int rows_local, int cl, int world) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * 8 + (threadIdx.x >> 5), q = blockIdx.y;
  if (c >= cl) return;
  const __half* s = (blockIdx.z ? st_im : st_re) + (static_cast<int64_t>(q) * cl + c) * rows_local;
  __half* d = (blockIdx.z ? dst_im : dst_re) + static_cast<int64_t>(c) * world * rows_local + static_cast<int64_t>(q) * rows_local;
  for (int k = lane; k < rows_local / 8; k += 32)
    reinterpret_cast<uint4*>(d)[k] = __ldcs(reinterpret_cast<const uint4*>(s) + k);
}

struct MgFlags {
  uint32_t* flags[kMgMaxRanks];   // flags[p] = rank p's flag array (world entries, peer-mapped): slot [rank] is ours to write
};

// One CTA, `world` threads.  Thread p: release-store `epoch` into our slot on rank p, then spin until rank p's slot in
// our own array has reached `epoch`.  The kernel runs after the data kernel in the same stream, so all of this rank's
// peer stores of the phase are complete before the flags go up.  The spin is bounded (timeout_ns): on expiry the
// kernel records the phase in *status (mapped host memory) and returns, so a lost rank cannot hang the device.
__global__ void mg_barrier(const MgFlags f, int rank, int world, uint32_t epoch, unsigned long long timeout_ns,
                           volatile int* status) {
  const int p = threadIdx.x;
  if (p >= world) return;
#if !defined(TFFT_MG_BARRIER_VARIANT) || TFFT_MG_BARRIER_VARIANT == 0
  __threadfence_system();
#endif
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.flags[p] + rank), "r"(epoch) : "memory");
  const uint32_t* mine = f.flags[rank] + p;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (static_cast<int32_t>(v - epoch) >= 0) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      *status = static_cast<int>(epoch);
      break;
    }
#if !defined(TFFT_MG_BARRIER_VARIANT) || TFFT_MG_BARRIER_VARIANT < 2
    __nanosleep(200);
#endif
  }
#if !defined(TFFT_MG_BARRIER_VARIANT) || TFFT_MG_BARRIER_VARIANT == 0
  __threadfence_system();
#endif
}

}  // namespace tfft
