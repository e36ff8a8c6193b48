// Multi-GPU 1-D transform (BASELINE config C4): exchange kernels that write straight into the peers'
// memory over NVLink (peer-mapped buffers, one process per GPU) and the flag barrier between the phases.  The reference
// has no working multi-GPU path (its ComputeFFTMultiGPU, src/base/ComputeFFT.h:295-557, is commented-out replica code).
//
//   mg_transpose_send : the distributed transpose of one exchange as ONE kernel: 64 x 64 fp16 tiles of the local slab go
//                       through shared memory and are stored transposed, as whole 128-byte lines, into the buffer of the
//                       rank that owns those columns -- pack, all-to-all and unpack of the NCCL version in one pass.
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
// [p*cl, (p+1)*cl), cl = cols / world, and receives them transposed: dst_p[c][rank*rows_local + r] = src[r][p*cl + c],
// dst row length = world * rows_local.  Grid: (cols/64, rows_local/64, 2 planes), 256 threads.
// Tiles are visited peer-interleaved (consecutive CTAs target different peers, starting at rank+1) so that all NVLink
// ports carry traffic at any moment.
__global__ void __launch_bounds__(256)
mg_transpose_send(const __half* __restrict__ src_re, const __half* __restrict__ src_im, const MgPeers peers,
                  int rows_local, int cols, int rank, int world, int64_t src_row_stride) {
  __shared__ __align__(16) __half tile[64][72];   // 144-byte rows: 16-byte column reads spread over the banks
  const int cl = cols / world, tiles_per_peer = cl / 64;
  const int bx = blockIdx.x;
  const int peer = (bx + rank + 1) % world;
  const int tx = bx / world;                       // column tile inside the peer's block
  (void)tiles_per_peer;
  const int plane = blockIdx.z;
  const __half* s = (plane ? src_im : src_re) + static_cast<int64_t>(blockIdx.y) * 64 * src_row_stride +
                    static_cast<int64_t>(peer) * cl + tx * 64;
  const int64_t d_row = static_cast<int64_t>(world) * rows_local;
  __half* d = (plane ? peers.im[peer] : peers.re[peer]) + static_cast<int64_t>(tx) * 64 * d_row +
              static_cast<int64_t>(rank) * rows_local + blockIdx.y * 64;
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (t >> 3) + 32 * i, ch = t & 7;
    // chunk ch of row r is kept at chunk position ch ^ (r / 8): the column reads below then hit 8 distinct bank groups
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(s + r * src_row_stride + ch * 8));
    *reinterpret_cast<uint4*>(&tile[r][(ch ^ ((r >> 3) & 7)) * 8]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = (t >> 3) + 32 * i, ch = t & 7;   // output row c (= source column), 8 consecutive source rows
    __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = tile[ch * 8 + j][(((c >> 3) ^ ch) << 3) + (c & 7)];
    *reinterpret_cast<uint4*>(d + c * d_row + ch * 8) = *reinterpret_cast<const uint4*>(v);
  }
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
  __threadfence_system();
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
    __nanosleep(200);
  }
  __threadfence_system();
}

}  // namespace tfft
