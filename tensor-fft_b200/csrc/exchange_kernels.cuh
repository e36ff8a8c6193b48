// Data-movement kernels around the all-to-all exchanges of the distributed six-step: the block
// transposes that pack / unpack the exchange buffers.  torch's generic strided copy moves these at a fraction of
// the HBM rate; here a 64 x 64 fp16 tile goes through shared memory with 16-byte global accesses on both sides.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tfft {

// dst[b][c][r] = src[b][r][c] for a batch of matrices; a matrix index b = b1 * nb0 + b0 addresses
// src + b0 * s_b0 + b1 * s_b1 (row stride s_row) and dst + b0 * d_b0 + b1 * d_b1 (row stride d_row).
// rows and cols are multiples of 64; all strides are multiples of 8 elements.
__global__ void __launch_bounds__(256)
transpose64_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int64_t s_row, int64_t d_row, int nb0,
                   int64_t s_b0, int64_t s_b1, int64_t d_b0, int64_t d_b1) {
  __shared__ __align__(16) __half tile[64][72];   // 144-byte rows: 16-byte column reads spread over the banks
  const int b0 = blockIdx.z % nb0, b1 = blockIdx.z / nb0;
  const __half* s = src + b0 * s_b0 + b1 * s_b1 + static_cast<int64_t>(blockIdx.y) * 64 * s_row + blockIdx.x * 64;
  __half* d = dst + b0 * d_b0 + b1 * d_b1 + static_cast<int64_t>(blockIdx.x) * 64 * d_row + blockIdx.y * 64;
  const int t = threadIdx.x;
  // 64 rows x 8 chunks of 8 halves: thread t loads rows (t >> 3) and (t >> 3) + 32, chunk t & 7
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (t >> 3) + 32 * i, ch = t & 7;
    // chunk ch of row r is kept at chunk position ch ^ (r / 8): the column reads below then hit 8 distinct bank groups
    *reinterpret_cast<uint4*>(&tile[r][(ch ^ ((r >> 3) & 7)) * 8]) = *reinterpret_cast<const uint4*>(s + r * s_row + ch * 8);
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

// dst[i2][i1][i0][0 .. run) = src[...] : a three-level strided copy of contiguous runs (run = multiple of 8 elements)
__global__ void __launch_bounds__(256)
copy_runs_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int64_t run, int n0, int n1,
                 int64_t s0, int64_t s1, int64_t s2, int64_t d0, int64_t d1, int64_t d2) {
  const int i0 = blockIdx.x % n0, i1 = (blockIdx.x / n0) % n1, i2 = blockIdx.x / (n0 * n1);
  const uint4* s = reinterpret_cast<const uint4*>(src + i0 * s0 + i1 * s1 + i2 * s2);
  uint4* d = reinterpret_cast<uint4*>(dst + i0 * d0 + i1 * d1 + i2 * d2);
  for (int64_t k = static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x; k < run / 8;
       k += static_cast<int64_t>(gridDim.y) * blockDim.x)
    d[k] = s[k];
}

}  // namespace tfft
