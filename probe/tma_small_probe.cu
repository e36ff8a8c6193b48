// Hardware probe (developer tool): how does a SWIZZLE_128B TMA tile land in shared memory when the inner box
// dimension is only 32 bytes (16 fp16)?  Box {16 rows, 4 transforms, 16 kappa, 2 groups} from a tensor whose element
// value encodes its coordinates; the kernel dumps 32 KiB of shared memory so the host can print where every
// element went and whether a 128-byte line holds 4 x 32 bytes (dense box) or one padded 32-byte row.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../tensor-fft_b200/csrc/sm100_ptx.cuh"
using namespace tfft::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmap, uint32_t tx_bytes, uint16_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t tbar;
  const int tid = threadIdx.x;
  for (int i = tid; i < 32768 / 2; i += 128) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
  if (tid == 0) { mbar_init(&tbar, 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(&tbar, tx_bytes);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %2, %2, %2}], [%3];"
                 :: "r"(smem_u32(smem)), "l"(&tmap), "r"(0), "r"(smem_u32(&tbar)) : "memory");
  }
  mbar_wait(&tbar, 0);
  __syncthreads();
  for (int i = tid; i < 32768 / 2; i += 128) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  // tensor: transforms t (stride 512 elements), element n = kappa*16 + m (L = 256, M = 16, R = 16)
  const int L = 256, M = 16, R = 16, T = 4, NT = 8, tstride = 512;
  std::vector<uint16_t> h(NT * tstride);
  for (int t = 0; t < NT; ++t) for (int n = 0; n < tstride; ++n) h[t * tstride + n] = (uint16_t)((t << 12) | (n & 0xFFF));
  uint16_t *d, *dout; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMalloc(&dout, 32768));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)M, (cuuint64_t)T, (cuuint64_t)R, (cuuint64_t)(NT / T)};
  cuuint64_t gstr[3] = {(cuuint64_t)tstride * 2, (cuuint64_t)M * 2, (cuuint64_t)tstride * T * 2};
  cuuint32_t box[4] = {(cuuint32_t)M, (cuuint32_t)T, (cuuint32_t)R, 2};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc = %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 0;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const uint32_t tx = M * T * R * 2 * 2;
  k<<<1, 128, 65536>>>(tm, tx, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s (expect_tx %u bytes)\n", cudaGetErrorString(e), tx);
  if (e != cudaSuccess) return 0;
  std::vector<uint16_t> o(16384);
  CK(cudaMemcpy(o.data(), dout, 32768, cudaMemcpyDeviceToHost));
  int written = 0, last = -1;
  for (int i = 0; i < 16384; ++i) if (o[i] != 0xFFFF) { ++written; last = i; }
  printf("halves written %d, last written half index %d (dense box would be %d)\n", written, last, M * T * R * 2 - 1);
  // print the first 4 lines of 128 bytes: (t, n) of each 16-byte chunk's first element
  for (int line = 0; line < 12; ++line) {
    printf("line %2d:", line);
    for (int c = 0; c < 8; ++c) { uint16_t v = o[line * 64 + c * 8]; if (v == 0xFFFF) printf("  ----  "); else printf(" t%d n%3d", v >> 12, v & 0xFFF); }
    printf("\n");
  }
  // check the assumed layout: element (group g, kappa, t_lo, m) at g*128R + kappa*128 + ((((t_lo*16+m)>>3) ^ (kappa&7))<<4) + ((t_lo*16+m)&7)*2
  int bad = 0;
  for (int g = 0; g < 2; ++g) for (int kap = 0; kap < R; ++kap) for (int tl = 0; tl < T; ++tl) for (int m = 0; m < M; ++m) {
    const int row = tl * 16 + m;
    const int off = g * 128 * R + kap * 128 + ((((row >> 3) & 7) ^ (kap & 7)) << 4) + (row & 7) * 2;
    const uint16_t want = (uint16_t)(((g * T + tl) << 12) | (kap * M + m));
    if (o[off / 2] != want) ++bad;
  }
  printf("assumed layout mismatches: %d of %d\n", bad, 2 * R * T * M);
  return 0;
}
