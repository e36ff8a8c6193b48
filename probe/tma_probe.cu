// Hardware probe (developer tool): SWIZZLE_128B MN-major A operand for tcgen05.mma, filled
//   (1) by the host with the assumed swizzled image, (2) by a 3-D TMA tensor load
// and checked against a CPU product.  Settles LBO/SBO roles for the swizzled layout and the
// tensor-map dimension order the FFT kernel's load phase uses.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../tensor-fft_b200/csrc/sm100_ptx.cuh"
using namespace tfft::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return make_smem_desc(saddr, lbo, sbo) | (uint64_t(2) << 61);
}

struct Args { uint32_t a_lbo, a_sbo, use_tma, image_bytes, a_bytes; };

// D[128 x 32] = A[128 x 16] * B[16 x 32]; A = rows m = 0..127 (two 64-row atoms), K = 16 (two K groups)
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmap, const uint8_t* image, Args args, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, tbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&tbar, 1); fence_mbar_init(); }
  __syncthreads();
  // B always from the host image (after the A region)
  for (uint32_t i = tid * 16; i < args.image_bytes; i += 128 * 16)
    if (i >= args.a_bytes || !args.use_tma) *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  if (args.use_tma && tid == 0) {
    mbar_arrive_expect_tx(&tbar, args.a_bytes);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(smem)), "l"(&tmap), "r"(0), "r"(0), "r"(0), "r"(smem_u32(&tbar)) : "memory");
  }
  if (args.use_tma) mbar_wait(&tbar, 0);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = slot;
  if (tid == 0) {
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = make_idesc_f16(128, 32, 1, 0);
    uint64_t da = make_desc_sw128(sb, args.a_lbo, args.a_sbo);
    uint64_t db = make_smem_desc(sb + args.a_bytes, 128, 256);
    umma_f16_ss(taddr, da, db, idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr + (uint32_t(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, 32);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int M = 128, N = 32, K = 16;
  // the "signal": x[n], n = kappa * M + m  (M = 128 rows contiguous, kappa = K index with stride M)
  std::vector<__half> x(M * K);
  std::vector<float> A(M * K), B(K * N), ref(M * N);
  srand(3);
  for (int kap = 0; kap < K; ++kap) for (int m = 0; m < M; ++m) { float v = (float)((rand() % 9) - 4); A[m * K + kap] = v; x[kap * M + m] = __float2half(v); }
  for (auto& v : B) v = (float)((rand() % 5) - 2) * 0.5f;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int kk = 0; kk < K; ++kk) s += A[m * K + kk] * B[kk * N + n]; ref[m * N + n] = s; }
  // host image of A in the assumed SW128 MN-major layout: atoms of 64 rows; atom stride = 2 K-groups * 1024
  const uint32_t atom_stride = 2048, a_bytes = 2 * atom_stride;
  std::vector<uint8_t> image(a_bytes + 1024, 0);
  for (int m = 0; m < M; ++m) for (int kk = 0; kk < K; ++kk) {
    uint32_t atom = m / 64, c = (m / 8) % 8, pos = m % 8, kg = kk / 8, l = kk % 8;
    uint32_t off = atom * atom_stride + kg * 1024 + l * 128 + ((c ^ l) * 16) + pos * 2;
    __half h = __float2half(A[m * K + kk]); memcpy(&image[off], &h, 2);
  }
  for (int kk = 0; kk < K; ++kk) for (int n = 0; n < N; ++n) {
    uint32_t off = a_bytes + (n >> 3) * 256 + (kk >> 3) * 128 + (n & 7) * 16 + (kk & 7) * 2;
    __half h = __float2half(B[kk * N + n]); memcpy(&image[off], &h, 2);
  }
  uint8_t* d_image; float* d_out; __half* d_x;
  CK(cudaMalloc(&d_image, image.size())); CK(cudaMalloc(&d_out, M * N * 4)); CK(cudaMalloc(&d_x, x.size() * 2));
  CK(cudaMemcpy(d_image, image.data(), image.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
  // tensor map: dims (fastest first) d0 = m_lo (64), d1 = kappa (K, stride M elems), d2 = m_hi (M/64, stride 64 elems)
  EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
  cuuint64_t gdim[3] = {64, (cuuint64_t)K, (cuuint64_t)(M / 64)};
  cuuint64_t gstride[2] = {(cuuint64_t)M * 2, 64 * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)K, (cuuint32_t)(M / 64)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d_x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("cuTensorMapEncodeTiled -> %d\n", (int)cr);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  for (int use_tma = 0; use_tma < 2; ++use_tma)
    for (int swap = 0; swap < 2; ++swap) {
      Args a; a.use_tma = use_tma; a.image_bytes = (uint32_t)image.size(); a.a_bytes = a_bytes;
      a.a_lbo = swap ? 1024 : atom_stride; a.a_sbo = swap ? atom_stride : 1024;
      CK(cudaMemset(d_out, 0, M * N * 4));
      k<<<1, 128, 32 * 1024>>>(tmap, d_image, a, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("use_tma=%d swap=%d: CUDA error %s\n", use_tma, swap, cudaGetErrorString(e)); return 1; }
      std::vector<float> out(M * N); CK(cudaMemcpy(out.data(), d_out, M * N * 4, cudaMemcpyDeviceToHost));
      int bad = 0; double mx = 0; for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (d > 1e-3) ++bad; if (d > mx) mx = d; }
      printf("SW128 use_tma=%d lbo=%u sbo=%u : bad=%d/%d maxerr=%g %s\n", use_tma, a.a_lbo, a.a_sbo, bad, M * N, mx, bad ? "mismatch" : "MATCH");
    }
  return 0;
}
