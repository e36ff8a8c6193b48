// Hardware probe (developer tool, not product): pins down the tcgen05 operand
// layouts the FFT kernels rely on, and measures a few pipe rates.
//   1. layout test: D[128x32] = Are*B1 + Aim*B2 with A MN-major / B K-major
//      SWIZZLE_NONE canonical layouts, for both LBO/SBO role assignments.
//   2. rate test: cycles per UMMA (M=128, K=16) for N in {32,64,128,256};
//      tcgen05.ld 32x32b.x32 cycles; STS.128 / LDS.128 cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../tensor-fft_b200/csrc/sm100_ptx.cuh"

using namespace tfft::ptx;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

struct ProbeArgs {
  uint32_t a_re_off, a_im_off, b1_off, b2_off;  // byte offsets in smem image
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;          // descriptor fields (bytes)
  uint32_t image_bytes;
  uint32_t idesc;
};

__global__ void __launch_bounds__(128) layout_kernel(const uint8_t* image, ProbeArgs args, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    tmem_alloc(&tmem_slot, 64);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (uint32_t i = tid * 16; i < args.image_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = tmem_slot;
  if (tid == 0) {
    const uint32_t sbase = smem_u32(smem);
    uint64_t da_re = make_smem_desc(sbase + args.a_re_off, args.a_lbo, args.a_sbo);
    uint64_t da_im = make_smem_desc(sbase + args.a_im_off, args.a_lbo, args.a_sbo);
    uint64_t db1 = make_smem_desc(sbase + args.b1_off, args.b_lbo, args.b_sbo);
    uint64_t db2 = make_smem_desc(sbase + args.b2_off, args.b_lbo, args.b_sbo);
    umma_f16_ss(taddr, da_re, db1, args.idesc, 0);
    umma_f16_ss(taddr, da_im, db2, args.idesc, 1);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr + (static_cast<uint32_t>(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, 64);
}

// ---------------------------------------------------------------- rate probes
// reps UMMAs of shape 128 x N x 16 over `ntile` distinct A tiles, one commit.
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int reps, int ntile, uint32_t a_tile_bytes,
                                                       uint32_t a_lbo, uint32_t a_sbo, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (uint32_t i = tid * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc = make_idesc_f16(128, N, 1, 0);
    const uint32_t b_off = 150 * 1024;
    uint64_t db = make_smem_desc(sbase + b_off, 128, 256);
    uint64_t da[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) da[t] = make_smem_desc(sbase + (ntile == 1 ? 0 : t) * a_tile_bytes, a_lbo, a_sbo);
    const uint32_t dstep = (uint32_t)(N & 255);
    t0 = clock64();
    for (int i = 0; i < reps; i += 8) {
#pragma unroll
      for (int t = 0; t < 8; ++t) umma_f16_ss(taddr + ((t * dstep) & 255), da[t], db, idesc, 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, 512);
}

// every warp of the CTA loads 32x32b.x32 `reps` times
__global__ void __launch_bounds__(256) tmem_ld_rate_kernel(int reps, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = tmem_slot;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((i * 32) & 511), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
  }
  long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cycles[0] = t1 - t0;
  sink[tid] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(taddr, 512);
}

// STS.128 / LDS.128 rate: each thread moves 16 B per op, conflict-free.
__global__ void __launch_bounds__(256) smem_rate_kernel(int reps, int do_store, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  uint4 v = make_uint4(tid, tid + 1, tid + 2, tid + 3);
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < reps; ++i) {
    uint32_t off = ((i * 256 + tid) * 16) & (64 * 1024 - 1);
    if (do_store) {
      *reinterpret_cast<uint4*>(smem + off) = v;
      v.x += i;
    } else {
      uint4 w = *reinterpret_cast<const uint4*>(smem + off);
      acc += w.x ^ w.y ^ w.z ^ w.w;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  sink[tid] = __uint_as_float(acc + v.x);
}

static float half_to_float_host(__half h) { return __half2float(h); }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d SMs=%d smem/block optin=%zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);

  // ---------------- layout test
  const int M = 128, N = 32, K = 16;
  std::vector<float> Are(M * K), Aim(M * K), B1(K * N), B2(K * N), ref(M * N);
  srand(7);
  for (auto& v : Are) v = (float)((rand() % 9) - 4);
  for (auto& v : Aim) v = (float)((rand() % 9) - 4);
  for (auto& v : B1) v = (float)((rand() % 5) - 2) * 0.5f;
  for (auto& v : B2) v = (float)((rand() % 5) - 2) * 0.25f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += Are[m * K + k] * B1[k * N + n] + Aim[m * K + k] * B2[k * N + n];
      ref[m * N + n] = s;
    }
  // smem image: A planes MN-major interleave: elem (m,k) at (m>>3)*S_mn + (k>>3)*S_kg + (k&7)*16 + (m&7)*2
  const uint32_t S_mn = 272, S_kg = 128;
  const uint32_t a_plane_bytes = 16 * S_mn;  // 4352
  // B K-major interleave: Bmath[k][n] at (n>>3)*S_ng + (k>>3)*S_kc + (n&7)*16 + (k&7)*2
  const uint32_t S_ng = 256, S_kc = 128;
  const uint32_t b_bytes = 4 * S_ng;
  ProbeArgs args;
  args.a_re_off = 0;
  args.a_im_off = 4608;
  args.b1_off = 9216;
  args.b2_off = 9216 + 1024;
  args.image_bytes = 9216 + 2048;
  std::vector<uint8_t> image(args.image_bytes, 0);
  auto put = [&](uint32_t off, float v) {
    __half h = __float2half(v);
    memcpy(&image[off], &h, 2);
  };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      uint32_t o = (m >> 3) * S_mn + (k >> 3) * S_kg + (k & 7) * 16 + (m & 7) * 2;
      put(args.a_re_off + o, Are[m * K + k]);
      put(args.a_im_off + o, Aim[m * K + k]);
    }
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      uint32_t o = (n >> 3) * S_ng + (k >> 3) * S_kc + (n & 7) * 16 + (k & 7) * 2;
      put(args.b1_off + o, B1[k * N + n]);
      put(args.b2_off + o, B2[k * N + n]);
    }
  (void)a_plane_bytes;
  (void)b_bytes;
  uint8_t* d_image;
  float* d_out;
  CK(cudaMalloc(&d_image, image.size()));
  CK(cudaMalloc(&d_out, M * N * sizeof(float)));
  CK(cudaMemcpy(d_image, image.data(), image.size(), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  args.idesc = make_idesc_f16(128, 32, /*a_mn=*/1, /*b_mn=*/0);
  for (int combo = 0; combo < 4; ++combo) {
    bool a_swap = combo & 1, b_swap = combo & 2;
    // hypothesis "plain": A: SBO = MN-chunk stride, LBO = K-group stride; B: SBO = N-group stride, LBO = K-chunk stride
    args.a_sbo = a_swap ? S_kg : S_mn;
    args.a_lbo = a_swap ? S_mn : S_kg;
    args.b_sbo = b_swap ? S_kc : S_ng;
    args.b_lbo = b_swap ? S_ng : S_kc;
    CK(cudaMemset(d_out, 0, M * N * sizeof(float)));
    layout_kernel<<<1, 128, 32 * 1024>>>(d_image, args, d_out);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(M * N);
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
      double e = fabs(out[i] - ref[i]);
      if (e > maxerr) maxerr = e;
      if (e > 1e-3) ++bad;
    }
    printf("LAYOUT combo a_swap=%d b_swap=%d : maxerr=%g bad=%d/%d  %s\n", (int)a_swap, (int)b_swap, maxerr, bad,
           M * N, bad == 0 ? "MATCH" : "mismatch");
    if (bad != 0 && combo == 0) {
      printf("  sample row0: got");
      for (int j = 0; j < 8; ++j) printf(" %g", out[j]);
      printf(" | want");
      for (int j = 0; j < 8; ++j) printf(" %g", ref[j]);
      printf("\n");
    }
  }

  // ---------------- rate probes
  long long* d_cyc;
  float* d_sink;
  CK(cudaMalloc(&d_cyc, 8));
  CK(cudaMalloc(&d_sink, 4096));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int N_ : {32, 64, 128, 256}) {
    for (int ntile : {1, 8}) {
      for (int pass = 0; pass < 2; ++pass) {
        int reps = 256;
        mma_rate_kernel<<<1, 128, 200 * 1024>>>(N_, reps, ntile, 16 * 272, 128, 272, d_cyc);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        if (pass == 1)
          printf("RATE umma M=128 N=%3d K=16 ntile=%d : %lld cycles / %d = %.1f cyc per UMMA\n", N_, ntile, c, reps,
                 (double)c / reps);
      }
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    int reps = 512;
    tmem_ld_rate_kernel<<<1, 256>>>(reps, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    if (pass == 1)
      printf("RATE tcgen05.ld 32x32b.x32, 8 warps: %lld cycles / %d reps = %.1f cyc per rep (8 warps x 4 KiB)\n", c, reps,
             (double)c / reps);
  }
  CK(cudaFuncSetAttribute(smem_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int st = 0; st < 2; ++st)
    for (int pass = 0; pass < 2; ++pass) {
      int reps = 1024;
      smem_rate_kernel<<<1, 256, 64 * 1024>>>(reps, st, d_cyc, d_sink);
      CK(cudaDeviceSynchronize());
      long long c;
      CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
      if (pass == 1)
        printf("RATE %s.128 256 thr: %lld cycles / %d reps = %.1f cyc per 4 KiB\n", st ? "STS" : "LDS", c, reps,
               (double)c / reps);
    }
  printf("probe done\n");
  return 0;
}
