// Hardware probe (developer tool): request rates of column-mode global traffic on one SM.
// A "plane unit" is W adjacent fp16 columns x 4096 rows of a [batch][4096][4096] array (row pitch 8 KiB), i.e. 4096
// pieces of 2W bytes, each on its own 128-byte line -- the access pattern of the 4096-point column passes (four-step
// 2^22 / 2^24, 2-D column pass).  148 persistent CTAs, one per SM, stream plane units in the kernel's unit order
// (unit = blockIdx.x + k * gridDim.x).  Modes:
//   0  TMA tile loads only (two 64 KiB buffers, always one load in flight behind the one being waited for)
//   1  TMA tile stores only (bulk groups, one store in flight behind the one being waited for)
//   2  LSU stores only: every thread stores 16-byte pieces (st.global.v4), one line per lane
//   3  TMA loads and TMA stores at the same time (independent issuing threads)
//   4  TMA loads and LSU stores at the same time
//   5  TMA stores and LSU stores at the same time, each on half of the units (split store path)
// Prints us per plane unit and SM, pieces per clock and SM, aggregate GB/s.
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

constexpr int kRows = 4096, kCols = 4096, kBatch = 32;
constexpr int kThreads = 512;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(0), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, uint32_t c0, uint32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(0), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void stg128(__half* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// W columns per unit; rows_per_box * W * 2 = 64 KiB
template <int MODE, int W>
__global__ void __launch_bounds__(kThreads, 1)
probe(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, __half* out, uint32_t n_units) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x;
  constexpr uint32_t kBoxRows = 65536 / (2 * W);          // rows per 64 KiB box
  constexpr uint32_t kBoxesPerCol = kRows / kBoxRows;      // boxes per column group and batch
  constexpr uint32_t kGroups = kCols / W;
  const uint32_t buf0 = smem_u32(smem), buf1 = buf0 + 65536, buf2 = buf0 + 131072;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_mbar_init(); }
  __syncthreads();
  // unit -> (column group, batch*boxes): consecutive units are adjacent column groups
  auto c0_of = [&](uint32_t u) { return (u % kGroups) * W; };
  auto c2_of = [&](uint32_t u) { return (u / kGroups) * (kBoxRows / 256); };
  const bool do_load = MODE == 0 || MODE == 3 || MODE == 4;
  const bool do_tstore = MODE == 1 || MODE == 3 || MODE == 5;
  const bool do_lstore = MODE == 2 || MODE == 4 || MODE == 5;
  if (do_load && tid == 0) {
    uint32_t ph[2] = {0, 0};
    uint32_t q = 0;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++q) {
      const uint32_t b = q & 1u;
      mbar_arrive_expect_tx(bar + b, 65536);
      tma_load_3d(b ? buf1 : buf0, &tin, c0_of(u), c2_of(u), bar + b);
      if (q >= 1) { mbar_wait(bar + (b ^ 1u), ph[b ^ 1u] & 1u); ph[b ^ 1u]++; }
    }
    if (q >= 1) { const uint32_t b = (q - 1) & 1u; mbar_wait(bar + b, ph[b] & 1u); }
  }
  if (do_tstore && tid == 32) {
    uint32_t q = 0;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++q) {
      if (MODE == 5 && (q & 1u)) continue;
      tma_store_3d(&tout, buf2, c0_of(u), c2_of(u));
      bulk_commit();
      bulk_wait_read<1>();
    }
    bulk_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (do_lstore && tid >= 64) {
    const int t = tid - 64, nt = kThreads - 64;
    uint32_t q = 0;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++q) {
      if (MODE == 5 && !(q & 1u)) continue;
      const uint32_t g = u % kGroups, bb = u / kGroups;   // bb: batch * boxes
      __half* base = out + static_cast<size_t>(bb) * kBoxRows * kCols + g * W;
      const uint4 v = make_uint4(u, t, 0, 0);
      // pieces of 16 bytes: W/8 per row
      for (uint32_t p = t; p < kBoxRows * (W / 8); p += nt)
        stg128(base + static_cast<size_t>(p / (W / 8)) * kCols + (p % (W / 8)) * 8, v);
    }
  }
  __syncthreads();
}


// In-place pipeline: load unit u (TMA tile, W_IN columns) into a ring of two buffers, store the same bytes back
// (STORE 0: TMA tile store, 1: LSU 16-byte stores after 16-byte shared loads) to `tout`/out with W_OUT columns per unit
// (in place when the caller passes the same array and W_IN == W_OUT).  SERIAL: no overlap (the next load is requested only
// after the store of the previous unit has finished reading shared memory), as in the single-buffer kernel.
template <int STORE, int W_IN, int W_OUT, bool SERIAL>
__global__ void __launch_bounds__(kThreads, 1)
pipe_probe(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, __half* out, uint32_t n_units) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2], empty[2];
  const int tid = threadIdx.x;
  constexpr uint32_t kInRows = 65536 / (2 * W_IN), kOutRows = 65536 / (2 * W_OUT);
  constexpr uint32_t kInGroups = kCols / W_IN, kOutGroups = kCols / W_OUT;
  const uint32_t buf[2] = {smem_u32(smem), smem_u32(smem) + 65536};
  if (tid == 0) { mbar_init(full, 1); mbar_init(full + 1, 1); mbar_init(empty, 1); mbar_init(empty + 1, 1); fence_mbar_init(); }
  __syncthreads();
  constexpr uint32_t kRing = SERIAL ? 1 : 2;
  if (tid == 0) {
    uint32_t q = 0;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++q) {
      const uint32_t b = q % kRing, use = q / kRing;
      if (use >= 1) mbar_wait(empty + b, (use - 1) & 1u);
      mbar_arrive_expect_tx(full + b, 65536);
      tma_load_3d(buf[b], &tin, (u % kInGroups) * W_IN, (u / kInGroups) * (kInRows / 256), full + b);
    }
  } else if (STORE == 0 && tid == 32) {
    uint32_t q = 0;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++q) {
      const uint32_t b = q % kRing, use = q / kRing;
      mbar_wait(full + b, use & 1u);
      tma_store_3d(&tout, buf[b], (u % kOutGroups) * W_OUT, (u / kOutGroups) * (kOutRows / 256));
      bulk_commit();
      bulk_wait_read<0>();
      mbar_arrive(empty + b);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (STORE == 1 && tid >= 64) {
    const int t = tid - 64, nt = kThreads - 64;
    uint32_t q = 0;
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x, ++q) {
      const uint32_t b = q % kRing, use = q / kRing;
      mbar_wait(full + b, use & 1u);
      const uint32_t g = u % kOutGroups, bb = u / kOutGroups;
      __half* base = out + static_cast<size_t>(bb) * kOutRows * kCols + g * W_OUT;
      for (uint32_t p = t; p < 4096; p += nt) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(buf[b] + p * 16));
        stg128(base + static_cast<size_t>(p / (W_OUT / 8)) * kCols + (p % (W_OUT / 8)) * 8, v);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kThreads - 64) : "memory");
      if (t == 0) mbar_arrive(empty + b);
    }
  }
  __syncthreads();
}

template <int STORE, int W_IN, int W_OUT, bool SERIAL>
static void run_pipe(__half* din, __half* dout, const char* what, double clock_ghz);

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static Enc g_enc;

template <int W>
static CUtensorMap make_map(__half* p) {
  CUtensorMap tm;
  const cuuint64_t rows_total = (cuuint64_t)kRows * kBatch;
  cuuint64_t gdim[3] = {(cuuint64_t)kCols, 256, rows_total / 256};
  cuuint64_t gstr[2] = {(cuuint64_t)kCols * 2, (cuuint64_t)kCols * 2 * 256};
  cuuint32_t box[3] = {(cuuint32_t)W, 256, (cuuint32_t)(65536 / (2 * W) / 256)};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, p, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode rc = %d\n", (int)r); exit(1); }
  return tm;
}

template <int MODE, int W>
static void run(__half* din, __half* dout, const char* what, double clock_ghz) {
  const CUtensorMap tin = make_map<W>(din), tout = make_map<W>(dout);
  constexpr uint32_t kBoxRows = 65536 / (2 * W);
  const uint32_t n_units = (uint32_t)((size_t)kRows * kBatch / kBoxRows * (kCols / W));
  auto fn = probe<MODE, W>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    fn<<<148, kThreads, 3 * 65536>>>(tin, tout, dout, n_units);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double units_per_sm = n_units / 148.0;
  const double us_unit = best * 1e3 / units_per_sm;
  const bool both = MODE == 3 || MODE == 4;
  const double bytes = (double)n_units * 65536 * (both ? 2 : 1);
  const double pieces = kBoxRows;   // pieces (lines touched) per unit and direction
  printf("mode %d W=%2d %-44s %8.3f ms  %6.2f us/unit/SM  %5.2f pieces/clk/SM/dir  %7.1f GB/s total\n", MODE, W, what, best,
         us_unit, pieces / (us_unit * 1e3 * clock_ghz) * (MODE == 5 ? 1 : 1), bytes / best * 1e-6);
}

template <int STORE, int W_IN, int W_OUT, bool SERIAL>
static void run_pipe(__half* din, __half* dout, const char* what, double clock_ghz) {
  const CUtensorMap tin = make_map<W_IN>(din), tout = make_map<W_OUT>(dout);
  const uint32_t n_units = 16384;
  auto fn = pipe_probe<STORE, W_IN, W_OUT, SERIAL>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    fn<<<148, kThreads, 2 * 65536>>>(tin, tout, dout, n_units);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double us_unit = best * 1e3 / (n_units / 148.0);
  printf("pipe %s in W=%2d out W=%2d %s %-34s %8.3f ms  %6.2f us/unit/SM  %7.1f GB/s (read+write)\n", STORE ? "LSU" : "TMA", W_IN,
         W_OUT, din == dout ? "in-place" : "2 arrays", what, best, us_unit, 2.0 * n_units * 65536 / best * 1e-6);
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  g_enc = (Enc)fn;
  const size_t halves = (size_t)kRows * kCols * kBatch;
  __half *din, *dout;
  CK(cudaMalloc(&din, halves * 2)); CK(cudaMalloc(&dout, halves * 2));
  CK(cudaMemset(din, 0, halves * 2)); CK(cudaMemset(dout, 0, halves * 2));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6;
  printf("array %zu MiB per direction, SM clock %.3f GHz (nominal)\n", halves * 2 >> 20, ghz);
  run<0, 8>(din, dout, "TMA load", ghz);
  run<1, 8>(din, dout, "TMA store", ghz);
  run<2, 8>(din, dout, "LSU store (16-byte pieces)", ghz);
  run<3, 8>(din, dout, "TMA load + TMA store", ghz);
  run<4, 8>(din, dout, "TMA load + LSU store", ghz);
  run<5, 8>(din, dout, "TMA store (even units) + LSU store (odd)", ghz);
  run<0, 16>(din, dout, "TMA load", ghz);
  run<1, 16>(din, dout, "TMA store", ghz);
  run<2, 16>(din, dout, "LSU store (2 x 16-byte pieces per row)", ghz);
  run<3, 16>(din, dout, "TMA load + TMA store", ghz);
  run<4, 16>(din, dout, "TMA load + LSU store", ghz);
  run<0, 64>(din, dout, "TMA load", ghz);
  run<1, 64>(din, dout, "TMA store", ghz);
  run<3, 64>(din, dout, "TMA load + TMA store", ghz);
  run_pipe<0, 8, 8, false>(din, din, "overlapped", ghz);
  run_pipe<1, 8, 8, false>(din, din, "overlapped", ghz);
  run_pipe<0, 8, 8, true>(din, din, "serial", ghz);
  run_pipe<1, 8, 8, true>(din, din, "serial", ghz);
  run_pipe<0, 8, 8, false>(din, dout, "overlapped", ghz);
  run_pipe<1, 8, 8, false>(din, dout, "overlapped", ghz);
  run_pipe<0, 16, 16, false>(din, din, "overlapped", ghz);
  run_pipe<1, 16, 16, false>(din, din, "overlapped", ghz);
  run_pipe<0, 16, 16, true>(din, din, "serial", ghz);
  run_pipe<0, 64, 8, false>(din, dout, "overlapped (pass-2 pattern)", ghz);
  run_pipe<1, 64, 8, false>(din, dout, "overlapped (pass-2 pattern)", ghz);
  run_pipe<0, 64, 16, false>(din, dout, "overlapped", ghz);
  run_pipe<0, 64, 64, false>(din, dout, "overlapped", ghz);
  run_pipe<0, 64, 64, false>(din, din, "overlapped", ghz);
  return 0;
}
