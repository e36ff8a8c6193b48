// C ABI of the B200-native fp16 FFT (include/tfft.h): plan construction, pass scheduling and
// kernel launches.  Host logic mirrors the reference's plan/execute split
// (src/base/Plan.h:77-194 CreatePlan, src/base/ComputeFFT.h:54-293 ComputeFFT) but one exec is
// one kernel launch per HBM pass for the WHOLE batch (the reference issues
// batch x (1 + r16 + 2^r2 - 1) launches on `batch` freshly created streams,
// ComputeFFT.h:167-284).  There is no CPU path: every exec needs an sm_100 device.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/tfft.h"
#include "fft_unit_kernel.cuh"
#include "kernel_table.h"
#include "harness_kernels.cuh"
#include "exchange_kernels.cuh"

namespace {

using namespace tfft;

// Everything a launch needs besides the destination pointers.  Building it costs two cuTensorMapEncodeTiled calls, the
// stride fill and an occupancy query; it only depends on (device, source pointers, strides, fused-twiddle arguments),
// so each pass keeps the last few in a small cache (a plan that is executed again on the same buffers -- the
// reference's DataHandler usage, src/base/DataHandler.h:22-82 -- pays for it once).
struct Prepared {
  int dev = -1;
  const void *sre = nullptr, *sim = nullptr;
  int64_t in_stride = 0, out_stride = 0, tw_first_col = 0, seg_stride = 0;
  int tw_log2 = 0, segs = 0;
  UnitPlan plan;
  alignas(64) CUtensorMap tmap_re, tmap_im;
  const void* fn = nullptr;
  const uint4* tables = nullptr;
  unsigned grid = 0, block = 0;
  uint32_t smem = 0;
  bool two_slot = false, cluster = false;   // two_slot: the two-slot calling convention (two-slot and landing-ring kernels)
};
constexpr size_t kPreparedSlots = 4;

struct Pass {
  UnitPlan plan;           // layout decisions (strides filled per exec)
  PlanBuildInfo info;
  UnitStrides strides;     // plan-time part of the addressing
  uint32_t n_units = 0;
  std::vector<uint8_t> tables;            // host image
  mutable uint4* d_tables[16] = {};       // per-device copies, uploaded at first use
  mutable int resident_ctas[16] = {};     // persistent grid size per device
  uint32_t smem = 0;
  int src = 0, dst = 0;    // 0 = user input planes, 1 = user output planes, 2 = plan workspace
  bool in_stride_is_user = false, out_stride_is_user = false;
  bool own_batch_strides = false;   // three-pass plans: the batch level of the unit addressing is internal
  bool row5 = false;                // last pass of a three-pass plan with TMA row tiles (5-D map, UnitPlan::tma_row5)
  bool outer_batch = false;         // batched three-pass plans: the transforms of the batch are an outer level of the unit
                                    // index (UnitPlan::b3_shift), in/out stride apart
  bool il_in = false, il_out = false;   // TFFT_INTERLEAVED: the pass reads / writes half2 elements
  int kind = 0;                     // 0: 1-D passes; 1: 2-D row pass (row mode, unit = U rows of one image); 2: 2-D column pass
  mutable std::vector<Prepared> prepared;   // launch cache (see prepare_launch), guarded by g_upload_mutex
  mutable size_t prepared_next = 0;
};

int ilog2_exact(int64_t n) {
  if (n <= 0 || (n & (n - 1))) return -1;
  int l = 0;
  while ((int64_t(1) << l) < n) ++l;
  return l;
}

// Host image of the per-plan constant tables the kernel stages into shared memory
// (layout: fft_unit_kernel.cuh "Device tables"): two-level twiddle table for unit angle 2*pi/L and
// the fp16 DFT matrices [Fr|Fi]/R, [-Fi|Fr]/R of every radix the plan uses.
std::vector<uint8_t> make_tables(const UnitPlan& plan, bool unscaled) {
  const TableLayout TL = table_layout(plan);
  std::vector<uint8_t> host(TL.total, 0);
  const double pi = 3.14159265358979323846264338327950288;
  const int64_t L = int64_t(1) << plan.log2_len;
  auto unit = [&](int64_t e, int64_t n, double* c, double* s) {   // exp(-2*pi*i*e/n), exact on the axes
    e %= n;
    if (e == 0) { *c = 1; *s = 0; return; }
    if (4 * e == n) { *c = 0; *s = -1; return; }
    if (2 * e == n) { *c = -1; *s = 0; return; }
    if (4 * e == 3 * n) { *c = 0; *s = 1; return; }
    const double a = -2.0 * pi * static_cast<double>(e) / static_cast<double>(n);
    *c = std::cos(a); *s = std::sin(a);
  };
  float2* tw = reinterpret_cast<float2*>(host.data());
  for (int j = 0; j < 64 && tw_table_bytes(plan); ++j) {
    double c, s;
    unit(j, L, &c, &s);
    tw[j] = make_float2(static_cast<float>(c), static_cast<float>(s));
  }
  for (int64_t j = 0; j < 512 && tw_table_bytes(plan); ++j) {
    double c, s;
    unit((64 * j) % L, L, &c, &s);
    tw[64 + j] = make_float2(static_cast<float>(c), static_cast<float>(s));
  }
  for (uint32_t t = 0; t < plan.stages; ++t) {
    const int rho = static_cast<int>(plan.log2_radix[t]), R = 1 << rho;
    __half* b1 = reinterpret_cast<__half*>(host.data() + TL.b_off[t]);
    // Kronecker last stage (2-D row pass): K index = (kappa_x low, kappa_y high), output column = (k_x low, k_y high),
    // F = F_x[kappa_x][k_x] * F_y[kappa_y][k_y]
    const int ybits = (plan.kron_bits && t + 1 == plan.stages) ? static_cast<int>(plan.kron_bits) : 0;
    const int Rx = R >> ybits, Ry = 1 << ybits;
#ifdef TFFT_TWO_MATRICES
    const int ncols = 4 * R;   // [Fr | Fi] then [-Fi | Fr], each 2R columns
#else
    const int ncols = 3 * R;   // [Fr | Fi | -Fr]
#endif
    for (int kap = 0; kap < R; ++kap)
      for (int n = 0; n < ncols; ++n) {
        double c, s;
        const int k = n % R;
        // phase / R = kap_x*k_x/Rx + kap_y*k_y/Ry
        unit(static_cast<int64_t>(kap % Rx) * (k % Rx) * Ry + static_cast<int64_t>(kap / Rx) * (k / Rx) * Rx, R, &c, &s);
        // 1/R per stage = the reference's "sequential scaling"; TFFT_UNSCALED (cuFFT convention) leaves it out
        const double sc = unscaled ? 1.0 : 1.0 / R;
        const float fr = static_cast<float>(c * sc), fi = static_cast<float>(s * sc);
#ifdef TFFT_TWO_MATRICES
        const int m = n / (2 * R), nn = n % (2 * R);
        const float v = m == 0 ? (nn < R ? fr : fi) : (nn < R ? -fi : fr);
        const uint32_t off = static_cast<uint32_t>(m) * 2 * R * R + (nn >> 3) * (8 * R) + (kap >> 3) * 64 + (nn & 7) * 8 + (kap & 7);
#else
        const float v = n < R ? fr : (n < 2 * R ? fi : -fr);
        const uint32_t off = (n >> 3) * (8 * R) + (kap >> 3) * 64 + (n & 7) * 8 + (kap & 7);   // in halves
#endif
        b1[off] = __float2half_rn(v);
      }
  }
  return host;
}

long long* g_trace = nullptr;   // developer phase trace buffer (tfft_debug_set_trace)
std::mutex g_upload_mutex;

// Developer knobs (A/B switches used by tools/ and a few tests) are environment variables that are honoured ONLY when
// TFFT_DEVELOPER is set: a production process cannot be steered by a stray TFFT_* variable.  The supported ways to
// configure a plan are the flags of tfft_plan_create and the tuner file (tfft_plan_create_from_file / TFFT_TUNER_FILE).
const char* dev_env(const char* name) {
  static const bool on = getenv("TFFT_DEVELOPER") != nullptr;
  return on ? getenv(name) : nullptr;
}
// Kernel instantiations: one per (unit size, radix schedule) the planner can produce (kernel_table.h / kernel_group.cu).
KernelFn kernel_for(const UnitPlan& p, int* threads) {
  if (p.cluster) {   // CTA-pair units: 512-thread kernels of kernel_group.cu group 5
    int count = 0;
    const ClusterEntry* e = kernel_cluster_group(&count);
    for (int i = 0; i < count; ++i)
      if (e[i].r0 == static_cast<int>(p.log2_radix[0]) && e[i].r1 == static_cast<int>(p.log2_radix[1]) &&
          e[i].r2 == static_cast<int>(p.stages == 3 ? p.log2_radix[2] : 0) && e[i].lm == static_cast<int>(p.tma_load)) {
        *threads = 512;
        return e[i].fn;
      }
    return nullptr;
  }
  static const bool narrow = dev_env("TFFT_NARROW_32K") != nullptr;   // developer A/B: 256-thread CTAs for 32K-element units
  typedef const KernelEntry* (*GroupFn)(int*);
  static const GroupFn groups[kKernelGroups] = {kernel_group_0, kernel_group_1, kernel_group_2, kernel_group_3, kernel_group_4};
  for (GroupFn g : groups) {
    int count = 0;
    const KernelEntry* e = g(&count);
    for (int i = 0; i < count; ++i) {
      const KernelEntry& k = e[i];
      if (k.log2e == static_cast<int>(p.log2_elems) && k.r0 == static_cast<int>(p.log2_radix[0]) &&
          k.r1 == static_cast<int>(p.log2_radix[1]) && k.r2 == static_cast<int>(p.stages == 3 ? p.log2_radix[2] : 0) &&
          !(narrow && k.threads == 512)) {
        *threads = k.threads;
        // tma_load 2 / 4: the entry's column-tile kernel; 1 / 3: its row-tile kernel
        return (p.tma_load == 5 || p.tma_load == 6) ? k.fn_tma_col64
               : (p.tma_load == 2 || p.tma_load == 4) ? k.fn_tma_col : (p.tma_load ? k.fn_tma : k.fn);
      }
    }
  }
  return nullptr;
}

// Two-slot variant (one CTA per SM, two units in flight, shared landing buffer) for 16K-element units
Kernel2Fn kernel2_for(const UnitPlan& p, bool allowed = true) {
  if (!allowed || p.cluster || p.tma_load != 1 || p.log2_elems != 14 || p.stages != 3) return nullptr;
  if (smem2_layout(p).total > 227 * 1024) return nullptr;   // e.g. three distinct DFT matrices
  int count = 0;
  const Kernel2Entry* e = kernel2_group(&count);
  for (int i = 0; i < count; ++i)
    if (e[i].r0 == static_cast<int>(p.log2_radix[0]) && e[i].r1 == static_cast<int>(p.log2_radix[1]) &&
        e[i].r2 == static_cast<int>(p.log2_radix[2]))
      return e[i].fn;
  return nullptr;
}

// Landing-ring variant (32K-element units, stage-1 operand in a separate ring; fft_unit_kernel_ring)
Kernel2Fn kernel_ring_for(const UnitPlan& p) {
  if (!p.ring) return nullptr;
  int count = 0;
  const RingEntry* e = kernel_ring_group(&count);
  for (int i = 0; i < count; ++i)
    if (e[i].r0 == static_cast<int>(p.log2_radix[0]) && e[i].r1 == static_cast<int>(p.log2_radix[1]) &&
        e[i].r2 == static_cast<int>(p.stages == 3 ? p.log2_radix[2] : 0) && e[i].lm == static_cast<int>(p.tma_load))
      return e[i].fn;
  return nullptr;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// Tensor map of one input plane for the TMA load: dims (fastest first) {64 rows, R kappa (stride M),
// M/64 (stride 64), transforms (stride tstride)}, box {64, R, M/64, U}, 128-byte swizzle.
int make_input_tensor_map(const UnitPlan& plan, const __half* base, int64_t tstride, int64_t n_transforms,
                          CUtensorMap* out, bool half_box = false) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return TFFT_E_UNSUPPORTED;
  const uint64_t L = uint64_t(1) << plan.log2_len, R = uint64_t(1) << plan.log2_radix[0], M = L / R;
  const uint64_t U = uint64_t(1) << plan.log2_units;
  cuuint64_t gdim[4] = {64, R, M / 64, static_cast<cuuint64_t>(n_transforms)};
  cuuint64_t gstride[3] = {M * 2, 128, static_cast<cuuint64_t>(tstride) * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(R), static_cast<cuuint32_t>(M / 64), static_cast<cuuint32_t>(U)};
  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
  if (plan.tma_load == 3) {   // M = 16 / 32 rows per K line: atoms of 16 rows, 32-byte swizzle
    gdim[0] = 16; gdim[2] = M / 16;
    gstride[1] = 32;
    box[0] = 16; box[2] = static_cast<cuuint32_t>(M / 16);
    swz = CU_TENSOR_MAP_SWIZZLE_32B;
  }
  if (half_box) {   // two-slot kernel: one box = half of a unit's stage-1 tiles
    if (U >= 2) box[3] = static_cast<cuuint32_t>(U / 2);
    else box[2] = static_cast<cuuint32_t>(M / 128);
  }
  if (plan.cluster) box[2] = static_cast<cuuint32_t>(M / 128);   // a cluster CTA loads the rows of one half of m
  if (plan.ring) {   // landing-ring units: one box = a quarter of the unit's stage-1 tiles
    if (U >= 4) box[3] = static_cast<cuuint32_t>(U / 4);
    else box[2] = static_cast<cuuint32_t>(M / 256);
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TFFT_OK : TFFT_E_INVALID_ARG;
}

// Row tiles of the last pass of a three-pass plan: transform t (a row of the N1 x N2 matrix, tstride = N2 apart) and column
// block cb (block_stride = Nb apart) start at base + t*tstride + cb*block_stride; element n = kappa*M + m of the length-Nb
// transform.  Dims {16 | 64 (m low), R kappa (stride M), M/16 | M/64, column blocks, transforms}, box {.., 1, U}: lands
// exactly like the 4-D tile of a contiguous batch.
int make_row5_tensor_map(const UnitPlan& plan, const __half* base, int64_t tstride, int64_t n_transforms, int64_t blocks,
                         int64_t block_stride, CUtensorMap* out) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return TFFT_E_UNSUPPORTED;
  const uint64_t L = uint64_t(1) << plan.log2_len, R = uint64_t(1) << plan.log2_radix[0], M = L / R;
  const uint64_t U = uint64_t(1) << plan.log2_units;
  const uint64_t atom = plan.tma_load == 3 ? 16 : 64;
  cuuint64_t gdim[5] = {atom, R, M / atom, static_cast<cuuint64_t>(blocks), static_cast<cuuint64_t>(n_transforms)};
  cuuint64_t gstride[4] = {M * 2, atom * 2, static_cast<cuuint64_t>(block_stride) * 2, static_cast<cuuint64_t>(tstride) * 2};
  cuuint32_t box[5] = {static_cast<cuuint32_t>(atom), static_cast<cuuint32_t>(R), static_cast<cuuint32_t>(M / atom), 1,
                       static_cast<cuuint32_t>(U)};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      plan.tma_load == 3 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TFFT_OK : TFFT_E_INVALID_ARG;
}

// Segmented input (tfft_exec_segmented): element n of transform t lies at base + (n / seg_len) * seg_stride + t * tstride +
// n % seg_len with seg_len = L / segs.  With n = kappa*M + m (M = L/R rows per K line, seg_len = kl*M, kl = R/segs):
// dims {64 (m low), kl (kappa low, stride M), segs (stride seg_stride), M/64 (stride 64), transforms (stride tstride)} --
// the box lands in shared memory exactly like the 4-D tile {64, R, M/64, U} because (kappa_lo, segment) enumerate kappa.
int make_seg_tensor_map(const UnitPlan& plan, const __half* base, int64_t tstride, int64_t n_transforms, int segs,
                        int64_t seg_stride, CUtensorMap* out, bool half_box) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return TFFT_E_UNSUPPORTED;
  const uint64_t L = uint64_t(1) << plan.log2_len, R = uint64_t(1) << plan.log2_radix[0], M = L / R;
  const uint64_t U = uint64_t(1) << plan.log2_units, kl = R / static_cast<uint64_t>(segs);
  cuuint64_t gdim[5] = {64, kl, static_cast<cuuint64_t>(segs), M / 64, static_cast<cuuint64_t>(n_transforms)};
  cuuint64_t gstride[4] = {M * 2, static_cast<cuuint64_t>(seg_stride) * 2, 128, static_cast<cuuint64_t>(tstride) * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(kl), static_cast<cuuint32_t>(segs), static_cast<cuuint32_t>(M / 64),
                       static_cast<cuuint32_t>(U)};
  if (half_box) {   // two-slot kernel: one box = half of a unit's stage-1 tiles
    if (U >= 2) box[4] = static_cast<cuuint32_t>(U / 2);
    else box[3] = static_cast<cuuint32_t>(M / 128);
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TFFT_OK : TFFT_E_INVALID_ARG;
}

// Kronecker units (2-D row pass): unit (image b, y_lo) reads rows y_lo + u*(ny/U), u < U, of image b.  The image
// stride is batch_step * (ny/U) rows (U for contiguous images, 2U for the [RE_b | IM_b] layout), so
// row = y_lo + (ny/U)*(u + batch_step*b): dims {64, R, M/64, y_lo (stride nx), u + batch_step*b}.
int make_kron_tensor_map(const UnitPlan& plan, const __half* base, int64_t ny, int64_t batch, int64_t batch_step,
                         CUtensorMap* out, bool half_box) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return TFFT_E_UNSUPPORTED;
  const uint64_t L = uint64_t(1) << plan.log2_len, R = uint64_t(1) << plan.log2_radix[0], M = L / R;
  const uint64_t U = uint64_t(1) << plan.log2_units;
  cuuint64_t gdim[5] = {64, R, M / 64, static_cast<cuuint64_t>(ny) / U,
                        static_cast<cuuint64_t>(batch_step) * static_cast<cuuint64_t>(batch - 1) + U};
  cuuint64_t gstride[4] = {M * 2, 128, L * 2, (static_cast<cuuint64_t>(ny) / U) * L * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(R), static_cast<cuuint32_t>(M / 64), 1,
                       static_cast<cuuint32_t>(half_box ? U / 2 : U)};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TFFT_OK : TFFT_E_INVALID_ARG;
}

// Column-mode input (four-step column pass, 2-D column pass): element n of column c of batch b at
// base + b*batch_stride + c + n*nstride, n = kappa*M + m.  Dims {columns, R kappa (stride M*nstride), M rows (stride
// nstride), batches}; box {8, R, M, 1}, no swizzle: the tile lands as dense 16-byte chunks [m][kappa][8 columns].
// outer > 0 (batched three-pass pass B): a fifth dimension of `outer` user transforms, outer_stride apart.
int make_col_tensor_map(const UnitPlan& plan, const __half* base, int64_t nstride, int64_t columns, int64_t batches,
                        int64_t batch_stride, CUtensorMap* out, int64_t outer = 0, int64_t outer_stride = 0) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return TFFT_E_UNSUPPORTED;
  const uint64_t L = uint64_t(1) << plan.log2_len, R = uint64_t(1) << plan.log2_radix[0], M = L / R;
  if (batches <= 1 || batch_stride <= 0) batch_stride = static_cast<int64_t>(L) * nstride;
  cuuint64_t gdim[5] = {static_cast<cuuint64_t>(columns), R, M, static_cast<cuuint64_t>(batches < 1 ? 1 : batches),
                        static_cast<cuuint64_t>(outer < 1 ? 1 : outer)};
  cuuint64_t gstride[4] = {M * static_cast<cuuint64_t>(nstride) * 2, static_cast<cuuint64_t>(nstride) * 2,
                           static_cast<cuuint64_t>(batch_stride) * 2,
                           static_cast<cuuint64_t>(outer_stride > 0 ? outer_stride : batch_stride) * 2};
  // mode 2: 8-column tiles, dense; mode 4: 16-column tiles (whole 32-byte sectors) as SWIZZLE_32B atoms
  cuuint32_t box[5] = {plan.tma_load == 5 ? 64u : plan.tma_load == 6 ? 32u : plan.tma_load == 4 ? 16u : 8u, static_cast<cuuint32_t>(R),
                       static_cast<cuuint32_t>(plan.cluster ? M / 2 : plan.ring ? M / 4 : M), 1, 1};   // cluster CTA: one half of m; ring unit: quarters
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, outer > 0 ? 5 : 4, const_cast<__half*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      plan.tma_load == 5 ? CU_TENSOR_MAP_SWIZZLE_128B : plan.tma_load == 6 ? CU_TENSOR_MAP_SWIZZLE_64B
                      : plan.tma_load == 4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TFFT_OK : TFFT_E_INVALID_ARG;
}

// All kernels are launched with programmatic stream serialization: a kernel's CTAs may become resident while its
// predecessor in the stream drains; the kernels themselves wait (griddepcontrol.wait) before they touch data.
cudaError_t launch_pdl(const void* fn, unsigned grid, unsigned block, void** args, size_t smem, cudaStream_t stream,
                       bool cluster = false) {
  static const bool no_pdl = dev_env("TFFT_NO_PDL") != nullptr;
  if (no_pdl && !cluster) return cudaLaunchKernel(fn, dim3(grid), dim3(block), args, smem, stream);
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster) {   // CTA-pair units (UnitPlan::cluster)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelExC(&cfg, fn, args);
}

}  // namespace

// Per-plan tuning knobs (tfft_plan_create_tuned / a tuner file); -1 = default (the environment variable of the same
// purpose, then the built-in choice).
struct Tuning {
  int tma = -1;            // TMA tile loads for row-mode units            (TFFT_NO_TMA)
  int pipe = -1;           // MMA / epilogue overlap of stage 2            (TFFT_NO_PIPE)
  int two_slot = -1;       // two-slot kernel for 16K-element TMA units    (TFFT_NO_2SLOT)
  int prefetch = -1;       // L2 prefetch of the next unit                 (TFFT_PREFETCH)
  int fourstep_lg1 = -1;   // log2 of the column-pass length, N > 2^15     (TFFT_FOURSTEP_LG1)
  int tma_col = -1;        // TMA column tiles in four-step column passes  (TFFT_NO_TMA_COL)
  int cluster = -1;        // CTA-pair units: N = 65536 in one pass, 16-column units for 4096-point column passes (off unless 1)
  int ring = -1;           // landing-ring kernel for 32K-element TMA units (TFFT_NO_RING)
};
int knob(int tuned, const char* env_off, int dflt) {   // env_off: variable whose presence switches the feature off
  if (tuned >= 0) return tuned;
  if (env_off && dev_env(env_off)) return 0;
  return dflt;
}

struct tfft_plan_s {
  Tuning tune;
  int64_t n = 0, batch = 0, ny = 0, nx = 0;
  uint32_t flags = 0;
  int lg = 0;
  std::vector<Pass> passes;
  std::vector<Pass> passes_strided;   // 2-D plans: the same passes without TMA loads, built at the first exec whose
                                      // images are not contiguous (the tensor maps need image stride == ny*nx)
  int ybits = 0;                      // 2-D plans: rows per Kronecker unit = 2^ybits (0: plain row pass)
  __half* workspace = nullptr;   // 2 * n * batch halves when TFFT_PRESERVE_INPUT on multi-pass sizes
  int64_t workspace_bytes = 0;
  bool loop_batch = false;           // three-pass plans run one transform at a time
  struct HostPath* host = nullptr;   // tfft_exec_host state, built at the first call
  std::mutex host_mutex;             // tfft_exec_host calls on one plan are serialised
  int device = 0;
};
constexpr int kHostRing = 4;   // device slots per direction of the tfft_exec_host pipeline
struct HostPath {
  __half* buf = nullptr;             // [slots input chunks | slots output chunks]
  std::vector<int64_t> sizes;        // transforms per chunk, in order
  std::map<int64_t, tfft_plan_s*> plans;   // one plan per distinct chunk size (the whole batch as one chunk: the plan itself)
  int64_t chunk = 0;                 // largest chunk = slot size
  int slots = 1;
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};   // upload, transform, download
  cudaEvent_t events[3 * kHostRing] = {};                  // per slot: uploaded, transformed, downloaded
};
static void destroy_host_path(HostPath* h) {
  if (!h) return;
  if (h->buf) cudaFree(h->buf);
  for (auto& kv : h->plans) tfft_plan_destroy(kv.second);
  for (cudaStream_t st : h->streams)
    if (st) cudaStreamDestroy(st);
  for (cudaEvent_t ev : h->events)
    if (ev) cudaEventDestroy(ev);
  delete h;
}

namespace {

// log2 of the unit size (elements resident per CTA) for transforms of length 2^lg: 16K elements
// (256 tensor-memory columns, two CTAs per SM) unless the length or its tables need otherwise.
int unit_log2_elems(int lg) {
  if (lg == 15) return 15;
  // 2048 = 32 * 64: with one 3R-column matrix per radix (30 KiB of tables) two 16K-element CTAs fit an SM; measured on
  // B200 0.389 ms per GiB against 0.434 ms with the 8K-element units of round 1 (developer A/B: TFFT_U2048_8K)
  if (lg == 11) return dev_env("TFFT_U2048_8K") ? 13 : 14;
  return 14;
}
// transforms per unit for a row/row pass, never padding a small batch beyond the 8K-element minimum
int pick_log2_units(int lg, int64_t batch) {
  int ups = unit_log2_elems(lg) - lg;
  while (ups > 0 && lg + ups > 13 && (int64_t(1) << (ups - 1)) >= batch) --ups;
  if (lg + ups < 13) ups = 13 - lg;
  return ups;
}

bool add_pass(tfft_plan_s* p, std::vector<Pass>* out, const UnitShape& shape, const UnitStrides& st, uint32_t n_units,
              int src, int dst, bool in_user, bool out_user) {
  Pass ps;
  if (!build_unit_plan(shape, &ps.plan, &ps.info)) {
    fprintf(stderr, "tfft: plan error: %s\n", ps.info.error.c_str());
    return false;
  }
  ps.strides = st;
  ps.strides.n_units = n_units;
  ps.n_units = n_units;
  ps.smem = ps.plan.ring ? smem_ring_layout(ps.plan).total : smem_layout(ps.plan).total;
  ps.tables = make_tables(ps.plan, (p->flags & TFFT_UNSCALED) != 0);
  ps.src = src;
  ps.dst = dst;
  ps.in_stride_is_user = in_user;
  ps.out_stride_is_user = out_user;
  out->push_back(ps);
  return true;
}
bool add_pass(tfft_plan_s* p, const UnitShape& shape, const UnitStrides& st, uint32_t n_units, int src, int dst,
              bool in_user, bool out_user) {
  return add_pass(p, &p->passes, shape, st, n_units, src, dst, in_user, out_user);
}

int build_1d(tfft_plan_s* p) {
  const int lg = p->lg;
  const int64_t n = p->n, batch = p->batch;
  if (lg <= 15) {
    UnitShape sh;
    sh.log2_len = lg;
    sh.log2_units = pick_log2_units(lg, batch);
    {
      int rho[kMaxStages];
      radix_schedule(lg, rho);
      // 16 or 32 rows per K line (N <= 1024): SWIZZLE_32B atoms (TFFT_NO_TMA_SMALL switches back to cp.async)
      const bool small_ok = (lg - rho[0]) >= 4 && dev_env("TFFT_NO_TMA_SMALL") == nullptr;
      sh.tma_load = ((lg - rho[0]) >= 6 || small_ok) && knob(p->tune.tma, "TFFT_NO_TMA", 1) && !(p->flags & TFFT_INTERLEAVED);
      sh.pipe_stage2 = sh.tma_load && lg >= 13 && knob(p->tune.pipe, "TFFT_NO_PIPE", 1);
      // N = 32768 (one 32K-element unit per CTA): the landing-ring kernel is opt-in (tuner key ring=2).  Measured on B200:
      // 0.6145 against 0.6111 ms per GiB -- the row-mode store phase (2.5 us) is too short to hide the chain "stage-1
      // UMMAs of the first half -> request the second half -> its latency -> its UMMAs" of a two-slot ring
      sh.ring = lg == 15 && sh.tma_load && sh.log2_units == 0 && p->tune.ring >= 2;
    }
    UnitStrides st;
    st.n_transforms = static_cast<uint32_t>(batch);
    st.units_per_batch = 0x7FFFFFFFu;   // unit base = unit * unit_stride
    const int64_t U = int64_t(1) << sh.log2_units;
    if (!add_pass(p, sh, st, static_cast<uint32_t>((batch + U - 1) / U), 0, 1, true, true)) return TFFT_E_UNSUPPORTED;
    p->passes.back().il_in = p->passes.back().il_out = (p->flags & TFFT_INTERLEAVED) != 0;
    return TFFT_OK;
  }
  // CTA-pair units are OFF by default: measured on B200 they are correct but not faster (N = 65536 x 4096: 1.08 ms in one
  // pass against 0.97 ms in two; 2^24: 2.24 against 2.19 ms).  Tuner key cluster=1 (or the developer variable
  // TFFT_CLUSTER) switches them on.
  const bool use_cluster = (p->tune.cluster > 0 || (p->tune.cluster < 0 && dev_env("TFFT_CLUSTER"))) && !(p->flags & TFFT_INTERLEAVED);
  if (lg == 16 && use_cluster) {
    // N = 65536 in ONE pass (BASELINE north_star (2)): a unit of 2^16 elements shared by a CTA pair -- each CTA loads the
    // rows of one half of m = n mod 4096, runs the radix-16 stage on them and hands the outputs of the other half of k_1
    // to its partner through distributed shared memory; the two radix-64 stages and the store are local
    UnitShape sh;
    sh.log2_len = 16;
    sh.log2_units = 0;
    sh.cluster = true;
    sh.tma_load = knob(p->tune.tma, "TFFT_NO_TMA", 1) != 0;
    UnitStrides st;
    st.n_transforms = static_cast<uint32_t>(batch);
    st.units_per_batch = 0x7FFFFFFFu;
    if (!add_pass(p, sh, st, static_cast<uint32_t>(batch), 0, 1, true, true)) return TFFT_E_UNSUPPORTED;
    return TFFT_OK;
  }
  // Three passes from 2^24 on (256 x 256 x 256: every piece of every pass is 128 bytes or more).  The four-step plan of 2^24
  // (4096 x 4096) moves 16-byte pieces in three of its four global access patterns, which the memory system sustains at
  // 2.4 - 3.2 TB/s only (probe/colio_probe.cu); measured on B200, batch 16: 2.11 ms in two passes, see DESIGN for three.
  // A tuner file that names a four-step split (lg1=) or CTA-pair units (cluster=1) for 2^24 keeps the two-pass plan;
  // developer knob TFFT_THREEPASS_LG=25.
  int three_from = (p->tune.fourstep_lg1 >= 0 || use_cluster) ? 25 : 24;
  if (const char* e = dev_env("TFFT_THREEPASS_LG")) three_from = std::min(25, std::max(24, atoi(e)));   // four-step covers <= 2^24 only
  if (lg >= three_from) {
    // TFFT_PRESERVE_INPUT: pass A writes into a plan-owned scratch of ONE transform (the batch is looped), passes B
    // and C work from there; without the flag passes A and B run in place on the input planes.  TFFT_INTERLEAVED: pass A
    // reads half2 pairs and pass C writes them, the scratch between them is planar
    const bool il3 = (p->flags & TFFT_INTERLEAVED) != 0;
    const bool preserve3 = (p->flags & TFFT_PRESERVE_INPUT) != 0 || il3;
    const int mid = preserve3 ? 2 : 0;
    // three passes: n = N1 * Na * Nb (each 2^8 .. 2^12), one transform at a time (exec loops over the batch).
    //   A: N2 = Na*Nb strided length-N1 transforms, times exp(-2*pi*i*k1*n2/n)            (column mode, in place)
    //   B: for every k1, Nb strided length-Na transforms of row k1, times exp(-2*pi*i*ka*b/N2)   (column mode, in place)
    //   C: for 8+ consecutive k1 and one ka, length-Nb transforms stored at k1 + N1*(ka + Na*kb)  (row in, column out)
    int lg1 = (lg + 2) / 3;
    if (lg1 < 8) lg1 = 8;
    const int lg2 = lg - lg1, la = (lg2 + 1) / 2, lb = lg2 - la;
    if (lg1 > 12 || la > 12 || lb < 8) return TFFT_E_INVALID_SIZE;
    const int64_t N1 = int64_t(1) << lg1, N2 = int64_t(1) << lg2, Na = int64_t(1) << la, Nb = int64_t(1) << lb;
    // Without a scratch (the default: passes A and B in place) the whole batch runs in three launches: the transform index
    // is an outer level of the unit index.  With TFFT_PRESERVE_INPUT / TFFT_INTERLEAVED the batch is looped (the scratch
    // holds one transform).
    const bool batched3 = !preserve3 && batch > 1 && batch * (N2 >> 3) < (int64_t(1) << 31) && dev_env("TFFT_THREEPASS_LOOP") == nullptr;
    p->loop_batch = !batched3;
    const uint32_t nb3 = batched3 ? static_cast<uint32_t>(batch) : 1u;
    {
      UnitShape sh;
      sh.log2_len = lg1; sh.log2_units = std::max(3, unit_log2_elems(lg1) - lg1);
      sh.in_mode = kColMode; sh.out_mode = kColMode;
      // column tiles by TMA as in the four-step pass: -5 % up to 2^27, +12 % at 2^28 / 2^29 (row strides of 1 MiB and
      // more), measured with tools/bench_large.py
      sh.tma_load = lg <= 27 && knob(p->tune.tma_col, "TFFT_NO_TMA_COL", 1) != 0 && !il3;
      sh.no_col64 = p->tune.tma_col == 2 || dev_env("TFFT_NO_COL64") != nullptr;
      const int64_t U = int64_t(1) << sh.log2_units;
      UnitStrides st;
      st.in_nstride = N2; st.out_nstride = N2; st.in_unit_stride = U; st.out_unit_stride = U;
      st.units_per_batch = static_cast<uint32_t>(N2 / U);
      st.col_base_stride = static_cast<uint32_t>(U);
      st.pass1_log2n = lg;
      if (batched3) st.b3_units = static_cast<uint32_t>(N2 / U);
      if (!add_pass(p, sh, st, nb3 * static_cast<uint32_t>(N2 / U), 0, mid, true, true)) return TFFT_E_UNSUPPORTED;
      p->passes.back().own_batch_strides = true;
      p->passes.back().outer_batch = batched3;
      p->passes.back().il_in = il3;
    }
    {
      UnitShape sh;
      sh.log2_len = la; sh.log2_units = std::max(3, unit_log2_elems(la) - la);
      sh.in_mode = kColMode; sh.out_mode = kColMode;
      // batched: the tile's batch coordinate is taken by k1, the transform of the batch is a fifth dimension of the map
      sh.tma_load = lg <= 27 && knob(p->tune.tma_col, "TFFT_NO_TMA_COL", 1) != 0;
      sh.no_col64 = p->tune.tma_col == 2 || dev_env("TFFT_NO_COL64") != nullptr;
      const int64_t U = int64_t(1) << sh.log2_units;
      UnitStrides st;
      st.in_nstride = Nb; st.out_nstride = Nb; st.in_unit_stride = U; st.out_unit_stride = U;
      st.in_batch_stride = N2; st.out_batch_stride = N2;
      if (batched3) st.b3_units = static_cast<uint32_t>(N1 * (Nb / U));
      st.units_per_batch = static_cast<uint32_t>(Nb / U);
      st.col_base_stride = static_cast<uint32_t>(U);
      st.pass1_log2n = lg2;
      if (!add_pass(p, sh, st, nb3 * static_cast<uint32_t>(N1 * (Nb / U)), mid, mid, true, true)) return TFFT_E_UNSUPPORTED;
      p->passes.back().own_batch_strides = true;
      p->passes.back().outer_batch = batched3;
    }
    {
      UnitShape sh;
      sh.log2_len = lb; sh.log2_units = std::max(3, unit_log2_elems(lb) - lb);
      sh.in_mode = kRowMode; sh.out_mode = kColMode;
      const int64_t U = int64_t(1) << sh.log2_units;
      UnitStrides st;
      st.in_tstride = N2; st.in_unit_stride = Nb; st.in_batch_stride = U * N2;
      st.out_nstride = N1 * Na; st.out_unit_stride = N1; st.out_batch_stride = U;
      st.units_per_batch = static_cast<uint32_t>(Na);
      if (batched3) st.b3_units = static_cast<uint32_t>((N1 / U) * Na);
      // Row tiles by TMA for rows up to 1024 points, as in the four-step row pass (5-D map: column block, matrix row).
      // Measured on B200 (ncu, 2^24 x 16): this pass 538 us with 16-byte cp.async against 409 us for the same shape as
      // the row pass of 2^16.  The batched plan needs the transform stride to be a whole number of matrix rows; tfft_exec
      // takes the cp.async twin (passes_strided) otherwise.
      const bool row_tma = lb <= 10 && knob(p->tune.tma, "TFFT_NO_TMA", 1) && dev_env("TFFT_NO_TMA_PASS2") == nullptr;
      if (row_tma) {
        const std::vector<Pass> head(p->passes.begin(), p->passes.end());
        if (!add_pass(p, sh, st, nb3 * static_cast<uint32_t>((N1 / U) * Na), mid, 1, true, true)) return TFFT_E_UNSUPPORTED;
        p->passes.back().own_batch_strides = true;
        p->passes.back().outer_batch = batched3;
        p->passes.back().il_out = il3;
        p->passes_strided = head;
        p->passes_strided.push_back(p->passes.back());
        p->passes.pop_back();
        sh.tma_load = true;
      }
      if (!add_pass(p, sh, st, nb3 * static_cast<uint32_t>((N1 / U) * Na), mid, 1, true, true)) return TFFT_E_UNSUPPORTED;
      p->passes.back().own_batch_strides = true;
      p->passes.back().outer_batch = batched3;
      p->passes.back().row5 = row_tma;
      p->passes.back().il_out = il3;
    }
    if (preserve3) {
      p->workspace_bytes = 2 * n * static_cast<int64_t>(sizeof(__half));   // one transform: exec loops over the batch
      if (cudaMalloc(&p->workspace, p->workspace_bytes) != cudaSuccess) {
        cudaGetLastError();
        return TFFT_E_NOMEM;
      }
    }
    return TFFT_OK;
  }
  // four-step: n = N1 * N2, element n1*N2 + n2.  Pass 1: N2 strided length-N1 transforms (column
  // mode, in place on the source), times exp(-2*pi*i*k1*n2/n).  Pass 2: N1 contiguous length-N2
  // transforms stored transposed: X[k1 + N1*k2].   
  // column-pass length 2^lg1, measured per size on B200 (tools/tune_fourstep.py): the balanced split except where it
  // would produce 2048-point units (40 KiB of DFT matrices -> one 16K-element CTA per SM): 2^19 = 512 x 1024 (-9 %),
  // 2^21 = 4096 x 512 (-3 %), 2^22 = 4096 x 1024 (-26 %); 2^23 = 2048 x 4096 with 16-column units (-16 %)
  // round 2 (tools/tune.py, gpurun_out/t27_tune.log -> profiles/r02_TunerResults.dat): 2^21 = 2048 x 1024 (1.24 ms against 1.42 ms
  // for 4096 x 512) now that two 16K-element CTAs of a 2048-point plan share an SM; the other sizes keep their split
  static const int kLg1[9] = {8, 9, 9, 9, 10, 11, 12, 11, 12};   // lg = 16 .. 24
  if (lg < 16 || lg > 24) return TFFT_E_INVALID_SIZE;
  int lg1 = kLg1[lg - 16];
  {   // tuner file / developer knob: length 2^lg1 of the column pass
    const char* e = dev_env("TFFT_FOURSTEP_LG1");
    const int v = p->tune.fourstep_lg1 >= 0 ? p->tune.fourstep_lg1 : (e ? atoi(e) : -1);
    if (v >= 8 && v <= 12 && lg - v >= 8 && lg - v <= 12) lg1 = v;
  }
  const int lg2 = lg - lg1;
  if (lg1 > 12 || lg2 < 8 || lg2 > 12) return TFFT_E_INVALID_SIZE;
  const int64_t N1 = int64_t(1) << lg1, N2 = int64_t(1) << lg2;
  // interleaved transforms go through a planar scratch between the passes (pass 1 reads half2, pass 2 writes half2)
  const bool interleaved = (p->flags & TFFT_INTERLEAVED) != 0;
  const bool preserve = (p->flags & TFFT_PRESERVE_INPUT) != 0 || interleaved;
  {
    UnitShape sh;
    sh.log2_len = lg1;
    sh.log2_units = std::max(3, unit_log2_elems(lg1) - lg1);
    // 2048-point columns: 16 columns per unit (32K elements, 16-column tiles) instead of 8 (measured at 2^23 = 2048 x 4096:
    // 2.23 -> 1.87 ms)
    if (lg1 == 11 && dev_env("TFFT_COL2048_U8") == nullptr) sh.log2_units = 4;
    // 4096-point columns: 16 columns per unit (32-byte pieces on both sides instead of 16-byte ones) as a CTA-pair unit
    if (lg1 == 12 && use_cluster) { sh.log2_units = 4; sh.cluster = true; }
    sh.in_mode = kColMode;
    sh.out_mode = kColMode;
    sh.tma_load = knob(p->tune.tma_col, "TFFT_NO_TMA_COL", 1) && !interleaved;   // column tiles {8 | 16 | 64 columns, R, M} by TMA
    // tuner key tma_col=2: 16-column tiles even where 64-column tiles (whole 128-byte lines, 256-point columns) exist
    sh.no_col64 = p->tune.tma_col == 2 || dev_env("TFFT_NO_COL64") != nullptr;
    // 4096-point columns (8 columns, 32K elements, one CTA per SM): landing-ring kernel
    sh.ring = lg1 == 12 && sh.log2_units == 3 && !sh.cluster && sh.tma_load && knob(p->tune.ring, "TFFT_NO_RING", 1);
    const int64_t U = int64_t(1) << sh.log2_units;
    UnitStrides st;
    st.in_nstride = N2; st.out_nstride = N2;
    st.in_unit_stride = U; st.out_unit_stride = U;
    st.units_per_batch = static_cast<uint32_t>(N2 / U);
    st.col_base_stride = static_cast<uint32_t>(U);
    st.pass1_log2n = lg;
    // batch stride: user's input stride (source) ; destination = source (in place) or workspace
    if (!add_pass(p, sh, st, static_cast<uint32_t>(batch * (N2 / U)), 0, preserve ? 2 : 0, true, !preserve))
      return TFFT_E_UNSUPPORTED;
    p->passes.back().il_in = interleaved;
  }
  {
    UnitShape sh;
    sh.log2_len = lg2;
    sh.log2_units = std::max(3, unit_log2_elems(lg2) - lg2);
    sh.in_mode = kRowMode;
    sh.out_mode = kColMode;
    const int64_t U = int64_t(1) << sh.log2_units;
    UnitStrides st;
    st.in_tstride = N2; st.in_unit_stride = U * N2;
    st.out_nstride = N1; st.out_unit_stride = U;
    st.units_per_batch = static_cast<uint32_t>(N1 / U);
    // the contiguous rows are loaded as TMA tiles (transform t of batch b at b*stride + t*N2: the batch stride must be
    // a whole number of rows; otherwise tfft_exec takes the cp.async twin of this pass from passes_strided)
    // Measured on B200 (C3 sizes): -8 .. -13 % for row lengths up to 1024 (2^16 .. 2^22), +10 .. +22 % for 2048 / 4096
    // (2^23, 2^24), so only the former use tiles.
    // 4096-point rows (8 per unit, 32K elements) through the landing-ring kernel: opt-in (tuner key ring=2).  Measured on
    // B200 it is much SLOWER (2^23: 2.31 against 1.84 ms, 2^24: 2.66 against 2.16 ms): this pass already runs at the rate
    // the memory system sustains for 128-byte reads + 16-byte scattered writes (probe/colio_probe.cu: 15 - 16 us per unit
    // for the bare traffic, 17 us measured), and overlapping its reads with its writes makes that traffic slower
    const bool row_ring = lg2 == 12 && !interleaved && p->tune.ring >= 2 && knob(p->tune.tma, "TFFT_NO_TMA", 1);
    const bool row_tma = (lg2 <= 10 || row_ring) && knob(p->tune.tma, "TFFT_NO_TMA", 1) && dev_env("TFFT_NO_TMA_PASS2") == nullptr;
    if (row_tma) {
      const Pass first = p->passes.back();
      if (!add_pass(p, sh, st, static_cast<uint32_t>(batch * (N1 / U)), preserve ? 2 : 0, 1, !preserve, true))
        return TFFT_E_UNSUPPORTED;
      p->passes.back().il_out = interleaved;
      p->passes_strided.push_back(first);
      p->passes_strided.push_back(p->passes.back());
      p->passes.pop_back();
      sh.tma_load = true;
      sh.ring = row_ring;
    }
    if (!add_pass(p, sh, st, static_cast<uint32_t>(batch * (N1 / U)), preserve ? 2 : 0, 1, !preserve, true))
      return TFFT_E_UNSUPPORTED;
    p->passes.back().il_out = interleaved;
  }
  if (preserve) {
    p->workspace_bytes = 2 * n * batch * static_cast<int64_t>(sizeof(__half));
    if (cudaMalloc(&p->workspace, p->workspace_bytes) != cudaSuccess) {
      cudaGetLastError();
      return TFFT_E_NOMEM;
    }
  }
  return TFFT_OK;
}

// 2-D transform of ny x nx row-major images (the reference has no 2-D path), two HBM passes:
//   rows    : every unit holds U = 2^yb rows y_lo + u*(ny/U) of one image (yb = log2(ny) - 12, 0 for ny <= 4096),
//             transforms them along x and -- when yb > 0 -- also across the U rows (first decimation-in-frequency
//             step of the column transform, folded into the last tensor stage as F_x (x) F_y) and multiplies row
//             k_y by exp(-2*pi*i*k_y*y_lo/ny); result row k_y of unit y_lo is stored at row y_lo*U + k_y of the output
//   columns : for every k_y, 8+ adjacent columns: length-(ny/U) transforms over rows k_y + U*y_lo, in place;
//             output index k' lands at row k_y + U*k' = the natural row of the 2-D transform.
int build_2d(tfft_plan_s* p, std::vector<Pass>* passes, bool allow_tma) {
  const int lgx = ilog2_exact(p->nx), lgy = ilog2_exact(p->ny);
  if (lgx < 8 || lgy < 8 || lgx > 15) return TFFT_E_INVALID_SIZE;
  auto yb_ok = [&](int yb) {
    if (yb == 0) return lgy <= 12;
    if (lgx + yb < 13 || lgx + yb > 15 || lgy - yb > 12 || lgy - yb < 8) return false;
    int rho[kMaxStages];
    const int s = radix_schedule(lgx + yb, rho);
    return rho[s - 1] - yb >= 3;   // the Kronecker stage keeps >= 3 bits of the x index (16-byte store chunks)
  };
  int yb = lgy > 12 ? lgy - 12 : 0;
  while (yb <= 3 && !yb_ok(yb)) ++yb;
  if (const char* e = dev_env("TFFT_2D_YBITS")) yb = atoi(e);   // developer override (e.g. 2 rows vs 4 rows per unit)
  if (yb < 0 || yb > 3 || !yb_ok(yb)) return TFFT_E_UNSUPPORTED;
  p->ybits = yb;
  const int64_t nx = p->nx, ny = p->ny, batch = p->batch;
  // Tiled intermediate (TFFT_2D_TILED=1, experiment kept for the record): the row pass writes, into a plan-owned
  // scratch, the operand order of the 8-column column units -- [k_y][x/8][y][x%8], one contiguous chunk per unit -- so
  // that the column pass loads contiguous 128-byte lines instead of 16-byte pieces 2*nx elements apart, and writes the
  // natural layout.  Correct (GPU suite passes) but measured SLOWER at C5 on B200: 1.33 ms against 0.78 ms -- the row
  // pass's stores become 16-byte pieces 64 KiB apart, which costs more than the column pass's loads gain.
  const int lg2c = lgy - yb;
  const bool tiled = dev_env("TFFT_2D_TILED") != nullptr && std::max(3, unit_log2_elems(lg2c) - lg2c) == 3;
  const int64_t rows2 = ny >> yb;
  passes->clear();
  bool ok = true;
  {
    UnitShape sh;
    sh.log2_len = lgx;
    sh.kron_bits = yb;
    sh.log2_units = yb ? yb : std::min(lgy, std::max(13 - lgx, unit_log2_elems(lgx) - lgx));
    int rho[kMaxStages];
    radix_schedule(yb ? lgx + yb : lgx, rho);
    const bool il2 = (p->flags & TFFT_INTERLEAVED) != 0;   // half2 images: register-split loads, no TMA tiles
    sh.tma_load = allow_tma && !il2 && (lgx - rho[0]) >= 4 && knob(p->tune.tma, "TFFT_NO_TMA", 1);   // 128-byte or 32-byte atoms
    sh.pipe_stage2 = sh.tma_load && lgx + yb >= 13 && knob(p->tune.pipe, "TFFT_NO_PIPE", 1);
    const int64_t U = int64_t(1) << sh.log2_units;
    UnitStrides st;
    st.in_tstride = yb ? (ny >> yb) * nx : nx;
    st.in_unit_stride = yb ? nx : U * nx;
    st.out_tstride = nx;
    st.out_unit_stride = U * nx;
    st.units_per_batch = static_cast<uint32_t>(ny / U);
    st.col_base_stride = 1;   // Kronecker row twiddle: col_base = y_lo
    st.kron_log2n = yb ? static_cast<uint32_t>(lgy) : 0u;
    if (tiled) {
      st.out_tstride = yb ? rows2 * nx : 8;      // output row k_y: next block of chunks / next row inside a chunk
      st.out_unit_stride = yb ? 8 : U * 8;       // unit = y_lo (Kronecker) or U consecutive rows
      st.out_hi_from = 3;
      st.out_hi_stride = rows2 * 8;              // x / 8 selects the chunk
    }
    ok = add_pass(p, passes, sh, st, static_cast<uint32_t>(batch * (ny / U)), 0, tiled ? 2 : 1, true, true);
    if (ok) passes->back().kind = 1;
    if (ok) passes->back().il_in = passes->back().il_out = il2;   // interleaved in, interleaved intermediate in the output array
  }
  if (ok) {
    const int lg2 = lgy - yb;
    UnitShape sh;
    sh.log2_len = lg2;
    sh.log2_units = std::max(3, unit_log2_elems(lg2) - lg2);
    sh.in_mode = kColMode;
    sh.out_mode = kColMode;
    // Column tiles by TMA are measured SLOWER here (C5 on B200: 1.27 ms against 0.77 ms with 16-byte cp.async): with a
    // row stride of 2*nx elements consecutive 16-byte pieces of a tile lie 32 KiB apart and the tile walks kappa
    // (2 MiB jumps) before m, while the cp.async units of neighbouring CTAs sweep the rows together.  Four-step column
    // passes (row stride <= 8 KiB) gain 1-4 % from the tiles and keep them.
    sh.tma_load = dev_env("TFFT_TMA_COL_2D") != nullptr;
    // developer knob: 4096-point columns through the landing-ring kernel (TMA column tiles, hidden under the store phase)
    if (lg2 == 12 && dev_env("TFFT_RING_2D") && !(p->flags & TFFT_INTERLEAVED)) { sh.tma_load = allow_tma; sh.ring = allow_tma; }
    if (lg2 == 11 && dev_env("TFFT_2D_COL_U16")) sh.log2_units = 4;   // developer knob: 16 columns x 2048 (32-byte pieces)
    // 4096-point columns: CTA-pair units of 16 columns (32-byte pieces)
    if (lg2 == 12 && dev_env("TFFT_CLUSTER")) { sh.log2_units = 4; sh.cluster = true; }   // developer knob, see build_1d
    const int64_t U = int64_t(1) << sh.log2_units;
    UnitStrides st;
    st.in_nstride = nx << yb; st.out_nstride = nx << yb;
    st.in_unit_stride = U; st.out_unit_stride = U;
    if (tiled) {   // unit uu = k_y * (nx/8) + x/8 is the uu-th contiguous chunk of rows2 x 8 elements
      st.in_nstride = 8;
      st.in_unit_stride = rows2 * 8;
    }
    st.units_per_batch = static_cast<uint32_t>((nx << yb) / U);
    ok = add_pass(p, passes, sh, st, static_cast<uint32_t>(batch * ((nx << yb) / U)), tiled ? 2 : 1, 1, true, true);
    if (ok) passes->back().kind = 2;
    if (ok) passes->back().il_in = passes->back().il_out = (p->flags & TFFT_INTERLEAVED) != 0;   // in place on the half2 output
  }
  if (ok && tiled && !p->workspace) {
    p->workspace_bytes = 2 * p->n * batch * static_cast<int64_t>(sizeof(__half));
    if (cudaMalloc(&p->workspace, p->workspace_bytes) != cudaSuccess) {
      cudaGetLastError();
      return TFFT_E_NOMEM;
    }
  }
  return ok ? TFFT_OK : TFFT_E_UNSUPPORTED;
}

// L2 prefetch of the next unit, measured per size with tools/tune.py on B200 (profiles/r01_TunerResults.dat): it pays
// for single-pass N = 2048 / 4096 (-9 % / -4 %) and for the 32K-element units of N = 32768 and N >= 2^23, and costs up to
// 8 % everywhere else (16K-element units whose loads already overlap compute; 2-D column passes; N = 2^21, 2^22)
bool prefetch_default(const tfft_plan_s* p, const Pass& ps, const UnitPlan& plan) {
  if (ps.kind != 0) return false;
  if (p->lg <= 15) return plan.tma_load == 1 && (p->lg == 11 || p->lg == 12 || p->lg == 15);
  return plan.log2_elems == 15 && p->lg >= 23;
}

// First use of a pass on a device (caller holds g_upload_mutex): opt in to > 48 KiB of dynamic shared memory -- a
// per-device attribute of the function (a process-wide call_once left every device but the first without it)
// --, upload the constant tables and size the persistent grid.  cudaFuncSetAttribute also forces the (lazily loaded)
// kernel module onto the device; tfft_plan_prepare runs this for every pass ahead of time.
int ensure_pass_on_device(const Pass& ps, int dev, const void* entry, bool two_slot, const void* fn, int threads,
                          uint32_t tmem_cols) {
  if (ps.d_tables[dev]) return TFFT_OK;
  cudaError_t e = cudaFuncSetAttribute(entry, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return e == cudaErrorInvalidDeviceFunction || e == cudaErrorNoKernelImageForDevice ? TFFT_E_NO_DEVICE : static_cast<int>(e);
  }
  uint4* d = nullptr;
  e = cudaMalloc(&d, ps.tables.size());
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? TFFT_E_NOMEM : static_cast<int>(e); }
  e = cudaMemcpy(d, ps.tables.data(), ps.tables.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return static_cast<int>(e); }
  int per_sm = 0, sms = 0, smem_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  if (ps.plan.cluster) {
    // resident CTA pairs: the pairs must fit the GPCs, so ask the runtime instead of multiplying
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(2 * sms));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = ps.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    e = cudaOccupancyMaxActiveClusters(&clusters, fn, &cfg);
    if (e != cudaSuccess || clusters < 1) { cudaGetLastError(); cudaFree(d); return TFFT_E_UNSUPPORTED; }
    ps.resident_ctas[dev] = 2 * clusters;
    ps.d_tables[dev] = d;
    return TFFT_OK;
  } else if (two_slot) {
    per_sm = 1;   // two units in flight inside one CTA
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, ps.smem);
    const int by_smem = smem_sm / static_cast<int>(ps.smem + 1024);
    if (per_sm < by_smem) per_sm = std::min(by_smem, 2);   // the API can under-report before the carve-out is set
    const int tmem_limit = 512 / static_cast<int>(tmem_cols);   // tensor memory: 512 columns per SM
    if (per_sm > tmem_limit) per_sm = tmem_limit;
  }
  if (per_sm < 1 || sms < 1) { cudaFree(d); return TFFT_E_UNSUPPORTED; }
  ps.resident_ctas[dev] = per_sm * sms;
  ps.d_tables[dev] = d;
  return TFFT_OK;
}

int prepare_launch(const tfft_plan_s* p, const Pass& ps, const __half* src_re, const __half* src_im, int64_t in_stride,
                   int64_t out_stride, int tw_log2, int64_t tw_first_col, int segs, int64_t seg_stride, int dev,
                   Prepared* out) {
  UnitStrides st = ps.strides;
  if (tw_log2) {   // fused output twiddle: column index of transform b = first_col + b
    st.pass1_log2n = static_cast<uint32_t>(tw_log2);
    st.col_base_stride = 1u << ps.plan.log2_units;
    st.col_div = 1;
  }
  const int64_t U = int64_t(1) << ps.plan.log2_units;
  int64_t tma_extent = 0;   // 2-D row pass: rows (4-D map) the tensor map spans
  if (ps.kind != 0) {   // 2-D passes: only the image strides come from the caller
    st.in_batch_stride = in_stride;
    st.out_batch_stride = out_stride;
    if (ps.kind == 1 && ps.plan.tma_load) {   // tfft_exec checked tma_ok_2d(): the image stride is a whole number of steps
      const int64_t step_elems = ps.plan.kron_bits ? (p->ny >> ps.plan.kron_bits) * p->nx : p->nx;
      st.tma_batch_step = static_cast<uint32_t>(p->batch > 1 ? in_stride / step_elems : 0);
      tma_extent = static_cast<int64_t>(st.tma_batch_step) * (p->batch - 1) + p->ny;
    }
  } else if (ps.plan.in_mode == kRowMode && ps.plan.out_mode == kRowMode) {   // batched 1-D, one pass
    st.in_tstride = in_stride; st.out_tstride = out_stride;
    st.in_unit_stride = U * in_stride; st.out_unit_stride = U * out_stride;
  } else if (!ps.own_batch_strides) {   // four-step passes: the batch level carries the user's (or workspace) transform stride
    st.in_batch_stride = in_stride;
    st.out_batch_stride = out_stride;
  } else if (ps.outer_batch) {   // batched three-pass plans: the outer level carries it
    st.in_b3_stride = in_stride;
    st.out_b3_stride = out_stride;
    // column tiles: pass A's own batch level is unused (batch coordinate = b3, 4-D map); pass B's is k1 (5-D map)
    st.tma_b3_step = st.units_per_batch == st.b3_units ? 1u : 2u;
  }
  if (ps.row5) {   // tfft_exec checked that in_stride is a whole number of matrix rows (in_tstride = N2)
    st.tma_row5 = true;
    st.tma_b3_step = ps.outer_batch ? static_cast<uint32_t>(in_stride / st.in_tstride) : 0u;
  }
  if (ps.kind == 0 && (ps.plan.tma_load == 1 || ps.plan.tma_load == 3) && st.units_per_batch != 0x7FFFFFFFu && !ps.row5) {
    // four-step row pass: transform t of batch b at b*batch_stride + t*tstride (tfft_exec checked divisibility)
    const int64_t batches = (ps.n_units + st.units_per_batch - 1) / st.units_per_batch;
    st.tma_batch_step = static_cast<uint32_t>(batches > 1 ? st.in_batch_stride / st.in_tstride : 0);
    tma_extent = static_cast<int64_t>(st.tma_batch_step) * (batches - 1) +
                 (static_cast<int64_t>(st.units_per_batch) << ps.plan.log2_units);
  }
  UnitPlan& plan = out->plan;
  plan = ps.plan;
  fill_strides(st, ps.info, &plan);
  {
    static const char* pf_env = dev_env("TFFT_PREFETCH");   // developer override: 0 / 1
    plan.il_in = ps.il_in ? 1u : 0u;
    plan.il_out = ps.il_out ? 1u : 0u;
    plan.il_swap = (p->flags & TFFT_INVERSE) ? 1u : 0u;
    plan.prefetch_next = ps.il_in ? 0u : p->tune.prefetch >= 0 ? static_cast<uint32_t>(p->tune.prefetch)
                         : pf_env ? static_cast<uint32_t>(atoi(pf_env))
                                : (prefetch_default(p, ps, plan) ? 1u : 0u);
  }
  plan.col_first = static_cast<uint32_t>(tw_first_col);
  bool ring_fallback = false;
  if (segs > 0 && plan.ring) {   // segmented input: the single-unit kernel (same plan; row-mode staging is the same)
    if (plan.out_mode != kRowMode) return TFFT_E_UNSUPPORTED;
    plan.ring = 0;
    ring_fallback = true;
  }
  if (segs > 0) {   // tfft_exec_segmented: only for the row tiles of 64-row atoms, whole K lines per segment
    const int R0 = 1 << plan.log2_radix[0];
    if (plan.tma_load != 1 || plan.cluster || plan.kron_bits || ps.kind != 0 || R0 % segs != 0) return TFFT_E_UNSUPPORTED;
    plan.tma_seg = 1;
    plan.prefetch_next = 0;
  }
  const bool allow2 = knob(p->tune.two_slot, "TFFT_NO_2SLOT", 1) != 0;
  int threads = kThreads;
  KernelFn fn = kernel_for(plan, &threads);
  if (!fn) return TFFT_E_UNSUPPORTED;
  Kernel2Fn fn2 = plan.ring ? kernel_ring_for(plan) : kernel2_for(plan, allow2);
  if (plan.ring && !fn2) return TFFT_E_UNSUPPORTED;
  const uint32_t smem = (fn2 && !plan.ring) ? smem2_layout(plan).total : ring_fallback ? smem_layout(plan).total : ps.smem;
  // a ring pass runs two different kernels (the ring kernel, and the single-unit kernel for segmented input), but
  // ensure_pass_on_device opts only the first one it sees in to > 48 KiB of shared memory: do it here for both
  if (ps.plan.ring && cudaFuncSetAttribute(fn2 ? reinterpret_cast<const void*>(fn2) : reinterpret_cast<const void*>(fn),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
    cudaGetLastError();
    return TFFT_E_UNSUPPORTED;
  }
  const void* entry = fn2 ? reinterpret_cast<const void*>(fn2) : reinterpret_cast<const void*>(fn);
  static const bool debug = dev_env("TFFT_DEBUG") != nullptr;
  {
    const int rc = ensure_pass_on_device(ps, dev, entry, fn2 != nullptr, reinterpret_cast<const void*>(fn), threads, plan.tmem_cols);
    if (rc != TFFT_OK) return rc;
  }
  out->tables = ps.d_tables[dev];
  out->grid = plan.cluster ? 2u * std::min<unsigned>(ps.n_units, static_cast<unsigned>(ps.resident_ctas[dev]) / 2u)
                           : std::min<unsigned>(ps.n_units, static_cast<unsigned>(ps.resident_ctas[dev]));
  out->cluster = plan.cluster != 0;
  out->block = plan.ring ? 512u : fn2 ? static_cast<unsigned>(kCta2Threads) : static_cast<unsigned>(threads);
  out->smem = smem;
  out->fn = entry;
  out->two_slot = fn2 != nullptr;
  std::memset(&out->tmap_re, 0, sizeof(CUtensorMap));
  std::memset(&out->tmap_im, 0, sizeof(CUtensorMap));
  if (plan.tma_load) {
    // row-mode input: transform t of the launch starts at src + t * tstride, or, for four-step row
    // passes, at src + (t / upb) * batch_stride + (t % upb) * tstride == t * tstride when contiguous
    const int64_t n_tr = tma_extent ? tma_extent : static_cast<int64_t>(ps.n_units) << plan.log2_units;
    const bool half_box = fn2 != nullptr;   // half-tile boxes: no other kernel may run these maps
    int rc;
    if (plan.tma_load == 2 || plan.tma_load >= 4) {
      const int64_t columns = static_cast<int64_t>(plan.units_per_batch) << plan.log2_units;
      const int64_t batches = (ps.n_units + plan.units_per_batch - 1) / plan.units_per_batch;
      // batched three-pass pass A: one "batch" per transform of the user's batch, in_stride apart
      if (ps.outer_batch && st.tma_b3_step == 2) {   // pass B: (k1, transform of the batch)
        const int64_t inner = st.b3_units / plan.units_per_batch, outer = ps.n_units / st.b3_units;
        rc = make_col_tensor_map(plan, src_re, st.in_nstride, columns, inner, plan.in_batch_stride, &out->tmap_re, outer, in_stride);
        if (rc == TFFT_OK)
          rc = make_col_tensor_map(plan, src_im, st.in_nstride, columns, inner, plan.in_batch_stride, &out->tmap_im, outer, in_stride);
      } else {
        const int64_t bstride = ps.outer_batch ? in_stride : plan.in_batch_stride;
        rc = make_col_tensor_map(plan, src_re, st.in_nstride, columns, batches, bstride, &out->tmap_re);
        if (rc == TFFT_OK) rc = make_col_tensor_map(plan, src_im, st.in_nstride, columns, batches, bstride, &out->tmap_im);
      }
    } else if (ps.row5) {
      // transforms along the matrix rows: N1 per user transform, user transforms tma_b3_step rows apart
      const int64_t per = static_cast<int64_t>(ps.outer_batch ? st.b3_units : ps.n_units) / plan.units_per_batch * U;
      const int64_t outer = ps.outer_batch ? ps.n_units / st.b3_units : 1;
      const int64_t total = static_cast<int64_t>(st.tma_b3_step) * (outer - 1) + per;
      rc = make_row5_tensor_map(plan, src_re, st.in_tstride, total, plan.units_per_batch, st.in_unit_stride, &out->tmap_re);
      if (rc == TFFT_OK)
        rc = make_row5_tensor_map(plan, src_im, st.in_tstride, total, plan.units_per_batch, st.in_unit_stride, &out->tmap_im);
    } else if (plan.kron_bits) {
      rc = make_kron_tensor_map(plan, src_re, p->ny, p->batch, plan.tma_batch_step, &out->tmap_re, half_box);
      if (rc == TFFT_OK) rc = make_kron_tensor_map(plan, src_im, p->ny, p->batch, plan.tma_batch_step, &out->tmap_im, half_box);
    } else if (plan.tma_seg) {
      const int64_t nt = plan.n_transforms ? plan.n_transforms : n_tr;
      rc = make_seg_tensor_map(plan, src_re, st.in_tstride, nt, segs, seg_stride, &out->tmap_re, half_box);
      if (rc == TFFT_OK) rc = make_seg_tensor_map(plan, src_im, st.in_tstride, nt, segs, seg_stride, &out->tmap_im, half_box);
    } else {
      rc = make_input_tensor_map(plan, src_re, st.in_tstride, plan.n_transforms ? plan.n_transforms : n_tr, &out->tmap_re, half_box);
      if (rc == TFFT_OK)
        rc = make_input_tensor_map(plan, src_im, st.in_tstride, plan.n_transforms ? plan.n_transforms : n_tr, &out->tmap_im, half_box);
    }
    if (rc != TFFT_OK) return rc;
  }
  if (debug)
    fprintf(stderr, "tfft: launch grid=%u units=%u smem=%u tmem=%u resident=%d two_slot=%d\n", out->grid, ps.n_units, smem,
            plan.tmem_cols, ps.resident_ctas[dev], out->two_slot ? 1 : 0);
  out->dev = dev;
  out->sre = src_re; out->sim = src_im;
  out->in_stride = in_stride; out->out_stride = out_stride;
  out->tw_log2 = tw_log2; out->tw_first_col = tw_first_col;
  out->segs = segs; out->seg_stride = seg_stride;
  return TFFT_OK;
}

int launch_pass(const tfft_plan_s* p, const Pass& ps, const __half* src_re, const __half* src_im, __half* dst_re,
                __half* dst_im, int64_t in_stride, int64_t out_stride, cudaStream_t stream, int tw_log2 = 0,
                int64_t tw_first_col = 0, int segs = 0, int64_t seg_stride = 0) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TFFT_E_NO_DEVICE : static_cast<int>(e); }
  if (dev < 0 || dev >= 16) return TFFT_E_UNSUPPORTED;
  Prepared L;
  {
    std::lock_guard<std::mutex> lock(g_upload_mutex);
    const Prepared* hit = nullptr;
    for (const Prepared& c : ps.prepared)
      if (c.dev == dev && c.sre == src_re && c.sim == src_im && c.in_stride == in_stride && c.out_stride == out_stride &&
          c.tw_log2 == tw_log2 && c.tw_first_col == tw_first_col && c.segs == segs && c.seg_stride == seg_stride) {
        hit = &c;
        break;
      }
    if (hit) {
      L = *hit;
    } else {
      const int rc = prepare_launch(p, ps, src_re, src_im, in_stride, out_stride, tw_log2, tw_first_col, segs, seg_stride, dev, &L);
      if (rc != TFFT_OK) return rc;
      if (ps.prepared.size() < kPreparedSlots) ps.prepared.push_back(L);
      else ps.prepared[ps.prepared_next++ % kPreparedSlots] = L;
    }
  }
  long long* trace = g_trace;
  if (L.two_slot) {
    void* args2[] = {&L.plan, &dst_re, &dst_im, &L.tables, &L.tmap_re, &L.tmap_im, &trace};
    e = launch_pdl(L.fn, L.grid, L.block, args2, L.smem, stream);
  } else {
    void* args[] = {&L.plan, &src_re, &src_im, &dst_re, &dst_im, &L.tables, &trace, &L.tmap_re, &L.tmap_im};
    e = launch_pdl(L.fn, L.grid, L.block, args, L.smem, stream, L.cluster);
  }
  return e == cudaSuccess ? TFFT_OK : static_cast<int>(e);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int tfft_version(void) { return 100; }

// developer hook (not in tfft.h): device buffer of gridDim*4*16 int64 for -DTFFT_TRACE builds
void tfft_debug_set_trace(long long* buf) { g_trace = buf; }

const char* tfft_error_string(int code) {
  switch (code) {
    case TFFT_OK: return "success";
    case TFFT_E_INVALID_SIZE: return "transform length must be a power of two in [256, 2^30]";
    case TFFT_E_INVALID_ARG: return "invalid argument (null / misaligned pointer or stride)";
    case TFFT_E_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU path)";
    case TFFT_E_UNSUPPORTED: return "unsupported configuration";
    case TFFT_E_NOMEM: return "device memory allocation failed";
    case TFFT_E_NOT_IN_FILE: return "tuner file holds no line for this transform length";
    case TFFT_E_TIMEOUT: return "multi-GPU plan: a peer rank did not reach a phase barrier in time";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown tfft error";
  }
}

static int plan_create_tuned(tfft_plan_t* out, int64_t n, int64_t batch, uint32_t flags, const Tuning& tune);

int tfft_plan_create(tfft_plan_t* out, int64_t n, int64_t batch, uint32_t flags) {
  // TFFT_TUNER_FILE: every plan looks its length up in that file first (see tfft_plan_create_from_file)
  if (const char* f = getenv("TFFT_TUNER_FILE")) {
    const int rc = tfft_plan_create_from_file(out, n, batch, flags, f);
    if (rc != TFFT_E_NOT_IN_FILE) return rc;
  }
  return plan_create_tuned(out, n, batch, flags, Tuning());
}

// Tuner file: one line per length, the reference's format `N mode base_warps r16_warps r2_block`
// (src/base/Plan.h:211-236, written by src/testing/FileWriter.h:250-269) optionally followed by the knobs of the new
// kernels as `key=value` words (tma, pipe, two_slot, prefetch, lg1, tma_col); tools/tune.py writes such files.  The five
// reference columns are accepted and ignored: they describe launch shapes of kernels that no longer exist.
int tfft_plan_create_from_file(tfft_plan_t* out, int64_t n, int64_t batch, uint32_t flags, const char* path) {
  if (!out || !path) return TFFT_E_INVALID_ARG;
  *out = nullptr;
  FILE* f = fopen(path, "r");
  if (!f) return TFFT_E_INVALID_ARG;
  char line[512];
  int rc = TFFT_E_NOT_IN_FILE;
  while (fgets(line, sizeof(line), f)) {
    char* save = nullptr;
    char* tok = strtok_r(line, " \t\r\n", &save);
    if (!tok || static_cast<int64_t>(atof(tok)) != n) continue;
    Tuning t;
    while ((tok = strtok_r(nullptr, " \t\r\n", &save)) != nullptr) {
      const char* eq = strchr(tok, '=');
      if (!eq) continue;   // one of the reference's four launch-shape columns
      const int v = atoi(eq + 1);
      const size_t kl = static_cast<size_t>(eq - tok);
      if (kl == 3 && !strncmp(tok, "tma", 3)) t.tma = v;
      else if (kl == 4 && !strncmp(tok, "pipe", 4)) t.pipe = v;
      else if (kl == 8 && !strncmp(tok, "two_slot", 8)) t.two_slot = v;
      else if (kl == 8 && !strncmp(tok, "prefetch", 8)) t.prefetch = v;
      else if (kl == 3 && !strncmp(tok, "lg1", 3)) t.fourstep_lg1 = v;
      else if (kl == 7 && !strncmp(tok, "tma_col", 7)) t.tma_col = v;
      else if (kl == 7 && !strncmp(tok, "cluster", 7)) t.cluster = v;
      else if (kl == 4 && !strncmp(tok, "ring", 4)) t.ring = v;
    }
    rc = plan_create_tuned(out, n, batch, flags, t);
    break;
  }
  fclose(f);
  return rc;
}

static int plan_create_tuned(tfft_plan_t* out, int64_t n, int64_t batch, uint32_t flags, const Tuning& tune) {
  if (!out) return TFFT_E_INVALID_ARG;
  *out = nullptr;
  const int lg = ilog2_exact(n);
  if (lg < 8 || lg > 30) return TFFT_E_INVALID_SIZE;
  if (batch < 1 || batch > (int64_t(1) << 30)) return TFFT_E_INVALID_ARG;
  tfft_plan_s* p = new (std::nothrow) tfft_plan_s;
  if (!p) return TFFT_E_NOMEM;
  p->tune = tune;
  p->n = n; p->batch = batch; p->flags = flags; p->lg = lg;
  int rc = build_1d(p);
  if (rc != TFFT_OK) {
    delete p;
    return rc;
  }
  *out = p;
  return TFFT_OK;
}

int tfft_plan_create_2d(tfft_plan_t* out, int64_t ny, int64_t nx, int64_t batch, uint32_t flags) {
  if (!out) return TFFT_E_INVALID_ARG;
  *out = nullptr;
  if (ilog2_exact(ny) < 8 || ilog2_exact(nx) < 8 || ny * nx > (int64_t(1) << 30)) return TFFT_E_INVALID_SIZE;
  if (batch < 1 || batch > (int64_t(1) << 20)) return TFFT_E_INVALID_ARG;
  tfft_plan_s* p = new (std::nothrow) tfft_plan_s;
  if (!p) return TFFT_E_NOMEM;
  p->n = ny * nx; p->ny = ny; p->nx = nx; p->batch = batch; p->flags = flags; p->lg = ilog2_exact(ny * nx);
  int rc = build_2d(p, &p->passes, true);
  // the twin without TMA loads (image strides that are not a whole number of tensor-map steps) is built now, so that
  // a plan is immutable after creation and concurrent tfft_exec calls never see a half-built pass list
  if (rc == TFFT_OK && p->passes.front().plan.tma_load) rc = build_2d(p, &p->passes_strided, false);
  if (rc != TFFT_OK) {
    tfft_plan_destroy(p);
    return rc;
  }
  *out = p;
  return TFFT_OK;
}

int tfft_plan_prepare(tfft_plan_t p) {
  if (!p) return TFFT_E_INVALID_ARG;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return TFFT_E_NO_DEVICE; }
  if (dev < 0 || dev >= 16) return TFFT_E_UNSUPPORTED;
  const bool allow2 = knob(p->tune.two_slot, "TFFT_NO_2SLOT", 1) != 0;
  std::lock_guard<std::mutex> lock(g_upload_mutex);
  for (const std::vector<Pass>* v : {&p->passes, &p->passes_strided})
    for (const Pass& ps : *v) {
      int threads = kThreads;
      KernelFn fn = kernel_for(ps.plan, &threads);
      if (!fn) return TFFT_E_UNSUPPORTED;
      Kernel2Fn fn2 = ps.plan.ring ? kernel_ring_for(ps.plan) : kernel2_for(ps.plan, allow2);
      if (ps.plan.ring && !fn2) return TFFT_E_UNSUPPORTED;
      const void* entry = fn2 ? reinterpret_cast<const void*>(fn2) : reinterpret_cast<const void*>(fn);
      const int rc = ensure_pass_on_device(ps, dev, entry, fn2 != nullptr, reinterpret_cast<const void*>(fn), threads,
                                           ps.plan.tmem_cols);
      if (rc != TFFT_OK) return rc;
    }
  return TFFT_OK;
}

int tfft_plan_info(tfft_plan_t p, tfft_plan_info_t* info) {
  if (!p || !info) return TFFT_E_INVALID_ARG;
  std::memset(info, 0, sizeof(*info));
  info->n = p->n;
  info->batch = p->batch;
  const Pass& first = p->passes.front();
  info->r16_stages = static_cast<int32_t>(first.plan.stages);
  info->tail_radix = 1 << (first.plan.log2_radix[first.plan.stages - 1] - 4);
  info->passes = static_cast<int32_t>(p->passes.size());
  info->results_in_results = 1;
  info->amount_of_r16_steps = p->lg / 4 - 1;
  info->amount_of_r2_steps = p->lg % 4;
  info->transforms_per_cta = 1 << first.plan.log2_units;
  for (const Pass& ps : p->passes) {
    if (static_cast<int32_t>(ps.smem) > info->smem_bytes) info->smem_bytes = static_cast<int32_t>(ps.smem);
    const int32_t tc = static_cast<int32_t>(ps.plan.tmem_cols);
    if (tc > info->tmem_columns) info->tmem_columns = tc;
  }
  info->grid = first.n_units;
  info->workspace_bytes = p->workspace_bytes;
  info->algorithmic_bytes = 8 * p->n * p->batch * static_cast<int64_t>(p->passes.size());
  return TFFT_OK;
}

int tfft_plan_destroy(tfft_plan_t p) {
  if (!p) return TFFT_E_INVALID_ARG;
  if (p->workspace) cudaFree(p->workspace);
  destroy_host_path(p->host);
  for (std::vector<Pass>* v : {&p->passes, &p->passes_strided})
    for (Pass& ps : *v)
      for (int d = 0; d < 16; ++d)
        if (ps.d_tables[d]) cudaFree(ps.d_tables[d]);
  delete p;
  return TFFT_OK;
}

int tfft_exec(tfft_plan_t p, const void* in_re, const void* in_im, void* out_re, void* out_im, int64_t in_stride,
              int64_t out_stride, void* stream_) {
  if (!p || !in_re || !out_re) return TFFT_E_INVALID_ARG;
  if (p->flags & TFFT_INTERLEAVED) {   // one half2 array each way; the imaginary-plane arguments are ignored
    in_im = in_re;
    out_im = out_re;
  }
  if (!in_im || !out_im) return TFFT_E_INVALID_ARG;
  if (!aligned16(in_re) || !aligned16(in_im) || !aligned16(out_re) || !aligned16(out_im)) return TFFT_E_INVALID_ARG;
  if (in_stride < 0 || out_stride < 0 || (in_stride & 7) || (out_stride & 7)) return TFFT_E_INVALID_ARG;
  if (p->batch > 1 && (in_stride < p->n || out_stride < p->n)) return TFFT_E_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return TFFT_E_NO_DEVICE;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if ((p->flags & TFFT_INVERSE) && !(p->flags & TFFT_INTERLEAVED)) {   // F^-1(x) = swap(F(swap(x))), swap = exchange of the
    // real and imaginary planes (interleaved plans exchange the halves of every pair inside the kernel instead)
    std::swap(in_re, in_im);
    std::swap(out_re, out_im);
  }
  const int64_t outer = p->loop_batch ? p->batch : 1;
  const std::vector<Pass>* passes = &p->passes;
  // 2-D row pass with TMA tiles: the image stride must be a whole number of tensor-map steps (rows, or groups of ny/U
  // rows for Kronecker units); otherwise the same passes with 16-byte asynchronous copies
  auto tma_ok_2d = [&]() {
    const UnitPlan& rp = p->passes.front().plan;
    const int64_t step = rp.kron_bits ? (p->ny >> rp.kron_bits) * p->nx : p->nx;
    return in_stride % step == 0 && in_stride / step < (int64_t(1) << 31);
  };
  if (!p->ny && p->passes.size() == 2 && !p->passes_strided.empty() && p->batch > 1 && p->passes[1].src == 0 &&
      in_stride % p->passes[1].strides.in_tstride != 0)
    passes = &p->passes_strided;   // four-step row pass: the batch stride is not a whole number of rows
  if (!p->ny && p->passes.size() == 3 && !p->passes_strided.empty() && p->passes[2].outer_batch &&
      in_stride % p->passes[2].strides.in_tstride != 0)
    passes = &p->passes_strided;   // batched three-pass row pass: the transform stride is not a whole number of matrix rows
  if (p->ny && p->batch > 1 && p->passes.front().plan.tma_load && !tma_ok_2d()) {
    if (p->passes_strided.empty()) return TFFT_E_UNSUPPORTED;
    passes = &p->passes_strided;
  }
  for (int64_t ob = 0; ob < outer; ++ob)
  for (const Pass& ps : *passes) {
    const int64_t il = (p->flags & TFFT_INTERLEAVED) ? 2 : 1;   // interleaved strides count complex elements = 2 halves
    const __half* ire = static_cast<const __half*>(in_re) + ob * in_stride * il;
    const __half* iim = static_cast<const __half*>(in_im) + ob * in_stride * il;
    __half* ore = static_cast<__half*>(out_re) + ob * out_stride * il;
    __half* oim = static_cast<__half*>(out_im) + ob * out_stride * il;
    const __half *sre, *sim;
    __half *dre, *dim;
    int64_t is, os;
    auto pick = [&](int which, const __half** re, const __half** im, int64_t* stride) {
      if (which == 0) { *re = ire; *im = iim; *stride = in_stride; }
      else if (which == 1) { *re = ore; *im = oim; *stride = out_stride; }
      else { *re = p->workspace; *im = p->workspace + p->n; *stride = 2 * p->n; }
    };
    const __half *tre, *tim;
    pick(ps.src, &sre, &sim, &is);
    pick(ps.dst, &tre, &tim, &os);
    dre = const_cast<__half*>(tre);
    dim = const_cast<__half*>(tim);
    int rc = launch_pass(p, ps, sre, sim, dre, dim, is, os, stream);
    if (rc != TFFT_OK) return rc;
  }
  return TFFT_OK;
}

int tfft_exec_twiddled(tfft_plan_t p, const void* in_re, const void* in_im, void* out_re, void* out_im,
                       int64_t in_stride, int64_t out_stride, int32_t log2_total, int64_t first_col, void* stream_) {
  if (!p || !in_re || !in_im || !out_re || !out_im) return TFFT_E_INVALID_ARG;
  if (p->passes.size() != 1 || log2_total < p->lg || log2_total > 30 || (p->flags & (TFFT_INVERSE | TFFT_INTERLEAVED)))
    return TFFT_E_UNSUPPORTED;
  if (first_col < 0 || first_col + p->batch > (int64_t(1) << (log2_total - p->lg))) return TFFT_E_INVALID_ARG;
  if (!aligned16(in_re) || !aligned16(in_im) || !aligned16(out_re) || !aligned16(out_im)) return TFFT_E_INVALID_ARG;
  if ((in_stride & 7) || (out_stride & 7) || in_stride < p->n || out_stride < p->n) return TFFT_E_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return TFFT_E_NO_DEVICE;
  }
  return launch_pass(p, p->passes[0], static_cast<const __half*>(in_re), static_cast<const __half*>(in_im),
                     static_cast<__half*>(out_re), static_cast<__half*>(out_im), in_stride, out_stride,
                     static_cast<cudaStream_t>(stream_), log2_total, first_col);
}

int tfft_exec_segmented(tfft_plan_t p, const void* in_re, const void* in_im, void* out_re, void* out_im, int64_t in_stride,
                        int64_t out_stride, int32_t segments, int64_t segment_stride, int32_t log2_total, int64_t first_col,
                        void* stream_) {
  if (!p || !in_re || !in_im || !out_re || !out_im || segments < 1) return TFFT_E_INVALID_ARG;
  if (p->passes.size() != 1 || (p->flags & (TFFT_INVERSE | TFFT_INTERLEAVED)) || p->n % segments) return TFFT_E_UNSUPPORTED;
  if (log2_total && (log2_total < p->lg || log2_total > 30)) return TFFT_E_UNSUPPORTED;
  if (log2_total && (first_col < 0 || first_col + p->batch > (int64_t(1) << (log2_total - p->lg)))) return TFFT_E_INVALID_ARG;
  if (!aligned16(in_re) || !aligned16(in_im) || !aligned16(out_re) || !aligned16(out_im)) return TFFT_E_INVALID_ARG;
  const int64_t seg_len = p->n / segments;
  if ((in_stride & 7) || (out_stride & 7) || (segment_stride & 7) || in_stride < seg_len || out_stride < p->n) return TFFT_E_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return TFFT_E_NO_DEVICE;
  }
  return launch_pass(p, p->passes[0], static_cast<const __half*>(in_re), static_cast<const __half*>(in_im),
                     static_cast<__half*>(out_re), static_cast<__half*>(out_im), in_stride, out_stride,
                     static_cast<cudaStream_t>(stream_), log2_total, first_col, segments, segment_stride);
}

// Host-buffer path.  The batch is cut into chunks of whole transforms; upload, transform and download of consecutive
// chunks run on three plan-owned streams (PCIe is full duplex) through a ring of kHostRing device slots per direction,
// so the device footprint is a few chunks, not the batch.  The first and last chunks are small (1/8, 1/4, 1/2 of the
// full chunk): only the first upload and the last download are not overlapped with the opposite direction, so they are
// kept short.  All state is built locally and published only when complete (a failed first call used to
// leave a half-initialised path behind); calls on one plan are serialised.
int tfft_exec_host(tfft_plan_t p, const void* host_in, void* host_out) {
  if (!p || !host_in || !host_out) return TFFT_E_INVALID_ARG;
  std::lock_guard<std::mutex> serial(p->host_mutex);
  if (!p->host) {
    HostPath* h = new (std::nothrow) HostPath;
    if (!h) return TFFT_E_NOMEM;
    int rc = TFFT_OK;
    // chunking: about 16 MiB per direction and chunk (measured on B200: 4 / 8 / 16 / 32 MiB -> 7.1 / 6.3 / 6.07 / 6.16 ms
    // for the 2 x 256 MiB of C2), whole transforms only
    const int64_t per_transform = 2 * p->n * static_cast<int64_t>(sizeof(__half));
    int64_t chunk_mb = 16;
    if (const char* e = dev_env("TFFT_HOST_CHUNK_MB")) chunk_mb = std::max(1, atoi(e));   // developer tuning knob
    int64_t chunk = std::max<int64_t>(1, (chunk_mb << 20) / per_transform);
    if (chunk >= p->batch || dev_env("TFFT_HOST_NO_PIPELINE")) chunk = p->batch;
    h->chunk = chunk;
    // schedule: ramp up, full chunks, ramp down
    {
      int64_t left = p->batch;
      std::vector<int64_t> head, tail;
      if (chunk < p->batch && chunk >= 8 && !dev_env("TFFT_HOST_NO_RAMP"))
        for (int64_t c = chunk / 8; c < chunk && left > 2 * c + chunk; c *= 2) { head.push_back(c); tail.push_back(c); left -= 2 * c; }
      h->sizes = head;
      while (left > 0) { const int64_t c = std::min(chunk, left); h->sizes.push_back(c); left -= c; }
      for (size_t i = tail.size(); i-- > 0;) h->sizes.push_back(tail[i]);
    }
    h->slots = static_cast<int>(std::min<size_t>(kHostRing, h->sizes.size()));
    const uint32_t fl = p->flags & ~uint32_t(TFFT_PRESERVE_INPUT);
    for (int64_t c : h->sizes) {
      if (rc != TFFT_OK || c == p->batch || h->plans.count(c)) continue;
      tfft_plan_t cp = nullptr;
      rc = p->ny ? tfft_plan_create_2d(&cp, p->ny, p->nx, c, fl) : tfft_plan_create(&cp, p->n, c, fl);
      if (rc == TFFT_OK) h->plans[c] = cp;
    }
    cudaError_t e = cudaSuccess;
    if (rc == TFFT_OK) {
      e = cudaMalloc(&h->buf, 2 * static_cast<size_t>(h->slots) * 2 * p->n * chunk * sizeof(__half));
      for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking);
      for (int i = 0; i < 3 * kHostRing && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->events[i], cudaEventDisableTiming);
      if (e != cudaSuccess) {
        cudaGetLastError();
        rc = e == cudaErrorMemoryAllocation ? TFFT_E_NOMEM
             : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? TFFT_E_NO_DEVICE : static_cast<int>(e);
      }
    }
    if (rc != TFFT_OK) {
      destroy_host_path(h);
      return rc;
    }
    p->host = h;
  }
  HostPath* h = p->host;
  const int64_t chunk_halves = 2 * p->n * h->chunk;
  __half* din = h->buf;
  __half* dout = h->buf + static_cast<int64_t>(h->slots) * chunk_halves;
  cudaStream_t s_up = h->streams[0], s_fft = h->streams[1], s_down = h->streams[2];
  const __half* hin = static_cast<const __half*>(host_in);
  __half* hout = static_cast<__half*>(host_out);
  const int64_t tstride = (p->flags & TFFT_INTERLEAVED) ? p->n : 2 * p->n;   // complex elements / plane elements
  int64_t b0 = 0;
  for (size_t ci = 0; ci < h->sizes.size(); ++ci) {
    const int64_t nb = h->sizes[ci];
    const int64_t off = 2 * p->n * b0, cnt = 2 * p->n * nb;
    b0 += nb;
    const int slot = static_cast<int>(ci % h->slots);
    __half* si = din + slot * chunk_halves;
    __half* so = dout + slot * chunk_halves;
    cudaEvent_t ev_up = h->events[3 * slot], ev_fft = h->events[3 * slot + 1], ev_down = h->events[3 * slot + 2];
    cudaError_t e = cudaSuccess;
    // slot reuse: the previous transform out of this input slot / download out of this output slot must be done
    if (ci >= static_cast<size_t>(h->slots)) e = cudaStreamWaitEvent(s_up, ev_fft, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(si, hin + off, cnt * sizeof(__half), cudaMemcpyHostToDevice, s_up);
    if (e == cudaSuccess) e = cudaEventRecord(ev_up, s_up);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s_fft, ev_up, 0);
    if (e == cudaSuccess && ci >= static_cast<size_t>(h->slots)) e = cudaStreamWaitEvent(s_fft, ev_down, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
    tfft_plan_s* cp = nb == p->batch ? p : h->plans[nb];
    const int rc = tfft_exec(cp, si, si + p->n, so, so + p->n, tstride, tstride, s_fft);
    if (rc != TFFT_OK) return rc;
    e = cudaEventRecord(ev_fft, s_fft);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s_down, ev_fft, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hout + off, so, cnt * sizeof(__half), cudaMemcpyDeviceToHost, s_down);
    if (e == cudaSuccess) e = cudaEventRecord(ev_down, s_down);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  cudaError_t e = cudaStreamSynchronize(s_down);
  return e == cudaSuccess ? TFFT_OK : static_cast<int>(e);
}

int tfft_transpose_blocks(const void* src, void* dst, int64_t rows, int64_t cols, int64_t src_row_stride,
                          int64_t dst_row_stride, int64_t nb0, int64_t nb1, int64_t src_b0, int64_t src_b1, int64_t dst_b0,
                          int64_t dst_b1, void* stream_) {
  if (!src || !dst || rows < 64 || cols < 64 || (rows & 63) || (cols & 63) || nb0 < 1 || nb1 < 1 || nb0 * nb1 > 65535)
    return TFFT_E_INVALID_ARG;
  if (((src_row_stride | dst_row_stride | src_b0 | src_b1 | dst_b0 | dst_b1) & 7) || !aligned16(src) || !aligned16(dst))
    return TFFT_E_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return TFFT_E_NO_DEVICE;
  }
  const dim3 grid(static_cast<unsigned>(cols / 64), static_cast<unsigned>(rows / 64), static_cast<unsigned>(nb0 * nb1));
  if (grid.y > 65535) return TFFT_E_INVALID_ARG;
  transpose64_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __half*>(src), static_cast<__half*>(dst), src_row_stride, dst_row_stride, static_cast<int>(nb0),
      src_b0, src_b1, dst_b0, dst_b1);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TFFT_OK : static_cast<int>(e);
}

int tfft_copy_runs(const void* src, void* dst, int64_t run, int64_t n0, int64_t n1, int64_t n2, int64_t s0, int64_t s1,
                   int64_t s2, int64_t d0, int64_t d1, int64_t d2, void* stream_) {
  if (!src || !dst || run < 8 || (run & 7) || n0 < 1 || n1 < 1 || n2 < 1 || n0 * n1 * n2 > (int64_t(1) << 30))
    return TFFT_E_INVALID_ARG;
  if (((s0 | s1 | s2 | d0 | d1 | d2) & 7) || !aligned16(src) || !aligned16(dst)) return TFFT_E_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return TFFT_E_NO_DEVICE;
  }
  const int64_t vec = run / 8;
  const unsigned gy = static_cast<unsigned>(std::min<int64_t>(std::max<int64_t>(1, vec / 1024), 64));
  const dim3 grid(static_cast<unsigned>(n0 * n1 * n2), gy);
  copy_runs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __half*>(src), static_cast<__half*>(dst), run, static_cast<int>(n0), static_cast<int>(n1), s0, s1, s2,
      d0, d1, d2);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TFFT_OK : static_cast<int>(e);
}

int tfft_fixture_sine(void* re, void* im, int64_t n, int64_t batch, int64_t stride, const float* w_re,
                      const float* w_im, int32_t cutoff, void* stream_) {
  if (!re || !im || !w_re || !w_im || n < 1 || batch < 1 || batch > 65535 || cutoff < 1 || stride < n) return TFFT_E_INVALID_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* d_w = nullptr;
  const size_t wbytes = static_cast<size_t>(batch) * cutoff * sizeof(float);
  cudaError_t e = cudaMalloc(&d_w, 2 * wbytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TFFT_E_NO_DEVICE
           : e == cudaErrorMemoryAllocation                            ? TFFT_E_NOMEM
                                                                       : static_cast<int>(e);
  }
  e = cudaMemcpyAsync(d_w, w_re, wbytes, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_w + static_cast<size_t>(batch) * cutoff, w_im, wbytes, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    const dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(batch));
    sine_fixture_kernel<<<grid, 256, 0, stream>>>(static_cast<__half*>(re), static_cast<__half*>(im), n, stride, d_w,
                                                  d_w + static_cast<size_t>(batch) * cutoff, cutoff);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // the weights staging buffer is freed below
  cudaFree(d_w);
  return e == cudaSuccess ? TFFT_OK : static_cast<int>(e);
}

int tfft_error_stats(const void* a_re, const void* a_im, const double* b_re, const double* b_im, int64_t count,
                     double* out4, void* stream_) {
  if (!a_re || !a_im || !b_re || !b_im || !out4 || count < 1) return TFFT_E_INVALID_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  double* d_acc = nullptr;
  cudaError_t e = cudaMalloc(&d_acc, 4 * sizeof(double));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TFFT_E_NO_DEVICE : static_cast<int>(e);
  }
  e = cudaMemsetAsync(d_acc, 0, 4 * sizeof(double), stream);
  if (e == cudaSuccess) {
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((count + 255) / 256, 148 * 8));
    deviation_sums_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(a_re), static_cast<const __half*>(a_im), b_re,
                                                    b_im, count, d_acc, reinterpret_cast<unsigned long long*>(d_acc + 3));
    e = cudaGetLastError();
  }
  double h[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_acc, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d_acc);
  if (e != cudaSuccess) return static_cast<int>(e);
  const double cnt = 2.0 * static_cast<double>(count), avg = h[0] / cnt;
  out4[0] = h[3];                                                             // largest deviation
  out4[1] = avg;                                                              // average deviation
  out4[2] = std::sqrt(std::max(0.0, h[1] - 2.0 * avg * h[0] + cnt * avg * avg) / (cnt - 1.0));   // sigma
  out4[3] = h[2] > 0.0 ? std::sqrt(h[1] / h[2]) : 0.0;                        // relative L2 error
  return TFFT_OK;
}

}  // extern "C"
