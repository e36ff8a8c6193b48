// Fused shared-memory-resident FFT kernel for sm_100a.  One CTA transforms one *unit*
// (U transforms of length L = T*16^s, E = U*L <= 32768 complex elements) and touches HBM
// exactly once: planar fp16 in, planar fp16 out.
//
//   load    16-byte cp.async global->shared straight into the stage-1 tensor-core operand layout
//           (replaces the reference's digit-reversal gather, src/base/TensorFFT256.cu:125-171)
//   stage   radix-16/32/64 DFT as tcgen05.mma (M=128 rows of data x N=2R [re|im] x K=16 steps),
//           fp32 accumulation in tensor memory; epilogue tcgen05.ld -> packed-fp32 twiddle ->
//           fp16 -> 16-byte STS into the next stage's operand layout (replaces the wmma stages
//           of src/base/TensorFFT256.cu:191-275 and src/base/TensorRadix16.cu:101-213, which
//           accumulate in fp16 and pay one HBM round trip per radix-16 step, and the
//           Radix2Kernel launches of src/base/Radix2.cu:20-77 / ComputeFFT.h:123-145: the
//           non-power-of-16 factor is folded into a radix-32/64 tensor stage)
//   store   16-byte LDS -> 8x8 in-register transpose -> 16-byte coalesced STG.
// CTAs are persistent (grid = min(units, resident CTAs)) and loop over units.
// All index maps come from the host-side UnitPlan (unit_plan.h).  Scaling: 1/R_t is folded
// into each fp16 DFT matrix (exact powers of two), total 1/L like the reference's "sequential
// scaling" (TensorFFT256.cu:167-171, TensorRadix16.cu:133-136, Radix2.cu:64-76).
#pragma once
#include <type_traits>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "sm100_ptx.cuh"
#include "unit_plan.h"

namespace tfft {

constexpr int kThreads = 256;
// Inter-stage twiddle seeds, two implementations chosen per plan shape (measured on B200, same box A/B):
//   3-stage plans (N = 8192 .. 32768, the two-slot kernel): per-thread seeds in registers times per-tile factors from
//     kernel-parameter space -- no shared-memory loads in the epilogue (C2 -5 %; the table lookups were the only
//     instructions with excess shared wavefronts and sat at the head of every item's dependency chain);
//   2-stage plans (N <= 4096 and every pass of a multi-pass plan): the round-1 two-level shared-memory table (two LDS.64
//     per seed).  With two CTAs per SM the lookup latency is hidden and the seed registers cost spills:
//     N = 1024 0.430 against 0.467 ms per GiB, 2^16 .. 2^20 2 - 4 % (profiles/r02_twiddle_seed_ab.txt).
constexpr uint32_t kTwTableBytesMax = 512 + 4096;   // [TWlo: 64 float2 | TWhi: 512 float2]
__host__ __device__ inline uint32_t tw_table_bytes(const UnitPlan& p) { return p.stages == 3 ? 0u : kTwTableBytesMax; }

// ---------------------------------------------------------------- packed fp32 pairs (FFMA2)
// Blackwell issues two fp32 FMAs per lane per instruction on 64-bit register pairs
// (fma.rn.f32x2); the kernel is issue-bound, so all twiddle arithmetic runs on pairs of
// neighbouring elements.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// negation through the scalar halves: ptxas folds it into the operand modifier of FFMA2 / FMUL2
__device__ __forceinline__ f32x2 neg2(f32x2 a) {
  float lo, hi;
  upk(a, lo, hi);
  return pk(-lo, -hi);
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { return add2(a, neg2(b)); }
__device__ __forceinline__ uint32_t pack_half2_pair(f32x2 v) {
  float lo, hi;
  upk(v, lo, hi);
  return ptx::pack_half2(lo, hi);
}
// two complex numbers, planar: re = (re0, re1), im = (im0, im1)
struct C2 {
  f32x2 re, im;
};
__device__ __forceinline__ C2 cmul2(C2 a, C2 b) {
  return {fma2(neg2(a.im), b.im, mul2(a.re, b.re)), fma2(a.im, b.re, mul2(a.re, b.im))};
}
__device__ __forceinline__ C2 cadd2(C2 a, C2 b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ C2 csub2(C2 a, C2 b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
__device__ __forceinline__ C2 cmul2_mi(C2 a) { return {a.im, neg2(a.re)}; }   // * (-i)
__device__ __forceinline__ C2 cbcast(float re, float im) { return {pk(re, re), pk(im, im)}; }

struct Cplx {
  float re, im;
};
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) {
  return {fmaf(a.re, b.re, -a.im * b.im), fmaf(a.re, b.im, a.im * b.re)};
}
// exp(-2*pi*i * x / 2^log2n) for an integer phase x (already reduced mod 2^log2n or small)
__device__ __forceinline__ Cplx twiddle(uint32_t x, uint32_t log2n) {
  float s, c;
  // sincospif is exact in its range reduction; -2x/2^log2n is an exact fp32 product
  sincospif(-static_cast<float>(x) * __uint_as_float((128u - log2n) << 23), &s, &c);  // 2^(1-log2n)
  return {c, s};
}

// same for phases up to 2^30: the low 12 bits and the rest are looked up separately so that both
// fractions are exact in fp32
__device__ __forceinline__ Cplx twiddle64(uint64_t x, uint32_t log2n) {
  if (log2n <= 24) return twiddle(static_cast<uint32_t>(x), log2n);
  const Cplx hi = twiddle(static_cast<uint32_t>(x >> 12), log2n - 12);
  const Cplx lo = twiddle(static_cast<uint32_t>(x & 0xFFFu), log2n);
  return cmul(hi, lo);
}

// Inter-pass twiddles (tw_mode 2: exp(-2*pi*i * o * column / N), two per work item in the last epilogue of a column
// pass): the phase fraction is exact in fp32 up to 24 bits (split in two exact parts above), so one range-reduced
// MUFU sine / cosine (absolute error < 5e-7, three orders below an fp16 ulp) replaces two sincospif polynomials --
// in the 4096 x 8-column pass those were 20 % of the executed instructions.  -DTFFT_EXACT_TWIDDLE keeps sincospif.
__device__ __forceinline__ Cplx twiddle_fast64(uint64_t x, uint32_t log2n) {
#ifdef TFFT_EXACT_TWIDDLE
  return twiddle64(x, log2n);
#else
  // f = x / 2^log2n in [0, 1): high and low 12-bit parts are exact, their sum is rounded once (6e-8)
  const float hi = static_cast<float>(static_cast<uint32_t>(x >> 12)) * __uint_as_float((127u + 12u - log2n) << 23);
  const float lo = static_cast<float>(static_cast<uint32_t>(x & 0xFFFu)) * __uint_as_float((127u - log2n) << 23);
  float f = hi + lo;
  f -= (f >= 0.5f) ? 1.0f : 0.0f;   // [-0.5, 0.5): the argument of the MUFU stays inside [-pi, pi]
  float s, c;
  __sincosf(-6.283185307179586f * f, &s, &c);
  return {c, s};
#endif
}

__device__ __forceinline__ uint32_t bit_sum(uint32_t q, const uint32_t* contrib, int first, int count) {
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < count; ++i)
    if ((q >> i) & 1u) s += contrib[first + i];
  return s;
}

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// The transform streams: every byte is read once and written once.  The TMA loads carry an L2 evict-first policy
// (C2 on B200: -1.6 %; -DTFFT_NO_L2_HINTS builds without).  The same policy on the global stores was measured to cost
// 1.7 % (123.2 -> 121.3 us without it on the same box), so stores carry no hint (-DTFFT_L2_HINT_STORE adds it).
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void stg128(__half* p, uint4 v) {
#if defined(TFFT_L2_HINT_STORE)
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "l"(l2_evict_first())
               : "memory");
#else
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
#endif
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ f32x2 unpack_half2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  float2 f = __half22float2(h);
  return pk(f.x, f.y);
}

// Device tables (built on the host, tfft_api.cu make_tables; staged into shared memory once per CTA):
//   [0, 512)    TWlo[j]  = exp(-2*pi*i * j / L),        j < 64
//   [512, 4608) TWhi[j]  = exp(-2*pi*i * 64*j / L),     j < L/64 <= 512
//   then for every distinct radix R of the plan ONE matrix T = [Fr | Fi | -Fr] / R of 3R columns (fp16, K-major
//   SWIZZLE_NONE: Tmath[kappa][n] at (n>>3)*16R + (kappa>>3)*128 + (n&7)*16 + (kappa&7)*2 bytes), F = exp(-2*pi*i*kappa*k/R).
//   The two B operands of a stage are windows of it: B1 = [Fr | Fi] = columns 0 .. 2R-1 and B2' = [Fi | -Fr] = columns
//   R .. 3R-1; D = A_re * B1 + (-A_im) * B2' with the tensor core negating A (instruction-descriptor bit 13), which is
//   A_re*[Fr|Fi] + A_im*[-Fi|Fr].  6 R^2 bytes per radix instead of 8 R^2 (24 KiB instead of 32 KiB for R = 64: two
//   16K-element CTAs of a 2048-point plan fit one SM).  -DTFFT_TWO_MATRICES builds the round-1 layout [Fr|Fi], [-Fi|Fr].
struct TableLayout {
  uint32_t b_off[kMaxStages];   // byte offset of B1 of stage t (B2 follows at + 4*R*R)
  uint32_t total;               // multiple of 16
};
__host__ __device__ inline TableLayout table_layout(const UnitPlan& p) {
  TableLayout l;
  uint32_t off = tw_table_bytes(p);
  for (uint32_t t = 0; t < kMaxStages; ++t) l.b_off[t] = 0;
  for (uint32_t t = 0; t < p.stages; ++t) {
    bool found = false;
    const bool kron = p.kron_bits && t + 1 == p.stages;   // F_x (x) F_y: never shared with a plain stage
    for (uint32_t u = 0; u < t && !kron; ++u)
      if (p.log2_radix[u] == p.log2_radix[t]) { l.b_off[t] = l.b_off[u]; found = true; break; }
    if (!found) {
      l.b_off[t] = off;
#ifdef TFFT_TWO_MATRICES
      off += 8u << (2 * p.log2_radix[t]);   // 2 matrices of 2R x R halves
#else
      off += 6u << (2 * p.log2_radix[t]);   // one matrix of 3R x R halves
#endif
    }
  }
  l.total = off;
  return l;
}

// Dynamic shared memory carve-up (bytes): [plane_re | plane_im | tables | mbar | tmem slot]
struct SmemLayout {
  uint32_t plane_stride, table_off, bar_off, load_bar_off, slot_off, ytw_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(const UnitPlan& p) {
  SmemLayout l;
  l.plane_stride = (p.plane_bytes + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-byte alignment
  l.table_off = 2 * l.plane_stride;
  l.bar_off = l.table_off + table_layout(p).total;   // two MMA barriers
  l.load_bar_off = l.bar_off + 16;
  l.slot_off = l.load_bar_off + 8;
  l.ytw_off = l.slot_off + 8;        // Kronecker units: 8 float2 row twiddles of the current unit
  l.total = l.ytw_off + 64;
  return l;
}

// 5-D variant for Kronecker units: {64 rows, R kappa, M/64, y_lo, u + U*image}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t c3, uint32_t c4,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %2, %2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(0), "r"(c3), "r"(c4), "r"(ptx::smem_u32(bar))
      : "memory");
}

// L2 prefetches of the NEXT unit's input (no shared memory needed), enabled per plan (UnitPlan::prefetch_next).
// Measured on B200: helps where nothing else overlaps the load phase (32K-element units, one CTA per SM: four-step
// sizes 2^22..2^24 -5 %; N = 2048 -9 %), hurts where loads already overlap compute (two-slot kernel at C2: 118.6 ->
// 124.6 us; 16K-element column passes), so the planner switches it on only for the former.
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, uint32_t c0, uint32_t c2, uint32_t c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(0),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, uint32_t c3, uint32_t c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %1, %1, %2, %3}];" ::"l"(map), "r"(0), "r"(c3),
               "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d_col(const CUtensorMap* map, uint32_t c0, uint32_t c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %2, %3}];" ::"l"(map), "r"(c0), "r"(0),
               "r"(c3)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// column-mode tile {8 columns, R kappa, M rows, 1 batch}: coordinates (first column, 0, 0, batch)
__device__ __forceinline__ void tma_load_4d_col(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c3,
                                                uint64_t* bar, uint32_t c2 = 0) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(0), "r"(c2), "r"(c3), "r"(ptx::smem_u32(bar))
      : "memory");
}

// 5-D column-mode tile {W columns, R kappa, M rows, 1 batch, 1 outer batch}: batched three-pass plans, whose passes have a
// batch level of their own (UnitPlan::tma_b3_step == 2)
__device__ __forceinline__ void tma_load_5d_col(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c2, uint32_t c3,
                                                uint32_t c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(0), "r"(c2), "r"(c3), "r"(c4), "r"(ptx::smem_u32(bar))
      : "memory");
}

// 4-D TMA tile load {64 rows, R kappa, M/64, U transforms} -> SWIZZLE_128B stage-1 operand plane
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t c2, uint32_t c3,
                                            uint64_t* bar) {
#ifndef TFFT_NO_L2_HINTS
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %2, %3, %4}], [%5], %6;"
      ::"r"(dst), "l"(map), "r"(0), "r"(c2), "r"(c3), "r"(ptx::smem_u32(bar)), "l"(l2_evict_first())
      : "memory");
#else
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(0), "r"(c2), "r"(c3), "r"(ptx::smem_u32(bar))
      : "memory");
#endif
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// exp(-2*pi*i * x / L) from the two-level shared-memory table, x < L
__device__ __forceinline__ Cplx tw_lookup(const float2* tw_table, uint32_t x) {
  const float2 lo = tw_table[x & 63u], hi = tw_table[64u + (x >> 6)];
  return cmul({lo.x, lo.y}, {hi.x, hi.y});
}

// Optional phase trace (developer builds, -DTFFT_TRACE): thread 0 of every CTA records clock64 at
// phase boundaries of its first units into a global buffer (tools/trace_phases.py).
#ifdef TFFT_TRACE
#ifndef TFFT_TRACE_THREAD
#define TFFT_TRACE_THREAD 0   // thread of each slot that records (0 = the UMMA-issuing thread)
#endif
#define TFFT_TRACE_TID0 ((threadIdx.x & 255) == TFFT_TRACE_THREAD && threadIdx.x < 512)
#define TFFT_TRACE_SLOT (threadIdx.x >> 8)
#ifndef TFFT_TRACE_FIRST
#define TFFT_TRACE_FIRST 0
#endif
#define TFFT_TRACE_MARK(slot)                                                                       \
  do {                                                                                               \
    if (TFFT_TRACE_TID0 && trace != nullptr && trace_unit >= TFFT_TRACE_FIRST &&            \
        trace_unit < TFFT_TRACE_FIRST + 4)                                                           \
      trace[((static_cast<size_t>(blockIdx.x) * 2 + TFFT_TRACE_SLOT) * 4 + trace_unit -           \
             TFFT_TRACE_FIRST) * 32 + (slot)] = clock64();                                           \
  } while (0)
#else
#define TFFT_TRACE_MARK(slot) do {} while (0)
#endif

// Programmatic dependent launch: every kernel lets its successor's CTAs become resident as soon as SMs free up
// (launch_dependents at entry) and itself waits for its predecessor's completion and memory flush only after its own
// prologue (tensor-memory allocation, barrier setup, constant tables), right before it first touches data.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- thread-block cluster (CTA pair) helpers
// Cluster units (UnitPlan::cluster): two CTAs share a 64K-element unit; the stage-1 epilogue stores into either CTA's
// shared memory (distributed shared memory), everything else is local.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, uint32_t rank) {   // shared::cta address -> CTA `rank`'s window
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs; release / acquire at cluster scope
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts128_cluster(uint32_t caddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// generic-proxy stores into either CTA's shared memory -> visible to the async proxy (tensor core) of the owning CTA
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }

struct KernelCtx {
  uint32_t cl_rank;        // cluster units: rank of this CTA in its pair (0 otherwise)
  uint32_t cl_delta[2];    // shared::cluster address of CTA r's window minus this CTA's shared::cta base
  uint32_t sbase, s_re, s_im, taddr, lane_row, wgroup, lane_base;
  const float2* tw_table;
  const float2* ytw;       // Kronecker units: exp(-2*pi*i * k_y * y_lo / ny), k_y < 8, of the current unit
  uint32_t col_base;
  uint32_t a_re, a_im;     // stage-1 operand planes (== s_re / s_im unless a separate landing buffer is used)
  uint32_t bar_id;         // 0: the 256 threads are the whole CTA (__syncthreads); else named barrier of a slot
  uint32_t sync_threads;   // threads of the slot's named barrier
  int mma_warp;            // warp that issues the UMMAs: 0 (it also runs epilogues) or a dedicated 9th warp (8)
};
__device__ __forceinline__ void group_sync(const KernelCtx& c) {
  if (c.bar_id == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(c.bar_id), "r"(c.sync_threads) : "memory");
}

// sum of the contributions of the set bits of a COMPILE-TIME index: folds into constant-bank adds
template <uint32_t Q, int COUNT>
__device__ __forceinline__ uint32_t bit_sum_c(const uint32_t* contrib, int first) {
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < COUNT; ++i)
    if ((Q >> i) & 1u) s += contrib[first + i];
  return s;
}

// Epilogue of work item (2*II + wgroup) of stage ST: one 128-row tile x 16 output columns.
//   RHO: log2 radix, LAST: final stage (no inter-stage twiddle), items are numbered
//   item = tile * G + g (G = R/16 column groups per tile); bit 0 of the item index is the warp group.
// Work items of a stage: item = tile * G + g (G = R/16 column groups per 128-row tile).  Warp group w of NG (2 or 4)
// processes items NG*II + w, II = 0, 1, ...: the low bits of the item index come from the warp group, the rest is
// compile-time.
template <int NG, uint32_t G>
__host__ __device__ constexpr uint32_t item_tile_hi(uint32_t ii) { return (NG * ii) / G; }       // multiple of max(1, NG/G)
template <int NG, uint32_t G>
__host__ __device__ constexpr uint32_t item_g_hi(uint32_t ii) { return G > NG ? (NG * ii) % G : 0; }   // multiple of NG
template <int NG, uint32_t G>
__device__ __forceinline__ uint32_t item_tile_lo(uint32_t wgroup) { return G >= NG ? 0u : wgroup / G; }
template <int NG, uint32_t G>
__device__ __forceinline__ uint32_t item_g_lo(uint32_t wgroup) { return G >= NG ? wgroup : wgroup % G; }

template <int RHO, uint32_t II, int NG = 2>
__device__ __forceinline__ void epilogue_load(const KernelCtx& c, uint32_t (&are)[16], uint32_t (&aim)[16]) {
  constexpr uint32_t R = 1u << RHO, G = R / 16;
  const uint32_t tile = item_tile_hi<NG, G>(II) + item_tile_lo<NG, G>(c.wgroup);
  const uint32_t g = item_g_hi<NG, G>(II) + item_g_lo<NG, G>(c.wgroup);
  const uint32_t tcol = c.taddr + c.lane_base + tile * 2 * R + g * 16;
  ptx::tmem_ld_32x32b_x16(tcol, are);
  ptx::tmem_ld_32x32b_x16(tcol + R, aim);
}

// Per-thread constant part of the bit-linear row maps of stage ST (7 lane-row bits + the warp-group bits), computed once
// per kernel: (dst >> 4) | (aux << 16).  dst is a multiple of 16 below 2^20, aux (twiddle integer) is below 2^16.
template <int ST, int RHO, int NG>
__device__ __forceinline__ uint32_t thread_map(const UnitPlan& P, const KernelCtx& c) {
  constexpr uint32_t G = (1u << RHO) / 16;
  const UnitPlan::Epi& E = P.epi[ST];
  uint32_t dst = bit_sum(c.lane_row, E.dst, 0, 7), aux = bit_sum(c.lane_row, E.aux, 0, 7);
  if (P.cluster && c.cl_rank) {   // the CTA's rank is an index bit: the input half in stage 1, the k_1 half afterwards
    if (ST == 0) { dst += P.cl_in_dst; aux += P.cl_in_aux; }
    else if (ST + 1 == static_cast<int>(P.stages)) aux += P.cl_out_aux;
  }
  // warp-group bits: first the g bits (k_t[4], k_t[5]), then tile bits (row bits 7, 8)
  constexpr int kGBits = G >= 4 ? 2 : (G == 2 ? 1 : 0), kWBits = NG == 4 ? 2 : 1;
#pragma unroll
  for (int b = 0; b < kWBits; ++b) {
    if (!((c.wgroup >> b) & 1u)) continue;
    if (b < kGBits) dst += E.dst_k[1 + b];
    else { dst += E.dst[7 + b - kGBits]; aux += E.aux[7 + b - kGBits]; }
  }
  return (dst >> 4) | (aux << 16);
}
// Per-thread twiddle seeds of a non-last stage: w = exp(-2*pi*i*x/L), w16 = w^16 for x = (per-thread part of the
// twiddle integer) << tw_shift.  sincospif on exact binary fractions, once per kernel.
struct TwSeed {
  Cplx w, w16;
};
__device__ __forceinline__ TwSeed thread_seed(const UnitPlan& P, int st, uint32_t tmap) {
  const uint32_t mask = (1u << P.log2_len) - 1u;
  const uint32_t x = ((tmap >> 16) << P.epi[st].tw_shift) & mask;
  TwSeed s;
  s.w = twiddle(x, P.log2_len);
  s.w16 = twiddle((x * 16u) & mask, P.log2_len);
  return s;
}

template <int ST, int RHO, int NG>
__device__ __forceinline__ uint32_t thread_col(const UnitPlan& P, const KernelCtx& c) {   // last stage, tw_mode 2
  constexpr uint32_t G = (1u << RHO) / 16;
  const UnitPlan::Epi& E = P.epi[ST];
  uint32_t col = bit_sum(c.lane_row, E.col, 0, 7);
  constexpr int kGBits = G >= 4 ? 2 : (G == 2 ? 1 : 0), kWBits = NG == 4 ? 2 : 1;
#pragma unroll
  for (int b = kGBits; b < kWBits; ++b)
    if ((c.wgroup >> b) & 1u) col += E.col[7 + b - kGBits];
  return col;
}

template <int ST, int RHO, bool LAST, uint32_t II, int NG = 2, bool CL = false, bool TWT = false>
__device__ __forceinline__ void epilogue_item(const UnitPlan& P, const KernelCtx& c, uint32_t dst_thr, uint32_t aux_thr,
                                              uint32_t col_thr, const TwSeed& seed, const uint32_t (&are)[16],
                                              const uint32_t (&aim)[16]) {
  using namespace ptx;
  constexpr uint32_t R = 1u << RHO, G = R / 16;
  const UnitPlan::Epi& E = P.epi[ST];
  // item = NG*II + wgroup;  tile = item / G, g = item % G
  constexpr uint32_t kTileHi = item_tile_hi<NG, G>(II);   // compile-time part of the tile index
  constexpr uint32_t kGHi = item_g_hi<NG, G>(II);         // compile-time part of g
  const uint32_t g = kGHi + item_g_lo<NG, G>(c.wgroup);
  // dst_thr / aux_thr already hold the per-thread part including the warp-group bits
  constexpr uint32_t kTileShift = G >= NG ? 0 : (NG / G == 2 ? 1 : 2);   // tile bits below this come from the warp group
  uint32_t dst = dst_thr + bit_sum_c<(kTileHi >> kTileShift), kMaxRowBits - 7 - kTileShift>(E.dst, 7 + kTileShift);
  if (G == 4 && NG == 2) dst += bit_sum_c<(kGHi >> 1), 1>(E.dst_k, 2);   // k_t[5]
  uint32_t pre[8], pim[8];
  bool plain = LAST && E.tw_mode == 0;
  if (plain) {
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      pre[k >> 1] = pack_half2(__uint_as_float(are[k]), __uint_as_float(are[k + 1]));
      pim[k >> 1] = pack_half2(__uint_as_float(aim[k]), __uint_as_float(aim[k + 1]));
    }
  } else {
    const uint32_t aux = aux_thr + bit_sum_c<(kTileHi >> kTileShift), kMaxRowBits - 7 - kTileShift>(E.aux, 7 + kTileShift);
    Cplx t0, t1, s2;
    if (LAST && E.tw_mode == 3) {
      // Kronecker stage: columns k of this item belong to output row k_y = k >> tw_shift (tw_shift >= 3, so
      // each 8-column chunk has one k_y); factor exp(-2*pi*i*k_y*y_lo/ny) from the per-unit table
      const float2 w0 = c.ytw[(16u * g) >> E.tw_shift], w1 = c.ytw[(16u * g + 8u) >> E.tw_shift];
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const f32x2 wr = k < 8 ? pk(w0.x, w0.x) : pk(w1.x, w1.x), wi = k < 8 ? pk(w0.y, w0.y) : pk(w1.y, w1.y);
        const f32x2 xr = pk(__uint_as_float(are[k]), __uint_as_float(are[k + 1]));
        const f32x2 xi = pk(__uint_as_float(aim[k]), __uint_as_float(aim[k + 1]));
        pre[k >> 1] = pack_half2_pair(fma2(neg2(xi), wi, mul2(xr, wr)));
        pim[k >> 1] = pack_half2_pair(fma2(xi, wr, mul2(xr, wi)));
      }
      sts128(c.s_re + dst, make_uint4(pre[0], pre[1], pre[2], pre[3]));
      sts128(c.s_re + dst + E.dst_k[0], make_uint4(pre[4], pre[5], pre[6], pre[7]));
      sts128(c.s_im + dst, make_uint4(pim[0], pim[1], pim[2], pim[3]));
      sts128(c.s_im + dst + E.dst_k[0], make_uint4(pim[4], pim[5], pim[6], pim[7]));
      return;
    }
    if (!LAST) {
     if constexpr (TWT) {   // two-level shared-memory table (2-stage plans)
      const uint32_t idx = aux << E.tw_shift;                       // unit angle 2*pi/L
      const Cplx w1 = tw_lookup(c.tw_table, idx);
      s2 = cmul(w1, w1);
      if (G == 1) {
        t0 = {1.f, 0.f};
        t1 = w1;
      } else {
        t0 = tw_lookup(c.tw_table, (idx * 16u * g) & ((1u << P.log2_len) - 1u));
        t1 = cmul(t0, w1);
      }
     } else {
      // w1 = exp(-2*pi*i*m/N_t) of this row = (per-thread seed) * (factor of the tile, warp-uniform, parameter space)
      (void)aux;
      const Cplx w1 = cmul(seed.w, {E.tile_tw[kTileHi][0], E.tile_tw[kTileHi][1]});
      s2 = cmul(w1, w1);
      if (G == 1) {
        t0 = {1.f, 0.f};
        t1 = w1;
      } else {
        // t0 = w1^(16 g), g < G <= 4: bit 0 of g selects w1^16, bit 1 another factor w1^32
        const Cplx w16 = cmul(seed.w16, {E.tile_tw16[kTileHi][0], E.tile_tw16[kTileHi][1]});
        const bool g1 = (g & 1u) != 0;
        const bool g2 = G == 4 && (NG == 2 ? (kGHi & 2u) != 0 : (g & 2u) != 0);
        t0 = {g1 ? w16.re : 1.f, g1 ? w16.im : 0.f};
        if (g2) t0 = cmul(t0, cmul(w16, w16));
        t1 = cmul(t0, w1);
      }
     }
    } else {
      const uint64_t col = col_thr + bit_sum_c<(kTileHi >> kTileShift), kMaxRowBits - 7 - kTileShift>(E.col, 7 + kTileShift) + c.col_base;
      const uint64_t mask = (uint64_t(1) << E.tw_log2n) - 1u;
      const Cplx w1 = twiddle_fast64((E.tw_kw * col) & mask, E.tw_log2n);
      t0 = twiddle_fast64(((aux + 16u * g * E.tw_kw) * col) & mask, E.tw_log2n);
      t1 = cmul(t0, w1);
      s2 = cmul(w1, w1);
    }
    // twiddles of columns (k, k+1) as packed pairs P_j = (t_2j, t_2j+1).  P_1 = P_0 * step^2 is a full
    // complex product; after that the three-term recurrence P_{j+1} = 2cos(2 theta) P_j - P_{j-1} (exact for
    // a geometric sequence on the unit circle) advances a pair with ONE packed FMA per component instead
    // of a packed complex product (two).  Seven steps from exact seeds: error <= ~50 ulp(fp32) ~ 3e-6.
    f32x2 tre = pk(t0.re, t1.re), tim = pk(t0.im, t1.im);
    const f32x2 sre = pk(s2.re, s2.re), sim = pk(s2.im, s2.im);
    const f32x2 c2 = pk(2.f * s2.re, 2.f * s2.re);
    f32x2 qre = tre, qim = tim;   // P_{j-1}
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const f32x2 xr = pk(__uint_as_float(are[k]), __uint_as_float(are[k + 1]));
      const f32x2 xi = pk(__uint_as_float(aim[k]), __uint_as_float(aim[k + 1]));
      pre[k >> 1] = pack_half2_pair(fma2(neg2(xi), tim, mul2(xr, tre)));
      pim[k >> 1] = pack_half2_pair(fma2(xi, tre, mul2(xr, tim)));
      if (k == 0) {
        tre = fma2(neg2(qim), sim, mul2(qre, sre));
        tim = fma2(qim, sre, mul2(qre, sim));
      } else if (k < 14) {
        const f32x2 nre = fma2(c2, tre, neg2(qre)), nim = fma2(c2, tim, neg2(qim));
        qre = tre;
        qim = tim;
        tre = nre;
        tim = nim;
      }
    }
  }
  if constexpr (CL && ST == 0) {
    // cluster unit, stage 1: the top bit of k_1 (bit RHO-4 of the 8-column chunk index 2g + h) selects the CTA whose
    // stage-2 operand receives the chunk
    const uint32_t da = (((2u * g) >> (RHO - 4)) & 1u) ? c.cl_delta[1] : c.cl_delta[0];
    const uint32_t db = (((2u * g + 1u) >> (RHO - 4)) & 1u) ? c.cl_delta[1] : c.cl_delta[0];
    sts128_cluster(c.s_re + dst + da, make_uint4(pre[0], pre[1], pre[2], pre[3]));
    sts128_cluster(c.s_re + dst + E.dst_k[0] + db, make_uint4(pre[4], pre[5], pre[6], pre[7]));
    sts128_cluster(c.s_im + dst + da, make_uint4(pim[0], pim[1], pim[2], pim[3]));
    sts128_cluster(c.s_im + dst + E.dst_k[0] + db, make_uint4(pim[4], pim[5], pim[6], pim[7]));
    return;
  }
  sts128(c.s_re + dst, make_uint4(pre[0], pre[1], pre[2], pre[3]));
  sts128(c.s_re + dst + E.dst_k[0], make_uint4(pre[4], pre[5], pre[6], pre[7]));
  sts128(c.s_im + dst, make_uint4(pim[0], pim[1], pim[2], pim[3]));
  sts128(c.s_im + dst + E.dst_k[0], make_uint4(pim[4], pim[5], pim[6], pim[7]));
}

// Software-pipelined item loop over items [II, END): the tensor-memory load of item II+1 is in flight
// while item II is processed (tcgen05.wait::ld waits for ALL outstanding loads, so the next load is
// issued right after the wait and before the arithmetic).
template <int ST, int RHO, bool LAST, uint32_t II, uint32_t END, int NG = 2, bool CL = false, bool TWT = false>
__device__ __forceinline__ void epilogue_range(const UnitPlan& P, const KernelCtx& c, uint32_t dst_thr,
                                               uint32_t aux_thr, uint32_t col_thr, const TwSeed& seed,
                                               uint32_t (&cre)[16], uint32_t (&cim)[16], uint32_t (&nre)[16],
                                               uint32_t (&nim)[16]) {
  if constexpr (II < END) {
    ptx::tmem_ld_wait();                                            // item II has landed in (cre, cim)
    if constexpr (II + 1 < END) epilogue_load<RHO, II + 1, NG>(c, nre, nim);
    epilogue_item<ST, RHO, LAST, II, NG, CL, TWT>(P, c, dst_thr, aux_thr, col_thr, seed, cre, cim);
    epilogue_range<ST, RHO, LAST, II + 1, END, NG, CL, TWT>(P, c, dst_thr, aux_thr, col_thr, seed, nre, nim, cre, cim);
  }
}

// every warp waits for the MMA barrier itself (one polling lane per warp)
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) ptx::mbar_wait(bar, parity);
  __syncwarp();
  ptx::tc_fence_after_sync();
}

// One tensor-core stage: all UMMAs (one thread), then the epilogue on all 8 warps.
// PIPE: the UMMAs are committed in two halves (by tile) and the epilogue of the first half overlaps
// the tensor-core work of the second half.  Only legal when the first-half epilogue cannot touch
// operand bytes the second-half MMAs still read: stage 1 reading a separate landing buffer, or a plan
// built with pipe_stage2 (checked on the CPU by tests/sim/plan_sim.cpp).
// Hooks of the MMA-issuing thread (warp 0, lane 0): before it issues the UMMAs of tile half h, and once it
// has observed their completion.  The two-slot kernel uses them for its landing-buffer ring.
// mid(): called by EVERY thread of a pipelined stage between the epilogues of its two tile halves.
struct NoHook {
  __device__ __forceinline__ void before_half(int) const {}
  __device__ __forceinline__ void after_half(int) const {}
  __device__ __forceinline__ void mid(int, int) const {}
};
template <int RHO, int LOG2E, bool PIPE, int NG = 2>
struct StageShape {
  static constexpr uint32_t R = 1u << RHO, G = R / 16, kSteps = R / 16;
  static constexpr uint32_t S = 16 * R + 16;                       // chunk stride of this stage's operand layout
  static constexpr uint32_t kTiles = (1u << LOG2E) / R / 128;
  static constexpr uint32_t kItemsPerGroup = (1u << LOG2E) / 2048 / NG;   // E/2048 work items per stage, NG warp groups
  static constexpr bool kPipe = PIPE && kTiles >= 2 && kItemsPerGroup >= 2;
};

// All UMMAs of one stage, issued by ONE thread (under elect_one()).
// PART: 0 = all tiles; 1 = only the first half of the tiles, committed to bar[0]; 2 = only the second half, committed
// to bar[1] (split issue of the last stage, see run_stage EARLY)
template <int ST, int RHO, int LOG2E, int LM, bool PIPE, class Hook, int NG = 2, int PART = 0>
__device__ __forceinline__ void stage_issue(const KernelCtx& c, uint32_t b1_saddr, uint64_t* bar, const Hook& hook,
                                            long long* trace, uint32_t trace_unit) {
  using namespace ptx;
  using SS = StageShape<RHO, LOG2E, PIPE, NG>;
  constexpr uint32_t R = SS::R, kSteps = SS::kSteps, kTiles = SS::kTiles;
  constexpr bool SW128 = (LM == 1 || LM == 5) && ST == 0;   // 5: column tiles of 64 columns, the same atoms of 64 rows
  constexpr bool SW32 = (LM == 3 || LM == 4) && ST == 0;   // 16-row atoms: LBO = atom stride 32R, SBO = K-group stride 256
  constexpr bool SW64 = LM == 6 && ST == 0;                // 32-row atoms (column tiles of 32 columns): LBO = 64R, SBO = 512
  constexpr uint32_t S = (LM == 2 && ST == 0) ? 16 * R : SS::S;   // column tiles loaded by TMA are dense: no padding
  constexpr bool kPipe = SS::kPipe;
  constexpr uint32_t idesc = make_idesc_f16(128, 2 * R, /*a_mn=*/1, /*b_mn=*/0);
  // descriptors differ only in the 14-bit start-address field (units of 16 bytes): add offsets there
  // A operand: SWIZZLE_NONE padded chunks (written by cp.async / epilogues), or for a TMA-loaded
  // stage 1 SWIZZLE_128B atoms of 64 rows (LBO = atom stride 128R, SBO = K-group stride 1024)
  constexpr uint64_t kSw128 = uint64_t(2) << 61;
  // an opaque zero pins the descriptor arithmetic below to this (single-thread) branch: without it the compiler
  // hoists ~100 integer instructions per stage into the code every warp executes
  uint32_t zero;
  asm volatile("mov.u32 %0, 0;" : "=r"(zero));
  b1_saddr += zero;
  const uint32_t pa_re = (ST == 0 ? c.a_re : c.s_re) + zero, pa_im = (ST == 0 ? c.a_im : c.s_im) + zero;
  const uint32_t taddr = c.taddr + zero;
  constexpr uint64_t kSw32 = uint64_t(6) << 61, kSw64 = uint64_t(4) << 61;
  const uint64_t da_re = SW128 ? (make_smem_desc(pa_re, 128 * R, 1024) | kSw128)
                         : SW64 ? (make_smem_desc(pa_re, 64 * R, 512) | kSw64)
                         : SW32 ? (make_smem_desc(pa_re, 32 * R, 256) | kSw32) : make_smem_desc(pa_re, kKGroupStride, S);
  const uint64_t da_im = SW128 ? (make_smem_desc(pa_im, 128 * R, 1024) | kSw128)
                         : SW64 ? (make_smem_desc(pa_im, 64 * R, 512) | kSw64)
                         : SW32 ? (make_smem_desc(pa_im, 32 * R, 256) | kSw32) : make_smem_desc(pa_im, kKGroupStride, S);
  constexpr uint32_t kTileStep = (SW128 || SW32 || SW64) ? (2 * 128 * R) / 16 : S;   // descriptor address units (16 B) per tile
  constexpr uint32_t kKStep = SW128 ? 2048 / 16 : SW64 ? 1024 / 16 : SW32 ? 512 / 16 : 16;      // ... per 16-wide K step
  const uint64_t db1 = make_smem_desc(b1_saddr, kKGroupStride, 16 * R);
#ifdef TFFT_TWO_MATRICES
  const uint64_t db2 = make_smem_desc(b1_saddr + 4 * R * R, kKGroupStride, 16 * R);
  constexpr uint32_t idesc2 = idesc;
#else
  const uint64_t db2 = make_smem_desc(b1_saddr + 2 * R * R, kKGroupStride, 16 * R);   // columns R .. 3R-1: [Fi | -Fr]
  constexpr uint32_t idesc2 = idesc | (1u << 13);                                       // negate A: (-A_im) * [Fi | -Fr]
#endif
  constexpr uint32_t kTileBegin = PART == 2 ? kTiles / 2 : 0, kTileEnd = PART == 1 ? kTiles / 2 : kTiles;
#pragma unroll
  for (uint32_t tile = kTileBegin; tile < kTileEnd; ++tile) {
    const uint32_t d = taddr + tile * 2 * R;
    if (tile == 0) {
      hook.before_half(0);
      TFFT_TRACE_MARK(16 + 4 * ST);
    }
    if (tile == (kTiles + 1) / 2) {
      hook.before_half(1);
      TFFT_TRACE_MARK(17 + 4 * ST);
    }
    if (kTiles == 1 && tile == 0) hook.before_half(1);
#pragma unroll
    for (uint32_t j = 0; j < kSteps; ++j)
      umma_f16_ss(d, da_re + (tile * kTileStep + j * kKStep), db1 + j * 16, idesc, j > 0 ? 1u : 0u);
#pragma unroll
    for (uint32_t j = 0; j < kSteps; ++j)
      umma_f16_ss(d, da_im + (tile * kTileStep + j * kKStep), db2 + j * 16, idesc2, 1u);
    if (PART == 0 && kPipe && tile + 1 == kTiles / 2) umma_commit(bar);     // first half of the tiles
  }
  if (PART == 1) umma_commit(bar);
  else umma_commit((kPipe || PART == 2) ? bar + 1 : bar);   // separate barriers: a parity wait must never fall two phases behind
  TFFT_TRACE_MARK(18 + 4 * ST);
}

// Dedicated UMMA warp: observe the completions of a stage (one lane) and run the landing-ring hooks.
template <int RHO, int LOG2E, bool PIPE, class Hook>
__device__ __forceinline__ void stage_observe(uint64_t* bar, uint32_t (&phase)[2], const Hook& hook) {
  constexpr bool kPipe = StageShape<RHO, LOG2E, PIPE>::kPipe;
  if (ptx::elect_one()) {
    ptx::mbar_wait(bar, phase[0] & 1u);
    hook.after_half(0);
    if (kPipe) ptx::mbar_wait(bar + 1, phase[1] & 1u);
    hook.after_half(1);
  }
  __syncwarp();
  phase[0]++;
  if (kPipe) phase[1]++;
}

// ROLE 0: every warp runs the epilogue and warp 0 also issues the UMMAs.
// ROLE 1: epilogue warp of a slot whose UMMAs are issued by a dedicated warp.  That warp issues the
//         stage-1 UMMAs of a unit ahead of time (during the previous unit's store phase), so stage 1 has
//         no leading barrier here.
// EARLY2: the first half of this stage's UMMAs was already issued (and committed to bar[0]) by the previous stage's
// mid() hook; only the second half is issued here (bar[1]) and the epilogue starts when both have completed.
// CL: cluster unit (CTA pair).  Stage 1 waits at a cluster barrier between its MMAs and its epilogue (the partner's MMAs
// must have consumed the partner's operand before this CTA stores into it); every later stage starts at a cluster barrier
// instead of the CTA barrier (the partner's stores into this CTA's operand must have landed).
template <int ST, int RHO, bool LAST, int LOG2E, int LM = 0, bool PIPE = false, class Hook = NoHook,
          int ROLE = 0, int NG = 2, bool EARLY2 = false, bool CL = false, bool TWT = false>
__device__ __forceinline__ void run_stage(const UnitPlan& P, const KernelCtx& c, uint32_t b1_saddr, uint64_t* bar,
                                          uint32_t (&phase)[2], int warp, int lane, long long* trace,
                                          uint32_t trace_unit, uint32_t tmap, uint32_t col_thr,
                                          const TwSeed& seed = TwSeed(), Hook hook = Hook()) {
  using namespace ptx;
  using SS = StageShape<RHO, LOG2E, PIPE, NG>;
  constexpr uint32_t kItemsPerGroup = SS::kItemsPerGroup;
  constexpr bool kPipe = SS::kPipe;
  if (ROLE == 0 || ST > 0) {
    if (CL && ST == 1) {
      fence_proxy_async_all();   // this CTA's stores went into both CTAs' operands
      tc_fence_before_sync();
      cluster_sync_all();
      fence_proxy_async_all();
    } else {
      fence_proxy_async_smem();   // generic-proxy / cp.async operand writes -> visible to the tensor core
      tc_fence_before_sync();
      group_sync(c);
    }
    tc_fence_after_sync();
  }
  TFFT_TRACE_MARK(9 + 2 * ST);
  if (ROLE == 0 && warp == 0 && elect_one())
    stage_issue<ST, RHO, LOG2E, LM, PIPE, Hook, NG, (EARLY2 ? 2 : 0)>(c, b1_saddr, bar, hook, trace, trace_unit);
  const bool hook_warp = ROLE == 0 && warp == 0;   // converged at every use below (after warp_wait)
  // per-thread parts of the bit-linear row maps (thread_map(): 7 lane-row bits + the warp-group bits)
  const uint32_t dst_thr = (tmap & 0xFFFFu) << 4, aux_thr = tmap >> 16;
  uint32_t ra[16], rb[16], rc[16], rd[16];
#if defined(TFFT_DEBUG_SKIP) || defined(TFFT_DEBUG_MMA_ONLY)
  if (true) {   // developer experiment: no epilogue work, only the barrier protocol
    warp_wait(bar, phase[0] & 1u, lane);
    if (hook_warp && elect_one()) hook.after_half(0);
    if (kPipe) warp_wait(bar + 1, phase[1] & 1u, lane);
    if (hook_warp && elect_one()) hook.after_half(1);
    phase[0]++;
    if (kPipe) phase[1]++;
    return;
  }
#endif
  if constexpr (kPipe) {
    // items [0, half) only touch tiles of the first half for every radix (item = NG*II + wgroup)
    constexpr uint32_t kHalf = kItemsPerGroup / 2;
    warp_wait(bar, phase[0] & 1u, lane);
    if (hook_warp && elect_one()) hook.after_half(0);
    TFFT_TRACE_MARK(10 + 2 * ST);
    epilogue_load<RHO, 0, NG>(c, ra, rb);
    epilogue_range<ST, RHO, LAST, 0, kHalf, NG, false, TWT>(P, c, dst_thr, aux_thr, col_thr, seed, ra, rb, rc, rd);
    hook.mid(warp, lane);
    warp_wait(bar + 1, phase[1] & 1u, lane);
    if (hook_warp && elect_one()) hook.after_half(1);
    epilogue_load<RHO, kHalf, NG>(c, ra, rb);
    epilogue_range<ST, RHO, LAST, kHalf, kItemsPerGroup, NG, false, TWT>(P, c, dst_thr, aux_thr, col_thr, seed, ra, rb, rc, rd);
    phase[0]++;
    phase[1]++;
  } else {
    warp_wait(bar, phase[0] & 1u, lane);
    if (EARLY2) warp_wait(bar + 1, phase[1] & 1u, lane);
    if (hook_warp && elect_one()) {
      hook.after_half(0);
      hook.after_half(1);
    }
    if (CL && ST == 0) cluster_sync_all();   // both CTAs' stage-1 MMAs have read their operands
    TFFT_TRACE_MARK(10 + 2 * ST);
    epilogue_load<RHO, 0, NG>(c, ra, rb);
    epilogue_range<ST, RHO, LAST, 0, kItemsPerGroup, NG, CL, TWT>(P, c, dst_thr, aux_thr, col_thr, seed, ra, rb, rc, rd);
    phase[0]++;
    if (EARLY2) phase[1]++;
  }
}

// Store phase: 16-byte shared loads of 8 staging chunks, 8x8 in-register transpose, 16-byte global stores.
// pump(): called by every thread, warp-converged, before each plane of each item (the ring kernel advances its loads and
// stage-1 UMMAs of the next unit from there).
struct NoPump {
  __device__ __forceinline__ void operator()() const {}
};
template <int LOG2E, int NT = kThreads, class Pump = NoPump>
__device__ __forceinline__ void store_phase(const UnitPlan& P, const KernelCtx& c, __half* gre, __half* gim, int tid,
                                            uint32_t st_s_lo, uint32_t st_g_lo, uint32_t st_u_lo, uint32_t u_limit,
                                            const Pump& pump = Pump()) {
  constexpr uint32_t kStoreBlocks = (1u << LOG2E) / 64;            // 8x8 blocks per plane
  constexpr uint32_t kStoreItems = kStoreBlocks >= NT ? kStoreBlocks / NT : 1;
  constexpr int TB = NT == 512 ? 9 : 8;                            // item q = tid + NT * i
#pragma unroll
  for (uint32_t i = 0; i < kStoreItems; ++i) {
    const uint32_t so = st_s_lo + bit_sum(i, P.store_sofs, TB, kMaxItemBits - TB);
    const uint32_t g = st_g_lo + bit_sum(i, P.store_gofs, TB, kMaxItemBits - TB);
    if (kStoreBlocks < NT && tid >= static_cast<int>(kStoreBlocks)) continue;
    if (st_u_lo + bit_sum(i, P.store_uval, TB, kMaxItemBits - TB) >= u_limit) continue;
    if (P.il_out) {
      // interleaved output (TFFT_INTERLEAVED): element k of the result is the half2 (re, im) at gre + 2*k
      uint4 ar[8], ai[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {   // il_swap (inverse transform): the planes hold (im, re)
        const uint32_t o = so + bit_sum(static_cast<uint32_t>(x), P.store_xs, 0, 3);
        ar[x] = lds128((P.il_swap ? c.s_im : c.s_re) + o);
        ai[x] = lds128((P.il_swap ? c.s_re : c.s_im) + o);
      }
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const uint32_t sel = (cc & 1) ? 0x7632u : 0x5410u;
        uint32_t wr[4], wi[4];
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          wr[i2] = __byte_perm(reinterpret_cast<const uint32_t*>(&ar[2 * i2])[cc >> 1],
                               reinterpret_cast<const uint32_t*>(&ar[2 * i2 + 1])[cc >> 1], sel);
          wi[i2] = __byte_perm(reinterpret_cast<const uint32_t*>(&ai[2 * i2])[cc >> 1],
                               reinterpret_cast<const uint32_t*>(&ai[2 * i2 + 1])[cc >> 1], sel);
        }
        __half* gp = gre + 2 * (static_cast<size_t>(g) + bit_sum(static_cast<uint32_t>(cc), P.store_cg, 0, 3));
        stg128(gp, make_uint4(__byte_perm(wr[0], wi[0], 0x5410u), __byte_perm(wr[0], wi[0], 0x7632u),
                              __byte_perm(wr[1], wi[1], 0x5410u), __byte_perm(wr[1], wi[1], 0x7632u)));
        stg128(gp + 8, make_uint4(__byte_perm(wr[2], wi[2], 0x5410u), __byte_perm(wr[2], wi[2], 0x7632u),
                                  __byte_perm(wr[3], wi[3], 0x5410u), __byte_perm(wr[3], wi[3], 0x7632u)));
      }
      continue;
    }
#pragma unroll
    for (int plane = 0; plane < 2; ++plane) {
      pump();
      const uint32_t sp = plane ? c.s_im : c.s_re;
      __half* gp = plane ? gim : gre;
      uint4 a[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) a[x] = lds128(sp + so + bit_sum(static_cast<uint32_t>(x), P.store_xs, 0, 3));
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
        stg128(gp + g + bit_sum(static_cast<uint32_t>(cc), P.store_cg, 0, 3), make_uint4(w[0], w[1], w[2], w[3]));
      }
    }
  }
}

// NT threads: 256 (two warp groups), or 512 (four warp groups) for the 32K-element units that run one CTA per SM
// LM: how the stage-1 operand is loaded: 0 = 16-byte cp.async (any mode), 1 = one SWIZZLE_128B TMA tile per plane (row
// mode), 2 = dense TMA tiles of 8 columns (column mode)
// CL = 1: cluster units -- launched with a cluster dimension of 2; CTA pair c = blockIdx.x / 2 works on units c, c + gridDim.x/2, ...
template <int LOG2E, int RHO0, int RHO1, int RHO2, int LM, int NT = kThreads, int CL = 0>
__global__ void __launch_bounds__(NT, (LOG2E == 15 ? 1 : 2))
fft_unit_kernel(const __grid_constant__ UnitPlan P, const __half* __restrict__ in_re,
                const __half* __restrict__ in_im, __half* __restrict__ out_re, __half* __restrict__ out_im,
                const uint4* __restrict__ tables, long long* __restrict__ trace,
                const __grid_constant__ CUtensorMap tmap_re, const __grid_constant__ CUtensorMap tmap_im) {
  using namespace ptx;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kStages = RHO2 ? 3 : 2;
  constexpr int NG = NT / 128, TB = NT == 512 ? 9 : 8;
  const SmemLayout SL = smem_layout(P);
  const TableLayout TL = table_layout(P);
  KernelCtx c;
  c.sbase = smem_u32(smem);
  c.s_re = c.sbase;
  c.s_im = c.sbase + SL.plane_stride;
  c.tw_table = reinterpret_cast<const float2*>(smem + SL.table_off);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SL.bar_off);
  uint64_t* load_bar = reinterpret_cast<uint64_t*>(smem + SL.load_bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SL.slot_off);
  const uint32_t table_base = c.sbase + SL.table_off;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  c.lane_row = static_cast<uint32_t>((warp & 3) * 32 + lane);
  c.lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  c.wgroup = static_cast<uint32_t>(warp >> 2);
  c.a_re = c.s_re;
  c.a_im = c.s_im;
  c.bar_id = 0;
  c.sync_threads = NT;
  c.mma_warp = 0;
  c.ytw = reinterpret_cast<const float2*>(smem + SL.ytw_off);
  c.cl_rank = CL ? cluster_ctarank() : 0u;
  c.cl_delta[0] = CL ? cluster_map(c.sbase, 0) - c.sbase : 0u;
  c.cl_delta[1] = CL ? cluster_map(c.sbase, 1) - c.sbase : 0u;
  const uint32_t first_unit = CL ? blockIdx.x >> 1 : blockIdx.x, unit_step = CL ? gridDim.x >> 1 : gridDim.x;
  uint32_t trace_unit = 0;
  (void)trace_unit;

  // ------------------------------------------------------------------ setup (once per CTA)
  pdl_launch_dependents();
  if (warp == 0) {
    tmem_alloc(tmem_slot, P.tmem_cols);
    tmem_relinquish();
  }
  // row-mode TMA tiles: the tile of the CTA's first unit is requested before the constant tables are staged, so that its
  // latency overlaps the table copy (matters for the latency of a single small transform, BASELINE config C1)
  constexpr bool kEarlyFirstTile = LM == 1 || LM == 3;
  // row-mode tile of a unit (one thread): one tensor tile per plane
  auto request_row_tile = [&](uint32_t unit) {
    const uint32_t b3 = unit >> P.b3_shift, unit_lo = unit & ((1u << P.b3_shift) - 1u);
    const uint32_t ub = unit_lo >> P.upb_shift, uu = unit_lo & ((1u << P.upb_shift) - 1u);
    mbar_arrive_expect_tx(load_bar, 4u << LOG2E);
    if (P.kron_bits) {   // 5-D map {64, R, M/64, y_lo, u + step * image}
      tma_load_5d(c.s_re, &tmap_re, uu, ub * P.tma_batch_step, load_bar);
      tma_load_5d(c.s_im, &tmap_im, uu, ub * P.tma_batch_step, load_bar);
    } else if (P.tma_row5) {   // last pass of a three-pass plan: 5-D map {16 | 64, R, M/16 | M/64, column block, transform}
      const uint32_t c4 = (ub << P.log2_units) + b3 * P.tma_b3_step;
      tma_load_5d(c.s_re, &tmap_re, uu, c4, load_bar);
      tma_load_5d(c.s_im, &tmap_im, uu, c4, load_bar);
    } else {
      const uint32_t c3 = ub * P.tma_batch_step + (uu << P.log2_units);
      if (P.tma_seg) {   // segmented input (multi-GPU staging planes): 5-D map {64, kappa_lo, segment, M/64, transform}
        tma_load_5d(c.s_re, &tmap_re, 0, c3, load_bar);
        tma_load_5d(c.s_im, &tmap_im, 0, c3, load_bar);
      } else {
        tma_load_4d(c.s_re, &tmap_re, c.cl_rank * P.cl_load_c2, c3, load_bar);
        tma_load_4d(c.s_im, &tmap_im, c.cl_rank * P.cl_load_c2, c3, load_bar);
      }
    }
  };
  if (tid == 32) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_init(load_bar, 1);
    fence_mbar_init();
    if (kEarlyFirstTile && first_unit < P.n_units) {
      pdl_wait();
      request_row_tile(first_unit);
    }
  }
  for (uint32_t o = tid * 16; o < TL.total; o += NT * 16) sts128(table_base + o, ldg128(tables + (o >> 4)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (CL) cluster_sync_all();   // the partner CTA is resident and its barriers are initialised before anything is stored into it
  pdl_wait();   // the predecessor in the stream may have produced this kernel's input (or still read its output)
  c.taddr = *tmem_slot;
  uint32_t phase[2] = {0, 0}, load_phase = 0;

  // per-thread constant parts of the bit-linear load / store maps (item q = tid + NT*i) and of the epilogue row maps
  constexpr uint32_t kLoadItems = (1u << LOG2E) / 8 / NT;    // 16-byte chunks per thread and plane
  const uint32_t ld_g_lo = bit_sum(tid, P.load_gofs, 0, TB), ld_s_lo = bit_sum(tid, P.load_sofs, 0, TB);
  const uint32_t ld_u_lo = bit_sum(tid, P.load_uval, 0, TB);
  const uint32_t st_s_lo = bit_sum(tid, P.store_sofs, 0, TB), st_g_lo = bit_sum(tid, P.store_gofs, 0, TB);
  const uint32_t st_u_lo = bit_sum(tid, P.store_uval, 0, TB);
  const uint32_t tmap0 = thread_map<0, RHO0, NG>(P, c), tmap1 = thread_map<1, RHO1, NG>(P, c);
  const uint32_t tmap2 = kStages == 3 ? thread_map<2, (RHO2 ? RHO2 : 4), NG>(P, c) : 0u;
  const uint32_t col_thr = kStages == 3 ? thread_col<2, (RHO2 ? RHO2 : 4), NG>(P, c) : thread_col<1, RHO1, NG>(P, c);
  constexpr bool kTwt = kStages == 2;   // 2-stage plans take their twiddle seeds from the shared-memory table
  const TwSeed seed0 = kTwt ? TwSeed() : thread_seed(P, 0, tmap0);
  const TwSeed seed1 = kStages == 3 ? thread_seed(P, 1, tmap1) : TwSeed();

  for (uint32_t unit = first_unit; unit < P.n_units; unit += unit_step) {
    // outer batch level (batched three-pass plans; b3_shift = 31 otherwise: b3 = 0), then (batch, unit in batch)
    const uint32_t b3 = unit >> P.b3_shift, unit_lo = unit & ((1u << P.b3_shift) - 1u);
    const uint32_t ub = unit_lo >> P.upb_shift, uu = unit_lo & ((1u << P.upb_shift) - 1u);
    const int64_t in_base = static_cast<int64_t>(ub) * P.in_batch_stride + static_cast<int64_t>(uu) * P.in_unit_stride +
                            static_cast<int64_t>(b3) * P.in_b3_stride + (CL ? static_cast<int64_t>(c.cl_rank) * P.cl_load_gofs : 0);
    const int64_t out_base = static_cast<int64_t>(ub) * P.out_batch_stride + static_cast<int64_t>(uu) * P.out_unit_stride +
                             static_cast<int64_t>(b3) * P.out_b3_stride + (CL ? static_cast<int64_t>(c.cl_rank) * P.cl_out_gofs : 0);
    // row/row passes with a ragged batch: transforms past the end are loaded as zeros, never stored
    const uint32_t u_limit =
        P.n_transforms ? P.n_transforms - min(P.n_transforms, unit << P.log2_units) : 0xFFFFFFFFu;
    c.col_base = (uu >> P.col_shift) * P.col_base_stride + P.col_first;
    if (P.kron_bits && tid < 8) {   // row twiddles of this unit (read by the last epilogue, two barriers from here)
      const Cplx w = twiddle((static_cast<uint32_t>(tid) * c.col_base) & ((1u << P.epi[kStages - 1].tw_log2n) - 1u),
                             P.epi[kStages - 1].tw_log2n);
      reinterpret_cast<float2*>(smem + SL.ytw_off)[tid] = make_float2(w.re, w.im);
    }

    // ---------------------------------------------------------------- load phase
    TFFT_TRACE_MARK(0);
    if constexpr (LM == 2 || LM == 4 || LM == 5 || LM == 6) {
      // per group of W = 8 (16, 32, 64) columns one tile {W columns, R kappa, M rows, 1 batch} per plane, dense: [group][m][kappa][W]
      if (tid == 0) {
        constexpr uint32_t W = LM == 5 ? 64 : LM == 6 ? 32 : LM == 4 ? 16 : 8;
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(load_bar, 4u << LOG2E);
        const uint32_t group_bytes = ((2 * W) << P.log2_len) >> CL;   // a cluster CTA loads half of the rows
        for (uint32_t ug = 0; ug < (1u << P.log2_units) / W; ++ug) {
          if (P.tma_b3_step == 2) {   // 5-D map: (batch of the pass, outer batch)
            tma_load_5d_col(c.s_re + ug * group_bytes, &tmap_re, (uu << P.log2_units) + W * ug, c.cl_rank * P.cl_load_c2, ub, b3,
                            load_bar);
            tma_load_5d_col(c.s_im + ug * group_bytes, &tmap_im, (uu << P.log2_units) + W * ug, c.cl_rank * P.cl_load_c2, ub, b3,
                            load_bar);
          } else {
            tma_load_4d_col(c.s_re + ug * group_bytes, &tmap_re, (uu << P.log2_units) + W * ug, ub + b3 * P.tma_b3_step,
                            load_bar, c.cl_rank * P.cl_load_c2);
            tma_load_4d_col(c.s_im + ug * group_bytes, &tmap_im, (uu << P.log2_units) + W * ug, ub + b3 * P.tma_b3_step,
                            load_bar, c.cl_rank * P.cl_load_c2);
          }
        }
      }
      TFFT_TRACE_MARK(1);
      mbar_wait(load_bar, load_phase & 1u);
      load_phase++;
    } else if constexpr (LM == 1 || LM == 3) {
      // one tensor tile per plane: {64 rows, R kappa, M/64, U transforms}; transforms past the end of
      // the batch are out of bounds of the tensor map and arrive as zeros
      if (tid == 0 && !(kEarlyFirstTile && unit == first_unit)) {   // the first tile was requested during the setup
        fence_proxy_async_smem();   // earlier generic-proxy reads of the planes precede the async-proxy writes
        request_row_tile(unit);
      }
      TFFT_TRACE_MARK(1);
      mbar_wait(load_bar, load_phase & 1u);
      load_phase++;
    } else {
      const __half* gre = in_re + in_base;
      const __half* gim = in_im + in_base;
      if (P.il_in) {
        // interleaved input (TFFT_INTERLEAVED): 8 consecutive elements are 32 bytes of (re, im) pairs at in_re + 2*k;
        // split into one chunk of real and one of imaginary parts in registers
        const __half* gil = in_re + 2 * in_base;
        const uint32_t sel_re = P.il_swap ? 0x7632u : 0x5410u, sel_im = P.il_swap ? 0x5410u : 0x7632u;
#pragma unroll
        for (uint32_t i = 0; i < kLoadItems; ++i) {
          const uint32_t g = ld_g_lo + bit_sum(i, P.load_gofs, TB, kMaxItemBits - TB);
          const uint32_t so = ld_s_lo + bit_sum(i, P.load_sofs, TB, kMaxItemBits - TB);
          const bool live = ld_u_lo + bit_sum(i, P.load_uval, TB, kMaxItemBits - TB) < u_limit;
          uint4 a = make_uint4(0, 0, 0, 0), b = a;
          if (live) {
            a = ldg128(gil + 2 * static_cast<size_t>(g));
            b = ldg128(gil + 2 * static_cast<size_t>(g) + 8);
          }
          sts128(c.s_re + so, make_uint4(__byte_perm(a.x, a.y, sel_re), __byte_perm(a.z, a.w, sel_re),
                                         __byte_perm(b.x, b.y, sel_re), __byte_perm(b.z, b.w, sel_re)));
          sts128(c.s_im + so, make_uint4(__byte_perm(a.x, a.y, sel_im), __byte_perm(a.z, a.w, sel_im),
                                         __byte_perm(b.x, b.y, sel_im), __byte_perm(b.z, b.w, sel_im)));
        }
      } else if (u_limit == 0xFFFFFFFFu) {
#pragma unroll
        for (uint32_t i = 0; i < kLoadItems; ++i) {
          const uint32_t g = ld_g_lo + bit_sum(i, P.load_gofs, TB, kMaxItemBits - TB);
          const uint32_t so = ld_s_lo + bit_sum(i, P.load_sofs, TB, kMaxItemBits - TB);
          cp_async16(c.s_re + so, gre + g, 16u);
          cp_async16(c.s_im + so, gim + g, 16u);
        }
      } else {
#pragma unroll
        for (uint32_t i = 0; i < kLoadItems; ++i) {
          const uint32_t g = ld_g_lo + bit_sum(i, P.load_gofs, TB, kMaxItemBits - TB);
          const uint32_t so = ld_s_lo + bit_sum(i, P.load_sofs, TB, kMaxItemBits - TB);
          const bool live = ld_u_lo + bit_sum(i, P.load_uval, TB, kMaxItemBits - TB) < u_limit;
          cp_async16(c.s_re + so, gre + (live ? g : 0u), live ? 16u : 0u);
          cp_async16(c.s_im + so, gim + (live ? g : 0u), live ? 16u : 0u);
        }
      }
      TFFT_TRACE_MARK(1);
      cp_async_wait_all();
    }
    TFFT_TRACE_MARK(2);
    // pull the next unit's input into L2 while this one is transformed and stored
    const bool pf = !CL && P.prefetch_next && P.b3_shift == 31 && unit + gridDim.x < P.n_units;
    if (pf) {
      const uint32_t un = unit + gridDim.x, nb = un >> P.upb_shift, nu = un & ((1u << P.upb_shift) - 1u);
      if constexpr (LM == 2 || LM == 4 || LM == 5 || LM == 6) {
        constexpr uint32_t W = LM == 5 ? 64 : LM == 6 ? 32 : LM == 4 ? 16 : 8;
        if (tid == 0)
          for (uint32_t ug = 0; ug < (1u << P.log2_units) / W; ++ug) {
            tma_prefetch_4d_col(&tmap_re, (nu << P.log2_units) + W * ug, nb);
            tma_prefetch_4d_col(&tmap_im, (nu << P.log2_units) + W * ug, nb);
          }
      } else if constexpr (LM == 1 || LM == 3) {
        if (tid == 0) {
          if (P.kron_bits) {
            tma_prefetch_5d(&tmap_re, nu, nb * P.tma_batch_step);
            tma_prefetch_5d(&tmap_im, nu, nb * P.tma_batch_step);
          } else {
            tma_prefetch_4d(&tmap_re, 0, 0, nb * P.tma_batch_step + (nu << P.log2_units));
            tma_prefetch_4d(&tmap_im, 0, 0, nb * P.tma_batch_step + (nu << P.log2_units));
          }
        }
      } else {
        const int64_t nbase = static_cast<int64_t>(nb) * P.in_batch_stride + static_cast<int64_t>(nu) * P.in_unit_stride;
        // row mode: 8 consecutive items share a 128-byte line; column mode: every 16-byte piece is its own line
        if (P.in_mode == kColMode || (tid & 7) == 0) {
#pragma unroll
          for (uint32_t i = 0; i < kLoadItems; ++i) {
            const uint32_t g = ld_g_lo + bit_sum(i, P.load_gofs, TB, kMaxItemBits - TB);
            prefetch_l2(in_re + nbase + g);
            prefetch_l2(in_im + nbase + g);
          }
        }
      }
    }

    // ---------------------------------------------------------------- tensor-core stages
    run_stage<0, RHO0, false, LOG2E, LM, false, NoHook, 0, NG, false, CL != 0, kTwt>(
        P, c, table_base + TL.b_off[0], bar, phase, warp, lane, trace, trace_unit, tmap0, 0u, seed0);
    TFFT_TRACE_MARK(3);
    // 3-stage plans built with pipe_stage2: the epilogue of stage 2's first tile half overlaps the UMMAs of its second
    if (kStages == 3 && P.pipe_stage2)
      run_stage<1, RHO1, kStages == 2, LOG2E, 0, true, NoHook, 0, NG, false, CL != 0>(
          P, c, table_base + TL.b_off[1], bar, phase, warp, lane, trace, trace_unit, tmap1, col_thr, seed1);
    else
      run_stage<1, RHO1, kStages == 2, LOG2E, 0, false, NoHook, 0, NG, false, CL != 0>(
          P, c, table_base + TL.b_off[1], bar, phase, warp, lane, trace, trace_unit, tmap1, col_thr, seed1);
    TFFT_TRACE_MARK(4);
    if constexpr (kStages == 3)
      run_stage<2, (RHO2 ? RHO2 : 4), true, LOG2E, 0, false, NoHook, 0, NG>(P, c, table_base + TL.b_off[2], bar, phase,
                                                                                warp, lane, trace, trace_unit, tmap2, col_thr);
    TFFT_TRACE_MARK(5);
    __syncthreads();
    TFFT_TRACE_MARK(6);

    // ---------------------------------------------------------------- store phase
    store_phase<LOG2E, NT>(P, c, out_re + (P.il_out ? 2 * out_base : out_base), out_im + out_base, tid, st_s_lo, st_g_lo,
                           st_u_lo, u_limit);
    TFFT_TRACE_MARK(7);
    __syncthreads();   // staging fully read before the next unit's loads overwrite it
    TFFT_TRACE_MARK(8);
    trace_unit++;
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before_sync();
  if (CL) cluster_sync_all();   // no CTA of a pair leaves while its partner could still address its shared memory
  __syncthreads();
  if (warp == 0) tmem_dealloc(c.taddr, P.tmem_cols);
}


// ==========================================================================================
// Landing-ring kernel (32K-element units, TMA input, one 512-thread CTA per SM; UnitShape::ring).
// The single-unit kernel above is a serial chain -- wait for the tile, stage 1, stage 2 (3), store -- and a 32K-element
// unit leaves no room for a second unit in shared memory.  Here the stage-1 operand does not land in the working planes
// but in a separate ring of two 32 KiB slots (a QUARTER of the unit each: re and im part of 16 KiB), and the stage-1
// accumulators of unit q+1 only need tensor memory, which is free as soon as the last epilogue of unit q has drained it.
// So while the 16 warps run the store phase of unit q, one thread (warp 0's elected lane, polling between its store
// items) issues the stage-1 UMMAs of unit q+1 part by part and re-requests each ring slot the moment its UMMAs have
// completed: the global loads and the stage-1 tensor work of the next unit disappear under the store phase, and the
// ring runs up to two parts ahead of the UMMAs (the first two parts of unit q+2 are in flight during stages 2.. of
// unit q+1).  Shared memory: [plane_re | plane_im | ring 64 KiB | tables | barriers]; the planes only hold the stage >= 2
// operands and the staging (dense for 8-column output, see unit_plan.h), which is what makes the ring fit.
// Part p of a unit = the top two row bits of the stage-1 operand = the UMMA tiles [p*T/4, (p+1)*T/4).
struct SmemRingLayout {
  uint32_t plane_stride, land_off, table_off, bar_off, total;
};
constexpr uint32_t kRingSlotBytes = 32768, kRingPartBytes = 16384;
__host__ __device__ inline SmemRingLayout smem_ring_layout(const UnitPlan& p) {
  SmemRingLayout l;
  l.plane_stride = (p.plane_bytes + 1023u) & ~1023u;
  l.land_off = 2 * l.plane_stride;
  l.table_off = l.land_off + 2 * kRingSlotBytes;
  l.bar_off = l.table_off + table_layout(p).total;
  l.total = l.bar_off + 128;   // [0,16) stage barriers, [16,32) full[2], [32,64) part_done[4], [64,72) tensor-memory slot
  return l;
}

// stage-1 UMMAs of part p (runtime) of a unit, A operand in a ring slot; one thread
template <int RHO, int LM>
__device__ __forceinline__ void ring_issue_part(uint32_t taddr, uint32_t land_re, uint32_t land_im, uint32_t b1_saddr,
                                                uint32_t p, uint64_t* done_bar) {
  using namespace ptx;
  constexpr uint32_t R = 1u << RHO, kSteps = R / 16, kTiles = 32768u / R / 128u, kTilesPerPart = kTiles / 4;
  static_assert(kTilesPerPart >= 1, "a part holds at least one tile");
  constexpr bool SW128 = LM == 1;
  constexpr uint32_t idesc = make_idesc_f16(128, 2 * R, /*a_mn=*/1, /*b_mn=*/0), idesc2 = idesc | (1u << 13);
  constexpr uint64_t kSw128 = uint64_t(2) << 61;
  const uint64_t da_re = SW128 ? (make_smem_desc(land_re, 128 * R, 1024) | kSw128) : make_smem_desc(land_re, kKGroupStride, 16 * R);
  const uint64_t da_im = SW128 ? (make_smem_desc(land_im, 128 * R, 1024) | kSw128) : make_smem_desc(land_im, kKGroupStride, 16 * R);
  constexpr uint32_t kTileStep = 16 * R;                    // 256R bytes per 128-row tile, in 16-byte descriptor units
  constexpr uint32_t kKStep = SW128 ? 2048 / 16 : 16;       // per 16-wide K step
  const uint64_t db1 = make_smem_desc(b1_saddr, kKGroupStride, 16 * R);
  const uint64_t db2 = make_smem_desc(b1_saddr + 2 * R * R, kKGroupStride, 16 * R);   // columns R .. 3R-1: [Fi | -Fr]
#pragma unroll
  for (uint32_t tt = 0; tt < kTilesPerPart; ++tt) {
    const uint32_t d = taddr + (p * kTilesPerPart + tt) * 2 * R;
#pragma unroll
    for (uint32_t j = 0; j < kSteps; ++j)
      umma_f16_ss(d, da_re + (tt * kTileStep + j * kKStep), db1 + j * 16, idesc, j > 0 ? 1u : 0u);
#pragma unroll
    for (uint32_t j = 0; j < kSteps; ++j)
      umma_f16_ss(d, da_im + (tt * kTileStep + j * kKStep), db2 + j * 16, idesc2, 1u);
  }
  umma_commit(done_bar);
}

// LM: 1 = row tiles of 64-row SWIZZLE_128B atoms, 2 = column tiles of 8 columns
template <int RHO0, int RHO1, int RHO2, int LM>
__global__ void __launch_bounds__(512, 1)
fft_unit_kernel_ring(const __grid_constant__ UnitPlan P, __half* __restrict__ out_re, __half* __restrict__ out_im,
                     const uint4* __restrict__ tables, const __grid_constant__ CUtensorMap tmap_re,
                     const __grid_constant__ CUtensorMap tmap_im, long long* __restrict__ trace) {
  using namespace ptx;
#ifdef TFFT_TWO_MATRICES
  static_assert(RHO0 < 0, "the ring kernel uses the 3R-column DFT matrices");
#endif
  constexpr int LOG2E = 15, NT = 512, NG = 4, TB = 9;
  constexpr int kStages = RHO2 ? 3 : 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemRingLayout SL = smem_ring_layout(P);
  const TableLayout TL = table_layout(P);
  KernelCtx c;
  c.sbase = smem_u32(smem);
  c.s_re = c.sbase;
  c.s_im = c.sbase + SL.plane_stride;
  c.a_re = c.s_re;
  c.a_im = c.s_im;
  c.tw_table = reinterpret_cast<const float2*>(smem + SL.table_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bar_off);
  uint64_t* bar = bars;             // UMMA completion of stages 2.. (two barriers: tile halves)
  uint64_t* full = bars + 2;        // full[s]: the part in ring slot s has landed
  uint64_t* part_done = bars + 4;   // part_done[p]: the stage-1 UMMAs of part p of a unit have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t table_base = c.sbase + SL.table_off;
  const uint32_t land = c.sbase + SL.land_off;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  c.lane_row = static_cast<uint32_t>((warp & 3) * 32 + lane);
  c.lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  c.wgroup = static_cast<uint32_t>(warp >> 2);
  c.bar_id = 0;
  c.sync_threads = NT;
  c.mma_warp = 0;
  c.ytw = nullptr;
  c.cl_rank = 0;
  c.cl_delta[0] = c.cl_delta[1] = 0;
  uint32_t trace_unit = 0;
  (void)trace_unit;
  (void)TB;

  const uint32_t first_unit = blockIdx.x, unit_step = gridDim.x;
  const uint32_t my_units = first_unit < P.n_units ? (P.n_units - first_unit + unit_step - 1) / unit_step : 0u;
  const uint32_t total_parts = 4 * my_units;

  // part g of this CTA: unit first_unit + (g >> 2) * unit_step, part g & 3, ring slot g & 1
  auto request = [&](uint32_t g) {
    const uint32_t unit = first_unit + (g >> 2) * unit_step, p = g & 3u, slot = g & 1u;
    const uint32_t ub = unit >> P.upb_shift, uu = unit & ((1u << P.upb_shift) - 1u);
    uint64_t* fb = full + slot;
    mbar_arrive_expect_tx(fb, kRingSlotBytes);
    const uint32_t dre = land + slot * kRingSlotBytes, dim = dre + kRingPartBytes;
    if constexpr (LM == 2) {   // {8 columns, R kappa, M/4 rows, 1 batch}; no L2 hint: the pass runs in place
      tma_load_4d_col(dre, &tmap_re, uu << P.log2_units, ub, fb, p * P.ring_c2_step);
      tma_load_4d_col(dim, &tmap_im, uu << P.log2_units, ub, fb, p * P.ring_c2_step);
    } else {                   // {64 rows, R kappa, a quarter of (M/64, U)}
      const uint32_t c3 = ub * P.tma_batch_step + (uu << P.log2_units) + p * P.ring_c3_step;
      tma_load_4d(dre, &tmap_re, p * P.ring_c2_step, c3, fb);
      tma_load_4d(dim, &tmap_im, p * P.ring_c2_step, c3, fb);
    }
  };

  // ------------------------------------------------------------------ setup (once per CTA)
  pdl_launch_dependents();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 32) {
    for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
    pdl_wait();
    if (total_parts) {   // the first two parts are requested before the constant tables are staged
      request(0);
      request(1);
    }
  }
  for (uint32_t o = tid * 16; o < TL.total; o += NT * 16) sts128(table_base + o, ldg128(tables + (o >> 4)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_wait();
  c.taddr = *tmem_slot;
  uint32_t phase[2] = {0, 0}, phase1[2] = {0, 0};

  const uint32_t st_s_lo = bit_sum(tid, P.store_sofs, 0, TB), st_g_lo = bit_sum(tid, P.store_gofs, 0, TB);
  const uint32_t st_u_lo = bit_sum(tid, P.store_uval, 0, TB);
  const uint32_t tmap0 = thread_map<0, RHO0, NG>(P, c), tmap1 = thread_map<1, RHO1, NG>(P, c);
  const uint32_t tmap2 = kStages == 3 ? thread_map<2, (RHO2 ? RHO2 : 4), NG>(P, c) : 0u;
  const uint32_t col_thr = kStages == 3 ? thread_col<2, (RHO2 ? RHO2 : 4), NG>(P, c) : thread_col<1, RHO1, NG>(P, c);
  constexpr bool kTwt = kStages == 2;
  const TwSeed seed0 = kTwt ? TwSeed() : thread_seed(P, 0, tmap0);
  const TwSeed seed1 = kStages == 3 ? thread_seed(P, 1, tmap1) : TwSeed();
  const uint32_t b_saddr0 = table_base + TL.b_off[0], b_saddr1 = table_base + TL.b_off[1],
                 b_saddr2 = table_base + TL.b_off[2];

  // Ring state, live in warp 0's elected lane only: parts requested / issued / observed complete (indices over all
  // parts of this CTA).  Invariants: dn <= is <= rq <= min(total, dn + 2), so a barrier is never tested more than
  // one phase behind.
  uint32_t rq = total_parts ? 2u : 0u, is = 0, dn = 0;
  auto step = [&](uint32_t issue_limit) -> bool {
    bool progress = false;
    if (dn < is && mbar_test(part_done + (dn & 3u), (dn >> 2) & 1u)) {
      ++dn;
      progress = true;
    }
    if (rq < total_parts && rq < dn + 2) {
      request(rq);
      ++rq;
      progress = true;
    }
    if (is < rq && is < issue_limit && mbar_test(full + (is & 1u), (is >> 1) & 1u)) {
      const uint32_t lre = land + (is & 1u) * kRingSlotBytes;
      ring_issue_part<RHO0, LM>(c.taddr, lre, lre + kRingPartBytes, b_saddr0, is & 3u, part_done + (is & 3u));
      ++is;
      progress = true;
    }
    return progress;
  };

  uint32_t q = 0;
  for (uint32_t unit = first_unit; unit < P.n_units; unit += unit_step, ++q) {
    const uint32_t ub = unit >> P.upb_shift, uu = unit & ((1u << P.upb_shift) - 1u);
    const int64_t out_base = static_cast<int64_t>(ub) * P.out_batch_stride + static_cast<int64_t>(uu) * P.out_unit_stride;
    const uint32_t u_limit =
        P.n_transforms ? P.n_transforms - min(P.n_transforms, unit << P.log2_units) : 0xFFFFFFFFu;
    c.col_base = (uu >> P.col_shift) * P.col_base_stride + P.col_first;

    // ---------------------------------------------------------------- stage 1
    // all four parts of this unit issued and complete (after the first unit that already happened under the previous
    // store phase), and the ring two parts ahead
    TFFT_TRACE_MARK(0);
    if (warp == 0) {
      __syncwarp();
      if (elect_one()) {
        const uint32_t need = 4 * (q + 1);
        while (dn < need) step(need);
        while (step(need)) {}
      }
      __syncwarp();
    }
    TFFT_TRACE_MARK(1);
    TFFT_TRACE_MARK(2);
    run_stage<0, RHO0, false, LOG2E, LM, false, NoHook, 1, NG, false, false, kTwt>(
        P, c, b_saddr0, part_done + 3, phase1, warp, lane, trace, trace_unit, tmap0, 0u, seed0);
    TFFT_TRACE_MARK(3);
    // ---------------------------------------------------------------- stages 2 (3): in place on the working planes
    if (kStages == 3 && P.pipe_stage2)
      run_stage<1, RHO1, kStages == 2, LOG2E, 0, true, NoHook, 0, NG>(P, c, b_saddr1, bar, phase, warp, lane, trace,
                                                                       trace_unit, tmap1, col_thr, seed1);
    else
      run_stage<1, RHO1, kStages == 2, LOG2E, 0, false, NoHook, 0, NG>(P, c, b_saddr1, bar, phase, warp, lane, trace,
                                                                        trace_unit, tmap1, col_thr, seed1);
    TFFT_TRACE_MARK(4);
    if constexpr (kStages == 3)
      run_stage<2, (RHO2 ? RHO2 : 4), true, LOG2E, 0, false, NoHook, 0, NG>(P, c, b_saddr2, bar, phase, warp, lane, trace,
                                                                             trace_unit, tmap2, col_thr);
    TFFT_TRACE_MARK(5);
    tc_fence_before_sync();   // the accumulators are drained: the next unit's stage-1 UMMAs may overwrite them
    __syncthreads();
    tc_fence_after_sync();
    TFFT_TRACE_MARK(6);

    // ---------------------------------------------------------------- store phase, with the next unit's stage 1 under it
    const uint32_t issue_limit = 4 * (q + 2);
    auto pump = [&]() {
      if (warp == 0) {
        __syncwarp();
        if (elect_one()) {
          while (step(issue_limit)) {}
        }
        __syncwarp();
      }
    };
    store_phase<LOG2E, NT>(P, c, out_re + (P.il_out ? 2 * out_base : out_base), out_im + out_base, tid, st_s_lo, st_g_lo,
                           st_u_lo, u_limit, pump);
    TFFT_TRACE_MARK(7);
    __syncthreads();   // staging fully read before the next unit's stage-1 epilogue overwrites it
    TFFT_TRACE_MARK(8);
    trace_unit++;
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(c.taddr, 512);
}

// ==========================================================================================
// Two-slot kernel (16K-element units, TMA input): one CTA per SM runs TWO units at a time.
// 512 threads = 2 slots x 256 threads; each slot is the single-unit program above on its own working
// planes, tensor-memory half and named barrier.  Both slots share the constant tables and ONE landing
// buffer: the TMA tile of a unit lands there, the unit's stage-1 MMAs read it, and as soon as those
// MMAs have completed (tcgen05.commit on `landing_free`) the next unit's tile is requested -- the
// global load of unit q+1 overlaps stages 2..s and the store of unit q instead of serialising with
// them.  Landing uses are strictly sequential: use q belongs to slot q & 1.
// Shared memory: [W0_re | W0_im | W1_re | W1_im | L_re | L_im | tables | barriers]
// By default the slot's warp 0 issues the UMMAs and also runs epilogues.  -DTFFT_DEDICATED_MMA_WARP builds
// the warp-specialised variant instead (a fifth warp group with one UMMA / TMA warp per slot, setmaxnreg
// moving its registers to the epilogue warps, and the stage-1 UMMAs of the next unit issued during the
// store phase).  Measured at C2 on B200: 127-130 us either way -- the slots are not latency-bound on the
// issuing warp -- so the simpler program is the default.
// -DTFFT_EARLY3 issues stage 3's first-half UMMAs from stage 2's mid() hook (Early3Hook below).  Correct (the GPU
// parity suite passes) but measured SLOWER at C2 on B200: 121.9 us against 117.1 us without -- the extra barrier stalls
// the issuing warp in the middle of its epilogue and the early UMMAs compete with the other slot's -- so it is off.
#ifdef TFFT_EARLY3
constexpr bool kNoEarly3 = false;
#else
constexpr bool kNoEarly3 = true;
#endif
#ifdef TFFT_DEDICATED_MMA_WARP
constexpr bool kDedicatedMmaWarp = true;
#else
constexpr bool kDedicatedMmaWarp = false;
#endif
constexpr int kSlotThreads = kThreads + (kDedicatedMmaWarp ? 32 : 0);
// dedicated mode: a fifth warp group (warps 16-19) holds the two UMMA-issuing warps (two warps idle) so
// that setmaxnreg can move its registers to the epilogue warp groups
constexpr int kCta2Threads = kDedicatedMmaWarp ? 640 : 512;
struct Smem2Layout {
  uint32_t plane_stride, land_off, land_stride, table_off, bar_off, ytw_off, total;
};
__host__ __device__ inline Smem2Layout smem2_layout(const UnitPlan& p) {
  Smem2Layout l;
  l.plane_stride = (p.plane_bytes + 1023u) & ~1023u;
  l.land_off = 4 * l.plane_stride;
  l.land_stride = 2u << p.log2_elems;   // dense SWIZZLE_128B plane: 2 bytes per element
  l.table_off = l.land_off + 2 * l.land_stride;
  l.bar_off = l.table_off + table_layout(p).total;
  l.ytw_off = l.bar_off + 128;   // [0,64) barriers, [64,72) tensor-memory slot, [128,256) row twiddles of the two slots
  l.total = l.bar_off + 256;
  return l;
}

template <int RHO0, int RHO1, int RHO2>
__global__ void __launch_bounds__(kCta2Threads, 1)
fft_unit_kernel_2slot(const __grid_constant__ UnitPlan P, __half* __restrict__ out_re, __half* __restrict__ out_im,
                      const uint4* __restrict__ tables, const __grid_constant__ CUtensorMap tmap_re,
                      const __grid_constant__ CUtensorMap tmap_im, long long* __restrict__ trace) {
  using namespace ptx;
  constexpr int LOG2E = 14;
  constexpr int kStages = RHO2 ? 3 : 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem2Layout SL = smem2_layout(P);
  const TableLayout TL = table_layout(P);
  const int tid_cta = threadIdx.x;
  // CTA warps 0-7: epilogue warps of slot 0, 8-15: of slot 1 (tcgen05.ld reaches the tensor-memory lane
  // quarter given by the CTA warp index mod 4, so each slot's epilogue warps start at a multiple of 4);
  // warps 16 / 17: the dedicated MMA warps of slot 0 / 1 (warp 8 of their slot)
  const uint32_t slot = tid_cta < 2 * kThreads ? static_cast<uint32_t>(tid_cta >> 8)
                                               : static_cast<uint32_t>(((tid_cta - 2 * kThreads) >> 5) & 1);
  const int tid = tid_cta < 2 * kThreads ? (tid_cta & (kThreads - 1)) : kThreads + (tid_cta & 31);
  const int warp = tid >> 5, lane = tid & 31;
  uint32_t trace_unit = 0;
  (void)trace_unit;

  KernelCtx c;
  c.sbase = smem_u32(smem);
  c.s_re = c.sbase + (2 * slot) * SL.plane_stride;
  c.s_im = c.s_re + SL.plane_stride;
  c.a_re = c.sbase + SL.land_off;
  c.a_im = c.a_re + SL.land_stride;
  c.tw_table = reinterpret_cast<const float2*>(smem + SL.table_off);
  c.lane_row = static_cast<uint32_t>((warp & 3) * 32 + lane);
  c.lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  c.wgroup = static_cast<uint32_t>(warp >> 2);
  c.bar_id = 1 + slot;
  c.sync_threads = kSlotThreads;
  c.mma_warp = 0;
  c.ytw = reinterpret_cast<const float2*>(smem + SL.ytw_off + 64 * slot);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SL.bar_off);
  uint64_t* mma_bar = bars + 2 * slot;      // per slot: UMMA completion (two barriers: tile halves)
  uint64_t* land_full = bars + 4;           // land_full[2*s + h]: tile half h for slot s has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t table_base = c.sbase + SL.table_off;

  // ------------------------------------------------------------------ setup (once per CTA)
  if (tid_cta < 32) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  uint32_t phase[2] = {0, 0};

  // landing use q (q = 0, 1, 2, ...) carries unit blockIdx.x + q * gridDim.x and belongs to slot q & 1.
  // The landing buffer is a ring of two half tiles (the first / second half of the stage-1 MMA tiles).
  // Half h of use q+1 is requested by the thread that observes the completion of the half-h MMAs of use q
  // (that half of the buffer is free from that moment), so loads are in flight almost all the time.
  auto unit_of = [&](uint32_t q) { return blockIdx.x + q * gridDim.x; };
  const uint32_t half_c2 = P.log2_units ? 0u : ((1u << (P.log2_len - P.log2_radix[0])) >> 7);   // (M/64)/2
  const uint32_t half_c3 = P.log2_units ? ((1u << P.log2_units) >> 1) : 0u;
  auto request = [&](uint32_t q, uint32_t h) {
    if (unit_of(q) >= P.n_units) return;
    uint64_t* full = land_full + 2 * (q & 1u) + h;
    mbar_arrive_expect_tx(full, 2u << LOG2E);
    const uint32_t off = h << LOG2E;   // half a plane: E bytes
    const uint32_t uq = unit_of(q), qb = uq >> P.upb_shift, qu = uq & ((1u << P.upb_shift) - 1u);
    // the same half of the NEXT use is pulled into L2 now
    const uint32_t un = unit_of(q + 1), nb = un >> P.upb_shift, nu = un & ((1u << P.upb_shift) - 1u);
    const bool pf = P.prefetch_next && un < P.n_units;
    if (P.kron_bits) {   // unit = (image, y_lo); tile halves split the U rows
      const uint32_t c4 = qb * P.tma_batch_step + h * half_c3;
      tma_load_5d(c.a_re + off, &tmap_re, qu, c4, full);
      tma_load_5d(c.a_im + off, &tmap_im, qu, c4, full);
      if (pf) {
        tma_prefetch_5d(&tmap_re, nu, nb * P.tma_batch_step + h * half_c3);
        tma_prefetch_5d(&tmap_im, nu, nb * P.tma_batch_step + h * half_c3);
      }
      return;
    }
    const uint32_t c3 = qb * P.tma_batch_step + (qu << P.log2_units) + h * half_c3;
    if (P.tma_seg) {   // segmented input (multi-GPU staging planes): 5-D map {64, kappa_lo, segment, M/64, transform}
      tma_load_5d(c.a_re + off, &tmap_re, h * half_c2, c3, full);
      tma_load_5d(c.a_im + off, &tmap_im, h * half_c2, c3, full);
      return;
    }
    tma_load_4d(c.a_re + off, &tmap_re, h * half_c2, c3, full);
    tma_load_4d(c.a_im + off, &tmap_im, h * half_c2, c3, full);
    if (pf) {
      const uint32_t n3 = nb * P.tma_batch_step + (nu << P.log2_units) + h * half_c3;
      tma_prefetch_4d(&tmap_re, 0, h * half_c2, n3);
      tma_prefetch_4d(&tmap_im, 0, h * half_c2, n3);
    }
  };
  // one thread sets up the barriers, waits for the predecessor kernel and requests the first tile; the constant
  // tables are staged meanwhile
  if (tid_cta == 32) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_init(bars + 2, 1);
    mbar_init(bars + 3, 1);
    for (int i = 0; i < 4; ++i) mbar_init(land_full + i, 1);
    mbar_init(bars + 10, 8);   // half_done of slot 0 / 1: one arrival per epilogue warp
    mbar_init(bars + 11, 8);
    fence_mbar_init();
    pdl_wait();
    request(0, 0);
    request(0, 1);
  }
  for (uint32_t o = tid_cta * 16; o < TL.total; o += kCta2Threads * 16) sts128(table_base + o, ldg128(tables + (o >> 4)));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_wait();
  c.taddr = *tmem_slot + slot * 256;
  struct LandingHook {
    uint64_t* full;
    uint32_t parity, q;
    decltype(request)& req;
    __device__ __forceinline__ void before_half(int h) const { mbar_wait(full + h, parity); }   // tile half landed
    __device__ __forceinline__ void after_half(int h) const { req(q + 1, static_cast<uint32_t>(h)); }
    __device__ __forceinline__ void mid(int, int) const {}
  };
  // Stage 2 of a pipe_stage2 plan: once EVERY warp has finished the epilogue of the first tile half, the operand
  // rows of stage 3's first tile half are complete (both stages keep k_1's top bit as the top row bit) and the
  // accumulators of those tiles are drained, so stage 3's first-half UMMAs are issued right away and run under
  // the second-half epilogue of stage 2.
  uint64_t* half_done = bars + 10 + slot;
  struct Early3Hook {
    const KernelCtx& c;
    uint64_t *half_done, *mma_bar;
    uint32_t parity, b_saddr2;
    long long* trace;
    uint32_t trace_unit;
    __device__ __forceinline__ void before_half(int) const {}
    __device__ __forceinline__ void after_half(int) const {}
    __device__ __forceinline__ void mid(int warp, int lane) const {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(half_done);
      if (warp == 0) {
        if (elect_one()) {
          mbar_wait(half_done, parity);
          tc_fence_after_sync();
          stage_issue<2, (RHO2 ? RHO2 : 4), LOG2E, 0, false, NoHook, 2, 1>(c, b_saddr2, mma_bar, NoHook(), trace, trace_unit);
        }
        __syncwarp();
      }
    }
  };
  const uint32_t b_saddr0 = table_base + TL.b_off[0], b_saddr1 = table_base + TL.b_off[1],
                 b_saddr2 = table_base + TL.b_off[2];
  // stage 1 reads the landing buffer and writes the working planes: MMA/epilogue overlap is always legal;
  // stage 2 works in place: overlap needs a plan built with pipe_stage2
  const bool pipe2 = kStages == 3 && P.pipe_stage2;

  if (kDedicatedMmaWarp && tid_cta >= 2 * kThreads) {
    // ---------------------------------------------------------------- UMMA / TMA warps (fifth warp group)
    // The stage-1 UMMAs of a unit only need its landed tile and a free tensor-memory half, so they are
    // issued as soon as the previous unit's last epilogue has drained the accumulators: the wait for the
    // tile and the stage-1 tensor work overlap the previous unit's store phase.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (tid_cta < 2 * kSlotThreads) {   // the other two warps of the group idle until the teardown
      auto issue0 = [&](uint32_t q) {
        if (unit_of(q) >= P.n_units) return;
        LandingHook hook{land_full + 2 * slot, (q >> 1) & 1u, q, request};
        if (elect_one())
          stage_issue<0, RHO0, LOG2E, 1, true, LandingHook>(c, b_saddr0, mma_bar, hook, trace, trace_unit);
        __syncwarp();
      };
      issue0(slot);
      for (uint32_t q = slot; unit_of(q) < P.n_units; q += 2) {
        LandingHook hook{land_full + 2 * slot, (q >> 1) & 1u, q, request};
        stage_observe<RHO0, LOG2E, true>(mma_bar, phase, hook);
#if !defined(TFFT_DEBUG_SKIP)
        tc_fence_before_sync();
        group_sync(c);   // stage-1 epilogue complete: stage-2 operands in place
        tc_fence_after_sync();
        if (pipe2) {
          if (elect_one()) stage_issue<1, RHO1, LOG2E, 0, true, NoHook>(c, b_saddr1, mma_bar, NoHook(), trace, trace_unit);
          __syncwarp();
          stage_observe<RHO1, LOG2E, true>(mma_bar, phase, NoHook());
        } else {
          if (elect_one()) stage_issue<1, RHO1, LOG2E, 0, false, NoHook>(c, b_saddr1, mma_bar, NoHook(), trace, trace_unit);
          __syncwarp();
          stage_observe<RHO1, LOG2E, false>(mma_bar, phase, NoHook());
        }
        if constexpr (kStages == 3) {
          tc_fence_before_sync();
          group_sync(c);
          tc_fence_after_sync();
          if (elect_one())
            stage_issue<2, (RHO2 ? RHO2 : 4), LOG2E, 0, false, NoHook>(c, b_saddr2, mma_bar, NoHook(), trace, trace_unit);
          __syncwarp();
          stage_observe<(RHO2 ? RHO2 : 4), LOG2E, false>(mma_bar, phase, NoHook());
        }
#endif
        tc_fence_before_sync();
        group_sync(c);   // last epilogue complete: the accumulators are free
        tc_fence_after_sync();
        issue0(q + 2);
        group_sync(c);   // store phase complete
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue / store warps
    if (kDedicatedMmaWarp) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    constexpr int ROLE = kDedicatedMmaWarp ? 1 : 0;
    const uint32_t st_s_lo = bit_sum(tid, P.store_sofs, 0, 8), st_g_lo = bit_sum(tid, P.store_gofs, 0, 8);
    const uint32_t st_u_lo = bit_sum(tid, P.store_uval, 0, 8);
    const uint32_t tmap0 = thread_map<0, RHO0, 2>(P, c), tmap1 = thread_map<1, RHO1, 2>(P, c);
    const uint32_t tmap2 = kStages == 3 ? thread_map<2, (RHO2 ? RHO2 : 4), 2>(P, c) : 0u;
    const uint32_t col_thr = kStages == 3 ? thread_col<2, (RHO2 ? RHO2 : 4), 2>(P, c) : thread_col<1, RHO1, 2>(P, c);
    const TwSeed seed0 = thread_seed(P, 0, tmap0);
    const TwSeed seed1 = kStages == 3 ? thread_seed(P, 1, tmap1) : TwSeed();
    uint32_t units_done = 0;   // parity of this slot's half_done barrier
    for (uint32_t q = slot; unit_of(q) < P.n_units; q += 2) {
      const uint32_t unit = unit_of(q);
      const uint32_t ub = unit >> P.upb_shift, uu = unit & ((1u << P.upb_shift) - 1u);
      const int64_t out_base =
          static_cast<int64_t>(ub) * P.out_batch_stride + static_cast<int64_t>(uu) * P.out_unit_stride;
      const uint32_t u_limit =
          P.n_transforms ? P.n_transforms - min(P.n_transforms, unit << P.log2_units) : 0xFFFFFFFFu;
      c.col_base = (uu >> P.col_shift) * P.col_base_stride + P.col_first;
      if (P.kron_bits && tid < 8) {   // row twiddles of this unit (read by the last epilogue, two barriers from here)
        const Cplx w = twiddle((static_cast<uint32_t>(tid) * c.col_base) & ((1u << P.epi[kStages - 1].tw_log2n) - 1u),
                               P.epi[kStages - 1].tw_log2n);
        reinterpret_cast<float2*>(smem + SL.ytw_off + 64 * slot)[tid] = make_float2(w.re, w.im);
      }
      TFFT_TRACE_MARK(0);
      TFFT_TRACE_MARK(1);
      TFFT_TRACE_MARK(2);
      LandingHook hook{land_full + 2 * slot, (q >> 1) & 1u, q, request};
      run_stage<0, RHO0, false, LOG2E, 1, true, LandingHook, ROLE>(P, c, b_saddr0, mma_bar, phase, warp, lane, trace,
                                                                      trace_unit, tmap0, 0u, seed0, hook);
      TFFT_TRACE_MARK(3);
#if !defined(TFFT_DEBUG_SKIP)
      constexpr bool kEarly3 = kStages == 3 && !kDedicatedMmaWarp && !kNoEarly3;
      if (pipe2 && kEarly3)
        run_stage<1, RHO1, kStages == 2, LOG2E, 0, true, Early3Hook, ROLE>(
            P, c, b_saddr1, mma_bar, phase, warp, lane, trace, trace_unit, tmap1, col_thr, seed1,
            Early3Hook{c, half_done, mma_bar, units_done & 1u, b_saddr2, trace, trace_unit});
      else if (pipe2)
        run_stage<1, RHO1, kStages == 2, LOG2E, 0, true, NoHook, ROLE>(P, c, b_saddr1, mma_bar, phase, warp, lane,
                                                                           trace, trace_unit, tmap1, col_thr, seed1);
      else
        run_stage<1, RHO1, kStages == 2, LOG2E, 0, false, NoHook, ROLE>(P, c, b_saddr1, mma_bar, phase, warp, lane,
                                                                            trace, trace_unit, tmap1, col_thr, seed1);
      TFFT_TRACE_MARK(4);
      if constexpr (kStages == 3) {
        if (pipe2 && kEarly3)
          run_stage<2, (RHO2 ? RHO2 : 4), true, LOG2E, 0, false, NoHook, ROLE, 2, true>(
              P, c, b_saddr2, mma_bar, phase, warp, lane, trace, trace_unit, tmap2, col_thr);
        else
          run_stage<2, (RHO2 ? RHO2 : 4), true, LOG2E, 0, false, NoHook, ROLE>(P, c, b_saddr2, mma_bar, phase, warp,
                                                                                   lane, trace, trace_unit, tmap2, col_thr);
      }
      units_done++;
#endif
      TFFT_TRACE_MARK(5);
      tc_fence_before_sync();
      group_sync(c);
      TFFT_TRACE_MARK(6);
      store_phase<LOG2E>(P, c, out_re + out_base, out_im + out_base, tid, st_s_lo, st_g_lo, st_u_lo, u_limit);
      TFFT_TRACE_MARK(7);
      group_sync(c);   // staging fully read before the next unit's epilogues overwrite the working planes
      TFFT_TRACE_MARK(8);
      trace_unit++;
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before_sync();
  __syncthreads();
  if (tid_cta < 32) tmem_dealloc(*tmem_slot, 512);
}

}  // namespace tfft
