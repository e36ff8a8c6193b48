// Table of the fused FFT kernel instantiations.  The kernels live in kernel_group.cu, compiled once per group
// (-DTFFT_GROUP=g) so that the 50-odd template instantiations build in parallel; tfft_api.cu only sees this table.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "unit_plan.h"

namespace tfft {

typedef void (*KernelFn)(const UnitPlan, const __half*, const __half*, __half*, __half*, const uint4*, long long*,
                         const CUtensorMap, const CUtensorMap);
// Two-slot variant (one CTA per SM, two units in flight, shared landing buffer) for 16K-element units
typedef void (*Kernel2Fn)(const UnitPlan, __half*, __half*, const uint4*, const CUtensorMap, const CUtensorMap,
                          long long*);
struct KernelEntry {
  int log2e, r0, r1, r2;
  KernelFn fn, fn_tma;   // fn_tma: stage-1 operand loaded by TMA (row-mode input)
  KernelFn fn_tma_col;   // column-mode input loaded by TMA column tiles
  KernelFn fn_tma_col64; // ... by tiles of 64 / 32 columns (UnitPlan::tma_load 5 / 6; 256- / 512-point columns of a 16K-element unit)
  int threads;
};
struct Kernel2Entry {
  int r0, r1, r2;
  Kernel2Fn fn;
};
// Cluster units (UnitPlan::cluster): 512-thread CTAs launched as pairs; lm = load mode (0 cp.async, 1 row TMA tile, 4 column
// TMA tiles of 16 columns)
struct ClusterEntry {
  int r0, r1, r2, lm;
  KernelFn fn;
};
const ClusterEntry* kernel_cluster_group(int* count);
// Landing-ring kernels (UnitPlan::ring): 32K-element units, 512 threads, two-slot calling convention; lm = load mode (1 row
// tiles of 64-row atoms, 2 column tiles of 8 columns)
struct RingEntry {
  int r0, r1, r2, lm;
  Kernel2Fn fn;
};
const RingEntry* kernel_ring_group(int* count);
constexpr int kKernelGroups = 5;
// entries of group g (kernel_group.cu built with -DTFFT_GROUP=g)
const KernelEntry* kernel_group_0(int* count);
const KernelEntry* kernel_group_1(int* count);
const KernelEntry* kernel_group_2(int* count);
const KernelEntry* kernel_group_3(int* count);
const KernelEntry* kernel_group_4(int* count);
const Kernel2Entry* kernel2_group(int* count);

}  // namespace tfft
