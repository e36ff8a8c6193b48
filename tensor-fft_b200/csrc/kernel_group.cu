// One group of instantiations of the fused FFT kernel (fft_unit_kernel.cuh); see kernel_table.h.
#include "fft_unit_kernel.cuh"
#include "kernel_table.h"

#ifndef TFFT_GROUP
#error "build with -DTFFT_GROUP=0..6"
#endif

namespace tfft {

#define TFFT_KT(E, A, B, C) {E, A, B, C, fft_unit_kernel<E, A, B, C, 0>, fft_unit_kernel<E, A, B, C, 1>, nullptr, nullptr, kThreads}
// N <= 1024: the row tile uses SWIZZLE_32B atoms (load mode 3)
#define TFFT_KS(E, A, B, C) {E, A, B, C, fft_unit_kernel<E, A, B, C, 0>, fft_unit_kernel<E, A, B, C, 3>, nullptr, nullptr, kThreads}
#define TFFT_KSC(E, A, B, C) /* column passes of these lengths have >= 16 columns per unit: 16-column tiles (mode 4) */ \
  {E, A, B, C, fft_unit_kernel<E, A, B, C, 0>, fft_unit_kernel<E, A, B, C, 3>, fft_unit_kernel<E, A, B, C, 4>, nullptr, kThreads}
// 256-point columns: 64 columns per 16K-element unit -> also the 64-column tiles (mode 5)
#define TFFT_KSC64(E, A, B, C) \
  {E, A, B, C, fft_unit_kernel<E, A, B, C, 0>, fft_unit_kernel<E, A, B, C, 3>, fft_unit_kernel<E, A, B, C, 4>, \
   fft_unit_kernel<E, A, B, C, 5>, kThreads}
#define TFFT_KTC(E, A, B, C) \
  {E, A, B, C, fft_unit_kernel<E, A, B, C, 0>, fft_unit_kernel<E, A, B, C, 1>, fft_unit_kernel<E, A, B, C, 2>, nullptr, kThreads}
// 32K-element units (one CTA per SM): 512 threads = four warp groups
#define TFFT_KW(E, A, B, C) \
  {E, A, B, C, fft_unit_kernel<E, A, B, C, 0, 512>, fft_unit_kernel<E, A, B, C, 1, 512>, fft_unit_kernel<E, A, B, C, 2, 512>, nullptr, 512}

#if TFFT_GROUP == 6
// landing-ring kernels: 4096 x 8 columns (four-step / 2-D column passes), 8 rows x 4096 (four-step row pass), N = 32768
static const RingEntry g_ring[] = {
    {6, 6, 0, 2, fft_unit_kernel_ring<6, 6, 0, 2>},
    {6, 6, 0, 1, fft_unit_kernel_ring<6, 6, 0, 1>},
    {5, 5, 5, 1, fft_unit_kernel_ring<5, 5, 5, 1>},
};
const RingEntry* kernel_ring_group(int* count) {
  *count = static_cast<int>(sizeof(g_ring) / sizeof(g_ring[0]));
  return g_ring;
}
#elif TFFT_GROUP == 5
// cluster units: a CTA pair shares 2^16 elements (N = 65536 in one pass; 16 columns x 4096 for column passes)
static const ClusterEntry g_cluster[] = {
    {4, 6, 6, 1, fft_unit_kernel<15, 4, 6, 6, 1, 512, 1>},
    {4, 6, 6, 0, fft_unit_kernel<15, 4, 6, 6, 0, 512, 1>},
    {6, 6, 0, 4, fft_unit_kernel<15, 6, 6, 0, 4, 512, 1>},
    {6, 6, 0, 0, fft_unit_kernel<15, 6, 6, 0, 0, 512, 1>},
};
const ClusterEntry* kernel_cluster_group(int* count) {
  *count = static_cast<int>(sizeof(g_cluster) / sizeof(g_cluster[0]));
  return g_cluster;
}
#else
static const KernelEntry g_entries[] = {
#if TFFT_GROUP == 0
    TFFT_KS(13, 4, 4, 0), TFFT_KSC64(14, 4, 4, 0),                  // L = 2^8
    TFFT_KS(13, 4, 5, 0),                                           // 2^9
    {14, 4, 5, 0, fft_unit_kernel<14, 4, 5, 0, 0>, fft_unit_kernel<14, 4, 5, 0, 3>, fft_unit_kernel<14, 4, 5, 0, 4>,
     fft_unit_kernel<14, 4, 5, 0, 6>, kThreads},                    //   32 columns per unit: also the 32-column tiles (mode 6)
#elif TFFT_GROUP == 1
    TFFT_KS(13, 5, 5, 0), TFFT_KSC(14, 5, 5, 0),                    // 2^10
    TFFT_KT(13, 5, 6, 0), TFFT_KTC(14, 5, 6, 0),                    // 2^11
#elif TFFT_GROUP == 2
    TFFT_KT(13, 6, 6, 0), TFFT_KT(14, 6, 6, 0), TFFT_KW(15, 6, 6, 0), TFFT_KTC(15, 6, 6, 0),  // 2^12
    {15, 5, 6, 0, fft_unit_kernel<15, 5, 6, 0, 0, 512>, nullptr, fft_unit_kernel<15, 5, 6, 0, 4, 512>, nullptr, 512},   // 16 columns x 2^11
#elif TFFT_GROUP == 3
    TFFT_KT(13, 4, 4, 5), TFFT_KT(14, 4, 4, 5),                     // 2^13
    TFFT_KT(14, 4, 5, 5),                                           // 2^14
    TFFT_KT(14, 5, 5, 4),                                           // 2 rows x 2^13 (2-D row pass, Kronecker last stage)
#else
    TFFT_KW(15, 5, 5, 5), TFFT_KT(15, 5, 5, 5),                     // 2^15
#endif
};

#define TFFT_CAT2(a, b) a##b
#define TFFT_CAT(a, b) TFFT_CAT2(a, b)
const KernelEntry* TFFT_CAT(kernel_group_, TFFT_GROUP)(int* count) {
  *count = static_cast<int>(sizeof(g_entries) / sizeof(g_entries[0]));
  return g_entries;
}

#if TFFT_GROUP == 4
static const Kernel2Entry g_entries2[] = {
    {4, 4, 5, fft_unit_kernel_2slot<4, 4, 5>},
    {4, 5, 5, fft_unit_kernel_2slot<4, 5, 5>},
    {5, 5, 4, fft_unit_kernel_2slot<5, 5, 4>},
};
const Kernel2Entry* kernel2_group(int* count) {
  *count = static_cast<int>(sizeof(g_entries2) / sizeof(g_entries2[0]));
  return g_entries2;
}
#endif
#endif   // TFFT_GROUP < 5

}  // namespace tfft
