// Device-side harness helpers: the reference builds its test signal and its error
// statistics around text files and host loops (src/testing/TestingDataCreation.h:89-117 fixture kernel,
// src/testing/AccuracyCalculator.h:86-148 max / average / sigma of the deviation).  Here both run on the device on
// planar buffers, so accuracy sweeps up to 2^30 points never round-trip through the host.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tfft {

// x_b[t] = sum_{i < cutoff} w_b[i] * sinf(2*pi*i*t / n): the reference's sine superposition (argument formed in
// fp64, evaluated with the fp32 sine, product in fp32, sum in fp64, rounded once to fp16)
__global__ void sine_fixture_kernel(__half* __restrict__ re, __half* __restrict__ im, int64_t n, int64_t stride,
                                    const float* __restrict__ w_re, const float* __restrict__ w_im, int cutoff) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  if (t >= n) return;
  const float* wr = w_re + b * cutoff;
  const float* wi = w_im + b * cutoff;
  const double step = 2.0 * 3.14159265358979323846 * static_cast<double>(t) / static_cast<double>(n);
  double acc_re = 0.0, acc_im = 0.0;
  for (int i = 0; i < cutoff; ++i) {
    const float s = sinf(static_cast<float>(step * i));
    acc_re += static_cast<double>(wr[i] * s);
    acc_im += static_cast<double>(wi[i] * s);
  }
  re[b * stride + t] = __double2half(acc_re);
  im[b * stride + t] = __double2half(acc_im);
}

// sums over the 2*count deviations d = |a - b| (real and imaginary parts alike):
// acc[0] = sum d, acc[1] = sum d^2, acc[2] = sum b^2, bits[0] = max d (as the bit pattern of a non-negative double)
__global__ void deviation_sums_kernel(const __half* __restrict__ a_re, const __half* __restrict__ a_im,
                                      const double* __restrict__ b_re, const double* __restrict__ b_im, int64_t count,
                                      double* __restrict__ acc, unsigned long long* __restrict__ bits) {
  double s1 = 0.0, s2 = 0.0, sb = 0.0, mx = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double br = b_re[i], bi = b_im[i];
    const double dr = fabs(static_cast<double>(__half2float(a_re[i])) - br);
    const double di = fabs(static_cast<double>(__half2float(a_im[i])) - bi);
    s1 += dr + di;
    s2 += dr * dr + di * di;
    sb += br * br + bi * bi;
    mx = fmax(mx, fmax(dr, di));
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
    s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    sb += __shfl_xor_sync(0xFFFFFFFFu, sb, o);
    mx = fmax(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc + 0, s1);
    atomicAdd(acc + 1, s2);
    atomicAdd(acc + 2, sb);
    atomicMax(bits, static_cast<unsigned long long>(__double_as_longlong(mx)));
  }
}

}  // namespace tfft
