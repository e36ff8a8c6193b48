// Drop-in shim: the reference's ComputeFFT.h interface (src/base/ComputeFFT.h:54-151 single transform,
// :162-293 batch) on top of the C ABI in include/tfft.h.  Link with libtfft.so.
//   * std::nullopt = success, otherwise an error string (the reference's convention)
//   * the single overload is asynchronous on the legacy default stream; the batch overload
//     synchronises the device before returning, like the reference (ComputeFFT.h:286)
//   * plan.results_in_results_ is always true; the input planes are preserved for N <= 32768 and
//     used as scratch above (the reference always overwrites them)
#pragma once

#include <iostream>
#include <map>
#include <optional>
#include <string>
#include <utility>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "DataHandler.h"
#include "Plan.h"
#include "tfft.h"

template <typename Integer>
Integer ExactPowerOf2(const int exponent) {   // reference: ComputeFFT.h:36-47
  if (exponent < 0) std::cout << "Error! Negative exponent not allowed." << std::endl;
  Integer result = 1;
  for (int i = 0; i < exponent; i++) result *= 2;
  return result;
}

namespace tfft_compat {
inline tfft_plan_t cached_plan(long long n, long long batch, int* rc) {
  static std::map<std::pair<long long, long long>, tfft_plan_t> cache;
  auto it = cache.find({n, batch});
  if (it != cache.end()) { *rc = TFFT_OK; return it->second; }
  tfft_plan_t p = nullptr;
  *rc = TFFT_E_NOT_IN_FILE;
  if (!tuner_file().empty()) *rc = tfft_plan_create_from_file(&p, n, batch, TFFT_DEFAULT, tuner_file().c_str());
  if (*rc == TFFT_E_NOT_IN_FILE) *rc = tfft_plan_create(&p, n, batch, TFFT_DEFAULT);
  if (*rc == TFFT_OK) cache[{n, batch}] = p;
  return p;
}
}  // namespace tfft_compat

template <typename Integer>
std::optional<std::string> ComputeFFT(Plan<Integer>& fft_plan, const DataHandler<Integer>& data,
                                      const int max_no_optin_shared_mem = 32768) {
  (void)max_no_optin_shared_mem;
  int rc = TFFT_OK;
  tfft_plan_t p = tfft_compat::cached_plan(static_cast<long long>(fft_plan.fft_length_), 1, &rc);
  if (rc == TFFT_OK)
    rc = tfft_exec(p, data.dptr_input_RE_, data.dptr_input_IM_, data.dptr_results_RE_, data.dptr_results_IM_,
                   2 * static_cast<long long>(fft_plan.fft_length_), 2 * static_cast<long long>(fft_plan.fft_length_),
                   nullptr);
  if (rc != TFFT_OK) return std::string(tfft_error_string(rc));
  if (cudaPeekAtLastError() != cudaSuccess) return cudaGetErrorString(cudaPeekAtLastError());
  return std::nullopt;
}

template <typename Integer>
std::optional<std::string> ComputeFFT(const Plan<Integer>& fft_plan, const DataBatchHandler<Integer>& data,
                                      const int max_no_optin_shared_mem) {
  (void)max_no_optin_shared_mem;
  int rc = TFFT_OK;
  const long long n = static_cast<long long>(fft_plan.fft_length_);
  tfft_plan_t p = tfft_compat::cached_plan(n, data.amount_of_ffts_, &rc);
  if (rc == TFFT_OK)
    rc = tfft_exec(p, data.dptr_input_RE_[0], data.dptr_input_IM_[0], data.dptr_results_RE_[0],
                   data.dptr_results_IM_[0], 2 * n, 2 * n, nullptr);
  if (rc != TFFT_OK) return std::string(tfft_error_string(rc));
  cudaDeviceSynchronize();
  if (cudaPeekAtLastError() != cudaSuccess) return cudaGetErrorString(cudaPeekAtLastError());
  return std::nullopt;
}
