// Drop-in shim: the reference's DataHandler.h interface (src/base/DataHandler.h:22-166).  Same layout:
// one device allocation [in_RE | in_IM | out_RE | out_IM] (single) or all inputs
// [RE_0|IM_0|RE_1|IM_1|...] followed by all results in the same order (batch).
#pragma once

#include <iostream>
#include <optional>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <typename Integer>
class DataHandler {
 public:
  DataHandler(const Integer fft_length) : fft_length_(fft_length) {
    if (cudaMalloc((void**)(&dptr_data_), 4 * sizeof(__half) * fft_length_) != cudaSuccess)
      std::cout << cudaGetErrorString(cudaPeekAtLastError()) << std::endl;
    dptr_input_RE_ = dptr_data_;
    dptr_input_IM_ = dptr_input_RE_ + fft_length_;
    dptr_results_RE_ = dptr_input_IM_ + fft_length_;
    dptr_results_IM_ = dptr_results_RE_ + fft_length_;
  }
  DataHandler(const DataHandler&) = delete;
  DataHandler& operator=(const DataHandler&) = delete;

  std::optional<std::string> PeakAtLastError() {
    if (cudaPeekAtLastError() != cudaSuccess) return cudaGetErrorString(cudaPeekAtLastError());
    return std::nullopt;
  }
  std::optional<std::string> CopyDataHostToDevice(__half* data) {
    if (cudaMemcpy(dptr_input_RE_, data, 2 * fft_length_ * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)
      return cudaGetErrorString(cudaPeekAtLastError());
    return std::nullopt;
  }
  std::optional<std::string> CopyResultsDeviceToHost(__half* data, bool results_in_results) {
    __half* results = results_in_results ? dptr_results_RE_ : dptr_input_RE_;
    if (cudaMemcpy(data, results, 2 * fft_length_ * sizeof(__half), cudaMemcpyDeviceToHost) != cudaSuccess)
      return cudaGetErrorString(cudaPeekAtLastError());
    return std::nullopt;
  }
  ~DataHandler() { cudaFree(dptr_data_); }

  Integer fft_length_;
  __half* dptr_data_ = nullptr;
  __half* dptr_input_RE_;
  __half* dptr_input_IM_;
  __half* dptr_results_RE_;
  __half* dptr_results_IM_;
};

template <typename Integer>
class DataBatchHandler {
 public:
  DataBatchHandler(const Integer fft_length, const int amount_of_ffts)
      : fft_length_(fft_length), amount_of_ffts_(amount_of_ffts) {
    if (cudaMalloc((void**)(&dptr_data_), amount_of_ffts_ * 4 * sizeof(__half) * fft_length_) != cudaSuccess)
      std::cout << cudaGetErrorString(cudaPeekAtLastError()) << std::endl;
    dptr_input_RE_.resize(amount_of_ffts_, nullptr);
    dptr_input_IM_.resize(amount_of_ffts_, nullptr);
    dptr_results_RE_.resize(amount_of_ffts_, nullptr);
    dptr_results_IM_.resize(amount_of_ffts_, nullptr);
    __half* results = dptr_data_ + static_cast<size_t>(2) * amount_of_ffts_ * fft_length_;
    for (int i = 0; i < amount_of_ffts_; i++) {
      dptr_input_RE_[i] = dptr_data_ + static_cast<size_t>(2) * i * fft_length_;
      dptr_input_IM_[i] = dptr_input_RE_[i] + fft_length_;
      dptr_results_RE_[i] = results + static_cast<size_t>(2) * i * fft_length_;
      dptr_results_IM_[i] = dptr_results_RE_[i] + fft_length_;
    }
  }
  DataBatchHandler(const DataBatchHandler&) = delete;
  DataBatchHandler& operator=(const DataBatchHandler&) = delete;

  std::optional<std::string> PeakAtLastError() {
    if (cudaPeekAtLastError() != cudaSuccess) return cudaGetErrorString(cudaPeekAtLastError());
    return std::nullopt;
  }
  std::optional<std::string> CopyDataHostToDevice(__half* data) {
    if (cudaMemcpy(dptr_input_RE_[0], data, amount_of_ffts_ * 2 * fft_length_ * sizeof(__half),
                   cudaMemcpyHostToDevice) != cudaSuccess)
      return cudaGetErrorString(cudaPeekAtLastError());
    cudaDeviceSynchronize();
    return std::nullopt;
  }
  std::optional<std::string> CopyResultsDeviceToHost(__half* data, bool results_in_results) {
    __half* results = results_in_results ? dptr_results_RE_[0] : dptr_input_RE_[0];
    if (cudaMemcpy(data, results, amount_of_ffts_ * 2 * fft_length_ * sizeof(__half), cudaMemcpyDeviceToHost) !=
        cudaSuccess)
      return cudaGetErrorString(cudaPeekAtLastError());
    return std::nullopt;
  }
  ~DataBatchHandler() { cudaFree(dptr_data_); }

  Integer fft_length_;
  int amount_of_ffts_;
  __half* dptr_data_ = nullptr;
  std::vector<__half*> dptr_input_RE_;
  std::vector<__half*> dptr_input_IM_;
  std::vector<__half*> dptr_results_RE_;
  std::vector<__half*> dptr_results_IM_;
};
