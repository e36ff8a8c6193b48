// Drop-in shim: the reference's Plan.h interface (CPestka/Tensor-FFT src/base/Plan.h) implemented for
// the B200-native library.  Same names, argument meaning and error behaviour (message on std::cout,
// std::nullopt) so that callers of the reference -- its own examples, tests and benchmarks --
// compile unchanged with this directory in place of src/base.  The launch-shape fields are kept
// because reference callers print them (src/testing/benchmarks/Bench.h:224-226); they do not steer
// the sm_100a kernels.
#pragma once

#include <cstdint>
#include <fstream>
#include <iostream>
#include <optional>
#include <sstream>
#include <string>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

enum BaseFFTMode { Mode_256, Mode_4096 };   // reference: Plan.h:14

template <typename Integer>
struct Plan {                               // reference: Plan.h:18-39
  Integer fft_length_;
  int amount_of_r16_steps_;
  int amount_of_r2_steps_;
  BaseFFTMode base_fft_mode_;
  bool results_in_results_;                 // always true here: results land in the results planes
  int base_fft_warps_per_block_;
  int base_fft_blocksize_;
  int base_fft_gridsize_;
  int base_fft_shared_mem_in_bytes_;
  int r16_warps_per_block_;
  int r16_blocksize_;
  int r16_gridsize_;
  int r16_shared_mem_in_bytes_;
  int r2_blocksize_;
};

template <typename Integer>
bool IsPowerOf2(const Integer x) {          // reference: Plan.h:41-47
  return x != 0 && (x & (x - 1)) == 0;
}

template <typename Integer>
int ExactLog2(const Integer x) {            // reference: Plan.h:50-67 (without its 32-bit truncation)
  int l = 0;
  for (long long v = static_cast<long long>(x); v > 1; v >>= 1) ++l;
  return l;
}

template <typename Integer>
std::optional<Plan<Integer>> CreatePlan(const Integer fft_length, const BaseFFTMode mode = Mode_256,
                                        const int base_fft_warps_per_block = 8,
                                        const int r16_warps_per_block = 8, const int r2_blocksize = 256) {
  // validation rules of the reference, Plan.h:85-190
  if (!IsPowerOf2(fft_length)) {
    std::cout << "Error! Input size has to be a power of 2!" << std::endl;
    return std::nullopt;
  }
  const int lg = ExactLog2(fft_length);
  if (lg < 8) {
    std::cout << "Error! Input size has to be larger than 256 i.e. 16^2" << std::endl;
    return std::nullopt;
  }
  if (mode == Mode_4096 && fft_length < 4096) {
    std::cout << "Error! Baselayer fft length cant be longer that fft_length." << std::endl;
    return std::nullopt;
  }
  Plan<Integer> p;
  p.fft_length_ = fft_length;
  p.amount_of_r16_steps_ = lg / 4 - 1;
  p.amount_of_r2_steps_ = lg % 4;
  p.base_fft_mode_ = mode;
  p.results_in_results_ = true;
  const long long warps = static_cast<long long>(fft_length) / 256;
  if (warps < base_fft_warps_per_block) {
    p.base_fft_warps_per_block_ = static_cast<int>(warps);
  } else {
    if (warps % base_fft_warps_per_block != 0) {
      std::cout << "Error! Total amount of warps (fft_length/256) has to be evenly devisable by "
                   "base_fft_warps_per_block." << std::endl;
      return std::nullopt;
    }
    p.base_fft_warps_per_block_ = mode == Mode_4096 ? 16 : base_fft_warps_per_block;
  }
  p.base_fft_blocksize_ = p.base_fft_warps_per_block_ * 32;
  p.base_fft_gridsize_ = static_cast<int>(warps / p.base_fft_warps_per_block_);
  p.base_fft_shared_mem_in_bytes_ = p.base_fft_warps_per_block_ * 1024 * static_cast<int>(sizeof(__half));
  if (warps < r16_warps_per_block) {
    p.r16_warps_per_block_ = static_cast<int>(warps);
  } else {
    if (warps % r16_warps_per_block != 0) {
      std::cout << "Error! Total amount of warps (fft_length/256) has to be evenly devisable by "
                   "amount_of_r16_warps_per_block." << std::endl;
      return std::nullopt;
    }
    p.r16_warps_per_block_ = r16_warps_per_block;
  }
  p.r16_blocksize_ = p.r16_warps_per_block_ * 32;
  p.r16_gridsize_ = static_cast<int>(warps / p.r16_warps_per_block_);
  p.r16_shared_mem_in_bytes_ = p.r16_warps_per_block_ * 768 * static_cast<int>(sizeof(__half));
  if ((static_cast<long long>(fft_length) >> p.amount_of_r2_steps_) % r2_blocksize != 0) {
    std::cout << "Error! smallest_r2_subfft_length i.e. pow(2,(log2_of_fft_lenght / 4)) has to be "
                 "evenly devisable by r2_blocksize." << std::endl;
    return std::nullopt;
  }
  p.r2_blocksize_ = r2_blocksize;
  return p;
}

namespace tfft_compat {
// the file the last successful CreatePlan(N, tuner_file) was given: ComputeFFT builds its library plans from it
// (tfft_plan_create_from_file reads the optional key=value knobs of the B200 kernels from the same lines)
inline std::string& tuner_file() { static std::string s; return s; }
}  // namespace tfft_compat

// Tuner-file overload, reference: Plan.h:197-255; line format `N mode base_warps r16_warps r2_block`
// (written by src/testing/FileWriter.h:250-269).
template <typename Integer>
std::optional<Plan<Integer>> CreatePlan(const Integer fft_length, const std::string tuner_results_file) {
  std::ifstream file(tuner_results_file);
  if (!file.is_open()) {
    std::cout << "Error! Failed to open tuner file." << std::endl;
    return std::nullopt;
  }
  std::string line;
  while (std::getline(file, line)) {
    std::stringstream ss(line);
    double n = 0;
    int mode = 0, bw = 0, rw = 0, r2 = 0;
    if (!(ss >> n >> mode >> bw >> rw >> r2)) continue;
    if (static_cast<Integer>(n) == fft_length) {
      tfft_compat::tuner_file() = tuner_results_file;
      return CreatePlan(fft_length, mode == 256 ? Mode_256 : Mode_4096, bw, rw, r2);
    }
  }
  std::cout << "Error! Tuner file didnt contain requested fft length." << std::endl;
  return std::nullopt;
}

template <typename Integer>
bool PlanWorksOnDevice(const Plan<Integer> my_plan, const int device_id) {   // reference: Plan.h:257-296
  (void)my_plan;
  cudaDeviceProp properties;
  if (cudaGetDeviceProperties(&properties, device_id) != cudaSuccess) {
    std::cout << "Error! No CUDA device." << std::endl;
    return false;
  }
  if (properties.major != 10) {
    std::cout << "Error! Compute capability 10.x (sm_100a) is required." << std::endl;
    return false;
  }
  return true;
}

inline int GetMaxNoOptInSharedMem(const int device_id) {   // reference: Plan.h:298-303
  cudaDeviceProp properties;
  cudaGetDeviceProperties(&properties, device_id);
  return static_cast<int>(properties.sharedMemPerBlock);
}
