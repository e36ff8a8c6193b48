// TEST / BASELINE INFRASTRUCTURE ONLY -- C-ABI driver around the UNMODIFIED reference.
//
// This file contains no reference code: it #includes the reference's own headers
// from where they lie (/root/reference/src/base, passed with -I by oracle/Makefile)
// and calls the reference's public API exactly as its examples do
// (src/testing/ExampleSingleFFT.cu:41-83, src/testing/ExampleBatchFFT.cu:32-69):
//   CreatePlan -> PlanWorksOnDevice -> (Batch)DataHandler -> CopyDataHostToDevice ->
//   ComputeFFT -> CopyResultsDeviceToHost(.., plan.results_in_results_).
// Output: oracle/_ref/libtfft_ref.so (git-ignored, travels to the GPU box).
// Used by tests (-m gpu) as the "reference's own fp16 tensor-core output" oracle, by
// tools/make_golden.py to produce tests/golden/, and by `bench.py --impl reference`.
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <vector>

#include "ComputeFFT.h"  // reference: src/base/ComputeFFT.h (pulls Plan.h, DataHandler.h, kernels)

namespace {
// The reference prints warnings/errors on std::cout; keep stdout clean for JSON.
struct CoutToCerr {
  std::streambuf* old;
  CoutToCerr() : old(std::cout.rdbuf(std::cerr.rdbuf())) {}
  ~CoutToCerr() { std::cout.rdbuf(old); }
};

std::optional<Plan<int>> make_plan(int n, int mode) {
  if (mode == 1) return CreatePlan<int>(n, Mode_4096, 16, 16, 256);
  return CreatePlan<int>(n);  // reference default: Mode_256, 8/8/256 (Plan.h:77-82)
}
}  // namespace

extern "C" {

// One call of the reference path with host buffers (its own H2D/D2H included).
// Layout of host_in/host_out: per transform [RE(n) | IM(n)] halves, transforms
// back to back (DataHandler.h:45-53,124-153).  use_batch_api: 0 = loop the single
// overload over the batch, 1 = the batch overload (one stream per transform).
int ref_fft(int n, int batch, int mode, int use_batch_api, const uint16_t* host_in, uint16_t* host_out) {
  CoutToCerr guard;
  auto plan_opt = make_plan(n, mode);
  if (!plan_opt) return -1;
  Plan<int> plan = plan_opt.value();
  int dev = 0;
  cudaGetDevice(&dev);
  if (!PlanWorksOnDevice(plan, dev)) return -2;
  const int smem_limit = GetMaxNoOptInSharedMem(dev);
  if (!use_batch_api) {
    DataHandler<int> h(n);
    if (h.PeakAtLastError()) return -3;
    for (int b = 0; b < batch; ++b) {
      __half* in = reinterpret_cast<__half*>(const_cast<uint16_t*>(host_in)) + size_t(2) * n * b;
      __half* out = reinterpret_cast<__half*>(host_out) + size_t(2) * n * b;
      if (h.CopyDataHostToDevice(in)) return -4;
      if (ComputeFFT(plan, h, smem_limit)) return -5;
      if (h.CopyResultsDeviceToHost(out, plan.results_in_results_)) return -6;
    }
    cudaDeviceSynchronize();
  } else {
    DataBatchHandler<int> h(n, batch);
    if (h.PeakAtLastError()) return -3;
    if (h.CopyDataHostToDevice(reinterpret_cast<__half*>(const_cast<uint16_t*>(host_in)))) return -4;
    if (ComputeFFT(plan, h, smem_limit)) return -5;
    if (h.CopyResultsDeviceToHost(reinterpret_cast<__half*>(host_out), plan.results_in_results_)) return -6;
    cudaDeviceSynchronize();
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -7;
}

// Timing of the reference on `batch` transforms per step.
//   kernel_ms[i]: CUDA-event time of ComputeFFT only, data resident (Bench.h:121-142
//                 times the same region with a host clock)
//   e2e_ms[i]   : H2D + ComputeFFT + D2H through the reference's handlers, host buffers
// The batch overload leaks `batch` streams per call (ComputeFFT.h:167-173), so keep
// batch * (steps + warmup) bounded.
int ref_bench(int n, int batch, int mode, int use_batch_api, int steps, int warmup, const uint16_t* host_in,
              uint16_t* host_out, double* kernel_ms, double* e2e_ms) {
  CoutToCerr guard;
  auto plan_opt = make_plan(n, mode);
  if (!plan_opt) return -1;
  Plan<int> plan = plan_opt.value();
  int dev = 0;
  cudaGetDevice(&dev);
  if (!PlanWorksOnDevice(plan, dev)) return -2;
  const int smem_limit = GetMaxNoOptInSharedMem(dev);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  __half* in = reinterpret_cast<__half*>(const_cast<uint16_t*>(host_in));
  __half* out = reinterpret_cast<__half*>(host_out);
  int rc = 0;
  if (use_batch_api) {
    DataBatchHandler<int> h(n, batch);
    if (h.PeakAtLastError()) return -3;
    for (int it = 0; it < warmup + steps && rc == 0; ++it) {
      float ms_e2e = 0, ms_k = 0;
      cudaDeviceSynchronize();
      cudaEventRecord(e0, 0);
      if (h.CopyDataHostToDevice(in)) rc = -4;
      if (ComputeFFT(plan, h, smem_limit)) rc = -5;
      if (h.CopyResultsDeviceToHost(out, plan.results_in_results_)) rc = -6;
      cudaEventRecord(e1, 0);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms_e2e, e0, e1);
      h.CopyDataHostToDevice(in);
      cudaDeviceSynchronize();
      cudaEventRecord(e0, 0);
      if (ComputeFFT(plan, h, smem_limit)) rc = -5;
      cudaEventRecord(e1, 0);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms_k, e0, e1);
      if (it >= warmup) {
        kernel_ms[it - warmup] = ms_k;
        e2e_ms[it - warmup] = ms_e2e;
      }
    }
  } else {
    DataHandler<int> h(n);
    if (h.PeakAtLastError()) return -3;
    for (int it = 0; it < warmup + steps && rc == 0; ++it) {
      float ms_e2e = 0, ms_k = 0;
      cudaDeviceSynchronize();
      cudaEventRecord(e0, 0);
      for (int b = 0; b < batch; ++b) {
        if (h.CopyDataHostToDevice(in + size_t(2) * n * b)) rc = -4;
        if (ComputeFFT(plan, h, smem_limit)) rc = -5;
        if (h.CopyResultsDeviceToHost(out + size_t(2) * n * b, plan.results_in_results_)) rc = -6;
      }
      cudaEventRecord(e1, 0);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms_e2e, e0, e1);
      h.CopyDataHostToDevice(in);
      cudaDeviceSynchronize();
      cudaEventRecord(e0, 0);
      for (int b = 0; b < batch; ++b)
        if (ComputeFFT(plan, h, smem_limit)) rc = -5;
      cudaEventRecord(e1, 0);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms_k, e0, e1);
      if (it >= warmup) {
        kernel_ms[it - warmup] = ms_k;
        e2e_ms[it - warmup] = ms_e2e;
      }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc == 0 && cudaGetLastError() != cudaSuccess) rc = -7;
  return rc;
}

// Plan fields of the reference for (n, mode): r16 steps, r2 steps, results_in_results,
// passes = 1 + #TensorRadix16 launches + #radix-2 steps.
int ref_plan_info(int n, int mode, int* out4) {
  CoutToCerr guard;
  auto p = make_plan(n, mode);
  if (!p) return -1;
  out4[0] = p->amount_of_r16_steps_;
  out4[1] = p->amount_of_r2_steps_;
  out4[2] = p->results_in_results_ ? 1 : 0;
  int r16_launches = p->amount_of_r16_steps_ - (mode == 1 ? 2 : 1);
  if (r16_launches < 0) r16_launches = 0;
  out4[3] = 1 + r16_launches + p->amount_of_r2_steps_;
  return 0;
}

}  // extern "C"
