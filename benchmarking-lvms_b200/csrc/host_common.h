// Host-side helpers shared by the translation units of libblvm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/blvm_b200.h"
#include "blvm_math.cuh"

namespace blvm_host {

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
const char* last_error();

// Per-device launch state (function attributes, occupancy, SM count) is cached per CUDA device: one process may drive
// several GPUs (the Python layer accepts tensors on a non-current device).
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// fp32 constants rounded from double exactly like torch rounds a Python scalar that meets an fp32 tensor.
inline blvm::DmolConsts make_consts(int num_bins, float log_epsilon) {
  blvm::DmolConsts C;
  C.h = static_cast<float>(1.0 / (num_bins - 1));
  C.log_two_h = static_cast<float>(log(2.0 / (num_bins - 1)));
  C.log_delta_thresh = static_cast<float>(log(static_cast<double>(blvm::kDeltaThresh)));
  C.lo_thresh = static_cast<float>(2.0 / num_bins - 1.0);
  C.hi_thresh = static_cast<float>(1.0 - 2.0 / num_bins);
  C.log_half_bins = static_cast<float>(log(num_bins / 2.0));
  C.log_eps = log_epsilon;
  const double log2e = 1.4426950408889634;
  C.log_half_bins2 = static_cast<float>(log(num_bins / 2.0) * log2e);
  C.neg_log_ratio2 = static_cast<float>(-log(static_cast<double>(num_bins) / (num_bins - 1)) * log2e);
  C.log_delta_thresh2 = static_cast<float>(log(static_cast<double>(blvm::kDeltaThresh)) * log2e);
  const double nb = num_bins, two_h = 2.0 / (nb - 1.0);
  C.two_h = static_cast<float>(two_h);
  C.fb = static_cast<float>(2.0 / nb);
  C.fd0 = static_cast<float>(two_h - 2.0 / nb);
  C.neg_two_h_sixth = static_cast<float>(-two_h / 6.0);
  C.neg_h2 = static_cast<float>(-1.0 / ((nb - 1.0) * (nb - 1.0)));
  return C;
}

// u = h / s <= h * exp(-log_epsilon): tiny for 16-bit bins with the -7 clamp (blvm_math.cuh: kUTiny)
inline bool u_is_tiny(const blvm::DmolConsts& C) {
  return static_cast<double>(C.h) * exp(-static_cast<double>(C.log_eps)) < static_cast<double>(blvm::kTinyU);
}

}  // namespace blvm_host
