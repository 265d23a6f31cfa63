// Fused sample() + mode() of the discretized logistic mixture: one read of the packed parameters
// produces the Gumbel-max mixture sample and the mode, instead of the reference's ~12 eager kernels
//   rsample_discretized_logistic_mixture  blvm/utils/variational.py:309-349  (uniform -> gumbel -> argmax -> gather ->
//                                          logistic inverse CDF -> clamp)
//   DiscretizedLogisticMixtureDense.mode   blvm/modules/distributions.py:363-368 (argmax of the logits -> gather loc)
// Random numbers: Philox4x32-10 keyed by (seed), counter = (sample index, draw block, offset): reproducible for a given
// seed/offset and independent of the launch geometry.  The stream differs from torch's generator, so parity with the
// reference is distributional (tests/test_gpu_parity.py: component frequencies, CDF, clamp, mode exactness).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "blvm_math.cuh"

namespace blvm {

struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}
// uniform in [lo, hi): 24 random bits, like torch's uniform_ (rand * (to - from) + from)
__device__ __forceinline__ float u01_to(uint32_t bits, float lo, float hi) {
  return fmaf(static_cast<float>(bits >> 8) * (1.0f / 16777216.0f), hi - lo, lo);
}

template <typename TP>
__device__ __forceinline__ float param_to_float(TP v);
template <>
__device__ __forceinline__ float param_to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float param_to_float<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float param_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

struct SampleArgs {
  const void* raw;      // (N, K(2D+1))
  int64_t N;
  int K, D;
  float log_eps;
  uint64_t seed, offset;
  float* sample;        // (N, D) nullable
  float* mode;          // (N, D) nullable
  int32_t* mode_index;  // (N) nullable: argmax of the logits (for the backward of mode())
};

template <typename TP>
__global__ void __launch_bounds__(256) dmol_sample_mode_kernel(const SampleArgs A) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (n >= A.N) return;
  const int K = A.K, D = A.D, P = K * (2 * D + 1);
  const TP* p = static_cast<const TP*>(A.raw) + n * P;
  // mixture indicator: Gumbel-max with u ~ U(1e-5, 1 - 1e-5)  (variational.py:338-339); mode: plain argmax (first max)
  int best_g = 0, best_m = 0;
  float vg = -INFINITY, vm = -INFINITY;
  const uint32_t k0 = static_cast<uint32_t>(A.seed), k1 = static_cast<uint32_t>(A.seed >> 32);
  const uint32_t c0 = static_cast<uint32_t>(n), c1 = static_cast<uint32_t>(static_cast<uint64_t>(n) >> 32);
  const uint32_t o0 = static_cast<uint32_t>(A.offset);
  for (int kb = 0; kb < K; kb += 4) {
    const Philox4 rnd = philox4x32_10(c0, c1, static_cast<uint32_t>(kb >> 2), o0, k0, k1);
    const uint32_t bits[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kb + j;
      if (k < K) {
        const float logit = param_to_float<TP>(p[k]);
        const float u = u01_to(bits[j], 1e-5f, 1.0f - 1e-5f);
        const float gumbel = -__logf(-__logf(u));
        const float s = logit + gumbel;
        if (s > vg) { vg = s; best_g = k; }
        if (logit > vm) { vm = logit; best_m = k; }
      }
    }
  }
  if (A.mode_index) A.mode_index[n] = best_m;
  // logistic inverse CDF on the chosen component, u ~ U(1e-8, 1 - 1e-8), clamp to [-1, 1]  (variational.py:282-306)
  const Philox4 rnd = philox4x32_10(c0, c1, 0x40000000u, o0, k0, k1);
  const uint32_t bits[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
  for (int d = 0; d < D; ++d) {
    const TP* pd = p + K + d * 2 * K;
    if (A.mode) A.mode[n * D + d] = param_to_float<TP>(pd[best_m]);
    if (A.sample) {
      const float loc = param_to_float<TP>(pd[best_g]);
      float ls = param_to_float<TP>(pd[K + best_g]);
      ls = (ls < A.log_eps) ? A.log_eps : ls;
      uint32_t b = bits[d & 3];
      if (d >= 4) b = philox4x32_10(c0, c1, 0x40000000u + static_cast<uint32_t>(d >> 2), o0, k0, k1).x;
      const float u = u01_to(b, 1e-8f, 1.0f - 1e-8f);
      const float x = loc + __expf(ls) * (__logf(u) - __logf(1.0f - u));
      A.sample[n * D + d] = fminf(fmaxf(x, -1.0f), 1.0f);
    }
  }
}

}  // namespace blvm
