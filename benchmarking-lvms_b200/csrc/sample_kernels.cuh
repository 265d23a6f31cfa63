// Fused sample() + mode() of the discretized logistic mixture: one read of the packed parameters
// produces the mixture sample and the mode, instead of the reference's ~12 eager kernels
//   rsample_discretized_logistic_mixture  blvm/utils/variational.py:309-349  (mixture indicator -> gather -> logistic
//                                          inverse CDF -> clamp)
//   DiscretizedLogisticMixtureDense.mode   blvm/modules/distributions.py:363-368 (argmax of the logits -> gather loc)
// The reference draws the indicator by Gumbel-max (K uniforms, 2K logs per sample); the same categorical distribution
// softmax(logits) is drawn here by inverse CDF from ONE uniform (K exps, a running sum), so a sample costs one
// Philox4x32-10 call instead of four at K = 10 — the kernel was issue-bound on the generator.
// Random numbers: Philox4x32-10 keyed by (seed), counter = (sample index, dimension block, offset): reproducible for a
// given seed/offset, independent of the launch geometry and of which of the two kernels below runs.  The stream differs
// from torch's generator, so parity with the reference is distributional (tests/test_gpu_parity.py: component
// frequencies, CDF, clamp, mode exactness).
//
// dmol_sample_mode_tile_kernel<K, TP>: register-kernel shapes with 16-byte aligned slabs — the tile's parameter slab comes
// in with one TMA bulk copy like in dmol_tile_kernel (every byte of the rows is fetched from HBM anyway: the logits and
// the gathered components touch almost every 32-byte sector), threads read logits and the two gathered components from
// shared memory.  dmol_sample_mode_kernel<TP>: any K, D, alignment, straight from global memory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "blvm_math.cuh"
#include "dmol_kernels.cuh"
#include "ptx_sm100.cuh"

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

struct SampleRng {
  uint32_t k0, k1, c0, c1, o0;
  __device__ __forceinline__ SampleRng(const SampleArgs& A, int64_t n)
      : k0(static_cast<uint32_t>(A.seed)), k1(static_cast<uint32_t>(A.seed >> 32)), c0(static_cast<uint32_t>(n)),
        c1(static_cast<uint32_t>(static_cast<uint64_t>(n) >> 32)), o0(static_cast<uint32_t>(A.offset)) {}
  __device__ __forceinline__ Philox4 block(uint32_t i) const { return philox4x32_10(c0, c1, i, o0, k0, k1); }
};

// Categorical draw by inverse CDF over softmax(logits) + first-max argmax.  `get(k)` returns logit k as float.  Two
// evaluation orders must not differ between the kernels: max first, then e_k = ex2((l_k - max) log2 e) summed in index order.
template <typename Get>
__device__ __forceinline__ void pick_components(int K, float u, Get get, int& pick, int& best_m) {
  float vm = -INFINITY;
  best_m = 0;
  for (int k = 0; k < K; ++k) {
    const float l = get(k);
    if (l > vm) { vm = l; best_m = k; }
  }
  float total = 0.f;
  for (int k = 0; k < K; ++k) total += fast_ex2((get(k) - vm) * kLog2e);
  const float t = u * total;          // u in [0, 1): t < total, so some prefix sum exceeds it (up to rounding: last component)
  float cum = 0.f;
  pick = K - 1;
  bool found = false;
  for (int k = 0; k < K; ++k) {
    cum += fast_ex2((get(k) - vm) * kLog2e);
    if (!found && cum > t) { pick = k; found = true; }
  }
}

// logistic inverse CDF on the chosen component, u ~ U(1e-8, 1 - 1e-8), clamp to [-1, 1]  (variational.py:282-306)
__device__ __forceinline__ float logistic_sample(float loc, float ls, float log_eps, uint32_t bits) {
  ls = (ls < log_eps) ? log_eps : ls;
  const float u = u01_to(bits, 1e-8f, 1.0f - 1e-8f);
  const float x = loc + __expf(ls) * (__logf(u) - __logf(1.0f - u));
  return fminf(fmaxf(x, -1.0f), 1.0f);
}

// word j of the per-sample random stream: word 0 draws the mixture indicator, word 1 + d the logistic of dimension d
__device__ __forceinline__ uint32_t rng_word(const SampleRng& R, const Philox4& first, int j) {
  if (j < 4) return j == 0 ? first.x : (j == 1 ? first.y : (j == 2 ? first.z : first.w));
  const Philox4 b = R.block(static_cast<uint32_t>(j >> 2));
  const int r = j & 3;
  return r == 0 ? b.x : (r == 1 ? b.y : (r == 2 ? b.z : b.w));
}

template <typename TP>
__global__ void __launch_bounds__(256) dmol_sample_mode_kernel(const SampleArgs A) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (n >= A.N) return;
  const int K = A.K, D = A.D, P = K * (2 * D + 1);
  const TP* p = static_cast<const TP*>(A.raw) + n * P;
  const SampleRng R(A, n);
  const Philox4 first = R.block(0);
  int pick, best_m;
  pick_components(K, static_cast<float>(first.x >> 8) * (1.0f / 16777216.0f), [&](int k) { return param_to_float<TP>(p[k]); }, pick, best_m);
  if (A.mode_index) A.mode_index[n] = best_m;
  for (int d = 0; d < D; ++d) {
    const TP* pd = p + K + d * 2 * K;
    if (A.mode) A.mode[n * D + d] = param_to_float<TP>(pd[best_m]);
    if (A.sample)
      A.sample[n * D + d] = logistic_sample(param_to_float<TP>(pd[pick]), param_to_float<TP>(pd[K + pick]), A.log_eps, rng_word(R, first, 1 + d));
  }
}

// D == 1, compile-time K: one tile of 128 * DmolSpt<K> samples per CTA, slab staged by TMA (see the header comment).
template <int K, typename TP>
__global__ void __launch_bounds__(128) dmol_sample_mode_tile_kernel(const SampleArgs A) {
  constexpr int P = 3 * K, TPB = 128, SPT = DmolSpt<K>::value, TILE = TPB * SPT;
  constexpr size_t kTileBytes = ((size_t(TILE) * P * sizeof(TP) + 15) / 16) * 16;
  extern __shared__ __align__(128) unsigned char smem[];
  TP* tile = reinterpret_cast<TP*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kTileBytes);
  const int tid = threadIdx.x;
  const int64_t s0 = static_cast<int64_t>(blockIdx.x) * TILE;
  const int n = static_cast<int>(min(static_cast<int64_t>(TILE), A.N - s0));
  if (tid == 0) {
    const uint32_t bytes = static_cast<uint32_t>(n) * P * sizeof(TP);   // 16-byte multiple: checked on the host
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(bar, bytes);
    ptx::bulk_g2s(tile, static_cast<const TP*>(A.raw) + s0 * P, bytes, bar, ptx::policy_evict_first());
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    const int i = j * TPB + tid;
    if (i >= n) break;
    const int64_t s = s0 + i;
    const TP* row = tile + i * P;
    float logit[K];
    if constexpr (sizeof(TP) == 4 && K % 2 == 0) {   // 64-bit shared loads: conflict-free for odd 3K/2 (dmol_kernels.cuh)
#pragma unroll
      for (int k = 0; k < K / 2; ++k) {
        const float2 v = reinterpret_cast<const float2*>(row)[k];
        logit[2 * k] = v.x;
        logit[2 * k + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) logit[k] = param_to_float<TP>(row[k]);
    }
    const SampleRng R(A, s);
    const Philox4 first = R.block(0);
    // same arithmetic, same order as pick_components (results must not depend on which kernel runs)
    float vm = -INFINITY;
    int best_m = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (logit[k] > vm) { vm = logit[k]; best_m = k; }
    float e[K], total = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      e[k] = fast_ex2((logit[k] - vm) * kLog2e);
      total += e[k];
    }
    const float t = (static_cast<float>(first.x >> 8) * (1.0f / 16777216.0f)) * total;
    float cum = 0.f;
    int pick = K - 1;
    bool found = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      cum += e[k];
      if (!found && cum > t) { pick = k; found = true; }
    }
    if (A.mode_index) A.mode_index[s] = best_m;
    if (A.mode) A.mode[s] = param_to_float<TP>(row[K + best_m]);
    if (A.sample) A.sample[s] = logistic_sample(param_to_float<TP>(row[K + pick]), param_to_float<TP>(row[2 * K + pick]), A.log_eps, first.y);
  }
}

}  // namespace blvm
