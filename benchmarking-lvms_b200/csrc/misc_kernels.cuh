// Small helpers around the hot path: the integer Quantize transform and an early-exit in-place scale.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "blvm_math.cuh"

namespace blvm {

// torch.bucketize(x, boundaries, right=False) (blvm/data/transforms.py:257): lower bound, i.e. the first index i with
// boundaries[i] >= x; the predicate is written `!(b >= x)` like ATen's so that NaN maps to n_bins.
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ x, int64_t n,
                                                       const float* __restrict__ boundaries, int64_t n_bins,
                                                       int64_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  int64_t lo = 0, hi = n_bins;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (!(__ldg(boundaries + mid) >= v)) lo = mid + 1;
    else hi = mid;
  }
  out[i] = lo;
}

// buf *= *scale, skipped entirely (no memory traffic beyond the scalar) when *scale == 1.
__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ buf, int64_t n, const double* __restrict__ scale) {
  const float s = static_cast<float>(*scale);
  if (s == 1.0f) return;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) buf[i] *= s;
}

// Elementwise gaussian_ll (blvm/utils/log_likelihoods.py:17-39) on separate y / mu / sd arrays: value, and with GRAD
// gout * d/d mu and gout * d/d sd (sd is detached when sd_floor > 0, like the reference's no_grad clamp).
template <bool GRAD>
__global__ void __launch_bounds__(256) gaussian_ll_kernel(const float* __restrict__ y, const float* __restrict__ mu,
                                                          const float* __restrict__ sd, const float* __restrict__ gout,
                                                          int64_t n, DmolConsts C, float* __restrict__ lp,
                                                          float* __restrict__ g_mu, float* __restrict__ g_sd) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) {
    float l, dmu = 0.f, dsd = 0.f;
    gauss_component<GRAD, false>(y[i], mu[i], sd[i], C, l, dmu, dsd);
    if (lp) lp[i] = l;
    if (GRAD) {
      const float g = gout ? gout[i] : 1.f;
      g_mu[i] = g * dmu;
      g_sd[i] = g * dsd;
    }
  }
}

// Rows of buf (B, row_elems) whose row_values[b] is NaN are multiplied by 0 (finite entries become 0, NaN stays NaN: what
// `0 * local derivative` gives in the reference when nansum's backward hands a NaN utterance a zero gradient,
// wavenet.py:145).  One CTA per kRowGateChunk elements of a row; CTAs of finite rows leave after reading one double.
constexpr int kRowGateChunk = 8192;
template <typename T>
__global__ void __launch_bounds__(256) row_gate_kernel(T* __restrict__ buf, int64_t row_elems, unsigned chunks,
                                                       const double* __restrict__ row_values) {
  const unsigned b = blockIdx.x / chunks, c = blockIdx.x - b * chunks;
  const double v = row_values[b];
  if (v == v) return;
  T* row = buf + static_cast<int64_t>(b) * row_elems;
  const int64_t e0 = static_cast<int64_t>(c) * kRowGateChunk;
  const int64_t e1 = e0 + kRowGateChunk < row_elems ? e0 + kRowGateChunk : row_elems;
  for (int64_t i = e0 + threadIdx.x; i < e1; i += 256) row[i] = static_cast<T>(static_cast<float>(row[i]) * 0.0f);
}

constexpr int kMaxScaleBuffers = 36;
struct ScaleMultiArgs {
  void* buf[kMaxScaleBuffers];
  int64_t n[kMaxScaleBuffers];
  int dtype[kMaxScaleBuffers];   // 0 fp32, 1 fp16, 2 bf16
  int count;
  const double* scale;
};
// One launch for all gradient buffers of a fused ELBO step.  The common case is scale == 1 (a plain loss.backward()):
// every CTA reads the scalar and leaves, so the grid is kept small (a few CTAs per SM, 1-D) and each CTA walks all
// buffers grid-stride; when a real factor arrives (AMP GradScaler) the walk is 128-bit vectorised where the buffer is
// 16-byte aligned (read + write: 8 B per fp32 element, HBM-bound).
template <typename T>
__device__ __forceinline__ float scale_load(const T* p);
template <> __device__ __forceinline__ float scale_load<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float scale_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void scale_store(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void scale_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__global__ void __launch_bounds__(256) scale_inplace_multi_kernel(const ScaleMultiArgs A) {
  const float s = static_cast<float>(*A.scale);
  if (s == 1.0f) return;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  for (int b = 0; b < A.count; ++b) {
    const int64_t n = A.n[b];
    if (A.dtype[b] == 0) {
      float* __restrict__ buf = static_cast<float*>(A.buf[b]);
      int64_t done = 0;
      if ((reinterpret_cast<uintptr_t>(buf) & 15u) == 0) {
        float4* __restrict__ b4 = reinterpret_cast<float4*>(buf);
        const int64_t n4 = n >> 2;
        for (int64_t i = i0; i < n4; i += stride) {
          float4 v = b4[i];
          v.x *= s; v.y *= s; v.z *= s; v.w *= s;
          b4[i] = v;
        }
        done = n4 << 2;
      }
      for (int64_t i = done + i0; i < n; i += stride) buf[i] *= s;
    } else if (A.dtype[b] == 1) {
      __half* __restrict__ buf = static_cast<__half*>(A.buf[b]);
      for (int64_t i = i0; i < n; i += stride) scale_store(buf + i, scale_load(buf + i) * s);
    } else {
      __nv_bfloat16* __restrict__ buf = static_cast<__nv_bfloat16*>(A.buf[b]);
      for (int64_t i = i0; i < n; i += stride) scale_store(buf + i, scale_load(buf + i) * s);
    }
  }
}

}  // namespace blvm
