// DMoL kernels for sm_100a.
//
// dmol_tile_kernel<K, TPB, GRAD, UMODE>: the hot kernel (UMODE: see blvm_math.cuh, chosen on the host).  One CTA = one tile of up to TPB consecutive waveform samples of ONE
// utterance (row b), one thread per sample.  The tile's K-mixture parameters are the contiguous slab
// raw[b, t0:t0+n, 0:3K] (3K*n floats): it is pulled into shared memory with one 1-D TMA bulk copy
// (cp.async.bulk -> SASS UBLKCP) signalled on an mbarrier, each thread lifts its own 3K-float row into registers with
// 64/128-bit LDS (row stride 3K floats; conflict-free for 3K/vec odd), evaluates value + gradient
// in registers (blvm_math.cuh), writes the gradient row back over its own slab row and the slab leaves with one bulk
// store.  y and the per-sample log-prob are plain coalesced 32-bit accesses.  The masked per-row sum is reduced in
// fp64 (warp shuffles, then shared memory) into partials[b, chunk] — no atomics, deterministic.
//
// Algorithmic traffic per sample: read 4 (y) + 12K (params), write 4 (log-prob) + 12K (grads) = 4(2+6K) bytes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "blvm_math.cuh"
#include "ptx_sm100.cuh"

namespace blvm {

enum : int {
  kFlagMaskOutput = 1,   // per-sample log-prob output is multiplied by the sequence mask (compute_elbo semantics)
  kFlagSkipPadded = 2,   // tiles that lie entirely in the padding are not read: outputs are exact zeros
};

struct DmolArgs {
  const float* y;          // (B*T*D)
  const void* raw;         // (B*T, P)  P = K(2D+1); fp32, or fp16/bf16 in the register kernels (AMP Linear output)
  const int64_t* x_sl;     // (B) valid samples per row, device; nullptr = all valid
  const float* gout;       // (B*T) upstream gradient per sample, nullptr = 1
  float gscale;            // scalar multiplier of every gradient (e.g. -1/sum(x_sl))
  const double* gscale_dev;  // optional device scalar multiplied into gscale (an upstream grad_output), nullptr = 1
  float* lp;               // (B*T) per-sample log-prob out, nullptr = not wanted
  void* graw;              // (B*T, P) gradient out (GRAD kernels), same element type as raw
  double* partials;        // (B, chunks) masked per-tile sums of log-prob, nullptr = not wanted
  int* err_flag;           // set to 1 if any y is outside [-1, 1] (the reference's assert, log_likelihoods.py:195)
  int64_t B, T, chunks;
  int64_t ctas_per_row;    // tile kernel: ceil(chunks / DmolGroup) CTAs per utterance (set by the launcher)
  int K, D;
  int flags;
  int lik;                 // kLikDmol / kLikGmmRaw / kLikGmmSd (generic kernel; the register kernels take it as a template argument)
  DmolConsts C;
};

template <int NWARPS>
__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) s += scratch[i];
  }
  return s;  // valid in thread 0
}

// Lift / drop one P-element row between shared memory and registers (as fp32) with the widest aligned vector the
// row stride allows.  fp32 rows: 128/64/32-bit; fp16/bf16 rows: 32-bit pairs when P is even.
template <typename TP, int P>
struct RowIO;

template <int P>
struct RowIO<float, P> {
  static __device__ __forceinline__ void load(const float* p, float (&r)[P]) {
    if constexpr (P % 4 == 0) {
#pragma unroll
      for (int i = 0; i < P / 4; ++i) {
        const float4 v = reinterpret_cast<const float4*>(p)[i];
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
    } else if constexpr (P % 2 == 0) {
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const float2 v = reinterpret_cast<const float2*>(p)[i];
        r[2 * i] = v.x; r[2 * i + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) r[i] = p[i];
    }
  }
  static __device__ __forceinline__ void store(float* p, const float (&r)[P]) {
    if constexpr (P % 4 == 0) {
#pragma unroll
      for (int i = 0; i < P / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    } else if constexpr (P % 2 == 0) {
#pragma unroll
      for (int i = 0; i < P / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(r[2 * i], r[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) p[i] = r[i];
    }
  }
};

// 16-bit rows: pairs travel as 32-bit words; a row whose byte size is a multiple of 16 / 8 is moved with 128 / 64-bit accesses
// (K = 8: three LDS.128 instead of twelve LDS.32 whose 12-word stride is 4-way bank conflicted; K = 12 / 20: LDS.64, odd strides).
template <typename TP>
struct Pair16;
template <>
struct Pair16<__half> {
  static __device__ __forceinline__ void unpack(uint32_t w, float& a, float& b) {
    asm("{\n.reg .f16 lo, hi;\nmov.b32 {lo, hi}, %2;\ncvt.f32.f16 %0, lo;\ncvt.f32.f16 %1, hi;\n}" : "=f"(a), "=f"(b) : "r"(w));
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    uint32_t w;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));   // first source -> upper half
    return w;
  }
};
template <>
struct Pair16<__nv_bfloat16> {
  static __device__ __forceinline__ void unpack(uint32_t w, float& a, float& b) {   // bf16 -> fp32 is a 16-bit shift: one SHL and one AND per pair
    a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};

template <typename TP, int P>
struct RowIO16 {
  static __device__ __forceinline__ void load(const TP* p, float (&r)[P]) {
    if constexpr (P % 8 == 0) {
#pragma unroll
      for (int i = 0; i < P / 8; ++i) {
        const uint4 v = reinterpret_cast<const uint4*>(p)[i];
        Pair16<TP>::unpack(v.x, r[8 * i], r[8 * i + 1]); Pair16<TP>::unpack(v.y, r[8 * i + 2], r[8 * i + 3]);
        Pair16<TP>::unpack(v.z, r[8 * i + 4], r[8 * i + 5]); Pair16<TP>::unpack(v.w, r[8 * i + 6], r[8 * i + 7]);
      }
    } else if constexpr (P % 4 == 0) {
#pragma unroll
      for (int i = 0; i < P / 4; ++i) {
        const uint2 v = reinterpret_cast<const uint2*>(p)[i];
        Pair16<TP>::unpack(v.x, r[4 * i], r[4 * i + 1]); Pair16<TP>::unpack(v.y, r[4 * i + 2], r[4 * i + 3]);
      }
    } else if constexpr (P % 2 == 0) {
#pragma unroll
      for (int i = 0; i < P / 2; ++i) Pair16<TP>::unpack(reinterpret_cast<const uint32_t*>(p)[i], r[2 * i], r[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) r[i] = static_cast<float>(p[i]);
    }
  }
  static __device__ __forceinline__ void store(TP* p, const float (&r)[P]) {
    if constexpr (P % 8 == 0) {
#pragma unroll
      for (int i = 0; i < P / 8; ++i)
        reinterpret_cast<uint4*>(p)[i] = make_uint4(Pair16<TP>::pack(r[8 * i], r[8 * i + 1]), Pair16<TP>::pack(r[8 * i + 2], r[8 * i + 3]),
                                                    Pair16<TP>::pack(r[8 * i + 4], r[8 * i + 5]), Pair16<TP>::pack(r[8 * i + 6], r[8 * i + 7]));
    } else if constexpr (P % 4 == 0) {
#pragma unroll
      for (int i = 0; i < P / 4; ++i)
        reinterpret_cast<uint2*>(p)[i] = make_uint2(Pair16<TP>::pack(r[4 * i], r[4 * i + 1]), Pair16<TP>::pack(r[4 * i + 2], r[4 * i + 3]));
    } else if constexpr (P % 2 == 0) {
#pragma unroll
      for (int i = 0; i < P / 2; ++i) reinterpret_cast<uint32_t*>(p)[i] = Pair16<TP>::pack(r[2 * i], r[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) p[i] = static_cast<TP>(r[i]);
    }
  }
};
template <int P>
struct RowIO<__half, P> : RowIO16<__half, P> {};
template <int P>
struct RowIO<__nv_bfloat16, P> : RowIO16<__nv_bfloat16, P> {};

// Rows made of an EVEN number NV of 16-byte vectors (fp32 K = 8, 16; 16-bit K = 16) start at bank groups NV * lane mod 8, which
// takes only 8 / gcd(NV, 8) distinct values over the 8 lanes of a 128-bit shared-memory wavefront: a gcd(NV, 8)-way bank conflict on
// every access (K = 16 fp32, 192-byte rows: 4-way; the kernel sat at 0.88 of the HBM roofline on shared-memory bandwidth).  The
// slab arrives by ONE bulk copy, so rows cannot be padded; instead each lane walks the vectors of every K-element section of its
// row (mixture logits | locations | log-scales) in an order rotated by rot = (lane / period) % ways: the lanes that share a
// starting bank group then touch different vectors of the section at every step.  Register slot k holds component
// (k + VE * rot) mod K of each section -- the same permutation in all three sections, and the likelihood is symmetric in the
// component index, so value and gradient are those of the unrotated walk up to the order of the K-term sums (gradients are
// written back through the same permutation).
constexpr int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }
template <typename TP, int K>
struct RowRot {
  static constexpr int P = 3 * K;
  static constexpr int VE = 16 / static_cast<int>(sizeof(TP));          // elements per 16-byte vector
  static constexpr bool vec16 = (P * sizeof(TP)) % 16 == 0 && K % VE == 0;
  static constexpr int NV = vec16 ? P / VE : 1;
  static constexpr int SV = vec16 ? K / VE : 1;                          // vectors per section
  static constexpr int ways = gcd_int(NV, 8);
  static constexpr int period = 8 / ways;
  static constexpr bool enabled = vec16 && NV % 2 == 0 && (SV & (SV - 1)) == 0 && SV >= ways;
  static __device__ __forceinline__ int rot(int lane) { return (lane / period) % ways; }

  static __device__ __forceinline__ void load(const TP* p, float (&r)[P], int rot) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
      for (int j = 0; j < SV; ++j) {
        const uint4 v = reinterpret_cast<const uint4*>(p)[s * SV + ((j + rot) & (SV - 1))];
        float* d = r + s * K + j * VE;
        if constexpr (sizeof(TP) == 4) {
          d[0] = __uint_as_float(v.x); d[1] = __uint_as_float(v.y); d[2] = __uint_as_float(v.z); d[3] = __uint_as_float(v.w);
        } else {
          Pair16<TP>::unpack(v.x, d[0], d[1]); Pair16<TP>::unpack(v.y, d[2], d[3]);
          Pair16<TP>::unpack(v.z, d[4], d[5]); Pair16<TP>::unpack(v.w, d[6], d[7]);
        }
      }
    }
  }
  static __device__ __forceinline__ void store(TP* p, const float (&r)[P], int rot) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
      for (int j = 0; j < SV; ++j) {
        const float* d = r + s * K + j * VE;
        uint4 v;
        if constexpr (sizeof(TP) == 4) {
          v = make_uint4(__float_as_uint(d[0]), __float_as_uint(d[1]), __float_as_uint(d[2]), __float_as_uint(d[3]));
        } else {
          v = make_uint4(Pair16<TP>::pack(d[0], d[1]), Pair16<TP>::pack(d[2], d[3]), Pair16<TP>::pack(d[4], d[5]), Pair16<TP>::pack(d[6], d[7]));
        }
        reinterpret_cast<uint4*>(p)[s * SV + ((j + rot) & (SV - 1))] = v;
      }
    }
  }
};
// Pair16 is only defined for the 16-bit types; fp32 rows never reach its branches
template <>
struct Pair16<float> {
  static __device__ __forceinline__ void unpack(uint32_t, float&, float&) {}
  static __device__ __forceinline__ uint32_t pack(float, float) { return 0u; }
};

// Samples per thread: small K means small rows, so a 128-sample tile would be a 1.5 KB slab (K = 1) and the per-CTA
// fixed cost (mbarrier, barriers, bulk-store drain) dominates; each thread then walks SPT samples of a 128*SPT tile.
#ifndef BLVM_SPT_1
#define BLVM_SPT_1 8    // K == 1
#endif
#ifndef BLVM_SPT_2
#define BLVM_SPT_2 4    // K == 2 (measured with the stream kernel: 47.5 vs 53.7 us fwd+grad, 37.2 vs 45.4 us fwd at T = 16000)
#define BLVM_SPT_5 4    // K <= 5
#define BLVM_SPT_8 2    // K <= 8
#define BLVM_SPT_12 1   // K <= 12
#endif
#ifndef BLVM_MINB_BIGK
#define BLVM_MINB_BIGK 4  // __launch_bounds__ min blocks/SM for K > 12: caps registers at 128 so that 4 CTAs (16 warps, 4 slabs
                          // of 46 KB at K = 30) are resident per SM; measured 543 -> 471 us at K = 30 (5.5 -> 6.3 TB/s)
#endif
template <int K>
struct DmolSpt {
  static constexpr int value = K <= 1 ? BLVM_SPT_1 : K <= 2 ? BLVM_SPT_2 : (K <= 5 ? BLVM_SPT_5 : (K <= 8 ? BLVM_SPT_8 : (K <= 12 ? BLVM_SPT_12 : 1)));
};
#ifndef BLVM_MINB_K20
#define BLVM_MINB_K20 5   // 16 < K <= 20: 96 registers, 5 CTAs/SM (K = 20: 311.7 -> 299.6 us, 6.67 TB/s)
#define BLVM_MINB_K16 6   // 12 < K <= 16: 80 registers, 6 CTAs/SM (K = 16: 324 -> 285 us; its 192-byte rows are 4-way bank conflicted)
#endif
#ifndef BLVM_MINB_MIDK
#define BLVM_MINB_MIDK 8  // 8 < K <= 12, gradient kernels: 64 registers (8 CTAs/SM; the packed-fp32 evaluation would otherwise take
                          // 79 and K = 10 fwd+grad drops from 156 to 162 us).  The forward-only kernels are left uncapped:
                          // K = 10 fwd 98.7 us capped, 86.3 us (6.07 TB/s) uncapped.
#endif
template <int K, bool GRAD = true, typename TP = float>
struct DmolMinBlocks {
  static constexpr int mid = BLVM_MINB_MIDK;
  static constexpr int value = K > 20 ? BLVM_MINB_BIGK : (K > 16 ? BLVM_MINB_K20 : (K > 12 ? BLVM_MINB_K16 : ((K > 8 && GRAD) ? mid : 0)));   // 0 = no constraint
};

// Chunks per CTA of the tile kernel.  The partial-sum layout (one fp64 per chunk of 128 * DmolSpt<K> samples, what
// blvm_dmol_chunks() counts) is fixed by K alone; a CTA may cover G consecutive chunks of an utterance with ONE slab, ONE
// mbarrier / bulk load / bulk store / block reduction: it writes its sum into the first chunk's slot and zeros into the
// others (the finalize kernel only ever adds a row's slots up).  The per-CTA fixed cost (index and bounds arithmetic,
// barriers, the fp64 block sum: ~240 warp instructions, ncu r2a) is what makes the 16-bit kernels issue-bound at one
// sample per thread; their slabs are half as large, so two chunks per CTA keep the shared-memory footprint of the fp32 kernel.
#ifndef BLVM_GROUP_H16
#define BLVM_GROUP_H16 2   // fp16 / bf16 parameters
#endif
#ifndef BLVM_GROUP_F32
#define BLVM_GROUP_F32 1
#endif
template <int K, typename TP>
struct DmolGroup {
  static constexpr int want = sizeof(TP) == 2 ? BLVM_GROUP_H16 : BLVM_GROUP_F32;
  // K <= 5 already walks 4-8 samples per thread (and normally runs the stream kernel); slabs are kept at or below 32 KB
  // (measured r2b, bf16: K = 10 107 -> 98.8 us, K = 16 170 -> 162 us with two chunks per CTA, but K = 30 (46 KB slabs, 4 CTAs
  // per SM either way) 320 -> 357 us: long serialised load / evaluate / store phases with too few CTAs to overlap them)
  static constexpr int value = (K > 5 && size_t(128) * DmolSpt<K>::value * want * 3 * K * sizeof(TP) <= 32 * 1024) ? want : 1;
};

// Linear-domain evaluation (blvm_math.cuh: dmol_sample_lin) pays where the kernel is bound by the MUFU unit and by issue slots: 16-bit
// parameters.  The fp32 kernels are HBM-bound and keep the single log-domain body (64 registers, no second copy of the sample code).
// fp16, 8 < K <= 12: next to the second (log-domain) copy of the sample body the compiler keeps the whole converted row live -- a bf16
// element is rematerialised from its packed word by one shift, an fp16 element is not -- and spills 136 bytes at 64 registers (72 at 72):
// K = 10 93.4 us (no gain over 94.5), K = 12 138 us (117 before).  K = 10 takes the packed-row variant below (88.8 us), K = 12 keeps the
// log-domain body.
#ifndef BLVM_LIN_F32_SMALLK
#define BLVM_LIN_F32_SMALLK 0   // fp32, K <= 5 (stream kernel): measured, no change (43.4 / 59.5 / 75.3 / 94.0 us for K = 2..5 either way): not issue-bound
#endif
template <typename TP, int K>
constexpr bool kLinTP = (sizeof(TP) == 2 && !(std::is_same<TP, __half>::value && K > 8)) || (BLVM_LIN_F32_SMALLK && sizeof(TP) == 4 && K <= 5);

// fp16, K = 10 (the reference's mixture size): the row stays PACKED in registers (3K/2 words) and every pair is converted where the evaluation uses it.
// With the row converted up front the compiler keeps all 3K floats live next to the second (log-domain) body -- a bf16 element is
// rematerialised from its packed word by one shift, an fp16 element is not -- and spills 136 bytes at 64 registers.
template <typename TP, int K>
struct PackedRowIn {
  static_assert(K % 2 == 0, "pairs must not straddle words");
  const uint32_t* w;
  __device__ __forceinline__ F2 pair(int e) const {
    F2 v;
    Pair16<TP>::unpack(w[e >> 1], v.x, v.y);
    return v;
  }
  __device__ __forceinline__ float one(int e) const {
    const F2 v = pair(e & ~1);
    return (e & 1) ? v.y : v.x;
  }
  __device__ __forceinline__ float logit(int k) const { return one(k); }
  __device__ __forceinline__ F2 logit2(int k) const { return pair(k); }
  __device__ __forceinline__ F2 mu2(int k) const { return pair(K + k); }
  __device__ __forceinline__ F2 ls2(int k) const { return pair(2 * K + k); }
  __device__ __forceinline__ float mu(int k) const { return one(K + k); }
  __device__ __forceinline__ float ls(int k) const { return one(2 * K + k); }
};
template <typename TP, int K>
constexpr bool kLinPackedTP = std::is_same<TP, __half>::value && K > 8 && K <= 10 && K % 2 == 0 && BLVM_LINEAR_DOMAIN;   // K = 12: 114.7 -> 123.5 us (192 bytes of spills)

// value + gradient of the sample whose 16-bit row starts at `row` (4-byte aligned: 3K even), gradient row written back in place
template <int K, bool GRAD, int UMODE, typename TP>
__device__ __forceinline__ float dmol_eval_packed_row(float y, TP* row, float g, const DmolConsts& C) {
  constexpr int P = 3 * K, NW = P / 2;
  uint32_t w[NW];
  constexpr int VW = (P * 2) % 16 == 0 ? 4 : ((P * 2) % 8 == 0 ? 2 : 1);   // words per shared-memory access
#pragma unroll
  for (int i = 0; i < NW / VW; ++i) {
    if constexpr (VW == 4) {
      const uint4 v = reinterpret_cast<const uint4*>(row)[i];
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    } else if constexpr (VW == 2) {
      const uint2 v = reinterpret_cast<const uint2*>(row)[i];
      w[2 * i] = v.x; w[2 * i + 1] = v.y;
    } else {
      w[i] = reinterpret_cast<const uint32_t*>(row)[i];
    }
  }
  float r[P];
  float L;
  if (!dmol_sample_lin_in<K, GRAD>(y, PackedRowIn<TP, K>{w}, r, g, C, L)) {
#pragma unroll
    for (int i = 0; i < NW; ++i) Pair16<TP>::unpack(w[i], r[2 * i], r[2 * i + 1]);
    L = dmol_sample<K, GRAD, UMODE, kLikDmol>(y, r, g, C);
  }
  if (GRAD) RowIO<TP, P>::store(row, r);
  return L;
}

template <int K, int TPB, typename TP>
constexpr size_t dmol_tile_smem_bytes() {
  return ((size_t(TPB) * DmolSpt<K>::value * DmolGroup<K, TP>::value * 3 * K * sizeof(TP) + 15) / 16) * 16 + 16 /*mbarrier*/ + (TPB / 32) * sizeof(double);
}

// The body of one tile; returns true if a bulk store is still reading the tile's shared memory (the caller must execute
// ptx::bulk_wait_read0() in thread 0 before the CTA exits or reuses it).  Shared by dmol_tile_kernel and elbo_step_kernel.
template <int K, int TPB, bool GRAD, int UMODE, typename TP, int LIK = kLikDmol>
__device__ __forceinline__ bool dmol_tile_body(const DmolArgs& A, const int64_t tile_id, unsigned char* smem) {
  constexpr int P = 3 * K;
  constexpr int NW = TPB / 32;
  constexpr int G = DmolGroup<K, TP>::value;            // chunks (partial-sum slots) per CTA
  constexpr int SPT = DmolSpt<K>::value * G;
  constexpr int TILE = TPB * SPT;
  constexpr size_t kTileBytes = ((size_t(TILE) * P * sizeof(TP) + 15) / 16) * 16;
  TP* tile = reinterpret_cast<TP*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kTileBytes);
  double* scratch = reinterpret_cast<double*>(smem + kTileBytes + 16);

  const int tid = threadIdx.x;
  constexpr bool kRot = RowRot<TP, K>::enabled && LIK == kLikDmol;   // rotated section walk (bank conflicts, see RowRot)
  const int rot = kRot ? RowRot<TP, K>::rot(tid & 31) : 0;
  // row-local arithmetic in 32 bits (the host rejects T >= 2^31 and more than 2^31 - 1 tiles); only the flat sample
  // offset of the tile is 64-bit (2.95 G parameter elements at the top of config 5)
  const unsigned tile32 = static_cast<unsigned>(tile_id), cpr = static_cast<unsigned>(A.ctas_per_row);
  const unsigned b = tile32 / cpr;
  const unsigned c = tile32 - b * cpr;                  // CTA index within the utterance
  const int T32 = static_cast<int>(A.T);
  const int t0 = static_cast<int>(c) * TILE;
  const int n = min(TILE, T32 - t0);                    // samples of this tile
  const int64_t s0 = static_cast<int64_t>(b) * A.T + t0;   // first flat sample
  const TP* gsrc = static_cast<const TP*>(A.raw) + s0 * P;
  TP* gdst = GRAD ? static_cast<TP*>(A.graw) + s0 * P : nullptr;
  const uint32_t bytes = static_cast<uint32_t>(n) * P * sizeof(TP);
  const bool bulk_in = ((reinterpret_cast<uintptr_t>(gsrc) | bytes) & 15u) == 0;
  const bool bulk_out = GRAD && ((reinterpret_cast<uintptr_t>(gdst) | bytes) & 15u) == 0;
  const uint64_t pol = ptx::policy_evict_first();
  // (K > 12 runs at its register cap with spills: there the early request costs more than it hides -- K = 30 bf16 324 -> 344 us)
  const bool late_load = K > 12 || (A.flags & kFlagSkipPadded) != 0;
  auto start_load = [&]() {
    if (bulk_in) {
      if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
        ptx::mbar_arrive_expect_tx(bar, bytes);
        ptx::bulk_g2s(tile, gsrc, bytes, bar, pol);
      }
    } else {  // unaligned slab (odd T with even K, or an offset view): coalesced element-wise loads
      for (int i = tid; i < n * P; i += TPB) tile[i] = gsrc[i];
    }
  };
  // Unless fully padded tiles are to be skipped, the slab is requested BEFORE the row length is read: the x_sl load (a dependent
  // global access, ~0.4 us) otherwise sits in front of every CTA's bulk copy.
  // sample j of this thread is tile-local index j*TPB + tid: coalesced y / log-prob accesses, conflict-free smem rows
  float yv[SPT];
  auto load_y = [&]() {   // in flight while the slab lands
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const int i = j * TPB + tid;
      yv[j] = (i < n) ? ptx::ldg_stream(A.y + s0 + i) : 0.f;
    }
  };
  if (!late_load) {
    start_load();
    load_y();
  }
  int64_t len64 = A.x_sl ? A.x_sl[b] : A.T;
  const int len = static_cast<int>(len64 < 0 ? 0 : (len64 > A.T ? A.T : len64));
  const int nvalid = max(0, min(n, len - t0));
  const bool skip = (A.flags & kFlagSkipPadded) && nvalid == 0;
  if (late_load) {
    if (!skip) start_load();
    load_y();
  }

  float gs = A.gscale;
  if (GRAD && A.gscale_dev) gs *= static_cast<float>(*A.gscale_dev);
  double acc = 0.0;
  // Interior tiles (full, fully valid, no per-sample upstream gradient) are almost all tiles: no per-sample bound / mask /
  // skip tests, the y-range check folded into one flag per thread.  Same arithmetic, bit-identical results.
  // (K > 12 runs at the 128-register cap: a second copy of the sample body costs more in spills than the tests it saves)
  const bool interior = (K <= 12) && !skip && n == TILE && nvalid == TILE && A.gout == nullptr;
  if (interior) {
    __syncthreads();  // mbarrier init / plain loads visible to everyone
    if (bulk_in) ptx::mbar_wait(bar, 0);
    bool bad = false;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const int i = j * TPB + tid;
      bad |= !(yv[j] <= 1.0f && yv[j] >= -1.0f);
      TP* row = tile + i * P;
      float L;
      if constexpr (kLinPackedTP<TP, K> && UMODE == kUTiny && LIK == kLikDmol) {
        L = dmol_eval_packed_row<K, GRAD, UMODE, TP>(yv[j], row, gs, A.C);
      } else {
        float r[P];
        if constexpr (kRot) RowRot<TP, K>::load(row, r, rot); else RowIO<TP, P>::load(row, r);
        L = dmol_eval<K, GRAD, UMODE, LIK, kLinTP<TP, K>>(yv[j], r, gs, A.C, [&](float (&rr)[P]) {
          if constexpr (kRot) RowRot<TP, K>::load(row, rr, rot); else RowIO<TP, P>::load(row, rr);
        });
        if constexpr (kRot) { if (GRAD) RowRot<TP, K>::store(row, r, rot); } else { if (GRAD) RowIO<TP, P>::store(row, r); }
      }
      if (A.lp) A.lp[s0 + i] = L;
      acc += static_cast<double>(L);
    }
    if (LIK == kLikDmol && bad && A.err_flag) atomicOr(A.err_flag, 1);
  } else {
    float g[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const int i = j * TPB + tid;
      g[j] = 0.f;
      if (i < n) {
        if (LIK == kLikDmol && !(yv[j] <= 1.0f && yv[j] >= -1.0f) && A.err_flag) atomicOr(A.err_flag, 1);
        if (GRAD) {
          g[j] = (i < nvalid) ? gs : 0.f;
          if (A.gout) g[j] *= ptx::ldg_stream(A.gout + s0 + i);
        }
      }
    }
    __syncthreads();  // mbarrier init / plain loads visible to everyone
    if (!skip && bulk_in) ptx::mbar_wait(bar, 0);

#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const int i = j * TPB + tid;
      if (i < n) {
        float L = 0.f;
        TP* row = tile + i * P;
        if constexpr (kLinPackedTP<TP, K> && UMODE == kUTiny && LIK == kLikDmol) {
          if (!skip) {
            L = dmol_eval_packed_row<K, GRAD, UMODE, TP>(yv[j], row, g[j], A.C);
          } else if (GRAD) {
#pragma unroll
            for (int q = 0; q < P / 2; ++q) reinterpret_cast<uint32_t*>(row)[q] = 0u;
          }
        } else {
          float r[P];
          if (!skip) {
            if constexpr (kRot) RowRot<TP, K>::load(row, r, rot); else RowIO<TP, P>::load(row, r);
            L = dmol_eval<K, GRAD, UMODE, LIK, kLinTP<TP, K>>(yv[j], r, g[j], A.C, [&](float (&rr)[P]) {
              if constexpr (kRot) RowRot<TP, K>::load(row, rr, rot); else RowIO<TP, P>::load(row, rr);
            });
          } else {
#pragma unroll
            for (int q = 0; q < P; ++q) r[q] = 0.f;
          }
          if constexpr (kRot) { if (GRAD) RowRot<TP, K>::store(row, r, rot); } else { if (GRAD) RowIO<TP, P>::store(row, r); }
        }
        // reference semantics: log_prob * mask (NaN/inf in the padding propagate like `* 0`), vrnn.py:268
        const float Lm = (i < nvalid) ? L : L * 0.0f;
        if (A.lp) A.lp[s0 + i] = (A.flags & kFlagMaskOutput) ? Lm : L;
        acc += static_cast<double>(Lm);
      }
    }
  }

  if (GRAD) {
    if (bulk_out) {
      ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA unit
      __syncthreads();
      if (tid == 0) {
        ptx::bulk_s2g(gdst, tile, bytes, pol);
        ptx::bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < n * P; i += TPB) gdst[i] = tile[i];
    }
  }
  if (A.partials) {
    const double s = block_sum_f64<NW>(acc, scratch);
    if (tid == 0) {
      double* slot = A.partials + static_cast<int64_t>(b) * A.chunks + static_cast<int64_t>(c) * G;
      slot[0] = s;
      if constexpr (G > 1) {
        const int64_t left = A.chunks - static_cast<int64_t>(c) * G;   // slots of this utterance from ours on
#pragma unroll
        for (int q = 1; q < G; ++q)
          if (q < left) slot[q] = 0.0;
      }
    }
  }
  return GRAD && bulk_out;
}

template <int K, int TPB, bool GRAD, int UMODE, typename TP, int LIK = kLikDmol>
__global__ void __launch_bounds__(TPB, DmolMinBlocks<K, GRAD, TP>::value) dmol_tile_kernel(const DmolArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  ptx::pdl_launch_dependents();   // a KL / finalize launch of the same step may fill this grid's tail (they wait for us to finish)
  const bool pending = dmol_tile_body<K, TPB, GRAD, UMODE, TP, LIK>(A, blockIdx.x, smem);
  if (pending && threadIdx.x == 0) ptx::bulk_wait_read0();  // smem must outlive the bulk store's read
}

// Generic shapes (any K, D >= 1): one thread per sample straight from global memory, two-pass recompute.
// Correctness path for configurations the register kernel is not instantiated for; not tuned.
template <int TPB, bool GRAD>
__global__ void __launch_bounds__(TPB) dmol_generic_kernel(const DmolArgs A) {
  __shared__ double scratch[TPB / 32];
  ptx::pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int64_t tile_id = blockIdx.x;
  const int64_t b = tile_id / A.chunks;
  const int64_t c = tile_id - b * A.chunks;
  const int64_t t0 = c * TPB;
  const int n = static_cast<int>(min(static_cast<int64_t>(TPB), A.T - t0));
  const int64_t s0 = b * A.T + t0;
  int64_t len = A.x_sl ? A.x_sl[b] : A.T;
  len = len < 0 ? 0 : (len > A.T ? A.T : len);
  const int nvalid = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(n), len - t0)));
  const int P = A.K * (2 * A.D + 1);
  const bool in_tile = tid < n, valid = tid < nvalid;
  const bool skip = (A.flags & kFlagSkipPadded) && nvalid == 0;
  float L = 0.f;
  if (in_tile) {
    const int64_t s = s0 + tid;
    const float* yv = A.y + s * A.D;
    if (A.lik == kLikDmol)
      for (int d = 0; d < A.D; ++d)
        if (!(yv[d] <= 1.0f && yv[d] >= -1.0f) && A.err_flag) atomicOr(A.err_flag, 1);
    float g = 0.f;
    if (GRAD) {
      g = valid ? A.gscale : 0.f;
      if (A.gscale_dev) g *= static_cast<float>(*A.gscale_dev);
      if (A.gout) g *= A.gout[s];
    }
    if (!skip) {
      L = dmol_sample_generic<GRAD>(yv, static_cast<const float*>(A.raw) + s * P, A.K, A.D, g, A.C,
                                    GRAD ? static_cast<float*>(A.graw) + s * P : nullptr, A.lik);
    } else if (GRAD) {
      for (int i = 0; i < P; ++i) static_cast<float*>(A.graw)[s * P + i] = 0.f;
    }
  }
  const float Lm = valid ? L : L * 0.0f;
  if (A.lp && in_tile) A.lp[s0 + tid] = (A.flags & kFlagMaskOutput) ? Lm : L;
  if (A.partials) {
    const double s = block_sum_f64<TPB / 32>(static_cast<double>(Lm), scratch);
    if (tid == 0) A.partials[tile_id] = s;
  }
}

// Single discretized logistic (DiscretizedLogisticDense, distributions.py:268-307): raw (B*T, 2) = [mu | log_scale].
// One CTA = kDlSpt * TPB consecutive samples of one utterance; each thread issues its kDlSpt coalesced 64-bit parameter
// loads and 32-bit target loads up front (memory-level parallelism), then evaluates and stores; same mask and
// partial-sum contract as the mixture kernel.  24 B/sample fwd+grad.
constexpr int kDlSpt = 8;

template <int TPB, bool GRAD>
__global__ void __launch_bounds__(TPB) dl_kernel(const DmolArgs A) {
  __shared__ double scratch[TPB / 32];
  ptx::pdl_launch_dependents();
  constexpr int TILE = TPB * kDlSpt;
  const int tid = threadIdx.x;
  const int64_t tile_id = blockIdx.x;
  const int64_t b = tile_id / A.chunks;
  const int64_t c = tile_id - b * A.chunks;
  const int64_t t0 = c * TILE;
  const int n = static_cast<int>(min(static_cast<int64_t>(TILE), A.T - t0));
  const int64_t s0 = b * A.T + t0;
  int64_t len = A.x_sl ? A.x_sl[b] : A.T;
  len = len < 0 ? 0 : (len > A.T ? A.T : len);
  const int nvalid = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(n), len - t0)));
  const bool skip = (A.flags & kFlagSkipPadded) && nvalid == 0;
  const float2* raw = static_cast<const float2*>(A.raw);
  float2* graw = static_cast<float2*>(A.graw);
  float gs = A.gscale;
  if (GRAD && A.gscale_dev) gs *= static_cast<float>(*A.gscale_dev);

  float yv[kDlSpt];
  float2 p[kDlSpt];
#pragma unroll
  for (int j = 0; j < kDlSpt; ++j) {
    const int i = j * TPB + tid;
    yv[j] = 0.f;
    p[j] = make_float2(0.f, 0.f);
    if (i < n) {
      yv[j] = ptx::ldg_stream(A.y + s0 + i);
      if (!skip) p[j] = raw[s0 + i];
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < kDlSpt; ++j) {
    const int i = j * TPB + tid;
    if (i < n) {
      const int64_t s = s0 + i;
      if (!(yv[j] <= 1.0f && yv[j] >= -1.0f) && A.err_flag) atomicOr(A.err_flag, 1);
      float L = 0.f, dmu = 0.f, dls = 0.f;
      if (!skip) dl_component<GRAD>(yv[j], dmol_edge(yv[j], A.C), p[j].x, p[j].y, A.C, L, dmu, dls);
      if (GRAD) {
        float g = (i < nvalid) ? gs : 0.f;
        if (A.gout) g *= A.gout[s];
        graw[s] = make_float2(g * dmu, g * dls);
      }
      const float Lm = (i < nvalid) ? L : L * 0.0f;
      if (A.lp) A.lp[s] = (A.flags & kFlagMaskOutput) ? Lm : L;
      acc += static_cast<double>(Lm);
    }
  }
  if (A.partials) {
    const double sum = block_sum_f64<TPB / 32>(acc, scratch);
    if (tid == 0) A.partials[tile_id] = sum;
  }
}

}  // namespace blvm
