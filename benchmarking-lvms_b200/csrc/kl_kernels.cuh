// Gaussian-KL + free-nats + sequence-mask kernels and the ELBO finalize kernel (sm_100a).
//
// kl_kernel<VEC, GRAD>: elementwise over the (B, Tz, Z) latent grid, 4 inputs in / up to 4 gradients out, 128-bit
// vectorised when the row length allows; one CTA = one chunk of one utterance so that the masked fp64 partial sums
// (raw KL and free-nats-discounted KL) land in partials[b, chunk] without atomics.
// Algorithmic traffic: 16 B read + 16 B written per latent element (fwd+bwd fused) = 32 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blvm_math.cuh"
#include "dmol_kernels.cuh"  // block_sum_f64
#include "ptx_sm100.cuh"

namespace blvm {

struct KlArgs {
  const float *mu_q, *sd_q, *mu_p, *sd_p;
  const int64_t* lens;   // (B) valid latent steps per row, nullptr = all
  const float* gout;     // per-element upstream gradient, nullptr = 1
  float gscale;          // scalar multiplier of every gradient (e.g. beta / sum(x_sl))
  const double* gscale_dev;  // optional device scalar multiplied into gscale (the GradScaler's loss scale under fp16 AMP), nullptr = 1
  float min_kl;          // free_nats / Z
  int fn_enabled;        // 0 = discount_free_nats is the identity (free_nats None/0, variational.py:107-108)
  float* kl;             // per-element raw KL out (nullable)
  float *g_mu_q, *g_sd_q, *g_mu_p, *g_sd_p;  // gradients out (GRAD)
  double* part_kl;       // (B, chunks) masked sums of kl (nullable)
  double* part_klfn;     // (B, chunks) masked sums of max(kl, min_kl) (nullable)
  int64_t B, row_elems, Z, chunks;
  const float* kl_in;    // non-null: the level's KL is already materialised (compute_elbo's `kld_twise`); reduce it
  float* gkl;            // with kl_in: d/d kl out (nullable)
  int vec;               // 1: 128-bit accesses are legal for every pointer and row (set by the host)
  const float* z;        // non-null: Monte-Carlo KL log q(z) - log p(z) at the given sample (bottom-up STCN) instead of the analytic KL
  float* g_z;            // with z and the other gradients: d/d z out (nullable)
};

#ifndef BLVM_KL_TPB
#define BLVM_KL_TPB 128   // 8 elements per thread: per-thread overhead (index arithmetic, fp64 block sums) amortised over twice the work
#endif
constexpr int kKlTPB = BLVM_KL_TPB;
constexpr int kKlVec = 4;
constexpr int kKlChunk = 1024;             // elements per CTA = per partial sum (BLVM_KL_TILE)
static_assert(kKlChunk % (kKlTPB * kKlVec) == 0, "a CTA covers its chunk with whole 128-bit vectors per thread");

// s_kl / s_fn accumulate the (at most 4) elements of one thread in fp32 — KL terms are non-negative, so the 4-term fp32
// sum is good to ~1e-7 relative — and are widened to fp64 once per thread for the block reduction (fp32->fp64
// conversions run on the 16-lane XU pipe; per element they were 2 of ~6 XU operations).
template <bool GRAD>
__device__ __forceinline__ void kl_element(const KlArgs& A, float gs, int64_t idx, bool valid, float mq, float sq, float mp, float sp,
                                           float& kl, float& gmq, float& gsq, float& gmp, float& gsp, float& s_kl,
                                           float& s_fn) {
  const KlTerms t = kl_gaussian_terms(mq, sq, mp, sp);
  kl = t.kl;
  const float klfn = (A.fn_enabled && kl < A.min_kl) ? A.min_kl : kl;  // torch.maximum; NaN propagates
  s_kl += valid ? kl : kl * 0.0f;                                      // `kld * mask` semantics (vrnn.py:272)
  s_fn += valid ? klfn : klfn * 0.0f;
  if (GRAD) {
    float g = valid ? gs * free_nats_gate(kl, A.min_kl, A.fn_enabled != 0) : 0.f;
    if (A.gout) g *= A.gout[idx];
    kl_gaussian_grads(t, sq, g, gmq, gsq, gmp, gsp);
  }
}

// One tile (kKlChunk elements of one utterance) with TPB threads; writes the two fp64 partial sums of the tile.
// `scratch` = 2 x (TPB/32) doubles of shared memory.
template <int TPB, bool GRAD>
__device__ __forceinline__ void kl_tile_body(const KlArgs& A, const int64_t tile_id, double* scratch) {
  constexpr int EPT = kKlChunk / TPB;                    // elements per thread (4 at 256 threads, 8 at 128)
  const int tid = threadIdx.x;
  // 32-bit division: the host rejects launches with more than 2^31 - 1 tiles, and a 64-bit divide costs ~100 instructions
  // per thread — a quarter of this kernel's issue slots when a thread only owns 4 elements
  const unsigned tile32 = static_cast<unsigned>(tile_id), chunks32 = static_cast<unsigned>(A.chunks);
  const int64_t b = tile32 / chunks32;
  const int64_t c = tile32 - static_cast<unsigned>(b) * chunks32;
  const int64_t e0 = c * kKlChunk;                       // first element of this chunk within the row
  const int64_t max_steps = A.row_elems / A.Z;
  int64_t len = A.lens ? A.lens[b] : max_steps;
  len = len < 0 ? 0 : (len > max_steps ? max_steps : len);
  const int64_t nvalid_row = len * A.Z;
  const int64_t base = b * A.row_elems;
  float s_kl = 0.f, s_fn = 0.f;
  float gs = A.gscale;
  if (A.gscale_dev) gs *= static_cast<float>(*A.gscale_dev);

  if (A.kl_in) {                                         // materialised KL: mask, free nats, sums, d/d kl
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int64_t e = e0 + static_cast<int64_t>(j) * TPB + tid;
      if (e < A.row_elems) {
        const int64_t i = base + e;
        const bool valid = e < nvalid_row;
        const float kl = A.kl_in[i];
        const float klfn = (A.fn_enabled && kl < A.min_kl) ? A.min_kl : kl;
        s_kl += valid ? kl : kl * 0.0f;
        s_fn += valid ? klfn : klfn * 0.0f;
        if (A.gkl) A.gkl[i] = valid ? gs * free_nats_gate(kl, A.min_kl, A.fn_enabled != 0) : 0.f;
      }
    }
  } else if (A.z) {                                      // Monte-Carlo KL at the sample z (signed terms; not a hot path)
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int64_t e = e0 + static_cast<int64_t>(j) * TPB + tid;
      if (e < A.row_elems) {
        const int64_t i = base + e;
        const bool valid = e < nvalid_row;
        const KlMcTerms t = kl_mc_terms(A.z[i], A.mu_q[i], A.sd_q[i], A.mu_p[i], A.sd_p[i]);
        const float kl = t.kl;
        const float klfn = (A.fn_enabled && kl < A.min_kl) ? A.min_kl : kl;
        s_kl += valid ? kl : kl * 0.0f;
        s_fn += valid ? klfn : klfn * 0.0f;
        if (A.kl) A.kl[i] = kl;
        if (GRAD) {
          float g = valid ? gs * free_nats_gate(kl, A.min_kl, A.fn_enabled != 0) : 0.f;
          if (A.gout) g *= A.gout[i];
          float gz;
          kl_mc_grads(t, g, A.g_mu_q[i], A.g_sd_q[i], A.g_mu_p[i], A.g_sd_p[i], gz);
          if (A.g_z) A.g_z[i] = gz;
        }
      }
    }
  } else if (A.vec) {
#pragma unroll
    for (int j = 0; j < EPT / kKlVec; ++j) {
      const int64_t e = e0 + (static_cast<int64_t>(j) * TPB + tid) * kKlVec;
      if (e < A.row_elems) {  // row_elems % 4 == 0 on this path, so the whole vector is inside the row
        const int64_t i = base + e;
        const float4 mq = ptx::ldg_stream4(reinterpret_cast<const float4*>(A.mu_q + i));
        const float4 sq = ptx::ldg_stream4(reinterpret_cast<const float4*>(A.sd_q + i));
        const float4 mp = ptx::ldg_stream4(reinterpret_cast<const float4*>(A.mu_p + i));
        const float4 sp = ptx::ldg_stream4(reinterpret_cast<const float4*>(A.sd_p + i));
        float4 kl, gmq, gsq, gmp, gsp;
        kl_element<GRAD>(A, gs, i + 0, e + 0 < nvalid_row, mq.x, sq.x, mp.x, sp.x, kl.x, gmq.x, gsq.x, gmp.x, gsp.x, s_kl, s_fn);
        kl_element<GRAD>(A, gs, i + 1, e + 1 < nvalid_row, mq.y, sq.y, mp.y, sp.y, kl.y, gmq.y, gsq.y, gmp.y, gsp.y, s_kl, s_fn);
        kl_element<GRAD>(A, gs, i + 2, e + 2 < nvalid_row, mq.z, sq.z, mp.z, sp.z, kl.z, gmq.z, gsq.z, gmp.z, gsp.z, s_kl, s_fn);
        kl_element<GRAD>(A, gs, i + 3, e + 3 < nvalid_row, mq.w, sq.w, mp.w, sp.w, kl.w, gmq.w, gsq.w, gmp.w, gsp.w, s_kl, s_fn);
        if (A.kl) ptx::stg_stream4(reinterpret_cast<float4*>(A.kl + i), kl);
        if (GRAD) {
          ptx::stg_stream4(reinterpret_cast<float4*>(A.g_mu_q + i), gmq);
          ptx::stg_stream4(reinterpret_cast<float4*>(A.g_sd_q + i), gsq);
          ptx::stg_stream4(reinterpret_cast<float4*>(A.g_mu_p + i), gmp);
          ptx::stg_stream4(reinterpret_cast<float4*>(A.g_sd_p + i), gsp);
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int64_t e = e0 + static_cast<int64_t>(j) * TPB + tid;
      if (e < A.row_elems) {
        const int64_t i = base + e;
        float kl, gmq, gsq, gmp, gsp;
        kl_element<GRAD>(A, gs, i, e < nvalid_row, A.mu_q[i], A.sd_q[i], A.mu_p[i], A.sd_p[i], kl, gmq, gsq, gmp, gsp, s_kl, s_fn);
        if (A.kl) A.kl[i] = kl;
        if (GRAD) {
          A.g_mu_q[i] = gmq;
          A.g_sd_q[i] = gsq;
          A.g_mu_p[i] = gmp;
          A.g_sd_p[i] = gsp;
        }
      }
    }
  }
  if (A.part_kl) {
    const double a = block_sum_f64<TPB / 32>(static_cast<double>(s_kl), scratch);
    const double f = block_sum_f64<TPB / 32>(static_cast<double>(s_fn), scratch + TPB / 32);
    if (tid == 0) {
      A.part_kl[tile_id] = a;
      A.part_klfn[tile_id] = f;
    }
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(kKlTPB) kl_kernel(const KlArgs A) {
  __shared__ double scratch[2 * (kKlTPB / 32)];
  ptx::pdl_launch_dependents();
  kl_tile_body<kKlTPB, GRAD>(A, blockIdx.x, scratch);
  // Launched with BLVM_FLAG_OVERLAP_PREV this grid runs next to the DMoL (or previous KL level's) grid of the same step,
  // none of whose outputs it reads.  Waiting for that grid HERE, before any CTA of this one retires, makes completion
  // transitive: when this grid is complete so is everything before it, which is what the finalize kernel's own wait relies on.
  ptx::pdl_wait();
}

// All latent levels of a hierarchical model (Clockwork-VAE: 3, STCN: up to 5) in ONE launch: the grid is the concatenation of
// the levels' tile ranges, a CTA finds its level by a scan over at most 8 prefix sums.  Same tile body, same partial-sum
// layout as one kl_kernel launch per level (bit-identical), 2 + 1 instead of 2 + L launches per step.
constexpr int kMaxLevels = 8;
struct KlMultiArgs {
  KlArgs level[kMaxLevels];
  int64_t tile_begin[kMaxLevels + 1];
  int n_levels;
};
template <bool GRAD>
__global__ void __launch_bounds__(kKlTPB) kl_multi_kernel(const __grid_constant__ KlMultiArgs M) {
  __shared__ double scratch[2 * (kKlTPB / 32)];
  ptx::pdl_launch_dependents();
  int l = 0;
  while (l + 1 < M.n_levels && static_cast<int64_t>(blockIdx.x) >= M.tile_begin[l + 1]) ++l;
  kl_tile_body<kKlTPB, GRAD>(M.level[l], static_cast<int64_t>(blockIdx.x) - M.tile_begin[l], scratch);
  ptx::pdl_wait();   // see kl_kernel
}

// ---- finalize: per-tile partials -> per-utterance sums -> loss / ELBO / bits-per-dim --------------------------------
struct FinalizeArgs {
  const double* logp_part;          // (B, logp_chunks), nullable (=> log p = 0)
  int64_t logp_chunks;
  const double* kl_part[kMaxLevels];    // per level (B, kl_chunks[l])
  const double* klfn_part[kMaxLevels];
  int64_t kl_chunks[kMaxLevels];
  int n_levels;
  const int64_t* x_sl;              // (B) device
  int64_t B;
  double beta;
  double denom;                     // normaliser of the loss; <= 0 means sum(x_sl) (data-parallel callers pass global/world)
  double* rows;                     // (4 + n_levels, B): logp, kl, kl_fn, elbo, kl_level_l...
  double* scalars;                  // (8): loss, sum logp, sum kl, sum kl_fn, sum elbo, sum x_sl, bpd, nan-safe loss
  int nansum_loss;                  // 1: scalars[0] is the nansum loss as well (WaveNet.compute_loss, wavenet.py:145)
};

constexpr int kFinTPB = 256;
constexpr int kFinWarps = kFinTPB / 32;

// ---- scalar-sum exchange over NVLink peer memory (fused into the finalize kernel) -----------------------------------
// Every rank owns one symmetric buffer (torch symmetric memory: the same allocation mapped into all ranks of the node):
//   values [kExNbuf][kExMaxWorld][8] fp64   slot (buf, r) = the 8 scalars rank r published for step seq, buf = seq % kExNbuf
//   flags  [kExNbuf][kExMaxWorld]    u64    = seq once the slot is complete
// The last CTA of finalize writes its scalars into slot (buf, my_rank) of EVERY rank's buffer with plain peer stores
// (NVLink P2P), fences system-wide, then releases the flags: the all-gather costs no extra launch and no host call.
// A one-warp consumer kernel (exchange_consume_kernel) waits for the W flags of a step and adds the slots in rank
// order (deterministic).  4 buffers + the consumer's wait bound the rank skew to < 4 steps.
constexpr int kExNbuf = 4;
constexpr int kExMaxWorld = 8;
constexpr int kExValues = 8;
constexpr size_t kExBufferBytes = sizeof(double) * kExNbuf * kExMaxWorld * kExValues + sizeof(unsigned long long) * kExNbuf * kExMaxWorld;

struct ExchangeArgs {
  int rank, world;                         // world == 0: no exchange
  double* peer_base[kExMaxWorld];          // every rank's buffer as mapped in this process (peer_base[rank] is local)
  unsigned long long* counters;            // local, not symmetric: [0] steps published, [1] steps consumed
  double* global_out;                      // nullable (8): the PREVIOUS step's global sums, consumed in the same kernel
  int* err;                                // nullable: bit 0 timeout, bit 1 slot overrun
};
__device__ __forceinline__ double* ex_slot(double* base, int buf, int r) { return base + (buf * kExMaxWorld + r) * kExValues; }
__device__ __forceinline__ unsigned long long* ex_flag(double* base, int buf, int r) {
  return reinterpret_cast<unsigned long long*>(base + kExNbuf * kExMaxWorld * kExValues) + buf * kExMaxWorld + r;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Wait for the W slots of step `seq` in the local buffer and add them in rank order.  Called by a full warp: lane r
// polls / reads rank r's slot (parallel over ranks), lane 0 accumulates through shuffles in a fixed order.
__device__ __forceinline__ void ex_consume_step(double* local_base, int world, unsigned long long seq, double beta, double* out,
                                                int* err) {
  const int lane = threadIdx.x & 31;
  const int buf = static_cast<int>(seq % kExNbuf);
  double v[kExValues];
#pragma unroll
  for (int i = 0; i < kExValues; ++i) v[i] = 0.0;
  if (lane < world) {
    const unsigned long long* f = ex_flag(local_base, buf, lane);
    unsigned long long got = ld_acquire_sys(f);
    long long spins = 0;
    while (got < seq) {
      __nanosleep(100);
      got = ld_acquire_sys(f);
      if (++spins > (1LL << 26)) {   // ~10 s: a peer died or never reached this step
        if (err) atomicOr(err, 1);
        break;
      }
    }
    if (got > seq && err) atomicOr(err, 2);
    const volatile double* src = ex_slot(local_base, buf, lane);
#pragma unroll
    for (int i = 0; i < kExValues; ++i) v[i] = src[i];
  }
  double s[kExValues];
#pragma unroll
  for (int i = 0; i < kExValues; ++i) s[i] = 0.0;
  for (int r = 0; r < world; ++r) {
#pragma unroll
    for (int i = 0; i < kExValues; ++i) s[i] += __shfl_sync(0xffffffffu, v[i], r);
  }
  if (lane == 0) {
    out[1] = s[1]; out[2] = s[2]; out[3] = s[3]; out[4] = s[4]; out[5] = s[5];
    out[0] = -(s[1] - beta * s[3]) / s[5];                        // global loss (vrnn.py:277)
    out[6] = -s[4] / 0.6931471805599453 / s[5];                   // global bits per dim
    out[7] = static_cast<double>(seq);                            // which step these sums belong to
  }
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Lane-partial sum of one utterance's chunk partials, lanes strided over the chunks (coalesced).  The first SLOTS loads are
// issued unconditionally up front (predicated, no dependent add in between) so that the loads of ALL arrays of the
// utterance are in flight together: the kernel is a chain of L2 latencies, not of bytes.  Fixed order: deterministic.
template <int SLOTS>
__device__ __forceinline__ void lane_first_loads(const double* __restrict__ p, int64_t n, int lane, double (&v)[SLOTS]) {
#pragma unroll
  for (int u = 0; u < SLOTS; ++u) {
    const int64_t c = lane + 32 * u;
    v[u] = (p != nullptr && c < n) ? __ldcg(p + c) : 0.0;   // written by other CTAs: read through L2
  }
}
template <int SLOTS>
__device__ __forceinline__ double lane_finish(const double* __restrict__ p, int64_t n, int lane, const double (&v)[SLOTS]) {
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < SLOTS; ++u) s += v[u];
  if (p != nullptr)
    for (int64_t c = lane + 32 * SLOTS; c < n; c += 32) s += __ldcg(p + c);
  return s;
}

// Reduce the per-tile partials of utterance b (called by ONE warp; lane 0 writes the row entries).
__device__ __forceinline__ void finalize_row(const FinalizeArgs& A, int64_t b, int lane) {
  double v_logp[4], v_kl[kMaxLevels][1], v_fn[kMaxLevels][1];
  lane_first_loads<4>(A.logp_part ? A.logp_part + b * A.logp_chunks : nullptr, A.logp_chunks, lane, v_logp);
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    const bool on = l < A.n_levels;
    lane_first_loads<1>(on ? A.kl_part[l] + b * A.kl_chunks[l] : nullptr, on ? A.kl_chunks[l] : 0, lane, v_kl[l]);
    lane_first_loads<1>(on ? A.klfn_part[l] + b * A.kl_chunks[l] : nullptr, on ? A.kl_chunks[l] : 0, lane, v_fn[l]);
  }
  const double logp = warp_sum_f64(lane_finish<4>(A.logp_part ? A.logp_part + b * A.logp_chunks : nullptr, A.logp_chunks, lane, v_logp));
  double kl = 0.0, fn = 0.0;
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    if (l < A.n_levels) {
      const double kl_l = warp_sum_f64(lane_finish<1>(A.kl_part[l] + b * A.kl_chunks[l], A.kl_chunks[l], lane, v_kl[l]));
      const double fn_l = warp_sum_f64(lane_finish<1>(A.klfn_part[l] + b * A.kl_chunks[l], A.kl_chunks[l], lane, v_fn[l]));
      if (lane == 0) A.rows[(4 + l) * A.B + b] = kl_l;
      kl += kl_l;   // sum over levels (clockwork_vae.py:155, stcn.py:290)
      fn += fn_l;
    }
  }
  if (lane == 0) {
    A.rows[0 * A.B + b] = logp;
    A.rows[1 * A.B + b] = kl;
    A.rows[2 * A.B + b] = fn;
    A.rows[3 * A.B + b] = logp - kl;                     // elbo (vrnn.py:273)
  }
}

// Rows -> scalars (+ the fused NVLink exchange), executed by ONE whole CTA of TPB threads after every row is final.
// `scratch` = 6 x (TPB/32) doubles of shared memory.
template <int TPB>
__device__ __forceinline__ void finalize_scalars(const FinalizeArgs& A, const ExchangeArgs& X, double* scratch) {
  constexpr int NW = TPB / 32;
  const int tid = threadIdx.x;
  double t_logp = 0, t_kl = 0, t_fn = 0, t_len = 0, t_nan_logp = 0, t_obj = 0;
  for (int64_t r = tid; r < A.B; r += TPB) {
    const double logp = __ldcg(A.rows + 0 * A.B + r), kl = __ldcg(A.rows + 1 * A.B + r), fn = __ldcg(A.rows + 2 * A.B + r);
    t_logp += logp;
    t_kl += kl;
    t_fn += fn;
    t_obj += logp - A.beta * fn;                         // vrnn.py:277 numerator
    t_nan_logp += (logp == logp) ? logp : 0.0;           // nansum (wavenet.py:145)
    t_len += static_cast<double>(A.x_sl[r]);
  }
  // six sums, one barrier: warp butterflies (independent chains), one shared-memory hop, thread 0 adds the warps in order
  double t6[6] = {t_logp, t_kl, t_fn, t_obj, t_nan_logp, t_len};
#pragma unroll
  for (int q = 0; q < 6; ++q) t6[q] = warp_sum_f64(t6[q]);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q) scratch[q * NW + (tid >> 5)] = t6[q];
  }
  __syncthreads();
  double s6[6] = {0, 0, 0, 0, 0, 0};
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
      for (int w = 0; w < NW; ++w) s6[q] += scratch[q * NW + w];
  }
  const double s_logp = s6[0], s_kl = s6[1], s_fn = s6[2], s_obj = s6[3], s_nan = s6[4], s_len = s6[5];
  if (tid == 0) {
    const double dn = A.denom > 0.0 ? A.denom : s_len;
    A.scalars[0] = A.nansum_loss ? -s_nan / dn : -s_obj / dn;   // loss (vrnn.py:277 / wavenet.py:145), consistent with the gradients' 1/denom
    A.scalars[1] = s_logp;
    A.scalars[2] = s_kl;
    A.scalars[3] = s_fn;
    A.scalars[4] = s_logp - s_kl;                        // sum elbo
    A.scalars[5] = s_len;
    A.scalars[6] = -(s_logp - s_kl) / 0.6931471805599453 / s_len;  // bits per dim (metrics.py:456)
    A.scalars[7] = -s_nan / dn;                          // WaveNet's nansum loss
  }
  // ---- fused exchange (warp 0): publish this step's scalars into every rank's buffer over NVLink peer memory, then
  // add up the previous step's slots.  Lane p talks to rank p.
  if (X.world > 0 && tid < 32) {
    __syncwarp();
    unsigned long long seq = 0;
    if (tid == 0) {
      __threadfence();                                   // scalars[] written above are visible to the lanes below
      seq = ++X.counters[0];
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    const int buf = static_cast<int>(seq % kExNbuf);
    if (tid < X.world) {
      volatile double* dst = ex_slot(X.peer_base[tid], buf, X.rank);
#pragma unroll
      for (int i = 0; i < kExValues; ++i) dst[i] = __ldcg(A.scalars + i);
      __threadfence_system();                            // the slot is complete before its flag
      st_relaxed_sys(ex_flag(X.peer_base[tid], buf, X.rank), seq);
    }
    // the PREVIOUS step's slots landed a whole step ago: no stall, no extra launch; this wait is also what bounds the
    // skew between ranks to < kExNbuf steps
    unsigned long long consumed = (tid == 0) ? X.counters[1] : 0;
    consumed = __shfl_sync(0xffffffffu, consumed, 0);
    if (X.global_out && seq >= 2 && consumed < seq - 1) {
      ex_consume_step(X.peer_base[X.rank], X.world, seq - 1, A.beta, X.global_out, X.err);
      if (tid == 0) X.counters[1] = seq - 1;
    }
  }
}

// One warp per utterance reduces its per-tile partials; the last CTA to finish (device counter, self-resetting)
// reduces the per-utterance rows to the scalars in a fixed order.  Grid = ceil(B / 8) CTAs.
static __global__ void __launch_bounds__(kFinTPB) elbo_finalize_kernel(const FinalizeArgs A, unsigned int* counter, const ExchangeArgs X) {
  __shared__ double scratch[6 * kFinWarps];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kFinWarps + warp;
  ptx::pdl_wait();   // launched as a programmatic dependent: the CTAs are resident before the partial sums are final
  if (b < A.B) finalize_row(A, b, lane);
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid == 0) *counter = 0u;                           // ready for the next launch on this stream
  finalize_scalars<kFinTPB>(A, X, scratch);
}

// Small batches (model-shaped steps: B = 32 / 64 utterances): ONE CTA of 32 warps reduces every utterance (warp w takes rows
// w, w + 32, ...) and then the scalars.  No inter-CTA hand-off (fence + atomic + second load phase through L2), which is
// the longest link of the multi-CTA kernel's latency chain when there are only a handful of rows.  Same per-row and
// per-scalar summation orders: bit-identical outputs.
constexpr int kFinSmallTPB = 1024;
static __global__ void __launch_bounds__(kFinSmallTPB) elbo_finalize_small_kernel(const FinalizeArgs A, const ExchangeArgs X) {
  __shared__ double scratch[6 * (kFinSmallTPB / 32)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ptx::pdl_wait();
  for (int64_t b = warp; b < A.B; b += kFinSmallTPB / 32) finalize_row(A, b, lane);
  __syncthreads();                                       // the rows (global memory, written by lane 0 of each warp) are visible CTA-wide
  finalize_scalars<kFinSmallTPB>(A, X, scratch);
}

// Consume one published step: wait until all W ranks' slots of step `seq = published - lag` have landed in the LOCAL
// buffer, add them in rank order and recompute the ratio entries.  No-op if that step does not exist yet or was already
// consumed.  err: bit 0 = timeout (a peer never published), bit 1 = slot overrun (a peer ran >= kExNbuf steps ahead).
static __global__ void __launch_bounds__(32) exchange_consume_kernel(double* local_base, int world, unsigned long long* counters,
                                                              int lag, double beta, double* out, int* err) {
  const int lane = threadIdx.x;
  const unsigned long long published = counters[0], consumed = counters[1];
  if (published < static_cast<unsigned long long>(lag) + 1) return;
  const unsigned long long seq = published - lag;
  if (seq <= consumed) return;
  ex_consume_step(local_base, world, seq, beta, out, err);
  if (lane == 0) counters[1] = seq;
}

}  // namespace blvm
