// Per-element math of the blvm DMoL / Gaussian-KL path, written once for device (sm_100a, MUFU approximations) and
// host (g++, libm) so that the closed forms can be checked on the CPU against the fp64 oracle before a GPU run
// (tests/hostsim/).  The host build is test infrastructure; the product only ever runs the device instantiation.
//
// Reference semantics restated here (paths under /root/reference):
//   blvm/modules/distributions.py:383-387   split of the Linear output, log_scale.clamp(min=log_epsilon)
//   blvm/utils/log_likelihoods.py:198-231   the four-branch discretized logistic, log_softmax, logsumexp
//   blvm/utils/variational.py:67-70,86-122  Gaussian KL (std-dev parametrisation), free nats
//
// Numerical design: the reference evaluates cdf_delta = sigmoid(a) - sigmoid(b) in fp32, which loses
// up to 3 digits to cancellation when the bin is narrow (16-bit audio: half width 1.5e-5).  We evaluate the same
// quantity without cancellation and without a divergent branch through
//     sigmoid(m+u) - sigmoid(m-u) = sinh(u) / (cosh(m) + cosh(u))            (a = m + u, b = m - u)
// (see dl_mid below), so results track the reference run in fp64 to ~3e-7 relative, well inside the fp32 reference's
// own error band (1e-3).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BLVM_HD __host__ __device__ __forceinline__
#else
#define BLVM_HD inline
#endif

namespace blvm {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kDeltaThresh = 1e-5f;   // log_likelihoods.py:222  `cdf_delta > 1e-5`   (fp32-rounded Python double)
constexpr float kDeltaFloor = 1e-10f;   // log_likelihoods.py:222  `clamp(cdf_delta, min=1e-10)`

// Constants derived on the host in double and rounded to fp32 exactly like torch rounds Python scalars that meet an
// fp32 tensor (blvm_b200.cu: make_consts).
struct DmolConsts {
  // Gaussian-mixture variant (sibling likelihood, same packed layout [logits | mu | log_sd]); unused by the DMoL path
  float sp_beta;        // softplus beta of the sd activation: ln2/initial_sd     distributions.py:167-170
  float sp_inv_beta;
  float sd_add;         // epsilon added after the softplus (AddConstant)          distributions.py:168
  float sd_floor;       // gaussian_ll's `epsilon` clamp under no_grad (0 = none)  log_likelihoods.py:33-35
  float h;              // 1/(num_bins-1)            log_likelihoods.py:206,208
  float log_two_h;      // log(2/(num_bins-1))
  float log_delta_thresh;  // log(float(1e-5))
  float lo_thresh;      // 2/num_bins - 1            log_likelihoods.py:226   y <  lo  -> lower edge bin
  float hi_thresh;      // 1 - 2/num_bins            log_likelihoods.py:227   y >  hi  -> upper edge bin
  float log_half_bins;  // log(num_bins/2)           log_likelihoods.py:222
  float log_eps;        // log_epsilon (-7)          distributions.py:386
  // the same constants in log2 units for dl_mid_pair_tiny (products formed in double on the host)
  float log_half_bins2;     // log2(num_bins/2)
  float neg_log_ratio2;     // -log2(num_bins/(num_bins-1)) = -[log(2h) + log(nb/2)] log2(e): first arm minus second arm at u = 0
  float log_delta_thresh2;  // log2(float(1e-5))
  // linear-domain constants of dl_mid_pair_lin
  float two_h;          // 2/(num_bins-1): the bin width of the first arm                 log_likelihoods.py:206-210
  float fb;             // 2/num_bins: exp(-log(num_bins/2)), the factor of the second arm  :222
  float fd0;            // two_h - fb = 2/(num_bins (num_bins-1)), formed in double
  float neg_two_h_sixth;  // -two_h/6
  float neg_h2;         // -h^2
};

// Host build (tests/hostsim): libm stands in for the MUFU unit.  With -DBLVM_HOSTSIM_MUFU_BITS=n the result is degraded
// to the precision class of the approximate instruction (rounded to 2^n ulps, i.e. a relative error up to 2^(n-24)),
// so that the CPU suite exercises the closed forms under MUFU-like error as well.
#if !defined(__CUDA_ARCH__) && defined(BLVM_HOSTSIM_MUFU_BITS)
inline float hostsim_degrade(float v, int bits) {
  uint32_t u;
  __builtin_memcpy(&u, &v, 4);
  u = (u + (1u << (bits - 1))) & ~((1u << bits) - 1u);   // nearest multiple of 2^bits ulps: |error| <= 2^(bits-1) ulps
  __builtin_memcpy(&v, &u, 4);
  return v;
}
// lg2.approx: absolute error 2^-22 for results in (-1, 1), relative 2^-22 outside
inline float hostsim_degrade_lg2(float v, int bits) {
  if (fabsf(v) < 1.0f) return ldexpf(rintf(ldexpf(v, 23 - bits)), bits - 23);
  return hostsim_degrade(v, bits);
}
#define BLVM_MUFU(v, bits) hostsim_degrade((v), (bits))
#define BLVM_MUFU_LG2(v, bits) hostsim_degrade_lg2((v), (bits))
#else
#define BLVM_MUFU(v, bits) (v)
#define BLVM_MUFU_LG2(v, bits) (v)
#endif

BLVM_HD float fast_ex2(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return BLVM_MUFU(exp2f(x), BLVM_HOSTSIM_MUFU_BITS);
#endif
}
BLVM_HD float fast_lg2(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return BLVM_MUFU_LG2(log2f(x), BLVM_HOSTSIM_MUFU_BITS);
#endif
}
BLVM_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return BLVM_MUFU(1.0f / x, BLVM_HOSTSIM_MUFU_BITS - 1);
#endif
}
BLVM_HD float fast_exp(float x) { return fast_ex2(x * kLog2e); }
// exp(x) with the rounding of the scaled argument compensated: x*log2(e) = p_hi + p_lo exactly (to ~2^-48),
// exp(x) = 2^p_hi * (1 + ln2 * p_lo).  Leaves only the MUFU.EX2 error (~2^-22.5); 3 extra FMA-pipe ops.
#ifndef BLVM_COMPENSATED_EXP
#define BLVM_COMPENSATED_EXP 2   // 2: compensated in the scalar paths (dl_mid / dl_edge), plain MUFU in dl_mid_pair_tiny; 1: everywhere; 0: nowhere
#endif
BLVM_HD float accurate_exp(float x) {
  constexpr float kLog2eLo = 1.925963033500011e-08f;  // log2(e) - float(log2(e))
  const float p_hi = x * kLog2e;
  const float p_lo = fmaf(x, kLog2eLo, fmaf(x, kLog2e, -p_hi));
  const float e = fast_ex2(p_hi);
  return fmaf(e, p_lo * kLn2, e);
}
BLVM_HD float fast_log(float x) { return fast_lg2(x) * kLn2; }

// exp(-log_scale): compensated by default (build with -DBLVM_COMPENSATED_EXP=0 to measure the plain MUFU path)
BLVM_HD float stable_exp(float x) {
#if BLVM_COMPENSATED_EXP
  return accurate_exp(x);
#else
  return fast_exp(x);
#endif
}

enum : int { kEdgeNone = 0, kEdgeLower = 1, kEdgeUpper = 2 };

// Which of the two y-predicates fires (bit-exact fp32 compares against fp32-rounded constants; upper wins because it
// is the last torch.where of log_likelihoods.py:226-227).
BLVM_HD int dmol_edge(float y, const DmolConsts& C) {
  return (y > C.hi_thresh) ? kEdgeUpper : ((y < C.lo_thresh) ? kEdgeLower : kEdgeNone);
}

// One (sample, component): log-prob of the discretized logistic and, if GRAD, d lp/d loc and d lp/d raw_log_scale
// (the clamp's pass-at-equality gate included).
//
// With m = (y-mu)/s (mid_in), u = h/s (half bin width in scale units), a = m+u, b = m-u and E = exp(-|m|):
//   cdf_delta = sigmoid(a) - sigmoid(b) = sinh(u) / (cosh(m) + cosh(u)) = 2 E sinh(u) / D,
//   D = (1+E)^2 + E w,  w = 2 cosh(u) - 2,
// hence  log cdf_delta = [-|m| - ls - 2 log(1+E)] + log(2h) + log(sinh(u)/u) - log1p(E w/(1+E)^2)
// while the reference's fallback arm is  [-|m| - ls - 2 log(1+E)] - log(nb/2):  both arms share the bracket (one EX2,
// one LG2, one RCP) and differ by O(u^2) corrections, so the `cdf_delta > 1e-5` selection is a branch-free select and
// nothing cancels.  Derivatives:  d/dm = -sgn(m) (1-E^2)/D,  d/du = coth(u) - cdf_delta  (DESIGN.md §4).
//
// UMODE selects how the u-terms are evaluated (the host picks it from num_bins and log_epsilon, which bound
// u <= h exp(-log_epsilon)):
//   kUTiny    u < 0.02 guaranteed (16-bit audio with the -7 clamp: u <= 0.0167): first-order series, O(u^4) ~ 1e-7 dropped
//   kUGeneral any u: second-order series below 1/8, exact evaluation from q = exp(-u) above (8-bit data, small scales)
enum : int { kUTiny = 0, kUGeneral = 1 };
constexpr float kTinyU = 0.02f;
constexpr float kSmallU = 0.125f;

// Non-edge sample (neither y-predicate fires): the common case, no divergent branch in kUTiny mode.
template <bool GRAD, int UMODE>
BLVM_HD void dl_mid(float y, float mu, float raw_ls, const DmolConsts& C, float& lp, float& dmu, float& dls) {
  const float ls = (raw_ls < C.log_eps) ? C.log_eps : raw_ls;  // clamp(min): NaN propagates like torch
  const float inv = stable_exp(-ls);                           // exp(-log_scale)            :203
  const float m = inv * (y - mu);                              // mid_in                     :202,219
  const float u = C.h * inv;
  const float am = fabsf(m);
  const float E = fast_ex2(-am * kLog2e);
  const float p1 = 1.f + E;
  const float r = fast_rcp(p1);
  const float common = (-am - ls) - (2.f * kLn2) * fast_lg2(p1);   // log_pdf_mid = m - ls - 2 softplus(m)  :220
  const float lp_fb = common - C.log_half_bins;                    // second arm of :221-223
  // tanh(|m|/2) = (1-E)/(1+E); odd series below 1/8 so that the bin-centre gradient does not cancel
  // Derivatives are carried as magnitudes: d lp/d m = -sgn(m) th, so d lp/d loc = -inv * d/dm = copysign(inv th, m) and
  // m * d/dm = -|m| th; the sign of m is applied once at the end.
  float th = 0.f;                                                  // tanh(|m|/2) = |1 - 2 sigmoid(m)|
  if (GRAD) {
    const float hx = 0.5f * am, hx2 = hx * hx;
    const float th_series = hx * fmaf(hx2, fmaf(hx2, 2.0f / 15.0f, -1.0f / 3.0f), 1.0f);
    th = (am < 0.25f) ? th_series : (1.f - E) * r;
  }
  float lp_d, th_d = 0.f, udu = 0.f;
  bool big;
  if (UMODE == kUTiny) {
    const float u2 = u * u;
    const float er2 = E * r * r;
    const float eps = er2 * u2;                                    // E w / (1+E)^2,  w = u^2 + O(u^4)
    lp_d = common + (fmaf(u2, 1.0f / 6.0f, C.log_two_h) - eps);    // log cdf_delta, first arm of :221-223
    big = lp_d > C.log_delta_thresh;                               // cdf_delta > 1e-5
    if (GRAD) {
      th_d = fmaf(-eps, th, th);                                   // th / (1 + eps)
      udu = fmaf(u2, fmaf(er2, -2.0f, 1.0f / 3.0f), 1.0f);         // u coth(u) - u cdf_delta
    }
  } else if (u < kSmallU) {
    const float u2 = u * u;
    const float er2 = E * r * r;
    const float eps = er2 * (u2 * fmaf(u2, 1.0f / 12.0f, 1.0f));
    const float corr = u2 * fmaf(u2, -1.0f / 180.0f, 1.0f / 6.0f) - eps * fmaf(eps, -0.5f, 1.0f);
    lp_d = common + (C.log_two_h + corr);
    big = lp_d > C.log_delta_thresh;
    if (GRAD) {
      const float inv1pe = fmaf(eps, eps - 1.0f, 1.0f);            // 1/(1+eps)
      th_d = th * inv1pe;
      const float udelta = 2.0f * u2 * er2 * fmaf(u2, 1.0f / 6.0f, 1.0f) * inv1pe;  // u * cdf_delta
      udu = fmaf(u2, fmaf(u2, -1.0f / 45.0f, 1.0f / 3.0f), 1.0f) - udelta;
    }
  } else {
    const float q = fast_ex2(-u * kLog2e);                         // exp(-u); everything below is overflow-free
    const float omq = 1.f - q, omq2 = omq * (1.f + q);
    const float rdq = fast_rcp(fmaf(q * p1, p1, E * omq * omq));   // q / D
    const float delta = E * omq2 * rdq;                            // cdf_delta                  :210
    big = delta > kDeltaThresh;
    lp_d = kLn2 * fast_lg2(fmaxf(delta, kDeltaFloor));
    if (GRAD) {
      th_d = th * (p1 * p1 * q * rdq);                             // (1-E^2)/D
      udu = u * (fmaf(q, q, 1.f) * fast_rcp(omq2) - delta);        // u (coth(u) - cdf_delta)
    }
  }
  lp = big ? lp_d : lp_fb;
  if (GRAD) {
    const float th_sel = big ? th_d : th;
    dmu = copysignf(inv * th_sel, m);                              // -inv * d lp/d m
    dls = fmaf(am, th_sel, big ? -udu : -1.0f);                    // -(m d/dm + u d/du)  resp.  -m d/dm - 1
    if (raw_ls < C.log_eps) dls = 0.f;  // clamp(min=eps) blocks the gradient strictly below eps, passes at equality
  }
}

// ---- two components per instruction: packed fp32x2 arithmetic (sm_100a FFMA2) -------------------------------------------
// Blackwell's FMA pipe executes fma/mul/add on a PAIR of fp32 values held in an aligned register pair in one instruction
// (PTX fma.rn.f32x2 -> SASS FFMA2).  The kernels are bound by issue slots, not by FMA-pipe lanes (ncu: issue-active 72 %,
// fma pipe 35 %), so evaluating components k and k+1 of a sample together removes about a fifth of all instructions; the
// MUFU evaluations, compares and selects stay scalar.  Each lane of a packed op rounds exactly like the scalar op, so the
// host build below (plain fmaf / * / +) computes the same values.
#ifndef BLVM_PACKED_FP32
#define BLVM_PACKED_FP32 1
#endif
struct F2 {
  float x, y;
};
BLVM_HD F2 f2(float a, float b) { return F2{a, b}; }
BLVM_HD F2 f2(float a) { return F2{a, a}; }
#if defined(__CUDA_ARCH__) && BLVM_PACKED_FP32
__device__ __forceinline__ unsigned long long f2_pack(F2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ F2 f2_unpack(unsigned long long v) {
  F2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(r);
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
#else
BLVM_HD F2 fma2(F2 a, F2 b, F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
BLVM_HD F2 mul2(F2 a, F2 b) { return F2{a.x * b.x, a.y * b.y}; }
BLVM_HD F2 add2(F2 a, F2 b) { return F2{a.x + b.x, a.y + b.y}; }
#endif
BLVM_HD F2 abs2(F2 a) { return F2{fabsf(a.x), fabsf(a.y)}; }
BLVM_HD F2 ex2_2(F2 a) { return F2{fast_ex2(a.x), fast_ex2(a.y)}; }
BLVM_HD F2 lg2_2(F2 a) { return F2{fast_lg2(a.x), fast_lg2(a.y)}; }
BLVM_HD F2 rcp_2(F2 a) { return F2{fast_rcp(a.x), fast_rcp(a.y)}; }

// dl_mid<GRAD, kUTiny> for two (sample, component) pairs at once: components k and k+1 of one sample (y.x == y.y), or the
// single component of two samples when K == 1.  In kUTiny mode EVERY mid-bin evaluation goes through this function (a
// leftover component duplicates its lane), so a value never depends on what it was paired with.
//
// The kernels that call this are bound by issue slots (ncu r2a, bf16 K = 10: 105 M warp instructions, issue-active 79 %,
// DRAM 46 %), so the formulas of dl_mid are arranged for the fewest instructions:
//  * the log-prob is carried in log2 units (lp2 = lp * log2(e)): the MUFU.EX2 arguments p = -ls log2(e) and t = -|m| log2(e)
//    ARE the two leading terms of lp2, and the mixture's exp / log work in base 2 anyway (dmol_sample converts once);
//  * selections are arithmetic blends with a 1 / 0 float (one FSET.BF each, packed FMAs afterwards) instead of predicates
//    that live across the whole evaluation: with 5-15 pairs in flight the compiler ran out of predicate registers and
//    spilled them into a bit mask (2-3 LOP3 per predicate);
//  * signs are folded into constants (-u^2 = inv^2 * (-h^2) is what is carried) because packed operands carry no negate
//    modifier in PTX;
//  * exp(-ls) is the plain MUFU result: its 2^-22.5 relative error plus the argument rounding (|ls| <= 7: 3e-7) stays
//    below 4 % of the parity tolerance in m, lp and the gradients (tools/hostsim_accuracy.py; -DBLVM_COMPENSATED_EXP=1
//    restores the compensated evaluation).
// Out: lp2 (log2 units); if GRAD, dmu = d lp/d loc and dls = d lp/d raw_log_scale in natural units.
constexpr float kInvLn2Sixth = kLog2e / 6.0f;
// single roundings the compiler may not contract into an FMA (results must not depend on which kernel inlines the code)
BLVM_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
BLVM_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

// max(x, lo) that keeps a NaN x (torch.clamp(min=) propagates NaN; fmaxf would drop it): one FMNMX.NAN
BLVM_HD float max_keep_nan(float x, float lo) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
  return r;
#else
  return (x < lo) ? lo : x;
#endif
}

template <bool GRAD>
BLVM_HD void dl_mid_pair_tiny(F2 y, F2 mu, F2 raw_ls, const DmolConsts& C, F2& lp2, F2& dmu, F2& dls) {
  const F2 ls = f2(max_keep_nan(raw_ls.x, C.log_eps), max_keep_nan(raw_ls.y, C.log_eps));   // clamp(min): NaN propagates like torch
  const F2 p = mul2(ls, f2(-kLog2e));                             // log2 of exp(-log_scale)     :203
#if BLVM_COMPENSATED_EXP == 1
  constexpr float kLog2eLo = 1.925963033500011e-08f;
  const F2 p_lo = fma2(ls, f2(-kLog2eLo), fma2(ls, f2(-kLog2e), mul2(p, f2(-1.f))));
  const F2 e0 = ex2_2(p);
  const F2 inv = fma2(e0, mul2(p_lo, f2(kLn2)), e0);
#else
  const F2 inv = ex2_2(p);
#endif
  const F2 m = mul2(inv, fma2(mu, f2(-1.f), y));                  // mid_in                     :202,219
  const F2 nu2 = mul2(mul2(inv, inv), f2(-C.h * C.h));            // -u^2, u = h / s (the sign rides on the constant)
  const F2 am = abs2(m);
  const F2 t = mul2(am, f2(-kLog2e));
  const F2 E = ex2_2(t);                                          // exp(-|m|)
  const F2 p1 = add2(E, f2(1.f));                                 // 1 + E
  const F2 r = rcp_2(p1);                                         // 1/(1+E)
  const F2 l = lg2_2(p1);                                         // log2(1+E)
  const F2 common2 = fma2(l, f2(-2.f), add2(t, p));               // [m - ls - 2 softplus(m)] log2(e)   :220
  const F2 lp_fb2 = add2(common2, f2(-C.log_half_bins2));         // second arm of :221-223
  const F2 Er = mul2(E, r);                                       // E / (1+E)
  const F2 er2 = mul2(Er, r);                                     // E / (1+E)^2
  const F2 neps = mul2(er2, nu2);                                 // -E w / (1+E)^2,  w = u^2 + O(u^4)
  // first arm minus second arm:  [log(2h) + log(nb/2) + u^2/6 - eps] log2(e)
  const F2 dlt2 = fma2(neps, f2(kLog2e), fma2(nu2, f2(-kInvLn2Sixth), f2(-C.neg_log_ratio2)));
  const F2 lp_d2 = add2(dlt2, lp_fb2);                            // log2 cdf_delta, first arm of :221-223
  const float thr2 = C.log_delta_thresh2;
  // [cdf_delta > 1e-5] as a 1.0 / 0.0 float: one FSET.BF per lane.  (A -1 / 0 blend compiles to FSETP + SEL + I2FP: the compiler
  // canonicalises any select of -1.0f / 0 into a sign-extended predicate converted to float, and I2FP competes with MUFU.)
  const F2 sel = f2(lp_d2.x > thr2 ? 1.f : 0.f, lp_d2.y > thr2 ? 1.f : 0.f);
  lp2 = fma2(sel, dlt2, lp_fb2);
  if (GRAD) {
    // tanh(|m|/2) = (1-E)/(1+E); odd series below 1/4 so that the bin-centre gradient does not cancel
    const F2 hx = mul2(am, f2(0.5f)), hx2 = mul2(hx, hx);
    const F2 th_series = mul2(hx, fma2(hx2, fma2(hx2, f2(2.0f / 15.0f), f2(-1.0f / 3.0f)), f2(1.0f)));
    const F2 th_exact = fma2(Er, f2(-2.f), f2(1.f));              // 1 - 2E/(1+E) = (1-E)/(1+E)
    const F2 th = f2(am.x < 0.25f ? th_series.x : th_exact.x, am.y < 0.25f ? th_series.y : th_exact.y);
    const F2 th_sel = fma2(mul2(sel, neps), th, th);              // th / (1 + eps) on the first arm, th on the second
    const F2 c = fma2(mul2(sel, nu2), fma2(er2, f2(-2.0f), f2(1.0f / 3.0f)), f2(-1.0f));   // -(u coth(u) - u cdf_delta)  resp.  -1
    // clamp(min=eps) blocks the gradient strictly below eps, passes at equality
    const F2 gate = f2(raw_ls.x >= C.log_eps ? 1.f : 0.f, raw_ls.y >= C.log_eps ? 1.f : 0.f);
    const F2 it = mul2(inv, th_sel);
    dmu = f2(copysignf(it.x, m.x), copysignf(it.y, m.y));         // -inv * d lp/d m
    dls = mul2(fma2(am, th_sel, c), gate);                        // -(m d/dm + u d/du)  resp.  -m d/dm - 1
  }
}

// y in the first / last bin (clipped samples): log sigmoid(a) resp. log(1 - sigmoid(b))   :213,216,226-227
template <bool GRAD>
BLVM_HD void dl_edge(float y, int edge, float mu, float raw_ls, const DmolConsts& C, float& lp, float& dmu, float& dls) {
  const float ls = (raw_ls < C.log_eps) ? C.log_eps : raw_ls;
  const float inv = stable_exp(-ls);
  const float c = y - mu;
  if (edge == kEdgeLower) {
    const float a = inv * (c + C.h);                                 // plus_in :206
    const float e = fast_exp(-fabsf(a));
    lp = fminf(a, 0.f) - fast_log(1.f + e);                          // a - softplus(a)
    if (GRAD) {
      const float da = ((a >= 0.f) ? e : 1.f) * fast_rcp(1.f + e);   // 1 - sigmoid(a)
      dmu = -inv * da;
      dls = -a * da;
    }
  } else {
    const float b = inv * (c - C.h);                                 // minus_in :208
    const float e = fast_exp(-fabsf(b));
    lp = -fmaxf(b, 0.f) - fast_log(1.f + e);                         // -softplus(b)
    if (GRAD) {
      const float db = -((b >= 0.f) ? 1.f : e) * fast_rcp(1.f + e);  // -sigmoid(b)
      dmu = -inv * db;
      dls = -b * db;
    }
  }
  if (GRAD) {
    if (raw_ls < C.log_eps) dls = 0.f;
  }
}

template <bool GRAD>
BLVM_HD void dl_component(float y, int edge, float mu, float raw_ls, const DmolConsts& C, float& lp, float& dmu,
                          float& dls) {
  if (edge != kEdgeNone) dl_edge<GRAD>(y, edge, mu, raw_ls, C, lp, dmu, dls);
  else dl_mid<GRAD, kUGeneral>(y, mu, raw_ls, C, lp, dmu, dls);
}

// 1/x to ~1 ulp: MUFU.RCP + one Newton step (2 FMA) instead of the ~10-instruction IEEE division sequence
BLVM_HD float rcp_nr(float x) {
  const float r = fast_rcp(x);
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// ---- Gaussian (mixture) likelihood: the sibling of the DMoL behind the same boundary --------------
//   gaussian_ll  blvm/utils/log_likelihoods.py:17-39:  -(y-mu)^2/(2 sd^2) - log sd - 0.5 log(2 pi)
// FROM_RAW: the third parameter is the Linear output; sd = softplus_beta(p) + sd_add (DiagonalGaussianMixtureDense.forward,
// distributions.py:198-203) is applied here and the chain rule d sd/d p = sigmoid(beta p) folded into the gradient.
// Otherwise it is sd itself; with sd_floor > 0 (the functional's `epsilon`) sd is clamped under no_grad, which in the
// reference detaches it: no gradient reaches sd at all in that case (kept).
constexpr float kHalfLog2Pi = 0.9189385332046727f;
template <bool GRAD, bool FROM_RAW>
BLVM_HD void gauss_component(float y, float mu, float p, const DmolConsts& C, float& lp, float& dmu, float& dp) {
  float sd, dsd_dp = 1.f;
  if (FROM_RAW) {
    const float bx = C.sp_beta * p;
    const float e = fast_exp(-fabsf(bx));                    // softplus(bx) = max(bx, 0) + log1p(exp(-|bx|)), threshold 20
    // log1p(e) on e in (0, 1] as e * P9(e) (minimax fit of log1p(e)/e, 1.3e-7 relative): sd ~ sd_add + exp(bx)/beta must keep
    // RELATIVE accuracy when e is small, which lg2(1 + e) cannot give (1 + e rounds) and a short series only gives below ~1/32
    float q = -0.003214032156392932f;
    q = fmaf(q, e, 0.019649142399430275f);
    q = fmaf(q, e, -0.05643497034907341f);
    q = fmaf(q, e, 0.10533220320940018f);
    q = fmaf(q, e, -0.15251445770263672f);
    q = fmaf(q, e, 0.19651488959789276f);
    q = fmaf(q, e, -0.24947808682918549f);
    q = fmaf(q, e, 0.3332909941673279f);
    q = fmaf(q, e, -0.4999985098838806f);
    const float l1p = e * fmaf(q, e, 1.0f);
    const float sp = (bx > 20.f) ? bx : fmaxf(bx, 0.f) + l1p;
    sd = fmaf(sp, C.sp_inv_beta, C.sd_add);
    if (GRAD) dsd_dp = (bx > 20.f) ? 1.f : ((bx >= 0.f) ? 1.f : e) * fast_rcp(1.f + e);   // sigmoid(bx)
  } else {
    sd = p;
    if (C.sd_floor > 0.f) {
      sd = (p < C.sd_floor) ? C.sd_floor : p;
      dsd_dp = 0.f;
    }
  }
  const float inv = rcp_nr(sd);
  const float z = (y - mu) * inv;
  lp = fmaf(-0.5f * z, z, -kLn2 * fast_lg2(sd) - kHalfLog2Pi);
  if (GRAD) {
    dmu = z * inv;                                           // (y - mu)/sd^2
    dp = fmaf(z, z, -1.f) * inv * dsd_dp;                    // ((y-mu)^2/sd^3 - 1/sd) d sd/d p
  }
}

// ---- Gaussian KL, std-dev parametrisation (variational.py:67-70), cancellation-free around q == p ----------------
//   kl = log sd_p - log sd_q + (sd_q^2 + (mu_q-mu_p)^2) / (2 sd_p^2) - 1/2
//      = -log1p(rho-1) + ((rho-1)(rho+1) + z^2)/2,   rho = sd_q/sd_p,  z = (mu_q-mu_p)/sd_p
struct KlTerms {
  float kl, z, rho_m1, rho_p1, inv_sp, q;  // q = (rho-1)(rho+1) + z^2
};
// log(rho), rho = sd_q/sd_p, as 2 atanh(s), s = (sd_q - sd_p)/(sd_q + sd_p): exact near rho = 1, where the KL's
// q/2 - log(rho) cancels; odd series to s^11 for |s| < 0.2 (truncation 3e-10 relative), MUFU.LG2 of the quotient elsewhere
// (|log rho| > 0.4).  d = sd_q - sd_p, ssum = sd_q + sd_p, inv_sp = 1/sd_p are passed in (the callers have them).
BLVM_HD float log_sd_ratio(float sd_q, float sd_p, float d, float ssum, float inv_sp) {
  const float s = d * rcp_nr(ssum), s2 = s * s;
  float p = 1.0f / 11.0f;
  p = fmaf(p, s2, 1.0f / 9.0f);
  p = fmaf(p, s2, 1.0f / 7.0f);
  p = fmaf(p, s2, 1.0f / 5.0f);
  p = fmaf(p, s2, 1.0f / 3.0f);
  p = fmaf(p, s2, 1.0f);
  const float log_series = 2.0f * s * p, log_mufu = kLn2 * fast_lg2(sd_q * inv_sp);   // both evaluated: a select, no branch
  (void)sd_p;
  return (fabsf(s) < 0.2f) ? log_series : log_mufu;
}
BLVM_HD KlTerms kl_gaussian_terms(float mu_q, float sd_q, float mu_p, float sd_p) {
  KlTerms t;
  t.inv_sp = rcp_nr(sd_p);
  t.z = (mu_q - mu_p) * t.inv_sp;
  const float d = sd_q - sd_p, ssum = sd_q + sd_p;
  t.rho_m1 = d * t.inv_sp;
  t.rho_p1 = ssum * t.inv_sp;
  t.q = fmaf(t.rho_m1, t.rho_p1, t.z * t.z);
  t.kl = 0.5f * t.q - log_sd_ratio(sd_q, sd_p, d, ssum, t.inv_sp);
  return t;
}
// d kl / d (mu_q, sd_q, mu_p, sd_p), each multiplied by g.
BLVM_HD void kl_gaussian_grads(const KlTerms& t, float sd_q, float g, float& g_mu_q, float& g_sd_q, float& g_mu_p,
                               float& g_sd_p) {
  g_mu_q = g * t.z * t.inv_sp;                       // (mu_q-mu_p)/sd_p^2
  g_mu_p = -g_mu_q;
  g_sd_q = g * (t.rho_m1 * t.rho_p1) * rcp_nr(sd_q);  // -1/sd_q + sd_q/sd_p^2
  g_sd_p = -g * t.q * t.inv_sp;                       // 1/sd_p - (sd_q^2 + d^2)/sd_p^3
}
// Monte-Carlo KL of one latent element (variational.py:73-83, bottom-up STCN stcn.py:288): log q(z) - log p(z) with both
// Gaussian log-densities of log_likelihoods.py:17-39 (epsilon = 0); the 0.5 log(2 pi) terms cancel.
//   kl = 0.5 (ap - aq)(ap + aq) - log(sd_q/sd_p),   aq = (z - mu_q)/sd_q,  ap = (z - mu_p)/sd_p
struct KlMcTerms {
  float kl, aq, ap, inv_sq, inv_sp;
};
BLVM_HD KlMcTerms kl_mc_terms(float z, float mu_q, float sd_q, float mu_p, float sd_p) {
  KlMcTerms t;
  t.inv_sq = rcp_nr(sd_q);
  t.inv_sp = rcp_nr(sd_p);
  t.aq = (z - mu_q) * t.inv_sq;
  t.ap = (z - mu_p) * t.inv_sp;
  t.kl = 0.5f * (t.ap - t.aq) * (t.ap + t.aq) - log_sd_ratio(sd_q, sd_p, sd_q - sd_p, sd_q + sd_p, t.inv_sp);
  return t;
}
// d kl / d (mu_q, sd_q, mu_p, sd_p, z), each multiplied by g
BLVM_HD void kl_mc_grads(const KlMcTerms& t, float g, float& g_mu_q, float& g_sd_q, float& g_mu_p, float& g_sd_p, float& g_z) {
  const float a = g * t.aq * t.inv_sq, b = g * t.ap * t.inv_sp;
  g_mu_q = a;                                          // d/d mu_q of -aq^2/2
  g_mu_p = -b;
  g_sd_q = g * fmaf(t.aq, t.aq, -1.0f) * t.inv_sq;     // (aq^2 - 1)/sd_q
  g_sd_p = g * fmaf(-t.ap, t.ap, 1.0f) * t.inv_sp;     // (1 - ap^2)/sd_p
  g_z = b - a;
}
// torch.maximum(kl, c) gradient routing: 1 above, 1/2 at the exact tie, 0 below.
BLVM_HD float free_nats_gate(float kl, float c, bool enabled) {
  if (!enabled) return 1.f;
  return (kl > c) ? 1.f : ((kl == c) ? 0.5f : 0.f);
}

}  // namespace blvm

namespace blvm {

// One waveform sample of the K-component mixture (D = 1), everything in registers.
//   in : r[0:K) logits, r[K:2K) locs, r[2K:3K) raw log-scales            (distributions.py:383-385 layout)
//   out: returns log p(y) = logsumexp_k(lp_k + log_softmax(logits)_k)    (log_likelihoods.py:229-231)
//        if GRAD, r[] is overwritten with g * d log p / d r[]
// d/d logit_k = resp_k - softmax_k,  d/d loc_k = resp_k * dlp_k/dloc,  d/d ls_k = resp_k * dlp_k/dls.
enum : int { kLikDmol = 0, kLikGmmRaw = 1, kLikGmmSd = 2 };

template <int K, bool GRAD, int UMODE = kUGeneral, int LIK = kLikDmol>
BLVM_HD float dmol_sample(float y, float (&r)[3 * K], float g, const DmolConsts& C) {
  const int edge = (LIK == kLikDmol) ? dmol_edge(y, C) : kEdgeNone;
  float v[K];   // lp_k + log_softmax numerator, in LOG2 units (the exp / log of the mixture algebra are MUFU.EX2 / LG2)
  // Logits centred on their maximum FIRST (like log_softmax in the reference, log_likelihoods.py:230) and scaled to log2
  // units in the same FMA: w_k = (logit_k - max logit) log2(e).  The rounding of max*log2(e) is a shift common to all
  // components, which cancels in logsumexp_k(lp_k + w_k) - logsumexp_k(w_k): large |logits| (~100) cost no absolute
  // precision in a log-prob that is itself close to 0 (wide bins).
  if constexpr (K > 1) {
    float m2 = r[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m2 = fmaxf(m2, r[k]);
    const float nm2 = -m2 * kLog2e;
    if constexpr (K % 2 == 0) {
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        const F2 w = fma2(f2(r[k], r[k + 1]), f2(kLog2e), f2(nm2));
        r[k] = w.x; r[k + 1] = w.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) r[k] = fmaf(r[k], kLog2e, nm2);
    }
  }
  if (LIK != kLikDmol) {   // Gaussian mixture: same layout and mixture algebra, different component density
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float lp, dmu = 0.f, dp = 0.f;
      gauss_component<GRAD, LIK == kLikGmmRaw>(y, r[K + k], r[2 * K + k], C, lp, dmu, dp);
      v[k] = (K == 1) ? lp * kLog2e : fmaf(lp, kLog2e, r[k]);
      if (GRAD) {
        r[K + k] = dmu;
        r[2 * K + k] = dp;
      }
    }
  } else if (edge == kEdgeNone) {   // hoisted out of the component loop: the hot loop below has no edge test
    constexpr int KP = (UMODE == kUTiny) ? (K / 2) * 2 : 0;   // components evaluated two at a time (packed fp32x2)
#pragma unroll
    for (int k = 0; k < KP; k += 2) {
      F2 lp2, dmu = f2(0.f), dls = f2(0.f);
      dl_mid_pair_tiny<GRAD>(f2(y), f2(r[K + k], r[K + k + 1]), f2(r[2 * K + k], r[2 * K + k + 1]), C, lp2, dmu, dls);
      const F2 vv = add2(lp2, f2(r[k], r[k + 1]));
      v[k] = vv.x;
      v[k + 1] = vv.y;
      if (GRAD) {
        r[K + k] = dmu.x;
        r[K + k + 1] = dmu.y;
        r[2 * K + k] = dls.x;
        r[2 * K + k + 1] = dls.y;
      }
    }
#pragma unroll
    for (int k = KP; k < K; ++k) {
      float lp2, dmu = 0.f, dls = 0.f;
      if constexpr (UMODE == kUTiny) {   // leftover component (odd K): the pair function with its lane duplicated
        F2 l2, dmu2 = f2(0.f), dls2 = f2(0.f);
        dl_mid_pair_tiny<GRAD>(f2(y), f2(r[K + k]), f2(r[2 * K + k]), C, l2, dmu2, dls2);
        lp2 = l2.x; dmu = dmu2.x; dls = dls2.x;
      } else {
        float lp;
        dl_mid<GRAD, UMODE>(y, r[K + k], r[2 * K + k], C, lp, dmu, dls);
        lp2 = lp * kLog2e;
      }
      v[k] = (K == 1) ? lp2 : lp2 + r[k];
      if (GRAD) {
        r[K + k] = dmu;
        r[2 * K + k] = dls;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) {  // (unrolled too: r[] must stay in registers)
      float lp, dmu = 0.f, dls = 0.f;
      dl_edge<GRAD>(y, edge, r[K + k], r[2 * K + k], C, lp, dmu, dls);
      v[k] = (K == 1) ? lp * kLog2e : fmaf(lp, kLog2e, r[k]);
      if (GRAD) {
        r[K + k] = dmu;
        r[2 * K + k] = dls;
      }
    }
  }
  if constexpr (K == 1) {
    // one component: log_softmax(logit) = 0, responsibility = softmax = 1, d/d logit = 0 (NaN/inf logits still propagate)
    const float z = r[0] - r[0];   // 0, or NaN for a non-finite logit
    const float L1 = fmaf(v[0], kLn2, z);
    if (GRAD) {
      r[0] = g * z;
      r[1] *= g;
      r[2] *= g;
    }
    return L1;
  }
  float m1 = v[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m1 = fmaxf(m1, v[k]);
  float s1 = 0.f, s2 = 0.f;
  if constexpr (K % 2 == 0) {   // two components per packed instruction (even / odd partial sums, added at the end)
    F2 s1p = f2(0.f), s2p = f2(0.f);
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      const F2 ev = ex2_2(add2(f2(v[k], v[k + 1]), f2(-m1)));               // exp(v_k - max)
      const F2 er = ex2_2(f2(r[k], r[k + 1]));                              // exp(w_k)
      v[k] = ev.x; v[k + 1] = ev.y;
      r[k] = er.x; r[k + 1] = er.y;
      s1p = add2(s1p, ev);
      s2p = add2(s2p, er);
    }
    s1 = s1p.x + s1p.y;
    s2 = s2p.x + s2p.y;
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      v[k] = fast_ex2(v[k] - m1);
      r[k] = fast_ex2(r[k]);
      s1 += v[k];
      s2 += r[k];
    }
  }
  const float L = kLn2 * (m1 + (fast_lg2(s1) - fast_lg2(s2)));
  if (GRAD) {
    const float g1 = g * fast_rcp(s1), g2 = g * fast_rcp(s2);
    if constexpr (K % 2 == 0) {
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        const F2 gr = mul2(f2(g1), f2(v[k], v[k + 1]));                     // g * responsibility_k
        const F2 gl = fma2(f2(r[k], r[k + 1]), f2(-g2), gr);                // g * (resp_k - softmax_k)
        const F2 gm = mul2(f2(r[K + k], r[K + k + 1]), gr);
        const F2 gs = mul2(f2(r[2 * K + k], r[2 * K + k + 1]), gr);
        r[k] = gl.x; r[k + 1] = gl.y;
        r[K + k] = gm.x; r[K + k + 1] = gm.y;
        r[2 * K + k] = gs.x; r[2 * K + k + 1] = gs.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float gr = g1 * v[k];            // g * responsibility_k
        r[k] = gr - g2 * r[k];                 // g * (resp_k - softmax_k)
        r[K + k] *= gr;
        r[2 * K + k] *= gr;
      }
    }
  }
  return L;
}

// ---- linear-domain evaluation (16-bit kernels are bound by the MUFU unit and by issue slots) ------------------------------------------
// dmol_sample carries log-probabilities: per (sample, component) that costs log2(1+E) for log_pdf_mid and exp2(lp + w - max) for the
// responsibility, 2 of its 6.4 transcendentals, plus the max chain.  But the likelihood of a component is a PRODUCT of factors the
// evaluation has anyway:
//   first arm   cdf_delta            = 2 E sinh(u) / D = [exp(-ls) E/(1+E)^2] * 2h (1 + u^2/6 - eps)      (O(u^4) ~ 1e-8 dropped)
//   second arm  exp(log_pdf_mid)/(nb/2) = [exp(-ls) E/(1+E)^2] * 2/nb                                         log_likelihoods.py:219-223
// so  p(y) = sum_k softmax_k lik_k  is formed directly: 4.4 transcendentals per component, no max chain, and the selection
// `cdf_delta > 1e-5` compares the quantity the reference compares.  What the product cannot do is represent a sample that is more
// than ~55 scales away from EVERY component (all E underflow; the log-domain value is still finite): dmol_sample_lin then
// returns false -- also for NaN / inf parameters and for the two edge bins -- and the caller re-reads the row and evaluates it with
// dmol_sample.  Gradient formulas are those of dl_mid_pair_tiny.
#ifndef BLVM_LINEAR_DOMAIN
#define BLVM_LINEAR_DOMAIN 1
#endif
constexpr float kLinMinSum = 1e-25f;   // below this the mixture sum has lost components to underflow (FLT_MIN 1.2e-38)
constexpr float kLinMaxSum = 1e30f;

template <bool GRAD>
BLVM_HD void dl_mid_pair_lin(F2 y, F2 mu, F2 raw_ls, const DmolConsts& C, F2& lik, F2& dmu, F2& dls) {
  const F2 ls = f2(max_keep_nan(raw_ls.x, C.log_eps), max_keep_nan(raw_ls.y, C.log_eps));   // clamp(min): NaN propagates like torch
  const F2 inv = ex2_2(mul2(ls, f2(-kLog2e)));                    // exp(-log_scale)            :203
  const F2 m = mul2(inv, fma2(mu, f2(-1.f), y));                  // mid_in                     :202,219
  const F2 nu2 = mul2(mul2(inv, inv), f2(C.neg_h2));              // -u^2, u = h / s
  const F2 am = abs2(m);
  const F2 E = ex2_2(mul2(am, f2(-kLog2e)));                      // exp(-|m|)
  const F2 r = rcp_2(add2(E, f2(1.f)));                           // 1/(1+E)
  const F2 Er = mul2(E, r);                                       // E / (1+E)
  const F2 er2 = mul2(Er, r);                                     // E / (1+E)^2
  const F2 base = mul2(inv, er2);                                 // exp(log_pdf_mid)
  const F2 neps = mul2(er2, nu2);                                 // -E w / (1+E)^2,  w = u^2 + O(u^4)
  const F2 fd = fma2(neps, f2(C.two_h), fma2(nu2, f2(C.neg_two_h_sixth), f2(C.fd0)));   // 2h (1 + u^2/6 - eps) - 2/nb
  const F2 la = mul2(base, add2(fd, f2(C.fb)));                   // cdf_delta, first arm of :221-223
  const F2 sel = f2(la.x > kDeltaThresh ? 1.f : 0.f, la.y > kDeltaThresh ? 1.f : 0.f);   // one FSET.BF per lane
  lik = mul2(base, fma2(sel, fd, f2(C.fb)));
  if (GRAD) {
    const F2 hx = mul2(am, f2(0.5f)), hx2 = mul2(hx, hx);
    const F2 th_series = mul2(hx, fma2(hx2, fma2(hx2, f2(2.0f / 15.0f), f2(-1.0f / 3.0f)), f2(1.0f)));
    const F2 th_exact = fma2(Er, f2(-2.f), f2(1.f));              // (1-E)/(1+E) = tanh(|m|/2)
    const F2 th = f2(am.x < 0.25f ? th_series.x : th_exact.x, am.y < 0.25f ? th_series.y : th_exact.y);
    const F2 th_sel = fma2(mul2(sel, neps), th, th);              // th / (1 + eps) on the first arm, th on the second
    const F2 c = fma2(mul2(sel, nu2), fma2(er2, f2(-2.0f), f2(1.0f / 3.0f)), f2(-1.0f));   // -(u coth(u) - u cdf_delta)  resp.  -1
    const F2 gate = f2(raw_ls.x >= C.log_eps ? 1.f : 0.f, raw_ls.y >= C.log_eps ? 1.f : 0.f);
    const F2 it = mul2(inv, th_sel);
    dmu = f2(copysignf(it.x, m.x), copysignf(it.y, m.y));         // -inv * d lp/d m
    dls = mul2(fma2(am, th_sel, c), gate);
  }
}

// One sample, K >= 2 components, 16-bit-audio mode (kUTiny).  Same in / out convention as dmol_sample; returns false -- with r[]
// clobbered -- when the sample has to be evaluated in the log domain instead (see above).
// `in` hands out the sample's parameters: RowIn reads them from r[] itself (every element is read before its slot is overwritten), the
// 16-bit tile kernel passes its packed row so that each pair is converted where it is used (dmol_kernels.cuh: PackedRowIn).
template <int K>
struct RowIn {
  const float* r;
  BLVM_HD float logit(int k) const { return r[k]; }
  BLVM_HD F2 logit2(int k) const { return f2(r[k], r[k + 1]); }
  BLVM_HD F2 mu2(int k) const { return f2(r[K + k], r[K + k + 1]); }
  BLVM_HD F2 ls2(int k) const { return f2(r[2 * K + k], r[2 * K + k + 1]); }
  BLVM_HD float mu(int k) const { return r[K + k]; }
  BLVM_HD float ls(int k) const { return r[2 * K + k]; }
};

template <int K, bool GRAD, typename In>
BLVM_HD bool dmol_sample_lin_in(float y, const In& in, float (&r)[3 * K], float g, const DmolConsts& C, float& L) {
  static_assert(K >= 2, "one component has no mixture algebra to save");
  if (dmol_edge(y, C) != kEdgeNone) return false;
  float m2 = in.logit(0);
#pragma unroll
  for (int k = 1; k < K; ++k) m2 = fmaxf(m2, in.logit(k));
  const float nm2 = -m2 * kLog2e;
  float n[K];   // softmax numerator times component likelihood
  F2 s1p = f2(0.f), s2p = f2(0.f);
  constexpr int KP = (K / 2) * 2;
#pragma unroll
  for (int k = 0; k < KP; k += 2) {
    const F2 ew = ex2_2(fma2(in.logit2(k), f2(kLog2e), f2(nm2)));   // exp(logit_k - max logit)
    F2 lik, dmu = f2(0.f), dls = f2(0.f);
    dl_mid_pair_lin<GRAD>(f2(y), in.mu2(k), in.ls2(k), C, lik, dmu, dls);
    const F2 nk = mul2(ew, lik);
    s1p = fma2(ew, lik, s1p);   // explicitly fused: ptxas contracts a packed mul + add into FFMA2 in some instantiations and not in others
    s2p = add2(s2p, ew);
    r[k] = ew.x; r[k + 1] = ew.y;
    n[k] = nk.x; n[k + 1] = nk.y;
    if (GRAD) {
      r[K + k] = dmu.x; r[K + k + 1] = dmu.y;
      r[2 * K + k] = dls.x; r[2 * K + k + 1] = dls.y;
    }
  }
  float s1 = add_rn(s1p.x, s1p.y), s2 = add_rn(s2p.x, s2p.y);
  if constexpr (KP < K) {   // leftover component (odd K): the pair function with its lane duplicated
    constexpr int k = K - 1;
    const float ew = fast_ex2(fmaf(in.logit(k), kLog2e, nm2));
    F2 lik, dmu = f2(0.f), dls = f2(0.f);
    dl_mid_pair_lin<GRAD>(f2(y), f2(in.mu(k)), f2(in.ls(k)), C, lik, dmu, dls);
    n[k] = mul_rn(ew, lik.x);   // explicit roundings: every kernel that inlines this must produce the same bits
    s1 = fmaf(ew, lik.x, s1);
    s2 = add_rn(s2, ew);
    r[k] = ew;
    if (GRAD) {
      r[K + k] = dmu.x;
      r[2 * K + k] = dls.x;
    }
  }
  if (!(s1 > kLinMinSum && s1 < kLinMaxSum && s2 < kLinMaxSum)) return false;   // underflow, inf or NaN: log domain
  L = mul_rn(kLn2, add_rn(fast_lg2(s1), -fast_lg2(s2)));
  if (GRAD) {
    const float g1 = mul_rn(g, fast_rcp(s1)), g2 = mul_rn(g, fast_rcp(s2));
    if constexpr (K % 2 == 0) {
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        const F2 gr = mul2(f2(g1), f2(n[k], n[k + 1]));                     // g * responsibility_k
        const F2 gl = fma2(f2(-g2), f2(r[k], r[k + 1]), gr);                // g * (resp_k - softmax_k)
        const F2 gm = mul2(f2(r[K + k], r[K + k + 1]), gr);
        const F2 gs = mul2(f2(r[2 * K + k], r[2 * K + k + 1]), gr);
        r[k] = gl.x; r[k + 1] = gl.y;
        r[K + k] = gm.x; r[K + k + 1] = gm.y;
        r[2 * K + k] = gs.x; r[2 * K + k + 1] = gs.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float gr = mul_rn(g1, n[k]);
        r[k] = fmaf(-g2, r[k], gr);
        r[K + k] = mul_rn(r[K + k], gr);
        r[2 * K + k] = mul_rn(r[2 * K + k], gr);
      }
    }
  }
  return true;
}

template <int K, bool GRAD>
BLVM_HD bool dmol_sample_lin(float y, float (&r)[3 * K], float g, const DmolConsts& C, float& L) {
  return dmol_sample_lin_in<K, GRAD>(y, RowIn<K>{r}, r, g, C, L);
}

// The evaluation the kernels call: linear domain where it applies, log domain otherwise.  `reload(r)` re-reads the sample's parameter
// row (shared memory / tensor memory / the caller's copy) after a failed linear-domain attempt.
template <int K, bool GRAD, int UMODE, int LIK>
struct DmolEvalTraits {
  // K > 12 runs at its register cap: the second copy of the sample body (the log-domain fallback) spills there (K = 16 fp32 239 -> 328 us)
  static constexpr bool kLin = BLVM_LINEAR_DOMAIN && K >= 2 && K <= 12 && UMODE == kUTiny && LIK == kLikDmol;
};
// The evaluation the kernels call.  LIN (chosen by the kernel, see DmolEvalTraits) tries the linear domain first; a sample it
// cannot represent is re-read through `reload(r)` (shared memory / the caller's copy) and evaluated by dmol_sample.  Measured
// alternatives for the fallback, all slower on the hot path: a __noinline__ call (calling-convention spills around the call site: bf16
// K = 10 84.5 -> 90.8 us) and the rolled generic evaluation on a local-memory copy (slow whenever it is taken: K = 2 44 -> 80 us on the
// bench's synthetic parameters, where 2 % of the samples are far from both components).
template <int K, bool GRAD, int UMODE, int LIK, bool LIN, typename Reload>
BLVM_HD float dmol_eval(float y, float (&r)[3 * K], float g, const DmolConsts& C, Reload&& reload) {
  if constexpr (LIN && DmolEvalTraits<K, GRAD, UMODE, LIK>::kLin) {
    float L;
    if (dmol_sample_lin<K, GRAD>(y, r, g, C, L)) return L;
    reload(r);
  }
  return dmol_sample<K, GRAD, UMODE, LIK>(y, r, g, C);
}

// K == 1 (a single discretized logistic with a dead logit column), 16-bit-audio mode: two SAMPLES of one thread evaluated
// together with packed instructions.  Either sample in an edge bin sends both through dmol_sample (rare: clipped audio).
// Bit-identical to calling dmol_sample<1, GRAD, kUTiny> on each sample.
template <bool GRAD>
BLVM_HD void dmol_k1_two_samples(float ya, float yb, float (&ra)[3], float (&rb)[3], float ga, float gb, const DmolConsts& C,
                                 float& La, float& Lb) {
  if (dmol_edge(ya, C) == kEdgeNone && dmol_edge(yb, C) == kEdgeNone) {
    F2 lp, dmu = f2(0.f), dls = f2(0.f);
    dl_mid_pair_tiny<GRAD>(f2(ya, yb), f2(ra[1], rb[1]), f2(ra[2], rb[2]), C, lp, dmu, dls);
    const float za = ra[0] - ra[0], zb = rb[0] - rb[0];   // 0, or NaN for a non-finite logit (log_softmax of one logit)
    La = fmaf(lp.x, kLn2, za);   // the pair function returns log2 units
    Lb = fmaf(lp.y, kLn2, zb);
    if (GRAD) {
      const F2 g2 = f2(ga, gb);
      const F2 gm = mul2(dmu, g2), gs = mul2(dls, g2);
      ra[0] = ga * za; rb[0] = gb * zb;
      ra[1] = gm.x; rb[1] = gm.y;
      ra[2] = gs.x; rb[2] = gs.y;
    }
  } else {
    La = dmol_sample<1, GRAD, kUTiny>(ya, ra, ga, C);
    Lb = dmol_sample<1, GRAD, kUTiny>(yb, rb, gb, C);
  }
}

// Generic (runtime K, D >= 1) sample, two passes with recomputation; `p` is the sample's K(2D+1) parameters laid
// out [logits K | d=0: locs K, log_scales K | d=1: ...] (distributions.py:383-385), `yv` its D targets, `o` the
// gradient row (may alias nothing; written only if GRAD).  Used for shapes the register kernel is not instantiated for.
// component dispatch for the generic (runtime K, D) path
template <bool GRAD>
BLVM_HD void any_component(int lik, float y, float mu, float p, const DmolConsts& C, float& lp, float& dmu, float& dp) {
  if (lik == kLikDmol) dl_component<GRAD>(y, dmol_edge(y, C), mu, p, C, lp, dmu, dp);
  else if (lik == kLikGmmRaw) gauss_component<GRAD, true>(y, mu, p, C, lp, dmu, dp);
  else gauss_component<GRAD, false>(y, mu, p, C, lp, dmu, dp);
}

template <bool GRAD>
BLVM_HD float dmol_sample_generic(const float* yv, const float* p, int K, int D, float g, const DmolConsts& C, float* o,
                                  int lik = kLikDmol) {
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int k = 0; k < K; ++k) m2 = fmaxf(m2, p[k]);   // logits centred on their maximum first (see dmol_sample)
  for (int k = 0; k < K; ++k) {
    float lpk = 0.f;
    for (int d = 0; d < D; ++d) {
      float lp, a, b;
      any_component<false>(lik, yv[d], p[K + d * 2 * K + k], p[K + d * 2 * K + K + k], C, lp, a, b);
      lpk += lp;
    }
    const float vk = lpk + (p[k] - m2);
    m1 = fmaxf(m1, vk);
    if (vk != vk) m1 = vk;  // NaN propagates
  }
  float s1 = 0.f, s2 = 0.f;
  for (int k = 0; k < K; ++k) {
    float lpk = 0.f;
    for (int d = 0; d < D; ++d) {
      float lp, a, b;
      any_component<false>(lik, yv[d], p[K + d * 2 * K + k], p[K + d * 2 * K + K + k], C, lp, a, b);
      lpk += lp;
    }
    s1 += fast_ex2((lpk + (p[k] - m2) - m1) * kLog2e);
    s2 += fast_ex2((p[k] - m2) * kLog2e);
  }
  const float L = m1 + kLn2 * (fast_lg2(s1) - fast_lg2(s2));
  if (GRAD) {
    const float g1 = g * fast_rcp(s1), g2 = g * fast_rcp(s2);
    for (int k = 0; k < K; ++k) {
      float lpk = 0.f;
      for (int d = 0; d < D; ++d) {
        float lp, dmu, dls;
        any_component<true>(lik, yv[d], p[K + d * 2 * K + k], p[K + d * 2 * K + K + k], C, lp, dmu, dls);
        lpk += lp;
        o[K + d * 2 * K + k] = dmu;
        o[K + d * 2 * K + K + k] = dls;
      }
      const float gr = g1 * fast_ex2((lpk + (p[k] - m2) - m1) * kLog2e);
      o[k] = gr - g2 * fast_ex2((p[k] - m2) * kLog2e);
      for (int d = 0; d < D; ++d) {
        o[K + d * 2 * K + k] *= gr;
        o[K + d * 2 * K + K + k] *= gr;
      }
    }
  }
  return L;
}

}  // namespace blvm
