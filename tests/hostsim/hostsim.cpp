// Host build of the per-element device math (benchmarking-lvms_b200/csrc/blvm_math.cuh) for CPU-side validation of
// the closed forms against the fp64 golden vectors.  TEST INFRASTRUCTURE: compiled by tests/test_hostsim_math.py
// with g++, loaded only by that test.  The MUFU approximations are replaced by libm (exp2f/log2f, 1/x), so this
// checks the algebra and the branch logic, not the last ulps of the GPU result (tests -m gpu do that).
#include <cmath>
#include <cstdint>
#include "blvm_math.cuh"

using namespace blvm;

static DmolConsts make_consts(int num_bins, float log_eps) {
  DmolConsts C;
  C.h = (float)(1.0 / (num_bins - 1));
  C.log_two_h = (float)std::log(2.0 / (num_bins - 1));
  C.log_delta_thresh = (float)std::log((double)kDeltaThresh);
  C.lo_thresh = (float)(2.0 / num_bins - 1.0);
  C.hi_thresh = (float)(1.0 - 2.0 / num_bins);
  C.log_half_bins = (float)std::log(num_bins / 2.0);
  C.log_eps = log_eps;
  const double log2e = 1.4426950408889634;
  C.log_half_bins2 = (float)(std::log(num_bins / 2.0) * log2e);
  C.neg_log_ratio2 = (float)(-std::log((double)num_bins / (num_bins - 1)) * log2e);
  C.log_delta_thresh2 = (float)(std::log((double)kDeltaThresh) * log2e);
  const double nb = num_bins, two_h = 2.0 / (nb - 1.0);
  C.two_h = (float)two_h;
  C.fb = (float)(2.0 / nb);
  C.fd0 = (float)(two_h - 2.0 / nb);
  C.neg_two_h_sixth = (float)(-two_h / 6.0);
  C.neg_h2 = (float)(-1.0 / ((nb - 1.0) * (nb - 1.0)));
  return C;
}

template <int K>
static void run_fixed(const float* y, const float* raw, const float* gout, int64_t N, const DmolConsts& C, int tiny,
                      float* lp, float* graw) {
  for (int64_t n = 0; n < N; ++n) {
    float r[3 * K];
    for (int i = 0; i < 3 * K; ++i) r[i] = raw[n * 3 * K + i];
    auto reload = [&](float (&rr)[3 * K]) {
      for (int i = 0; i < 3 * K; ++i) rr[i] = raw[n * 3 * K + i];
    };
    // the evaluation the kernels call (linear domain where it applies, log domain otherwise)
    lp[n] = tiny ? dmol_eval<K, true, kUTiny, kLikDmol, true>(y[n], r, gout ? gout[n] : 1.f, C, reload)
                 : dmol_eval<K, true, kUGeneral, kLikDmol, true>(y[n], r, gout ? gout[n] : 1.f, C, reload);
    for (int i = 0; i < 3 * K; ++i) graw[n * 3 * K + i] = r[i];
  }
}

extern "C" int hostsim_dmol(const float* y, const float* raw, const float* gout, int64_t N, int K, int D, int num_bins,
                            float log_eps, int force_generic, float* lp, float* graw) {
  const DmolConsts C = make_consts(num_bins, log_eps);
  // same rule as blvm_b200.cu: the tiny-u specialisation is used iff h * exp(-log_eps) < kTinyU
  const int tiny = (double)C.h * std::exp(-(double)log_eps) < (double)kTinyU;
  if (D == 1 && !force_generic) {
    switch (K) {
      case 1: run_fixed<1>(y, raw, gout, N, C, tiny, lp, graw); return tiny;
      case 2: run_fixed<2>(y, raw, gout, N, C, tiny, lp, graw); return tiny;
      case 5: run_fixed<5>(y, raw, gout, N, C, tiny, lp, graw); return tiny;
      case 10: run_fixed<10>(y, raw, gout, N, C, tiny, lp, graw); return tiny;
      case 30: run_fixed<30>(y, raw, gout, N, C, tiny, lp, graw); return tiny;
      default: break;
    }
  }
  const int P = K * (2 * D + 1);
  for (int64_t n = 0; n < N; ++n)
    lp[n] = dmol_sample_generic<true>(y + n * D, raw + n * P, K, D, gout ? gout[n] : 1.f, C, graw + n * P);
  return 2;
}

extern "C" void hostsim_dl(const float* y, const float* raw, const float* gout, int64_t N, int num_bins, float log_eps,
                           float* lp, float* graw) {
  const DmolConsts C = make_consts(num_bins, log_eps);
  for (int64_t n = 0; n < N; ++n) {
    float dmu, dls;
    dl_component<true>(y[n], dmol_edge(y[n], C), raw[2 * n], raw[2 * n + 1], C, lp[n], dmu, dls);
    const float g = gout ? gout[n] : 1.f;
    graw[2 * n] = g * dmu;
    graw[2 * n + 1] = g * dls;
  }
}

extern "C" void hostsim_kl(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, const float* gout,
                           int64_t n, float min_kl, int fn_enabled, float* kl, float* kl_fn, float* g_mu_q,
                           float* g_sd_q, float* g_mu_p, float* g_sd_p) {
  for (int64_t i = 0; i < n; ++i) {
    KlTerms t = kl_gaussian_terms(mu_q[i], sd_q[i], mu_p[i], sd_p[i]);
    kl[i] = t.kl;
    kl_fn[i] = (fn_enabled && t.kl < min_kl) ? min_kl : t.kl;
    const float g = (gout ? gout[i] : 1.f) * free_nats_gate(t.kl, min_kl, fn_enabled != 0);
    kl_gaussian_grads(t, sd_q[i], g, g_mu_q[i], g_sd_q[i], g_mu_p[i], g_sd_p[i]);
  }
}


// ---- Gaussian mixture (sibling likelihood) ---------------------------------------------------------------------------
template <int K>
static void run_gmm_fixed(const float* y, const float* raw, const float* gout, int64_t N, const DmolConsts& C, int from_raw,
                          float* lp, float* graw) {
  for (int64_t n = 0; n < N; ++n) {
    float r[3 * K];
    for (int i = 0; i < 3 * K; ++i) r[i] = raw[n * 3 * K + i];
    const float g = gout ? gout[n] : 1.f;
    lp[n] = from_raw ? dmol_sample<K, true, kUGeneral, kLikGmmRaw>(y[n], r, g, C)
                     : dmol_sample<K, true, kUGeneral, kLikGmmSd>(y[n], r, g, C);
    for (int i = 0; i < 3 * K; ++i) graw[n * 3 * K + i] = r[i];
  }
}

extern "C" int hostsim_gmm(const float* y, const float* raw, const float* gout, int64_t N, int K, float beta, float sd_add,
                           float sd_floor, int from_raw, int force_generic, float* lp, float* graw) {
  DmolConsts C = make_consts(256, -7.0f);
  C.sp_beta = beta; C.sp_inv_beta = 1.0f / beta; C.sd_add = sd_add; C.sd_floor = sd_floor;
  if (!force_generic) {
    switch (K) {
      case 1: run_gmm_fixed<1>(y, raw, gout, N, C, from_raw, lp, graw); return 0;
      case 5: run_gmm_fixed<5>(y, raw, gout, N, C, from_raw, lp, graw); return 0;
      case 10: run_gmm_fixed<10>(y, raw, gout, N, C, from_raw, lp, graw); return 0;
      case 20: run_gmm_fixed<20>(y, raw, gout, N, C, from_raw, lp, graw); return 0;
      default: break;
    }
  }
  const int P = 3 * K;
  for (int64_t n = 0; n < N; ++n)
    lp[n] = dmol_sample_generic<true>(y + n, raw + n * P, K, 1, gout ? gout[n] : 1.f, C, graw + n * P,
                                      from_raw ? kLikGmmRaw : kLikGmmSd);
  return 2;
}

extern "C" void hostsim_gauss(const float* y, const float* mu, const float* sd, const float* gout, int64_t n, float sd_floor,
                              float* lp, float* g_mu, float* g_sd) {
  DmolConsts C = make_consts(256, -7.0f);
  C.sd_floor = sd_floor;
  for (int64_t i = 0; i < n; ++i) {
    float dmu, dsd;
    gauss_component<true, false>(y[i], mu[i], sd[i], C, lp[i], dmu, dsd);
    const float g = gout ? gout[i] : 1.f;
    g_mu[i] = g * dmu;
    g_sd[i] = g * dsd;
  }
}

// K == 1, 16-bit mode: samples n and n + N/2 evaluated together (two samples per packed instruction, blvm_math.cuh:
// dmol_k1_two_samples) — must equal the per-sample evaluation bit for bit.
extern "C" void hostsim_k1_pairs(const float* y, const float* raw, const float* gout, int64_t N, int num_bins, float log_eps,
                                 float* lp, float* graw) {
  const DmolConsts C = make_consts(num_bins, log_eps);
  const int64_t H = N / 2;
  for (int64_t n = 0; n < H; ++n) {
    float ra[3], rb[3];
    for (int i = 0; i < 3; ++i) { ra[i] = raw[n * 3 + i]; rb[i] = raw[(n + H) * 3 + i]; }
    dmol_k1_two_samples<true>(y[n], y[n + H], ra, rb, gout ? gout[n] : 1.f, gout ? gout[n + H] : 1.f, C, lp[n], lp[n + H]);
    for (int i = 0; i < 3; ++i) { graw[n * 3 + i] = ra[i]; graw[(n + H) * 3 + i] = rb[i]; }
  }
}

// Monte-Carlo KL log q(z) - log p(z) of one latent element and its five gradients (blvm_math.cuh: kl_mc_terms / kl_mc_grads)
extern "C" void hostsim_kl_mc(const float* z, const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p,
                              const float* gout, int64_t n, float* kl, float* g_mu_q, float* g_sd_q, float* g_mu_p, float* g_sd_p,
                              float* g_z) {
  for (int64_t i = 0; i < n; ++i) {
    const KlMcTerms t = kl_mc_terms(z[i], mu_q[i], sd_q[i], mu_p[i], sd_p[i]);
    kl[i] = t.kl;
    kl_mc_grads(t, gout ? gout[i] : 1.f, g_mu_q[i], g_sd_q[i], g_mu_p[i], g_sd_p[i], g_z[i]);
  }
}
