/* Body of the C oracle, included twice (REAL = float / double).  TEST INFRASTRUCTURE — see blvm_oracle.c. */

static inline REAL FN(sigmoid)(REAL x) { return (REAL)1 / ((REAL)1 + EXP(-x)); }
/* F.softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x)) */
static inline REAL FN(softplus)(REAL x) { return x > (REAL)20 ? x : LOG1P(EXP(x)); }

/* One (sample, component) of blvm/utils/log_likelihoods.py:202-227 followed by the closed-form derivatives that
 * torch autograd produces for those lines: returns log-prob, writes d/d loc and d/d log_scale. */
static inline REAL FN(component)(REAL y, REAL loc, REAL log_scale, int num_bins, REAL* dloc, REAL* dls) {
  const REAL half = (REAL)(1.0 / (num_bins - 1));
  const REAL centered_y = y - loc;                                  /* :202 */
  const REAL inv_stdv = EXP(-log_scale);                            /* :203 */
  const REAL plus_in = inv_stdv * (centered_y + half);              /* :206 */
  const REAL cdf_plus = FN(sigmoid)(plus_in);                       /* :207 */
  const REAL minus_in = inv_stdv * (centered_y - half);             /* :208 */
  const REAL cdf_minus = FN(sigmoid)(minus_in);                     /* :209 */
  const REAL cdf_delta = cdf_plus - cdf_minus;                      /* :210 */
  const REAL mid_in = inv_stdv * centered_y;                        /* :219 */
  REAL lp, da = 0, db = 0, dm = 0, direct = 0;
  if (y > (REAL)(1.0 - 2.0 / num_bins)) {                           /* :227 upper edge wins */
    lp = -FN(softplus)(minus_in);                                   /* :216 */
    db = -cdf_minus;
  } else if (y < (REAL)(2.0 / num_bins - 1.0)) {                    /* :226 lower edge */
    lp = plus_in - FN(softplus)(plus_in);                           /* :213 */
    da = (REAL)1 - cdf_plus;
  } else if (cdf_delta > (REAL)1e-5) {                              /* :222 */
    const REAL d = cdf_delta > (REAL)1e-10 ? cdf_delta : (REAL)1e-10;
    lp = LOG(d);
    da = cdf_plus * ((REAL)1 - cdf_plus) / d;
    db = -cdf_minus * ((REAL)1 - cdf_minus) / d;
  } else {
    lp = mid_in - log_scale - (REAL)2 * FN(softplus)(mid_in) - (REAL)log(num_bins / 2.0);   /* :220,222 */
    dm = (REAL)1 - (REAL)2 * FN(sigmoid)(mid_in);
    direct = (REAL)-1;
  }
  *dloc = -inv_stdv * (da + db + dm);
  *dls = -(plus_in * da + minus_in * db + mid_in * dm) + direct;
  return lp;
}

/* DMoL over a (B, T) batch from the packed Linear output raw (B, T, 3K) = [logits | locs | log_scales]
 * (distributions.py:383-387), masked by x_sl, with d(gscale * sum_masked lp)/d raw.  row_logp (B) in double. */
void FN(oracle_dmol)(const REAL* y, const REAL* raw, const int64_t* x_sl, int64_t B, int64_t T, int K, int num_bins,
                     REAL log_eps, REAL gscale, REAL* lp_out, REAL* graw, double* row_logp) {
  const int P = 3 * K;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    double acc = 0.0;
    REAL v[64], dmu[64], dls[64];
    for (int64_t t = 0; t < T; ++t) {
      const int64_t s = b * T + t;
      const REAL* p = raw + s * P;
      const int valid = t < x_sl[b];
      REAL m1 = -INFINITY, m2 = -INFINITY;
      for (int k = 0; k < K; ++k) {
        const REAL raw_ls = p[2 * K + k];
        const REAL ls = raw_ls < log_eps ? log_eps : raw_ls;        /* clamp(min), distributions.py:386 */
        const REAL l = FN(component)(y[s], p[K + k], ls, num_bins, &dmu[k], &dls[k]);
        if (raw_ls < log_eps) dls[k] = 0;                           /* clamp passes gradient at equality */
        v[k] = l + p[k];
        if (v[k] > m1) m1 = v[k];
        if (p[k] > m2) m2 = p[k];
      }
      REAL s1 = 0, s2 = 0;
      for (int k = 0; k < K; ++k) { s1 += EXP(v[k] - m1); s2 += EXP(p[k] - m2); }
      const REAL L = (m1 + LOG(s1)) - (m2 + LOG(s2));               /* :230-231 log_softmax + logsumexp */
      const REAL Lm = valid ? L : L * (REAL)0;                      /* `* seq_mask`, vrnn.py:268 */
      if (lp_out) lp_out[s] = Lm;
      acc += (double)Lm;
      if (graw) {
        const REAL g = valid ? gscale : (REAL)0;
        REAL* o = graw + s * P;
        for (int k = 0; k < K; ++k) {
          const REAL resp = EXP(v[k] - m1) / s1, pi = EXP(p[k] - m2) / s2;
          o[k] = g * (resp - pi);
          o[K + k] = g * resp * dmu[k];
          o[2 * K + k] = g * resp * dls[k];
        }
      }
    }
    row_logp[b] = acc;
  }
}

/* Gaussian KL (variational.py:67-70) + free nats (:86-122, shared_dims=-1) + mask + per-row sums + gradients. */
void FN(oracle_kl)(const REAL* mu_q, const REAL* sd_q, const REAL* mu_p, const REAL* sd_p, const int64_t* lens, int64_t B,
                   int64_t Tz, int64_t Z, double free_nats, REAL gscale, REAL* g_mu_q, REAL* g_sd_q, REAL* g_mu_p,
                   REAL* g_sd_p, double* row_kl, double* row_klfn) {
  const REAL c = (REAL)(free_nats / (double)Z);
  const int fn = free_nats != 0.0;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    double a_kl = 0.0, a_fn = 0.0;
    for (int64_t i = 0; i < Tz * Z; ++i) {
      const int64_t e = b * Tz * Z + i;
      const int valid = (i / Z) < lens[b];
      const REAL d = mu_q[e] - mu_p[e];
      const REAL kl = LOG(sd_p[e]) - LOG(sd_q[e]) + (sd_q[e] * sd_q[e] + d * d) / ((REAL)2 * sd_p[e] * sd_p[e]) - (REAL)0.5;
      const REAL klfn = (fn && kl < c) ? c : kl;
      a_kl += (double)(valid ? kl : kl * (REAL)0);
      a_fn += (double)(valid ? klfn : klfn * (REAL)0);
      if (g_mu_q) {
        const REAL gate = !fn ? (REAL)1 : (kl > c ? (REAL)1 : (kl == c ? (REAL)0.5 : (REAL)0));  /* torch.maximum ties */
        const REAL g = valid ? gscale * gate : (REAL)0;
        const REAL sp2 = sd_p[e] * sd_p[e];
        g_mu_q[e] = g * d / sp2;
        g_mu_p[e] = -g_mu_q[e];
        g_sd_q[e] = g * (-(REAL)1 / sd_q[e] + sd_q[e] / sp2);
        g_sd_p[e] = g * ((REAL)1 / sd_p[e] - (sd_q[e] * sd_q[e] + d * d) / (sp2 * sd_p[e]));
      }
    }
    row_kl[b] = a_kl;
    row_klfn[b] = a_fn;
  }
}
