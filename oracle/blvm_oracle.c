/*
 * C restatement of the reference's DMoL + Gaussian-KL + masked-ELBO path.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent oracle next to oracle/blvm_oracle.py: the reference's arithmetic written out naively, op by op
 * (sigmoid differences, softplus with threshold 20, where-chains; citations = blvm/utils/log_likelihoods.py lines),
 * in fp32 (the reference's native precision) and fp64, multi-threaded over utterances with OpenMP.  It is (1) pinned
 * against the golden vectors generated from the reference (tests/test_oracle_c.py) and (2) the `cpu_baseline` /
 * `--impl reference` arm of bench.py ("port": the reference is Python/PyTorch and does not travel to the GPU box).
 * Nothing in benchmarking-lvms_b200/ links or loads this file.
 *
 * Build: make -C oracle   (-> oracle/_build/libblvm_oracle.so)
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FN(name) CAT(name, _f32)
#define EXP expf
#define LOG logf
#define LOG1P log1pf
#include "oracle_impl.h"
#undef REAL
#undef FN
#undef EXP
#undef LOG
#undef LOG1P

#define REAL double
#define FN(name) CAT(name, _f64)
#define EXP exp
#define LOG log
#define LOG1P log1p
#include "oracle_impl.h"

/* loss = -sum_b(logp_b - beta * klfn_b) / sum_b x_sl_b  (vrnn.py:277); also elbo_b = logp_b - kl_b (:273). */
double oracle_elbo_loss(const double* row_logp, const double* row_kl, const double* row_klfn, const int64_t* x_sl,
                        int64_t B, double beta, double* elbo) {
  double obj = 0.0, len = 0.0;
  for (int64_t b = 0; b < B; ++b) {
    const double fn = row_klfn ? row_klfn[b] : 0.0, kl = row_kl ? row_kl[b] : 0.0;
    obj += row_logp[b] - beta * fn;
    len += (double)x_sl[b];
    if (elbo) elbo[b] = row_logp[b] - kl;
  }
  return -obj / len;
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
