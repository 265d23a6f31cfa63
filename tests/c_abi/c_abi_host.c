/*
 * A torch-free, Python-free host of the C ABI (include/blvm_b200.h): plain C, cudaMalloc'd buffers, one DMoL + KL + ELBO
 * step through libblvm_b200.so, checked against the C oracle (oracle/blvm_oracle.c) run in fp64 on the same inputs.
 * TEST INFRASTRUCTURE (tests/test_gpu_c_abi.py builds and runs it on the GPU box): it shows that the boundary a
 * maintainer binds is exactly the header — raw device pointers, sizes, a stream — and nothing of PyTorch.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include tests/c_abi/c_abi_host.c -o c_abi_host \
 *       benchmarking-lvms_b200/lib/libblvm_b200.so oracle/_build/libblvm_oracle.so -L/usr/local/cuda/lib64 -lcudart -lm
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "blvm_b200.h"

/* the checker (fp64 restatement of the reference path) */
void oracle_dmol_f64(const double* y, const double* raw, const int64_t* x_sl, int64_t B, int64_t T, int K, int num_bins,
                     double log_eps, double gscale, double* lp_out, double* graw, double* row_logp);
void oracle_kl_f64(const double* mu_q, const double* sd_q, const double* mu_p, const double* sd_p, const int64_t* lens,
                   int64_t B, int64_t Tz, int64_t Z, double free_nats, double gscale, double* g_mu_q, double* g_sd_q,
                   double* g_mu_p, double* g_sd_p, double* row_kl, double* row_klfn);
double oracle_elbo_loss(const double* row_logp, const double* row_kl, const double* row_klfn, const int64_t* x_sl, int64_t B,
                        double beta, double* elbo);

#define CK(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define BK(call)                                                                               \
  do {                                                                                         \
    int rc_ = (call);                                                                          \
    if (rc_ != BLVM_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, blvm_last_error_string()); return 3; } \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand(void) {   /* xorshift64*, uniform in (0, 1) */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return ((rng_state * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) + 1e-12;
}
static double nrand(void) { return sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand()); }

static void* dev_copy(const void* host, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
  if (host) cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
  else cudaMemset(d, 0, bytes);
  return d;
}

int main(void) {
  const int64_t B = 5, T = 3000, S = 64, Z = 16, Tz = (T + S - 1) / S;
  const int K = 10, nb = 65536, P = 3 * K;
  const double beta = 0.5, free_nats = 0.0625;
  const int64_t x_sl[5] = {3000, 2999, 1501, 64, 1};
  int64_t lens[5];
  double total = 0;
  for (int b = 0; b < B; ++b) { lens[b] = (x_sl[b] + S - 1) / S; total += (double)x_sl[b]; }

  const int64_t N = B * T, L = B * Tz * Z;
  float* y = malloc(N * 4); float* raw = malloc(N * P * 4); float* kl[4];
  double* y64 = malloc(N * 8); double* raw64 = malloc(N * P * 8); double* kl64[4];
  for (int64_t i = 0; i < N; ++i) {
    y[i] = (float)((double)(int64_t)(urand() * nb) / (nb - 1) * 2 - 1);
    if (y[i] > 1.f) y[i] = 1.f;
    for (int k = 0; k < K; ++k) {
      raw[i * P + k] = (float)nrand();
      raw[i * P + K + k] = y[i] + 0.1f * (float)nrand();
      raw[i * P + 2 * K + k] = (float)(nrand() * 2 - 4);
    }
  }
  y[7] = -1.f; y[8] = 1.f;   /* both edge bins */
  for (int64_t i = 0; i < N; ++i) y64[i] = y[i];
  for (int64_t i = 0; i < N * P; ++i) raw64[i] = raw[i];
  for (int j = 0; j < 4; ++j) {
    kl[j] = malloc(L * 4); kl64[j] = malloc(L * 8);
    for (int64_t i = 0; i < L; ++i) {
      const double v = nrand();
      kl[j][i] = (float)((j & 1) ? log1p(exp(v)) + 1e-3 : v);
      kl64[j][i] = kl[j][i];
    }
  }

  /* ---- device side: only what the header declares ---- */
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const int64_t chunks = blvm_dmol_chunks(T, K, 1), kchunks = blvm_kl_chunks(Tz * Z);
  float* d_y = dev_copy(y, N * 4); float* d_raw = dev_copy(raw, N * P * 4);
  float* d_lp = dev_copy(NULL, N * 4); float* d_graw = dev_copy(NULL, N * P * 4);
  int64_t* d_xsl = dev_copy(x_sl, B * 8); int64_t* d_lens = dev_copy(lens, B * 8);
  double* d_part = dev_copy(NULL, B * chunks * 8);
  double* d_pk = dev_copy(NULL, B * kchunks * 8); double* d_pf = dev_copy(NULL, B * kchunks * 8);
  float *d_kl[4], *d_gkl[4];
  for (int j = 0; j < 4; ++j) { d_kl[j] = dev_copy(kl[j], L * 4); d_gkl[j] = dev_copy(NULL, L * 4); }
  double* d_rows = dev_copy(NULL, (4 + 1) * B * 8); double* d_scalars = dev_copy(NULL, 8 * 8);
  unsigned int* d_ctr = dev_copy(NULL, 4); int* d_err = dev_copy(NULL, 4);
  if (!d_y || !d_raw || !d_lp || !d_graw || !d_xsl || !d_lens || !d_part || !d_pk || !d_pf || !d_rows || !d_scalars || !d_ctr || !d_err) {
    fprintf(stderr, "cudaMalloc failed\n");
    return 2;
  }

  BK(blvm_dmol_fwd_grad(d_y, d_raw, BLVM_DTYPE_F32, d_xsl, NULL, (float)(-1.0 / total), NULL, B, T, K, 1, nb, -7.0f,
                        BLVM_FLAG_MASK_OUTPUT, d_lp, d_graw, d_part, d_err, st));
  BK(blvm_kl_elbo_fwd_grad(d_kl[0], d_kl[1], d_kl[2], d_kl[3], d_lens, B, Tz, Z, free_nats, (float)(beta / total), NULL,
                           d_gkl[0], d_gkl[1], d_gkl[2], d_gkl[3], d_pk, d_pf, BLVM_FLAG_OVERLAP_PREV, st));
  const double* pk_arr[1] = {d_pk}; const double* pf_arr[1] = {d_pf}; const int64_t kc_arr[1] = {kchunks};
  BK(blvm_elbo_finalize(d_part, chunks, pk_arr, pf_arr, kc_arr, 1, d_xsl, B, beta, total, d_rows, d_scalars, d_ctr, st));
  CK(cudaStreamSynchronize(st));

  double scalars[8]; double rows[5 * 5]; int err = 0;
  float* lp = malloc(N * 4); float* graw = malloc(N * P * 4); float* gkl0 = malloc(L * 4);
  CK(cudaMemcpy(scalars, d_scalars, sizeof scalars, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rows, d_rows, sizeof rows, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lp, d_lp, N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(graw, d_graw, N * P * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gkl0, d_gkl[0], L * 4, cudaMemcpyDeviceToHost));

  /* ---- the checker, fp64 ---- */
  double* o_lp = malloc(N * 8); double* o_graw = malloc(N * P * 8); double o_logp[5], o_kl[5], o_fn[5], o_elbo[5];
  double* o_g[4];
  for (int j = 0; j < 4; ++j) o_g[j] = malloc(L * 8);
  oracle_dmol_f64(y64, raw64, x_sl, B, T, K, nb, -7.0, -1.0 / total, o_lp, o_graw, o_logp);
  oracle_kl_f64(kl64[0], kl64[1], kl64[2], kl64[3], lens, B, Tz, Z, free_nats, beta / total, o_g[0], o_g[1], o_g[2], o_g[3], o_kl, o_fn);
  const double o_loss = oracle_elbo_loss(o_logp, o_kl, o_fn, x_sl, B, beta, o_elbo);

  int bad = 0;
  double worst_lp = 0, worst_g = 0, worst_gkl = 0;
  for (int64_t i = 0; i < N; ++i) {
    const double e = fabs(lp[i] - o_lp[i]) / (1e-5 * fabs(o_lp[i]) + 1e-6);
    if (e > worst_lp) worst_lp = e;
  }
  for (int64_t i = 0; i < N; ++i)
    for (int g0 = 0; g0 < P; g0 += K) {   /* relative to the sample's parameter-group scale (tests/parity.py) */
      double gmax = 0;
      for (int k = 0; k < K; ++k) gmax = fmax(gmax, fabs(o_graw[i * P + g0 + k]));
      for (int k = 0; k < K; ++k) {
        const double r = o_graw[i * P + g0 + k];
        const double e = fabs(graw[i * P + g0 + k] - r) / (1e-5 * fabs(r) + 1e-5 * gmax + 1e-6 / total + 1e-30);
        if (e > worst_g) worst_g = e;
      }
    }
  double gklmax = 0;
  for (int64_t i = 0; i < L; ++i) gklmax = fmax(gklmax, fabs(o_g[0][i]));
  for (int64_t i = 0; i < L; ++i) {
    const double e = fabs(gkl0[i] - o_g[0][i]) / (1e-5 * fabs(o_g[0][i]) + 1e-7 * gklmax);
    if (e > worst_gkl) worst_gkl = e;
  }
  if (worst_lp > 1 || worst_g > 1 || worst_gkl > 1) bad |= 1;
  for (int b = 0; b < B; ++b) {
    if (fabs(rows[0 * B + b] - o_logp[b]) > 1e-6 * fabs(o_logp[b]) + 1e-9) bad |= 2;
    if (fabs(rows[1 * B + b] - o_kl[b]) > 1e-6 * fabs(o_kl[b]) + 1e-9) bad |= 4;
    if (fabs(rows[2 * B + b] - o_fn[b]) > 1e-6 * fabs(o_fn[b]) + 1e-9) bad |= 8;
    if (fabs(rows[3 * B + b] - o_elbo[b]) > 1e-6 * fabs(o_elbo[b]) + 1e-9) bad |= 16;
  }
  if (fabs(scalars[0] - o_loss) > 1e-6 * fabs(o_loss)) bad |= 32;
  if (scalars[5] != total) bad |= 64;
  if (err != 0) bad |= 128;
  /* error behaviour: validation failures return a code and a message, nothing is launched */
  if (blvm_dmol_fwd_grad(d_y, d_raw, BLVM_DTYPE_F32, d_xsl, NULL, 1.f, NULL, B, T, 0, 1, nb, -7.0f, 0, d_lp, d_graw, d_part, d_err, st) == BLVM_OK) bad |= 256;
  if (strlen(blvm_last_error_string()) == 0) bad |= 512;

  printf("c_abi_host: version %d  loss %.12g (oracle %.12g)  worst err/tol: lp %.3g  graw %.3g  gkl %.3g  flags %d -> %s\n",
         blvm_version(), scalars[0], o_loss, worst_lp, worst_g, worst_gkl, bad, bad ? "FAIL" : "PASS");
  return bad ? 1 : 0;
}
