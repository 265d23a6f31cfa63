/*
 * blvm_b200 — C ABI of the B200-native (sm_100a) DMoL + Gaussian-KL + masked-ELBO kernels.
 *
 * This is the drop-in boundary for the one hot path of JakobHavtorn/benchmarking-lvms that this repo replaces.
 * The reference has no FFI layer (it is pure PyTorch); each entry point below names the reference Python function
 * (path:line under the reference tree) whose eager-op chain it replaces.  The Python package
 * `benchmarking-lvms_b200/` binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - all tensors are dense row-major fp32 unless stated; row sums / scalars are fp64;
 *  - `stream` is a `cudaStream_t` (CUstream) passed as an opaque pointer; every call is asynchronous, stream-ordered,
 *    never allocates and never synchronises;
 *  - return value 0 = OK, otherwise a BLVM_ERR_* code; `blvm_last_error_string()` (thread-local) describes it;
 *  - nullable arguments are marked `nullable`.
 *  - the caller owns all buffers and must keep them alive until the stream has passed the call.
 */
#ifndef BLVM_B200_H_
#define BLVM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLVM_B200_VERSION 201 /* 0.2.1 */

enum {
  BLVM_OK = 0,
  BLVM_ERR_INVALID_ARGUMENT = 1,
  BLVM_ERR_UNSUPPORTED = 2,
  BLVM_ERR_CUDA = 3,
};

/* flags for the DMoL / DL entry points */
enum {
  BLVM_FLAG_MASK_OUTPUT = 1, /* per-sample log-prob output is multiplied by the sequence mask (vrnn.py:268) */
  BLVM_FLAG_SKIP_PADDED = 2, /* tiles entirely inside the padding are not read; their outputs are exact zeros */
  BLVM_FLAG_OVERLAP_PREV = 4, /* KL launches only.  The caller guarantees that the kernel launched IMMEDIATELY BEFORE on this
                                 stream is a blvm DMoL / DL / GMM / KL launch of the same step and produces none of this
                                 call's inputs: this launch may then start while that kernel is still draining
                                 (programmatic dependent launch) and does not complete before it has.  Ordering with respect
                                 to everything earlier on the stream is unchanged. */
  BLVM_FLAG_NANSUM_LOSS = 8, /* blvm_elbo_step only: scalars[0] is WaveNet's -nansum(logp)/sum(x_sl) (wavenet.py:145) and the
                                likelihood gradient rows of utterances whose log-prob is NaN are multiplied by 0 afterwards
                                (what nansum's backward hands them in the reference) */
};

/* element type of the likelihood parameters `raw` (and of their gradient): the AMP Linear output can be consumed as is */
enum {
  BLVM_DTYPE_F32 = 0,
  BLVM_DTYPE_F16 = 1,
  BLVM_DTYPE_BF16 = 2,
};

#define BLVM_MAX_KL_LEVELS 8
#define BLVM_DMOL_TILE 128   /* threads per CTA of the DMoL / DL kernels (tile = 128 x samples-per-thread) */
#define BLVM_KL_TILE 1024    /* latent elements per partial sum of the KL kernels */

typedef void* blvm_stream_t;

int blvm_version(void);
const char* blvm_last_error_string(void);

/* Number of fp64 partial sums per utterance the kernels write (one per CTA tile; the DMoL tile is BLVM_DMOL_TILE
 * samples times a K-dependent samples-per-thread factor, the DL tile BLVM_DMOL_TILE, the KL tile BLVM_KL_TILE). */
int64_t blvm_dmol_chunks(int64_t T, int K, int D);
int64_t blvm_dl_chunks(int64_t T);
/* 1 if (K, D) has a register-resident TMA kernel (D == 1 and K in {1,2,3,4,5,6,8,10,12,16,20,30}); other shapes run the
 * generic fp32 kernel. */
int blvm_dmol_has_fast_path(int K, int D);
/* Kernel selection knob for A/B measurements and tests (process-wide, not thread-safe against concurrent launches):
 * 1 = small-K shapes whose tile slabs are 16-byte aligned run the persistent pipelined kernel (dmol_stream_kernel),
 * 0 = always one tile per CTA (dmol_tile_kernel), -1 = default (env BLVM_B200_STREAM, else 1).  Returns the previous
 * mode.  Both kernels produce bit-identical outputs. */
int blvm_set_stream_mode(int mode);
int64_t blvm_kl_chunks(int64_t row_elems);

/*
 * Discretized mixture of logistics, forward only.
 * Replaces DiscretizedLogisticMixtureDense.forward's split+clamp (blvm/modules/distributions.py:383-387) and
 * discretized_logistic_mixture_ll (blvm/utils/log_likelihoods.py:170-231), plus `* seq_mask` and the per-utterance
 * sum of compute_elbo (blvm/models/vrnn.py:266-269).
 *   y        (B, T, D)         targets in [-1, 1]
 *   raw      (B, T, K(2D+1))   the Linear output: [logits K | per d: locs K, log_scales K]; log-scales are clamped
 *                              at log_epsilon inside the kernel; element type `raw_dtype` (BLVM_DTYPE_*; fp16/bf16
 *                              only where blvm_dmol_has_fast_path), arithmetic is fp32 either way
 *   x_sl     (B) int64, nullable: valid samples per utterance (None = all T)
 *   lp       (B, T), nullable: per-sample log-prob out
 *   partials (B, blvm_dmol_chunks(T, K, D)) fp64, nullable: masked per-tile sums of log-prob
 *   err_flag int32, nullable: set to 1 if some y is outside [-1, 1] (the reference's assert, log_likelihoods.py:195)
 */
int blvm_dmol_fwd(const float* y, const void* raw, int raw_dtype, const int64_t* x_sl, int64_t B, int64_t T, int K,
                  int D, int num_bins, float log_epsilon, int flags, float* lp, double* partials, int* err_flag,
                  blvm_stream_t stream);

/*
 * Single pass forward + gradient: additionally writes (in the element type of raw)
 *   graw (B, T, K(2D+1)) = gscale * (*gscale_dev) * gout[b,t] * mask[b,t] * d log p(y[b,t]) / d raw[b,t,:]
 * which is what autograd produces through log_likelihoods.py:198-231 and the clamp of distributions.py:386
 * (gradient passes at raw == log_epsilon).  `gout` (B, T) nullable = 1; `gscale_dev` nullable fp64 device scalar (an
 * upstream grad_output, e.g. an AMP loss scale, applied without a host sync).  With gscale = -1/sum(x_sl) this is
 * d loss / d raw of vrnn.py:277.  Also serves as the backward of a generic log_prob call (lp = NULL).
 */
int blvm_dmol_fwd_grad(const float* y, const void* raw, int raw_dtype, const int64_t* x_sl, const float* gout,
                       float gscale, const double* gscale_dev, int64_t B, int64_t T, int K, int D, int num_bins,
                       float log_epsilon, int flags, float* lp, void* graw, double* partials, int* err_flag,
                       blvm_stream_t stream);

/*
 * Single discretized logistic (K = 1, no mixture weights), packed raw (B, T, 2) = [mu | log_scale].
 * Replaces DiscretizedLogisticDense.forward/log_prob (distributions.py:298-307) and discretized_logistic_ll
 * (log_likelihoods.py:98-166).  graw nullable (forward only).
 */
int blvm_dl_fwd_grad(const float* y, const float* raw, const int64_t* x_sl, const float* gout, float gscale, int64_t B,
                     int64_t T, int num_bins, float log_epsilon, int flags, float* lp, float* graw, double* partials,
                     int* err_flag, blvm_stream_t stream);

/*
 * Sibling likelihood: Gaussian mixture with the DMoL's packed layout raw (B, T, K(2D+1)) = [logits | per d: mu K, s K]
 * (SURVEY.md §8f row 4).  Replaces DiagonalGaussianMixtureDense.forward's split + sd activation
 * (blvm/modules/distributions.py:198-203) and gaussian_mixture_ll (blvm/utils/log_likelihoods.py:42-60) + autograd.
 *   from_raw = 1: s is the Linear output, sd = softplus_{softplus_beta}(s) + sd_add is applied inside (chain rule folded
 *                 into graw);  from_raw = 0: s is sd itself, clamped at sd_floor if > 0 (then detached, like the
 *                 reference's no_grad clamp, log_likelihoods.py:33-35)
 *   graw nullable (forward only); partials (B, blvm_gmm_chunks(T, K, D)) nullable; mask / gout / gscale as for the DMoL
 */
int64_t blvm_gmm_chunks(int64_t T, int K, int D);
int blvm_gmm_fwd_grad(const float* y, const float* raw, const int64_t* x_sl, const float* gout, float gscale,
                      const double* gscale_dev, int64_t B, int64_t T, int K, int D, int from_raw, double softplus_beta,
                      double sd_add, double sd_floor, int flags, float* lp, float* graw, double* partials,
                      blvm_stream_t stream);
/* Elementwise gaussian_ll (log_likelihoods.py:17-39): lp nullable; g_mu, g_sd nullable together (= gout * d lp/d .). */
int blvm_gaussian_ll(const float* y, const float* mu, const float* sd, const float* gout, int64_t n, double sd_floor,
                     float* lp, float* g_mu, float* g_sd, blvm_stream_t stream);

/*
 * Diagonal-Gaussian KL(q||p), std-dev parametrisation, elementwise over n elements.
 * Replaces kl_divergence_gaussian (blvm/utils/variational.py:67-70).
 */
int blvm_kl_gaussian_fwd(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, int64_t n, float* kl,
                         blvm_stream_t stream);
/* its autograd backward: g_* = gout * d kl / d *. */
int blvm_kl_gaussian_bwd(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, const float* gout,
                         int64_t n, float* g_mu_q, float* g_sd_q, float* g_mu_p, float* g_sd_p, blvm_stream_t stream);

/*
 * Fused KL + free nats + sequence mask + per-utterance sums + gradients, one latent level (B, Tz, Z).
 * Replaces kl_divergence_gaussian + discount_free_nats(shared_dims=-1) (variational.py:86-122) + the masked sums of
 * compute_elbo (vrnn.py:271-276, srnn.py:150-156, clockwork_vae.py:147-153, stcn.py:284-292).
 *   lens        (B) int64, nullable: valid latent steps per utterance (= ceil(x_sl / stride))
 *   free_nats   budget per latent step (shared over Z); 0 disables (variational.py:107)
 *   gscale      multiplier of the gradients (beta / sum(x_sl) for vrnn.py:277)
 *   kl          (B, Tz, Z) nullable: raw elementwise KL out
 *   g_*         (B, Tz, Z) nullable together: d(gscale * sum_masked max(kl, free_nats/Z)) / d inputs,
 *               torch.maximum's 1/2-1/2 rule at exact ties
 *   part_kl, part_klfn (B, blvm_kl_chunks(Tz*Z)) fp64: masked per-tile sums of kl and of max(kl, free_nats/Z)
 *   flags       0 or BLVM_FLAG_OVERLAP_PREV
 */
int blvm_kl_elbo_fwd_grad(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p,
                          const int64_t* lens, int64_t B, int64_t Tz, int64_t Z, double free_nats, float gscale,
                          float* kl, float* g_mu_q, float* g_sd_q, float* g_mu_p, float* g_sd_p, double* part_kl,
                          double* part_klfn, int flags, blvm_stream_t stream);

/* Same reduction when the caller already holds the elementwise KL (compute_elbo's `kld_twise` argument). gkl nullable. */
int blvm_kl_reduce_fwd_grad(const float* kl, const int64_t* lens, int64_t B, int64_t Tz, int64_t Z, double free_nats,
                            float gscale, float* gkl, double* part_kl, double* part_klfn, int flags, blvm_stream_t stream);

/*
 * All latent levels of a step in ONE launch (Clockwork-VAE: 3 levels, clockwork_vae.py:147-155; STCN: n_latents levels,
 * stcn.py:284-292).  Per level either the four Gaussian parameter tensors (fully fused, like blvm_kl_elbo_fwd_grad) or the
 * materialised elementwise KL (like blvm_kl_reduce_fwd_grad); outputs, partial-sum layout and values are identical to one
 * call per level.  `levels_host` is a HOST array of n_levels (<= BLVM_MAX_KL_LEVELS) descriptors holding device pointers.
 */
typedef struct blvm_kl_level {
  const float *mu_q, *sd_q, *mu_p, *sd_p; /* (B, Tz, Z) each; all NULL when `kl` is given */
  const float* kl;                        /* (B, Tz, Z) materialised elementwise KL, or NULL */
  const int64_t* lens;                    /* (B) valid latent steps, nullable */
  int64_t Tz, Z;
  double free_nats;                       /* budget per latent step of this level; 0 disables */
  float *g_mu_q, *g_sd_q, *g_mu_p, *g_sd_p; /* gradients out, nullable together (with the four inputs) */
  float* g_kl;                            /* d/d kl out, nullable (with `kl`) */
  double *part_kl, *part_klfn;            /* (B, blvm_kl_chunks(Tz*Z)) fp64 */
  const float* z;                         /* nullable (B, Tz, Z): with the four parameter tensors, the level's KL is the Monte-Carlo
                                             estimate log q(z) - log p(z) (blvm/utils/variational.py:73-83, bottom-up STCN
                                             stcn.py:288) instead of the analytic KL */
  float* g_z;                             /* nullable: d/d z out (with z and the four gradient outputs) */
} blvm_kl_level_t;
int blvm_kl_elbo_levels_fwd_grad(const blvm_kl_level_t* levels_host, int n_levels, int64_t B, float gscale, int flags,
                                 blvm_stream_t stream);
/* The same with `gscale_dev`, a nullable fp64 DEVICE scalar multiplied into gscale by the kernel (the GradScaler's loss scale of
 * experiments/experiment_vrnn_audio.py:222-226 under fp16 AMP: the KL gradients then carry it like the likelihood's do and the
 * backward needs no rescale pass over them). */
int blvm_kl_elbo_levels_fwd_grad_scaled(const blvm_kl_level_t* levels_host, int n_levels, int64_t B, float gscale,
                                        const double* gscale_dev, int flags, blvm_stream_t stream);

/*
 * Partials -> per-utterance log p(x|z), KL, free-nats KL, ELBO -> loss and bits-per-dim.
 * Replaces vrnn.py:269-277 / srnn.py:149-158 / clockwork_vae.py:155-159 / stcn.py:290-297 / wavenet.py:143-145 and
 * the BitsPerDimMetric arithmetic (blvm/evaluation/metrics.py:443-468).
 *   kl_part_host / klfn_part_host / kl_chunks_host: HOST arrays of n_levels device pointers / chunk counts
 *   rows     (4 + n_levels, B) fp64: logp, kl (sum over levels), kl_fn, elbo, then kl of each level
 *   denom    normaliser of the loss (<= 0: sum(x_sl)); data-parallel ranks pass sum_global(x_sl)/world, the value their
 *            gradients were scaled with
 *   scalars  (8) fp64: loss = -sum_b(logp - beta kl_fn)/denom, sum logp, sum kl, sum kl_fn, sum elbo, sum x_sl,
 *            bits-per-dim = -sum elbo / ln 2 / sum x_sl, nansum-loss = -nansum(logp)/sum x_sl (wavenet.py:145)
 *   sync_counter  one uint32 the caller zero-initialises ONCE per device/stream; the kernel leaves it at zero
 *            (inter-CTA "last block reduces" handshake; launches sharing a counter must be stream-ordered)
 */
int blvm_elbo_finalize(const double* logp_part, int64_t logp_chunks, const double* const* kl_part_host,
                       const double* const* klfn_part_host, const int64_t* kl_chunks_host, int n_levels,
                       const int64_t* x_sl, int64_t B, double beta, double denom, double* rows, double* scalars,
                       unsigned int* sync_counter, blvm_stream_t stream);

/*
 * Multi-GPU: the same finalize with the scalar exchange fused in.  After computing `scalars` the last
 * CTA stores them into this rank's slot of EVERY rank's exchange buffer over NVLink peer memory and releases a
 * per-slot flag — an all-gather of 8 fp64 values with no extra launch and no host call.
 *   peer_bases_host   HOST array of `world` device pointers: every rank's exchange buffer (blvm_exchange_buffer_bytes()
 *                     bytes, zero-initialised, symmetric memory / peer-mapped) as mapped in THIS process; entry `rank`
 *                     is the local buffer
 *   exchange_counters 2 x uint64 device memory, zero-initialised once: steps published / consumed by this rank
 *   prev_global_sums  nullable (8) fp64: the same kernel also consumes the PREVIOUS step (all ranks' slots landed a step
 *                     ago) -> [global loss, sum log_prob, sum kl, sum kl_fn, sum elbo, sum x_sl, global bpd, step number]
 *   err_flag          nullable int32: bit 0 timeout waiting for a peer, bit 1 slot overrun
 */
int blvm_elbo_finalize_publish(const double* logp_part, int64_t logp_chunks, const double* const* kl_part_host,
                               const double* const* klfn_part_host, const int64_t* kl_chunks_host, int n_levels,
                               const int64_t* x_sl, int64_t B, double beta, double denom, double* rows, double* scalars,
                               unsigned int* sync_counter, void* const* peer_bases_host, int rank, int world,
                               unsigned long long* exchange_counters, double* prev_global_sums, int* err_flag,
                               blvm_stream_t stream);
int64_t blvm_exchange_buffer_bytes(void);

/*
 * One step of the whole path from ONE call: likelihood (value + gradient + masked row partials), the KL of every latent
 * level in one launch, finalize (+ the NVLink exchange when `world > 0`).  Same kernels, same partial-sum layout and
 * bit-identical results as the separate entry points above; what it removes is host time (one FFI crossing, no
 * per-level calls), which is what bounds the model-shaped steps (BASELINE configs 2-4: 29-115 us of GPU time).
 * Replaces, in one go, the tail of every reference model: `likelihood.log_prob` + `kl_divergence_gaussian` +
 * `compute_elbo` / `compute_loss` (vrnn.py:255-279, srnn.py:137-160, clockwork_vae.py:132-161, stcn.py:256-297,
 * wavenet.py:128-146).
 *   likelihood  BLVM_LIK_*; with BLVM_LIK_NONE the step is KL-only (log p = 0)
 *   graw        nullable: NULL = forward only (no gradient is written anywhere: the levels' g_* must be NULL too)
 *   loss_scale  nullable fp64 device scalar multiplied into EVERY gradient the step writes, likelihood and KL (fp16 parameters
 *               under a GradScaler: the scaled gradients are representable and the backward only multiplies by grad_output / scale)
 *   levels      n_levels descriptors; their part_kl / part_klfn members are IGNORED (placed in the workspace)
 *   workspace   blvm_elbo_step_workspace_doubles(desc) fp64 values: [scalars 8 | rows (4 + n_levels) x B | partial sums];
 *               scalars and rows are the outputs documented at blvm_elbo_finalize
 *   world       0 = no exchange; otherwise the exchange arguments of blvm_elbo_finalize_publish
 */
enum { BLVM_LIK_NONE = 0, BLVM_LIK_DMOL = 1, BLVM_LIK_DL = 2, BLVM_LIK_GMM = 3 };
typedef struct blvm_elbo_step {
  int likelihood, raw_dtype, K, D, num_bins, flags, n_levels, rank, world;
  float log_epsilon;
  int64_t B, T;
  const float* y;              /* (B, T, D) */
  const void* raw;             /* (B, T, P) */
  const int64_t* x_sl;         /* (B) int64, device */
  float* lp_twise;             /* nullable (B, T): masked per-sample log-prob */
  void* graw;                  /* nullable (B, T, P), element type raw_dtype */
  const double* loss_scale;    /* nullable */
  double gmm_softplus_beta, gmm_sd_add;   /* BLVM_LIK_GMM: the sd activation (distributions.py:167-170) */
  double beta, denom;          /* denom <= 0: sum(x_sl); the gradients are scaled with -1/denom resp. beta/denom */
  blvm_kl_level_t levels[BLVM_MAX_KL_LEVELS];
  double* workspace;
  unsigned int* sync_counter;  /* as for blvm_elbo_finalize */
  int* err_flag;               /* nullable: y-range flag of the DMoL / DL kernels */
  void* const* peer_bases_host;
  unsigned long long* exchange_counters;
  double* prev_global_sums;
  int* exchange_err;
} blvm_elbo_step_t;
int64_t blvm_elbo_step_workspace_doubles(const blvm_elbo_step_t* desc_host);
int blvm_elbo_step(const blvm_elbo_step_t* desc_host, blvm_stream_t stream);
/* Kernels the last successful blvm_elbo_step of THIS thread launched: 3 (likelihood, KL of all levels, finalize; fewer when a
 * part is absent), + 1 for BLVM_FLAG_NANSUM_LOSS's row gate.  (What the Python layer reports as its launch count.) */
int blvm_last_step_launches(void);

/* In-place: rows b of buf (B, row_elems) (element type `dtype`) whose row_values[b] (fp64) is NaN are multiplied by 0;
 * CTAs of finite rows exit after reading one double.  The backward of nansum over utterances (wavenet.py:145). */
int blvm_row_gate_inplace(void* buf, int dtype, int64_t B, int64_t row_elems, const double* row_values, blvm_stream_t stream);
/*
 * Consume step `published - lag` (no-op if it does not exist or was consumed): wait for all ranks' slots in the LOCAL
 * buffer, add them in rank order -> out_sums (8) fp64 = [global loss, sum log_prob, sum kl, sum kl_fn, sum elbo,
 * sum x_sl, global bits-per-dim, step number].  err_flag (nullable int32): bit 0 timeout, bit 1 slot overrun.
 */
int blvm_exchange_consume(void* local_base, int world, unsigned long long* exchange_counters, int lag, double beta,
                          double* out_sums, int* err_flag, blvm_stream_t stream);

/*
 * Quantize: torch.bucketize(x, boundaries, right=False) (blvm/data/transforms.py:257) -> int64 bin index,
 * bit-exact: first i with boundaries[i] >= x.  boundaries (n_bins) fp32 ascending, device.
 */
int blvm_quantize(const float* x, int64_t n, const float* boundaries, int64_t n_bins, int64_t* out, blvm_stream_t stream);

/*
 * Fused sample() + mode() of the mixture, one read of the parameters.
 * Replaces rsample_discretized_logistic_mixture (blvm/utils/variational.py:309-349: Gumbel-max over the logits with
 * u ~ U(1e-5, 1-1e-5), gather, logistic inverse CDF with u ~ U(1e-8, 1-1e-8), clamp to [-1, 1]) and
 * DiscretizedLogisticMixtureDense.mode (blvm/modules/distributions.py:363-368: loc of the arg-max logit).
 *   raw (N, K(2D+1)) in raw_dtype; sample, mode (N, D) fp32, nullable; mode_index (N) int32, nullable
 *   seed / offset: Philox4x32-10 key and stream offset (same seed + offset => same samples)
 */
int blvm_dmol_sample_mode(const void* raw, int raw_dtype, int64_t N, int K, int D, float log_epsilon, uint64_t seed,
                          uint64_t offset, float* sample, float* mode, int32_t* mode_index, blvm_stream_t stream);

/* In-place `buf *= (float)*scale` (fp64 device scalar) that exits early when *scale == 1: the autograd backward of the
 * fused ELBO op uses it to apply an upstream grad_output (e.g. an AMP loss scale) without a host sync. */
int blvm_scale_inplace(float* buf, int64_t n, const double* scale, blvm_stream_t stream);
/* Same for up to BLVM_MAX_SCALE_BUFFERS buffers in ONE launch (bufs_host / ns_host / dtypes_host are HOST arrays of
 * length count; dtypes_host nullable = all fp32, else BLVM_DTYPE_* per buffer). */
#define BLVM_MAX_SCALE_BUFFERS 36
int blvm_scale_inplace_multi(void* const* bufs_host, const int64_t* ns_host, const int* dtypes_host, int count,
                             const double* scale, blvm_stream_t stream);

/*
 * The likelihood HEAD in one kernel: nn.Linear(x_dim -> 3K) on the tcgen05 tensor cores (accumulator in
 * tensor memory), DMoL value + gradient in registers, and the Linear's backward (dx, dW, db) on the tensor cores again; the
 * (B, T, 3K) parameter tensor and its gradient never touch HBM.  Replaces DiscretizedLogisticMixtureDense.forward
 * (blvm/modules/distributions.py:381-387: `self.params(x)`, split, clamp) + discretized_logistic_mixture_ll
 * (blvm/utils/log_likelihoods.py:170-231) + mask / row sums (vrnn.py:266-269) + the autograd backward of all of them, for
 * 16-bit (AMP) activations: x and W are fp16 / bf16 (W: the autocast copy of `params.weight`), accumulation is fp32.
 *   x (B, T, Din), W (3K, Din) element type `dtype` (BLVM_DTYPE_F16 / BF16); bias (3K) fp32 nullable; y (B, T) fp32
 *   supported: K == 10, even Din <= 79 (blvm_linear_dmol_padded_dim() != 0)
 *   lp (B, T) nullable; partials (B, ceil(T / 128)) fp64 nullable: the layout blvm_elbo_finalize expects for K = 10
 *   dx (B, T, Din) in `dtype` and dw_partial (dw_partial_ctas >= blvm_linear_dmol_max_ctas(), 32, padded_dim) fp32: nullable
 *     TOGETHER (forward only).  dx = gscale * (*gscale_dev) * mask * d log p / d x; blvm_linear_dmol_reduce_dw() turns the per-CTA
 *     partials into dW (3K, Din) and db (3K) (fixed order: bit-reproducible); *ctas_used_host (host, nullable) = CTAs launched
 *   raw_debug (B*T, 32) fp32 nullable: the tensor-core result x W^T + b per sample (tests)
 */
int blvm_linear_dmol_padded_dim(int K, int64_t Din);
int64_t blvm_linear_dmol_max_ctas(void);
int blvm_linear_dmol_fwd_grad(const float* y, const void* x, const void* W, const float* bias, int dtype, const int64_t* x_sl,
                              float gscale, const double* gscale_dev, int64_t B, int64_t T, int64_t Din, int K, int num_bins,
                              float log_epsilon, int flags, float* lp, void* dx, float* dw_partial, int64_t dw_partial_ctas,
                              double* partials, int* err_flag, float* raw_debug, int64_t* ctas_used_host, blvm_stream_t stream);
int blvm_linear_dmol_reduce_dw(const float* dw_partial, int64_t ctas, int64_t Din, int K, float* dW, float* db, blvm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BLVM_B200_H_ */
