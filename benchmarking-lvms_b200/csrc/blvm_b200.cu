// C ABI of the blvm_b200 kernels (include/blvm_b200.h).  Validation + launch only: no allocation, no synchronisation.
#include "../../include/blvm_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "host_common.h"

#include "dmol_dispatch.cuh"
#include "kl_kernels.cuh"
#include "misc_kernels.cuh"
#include "sample_kernels.cuh"

using namespace blvm;

static_assert(BLVM_DMOL_TILE == 128, "tile constant mirrors the kernel template argument");
static_assert(BLVM_KL_TILE == kKlChunk, "tile constant mirrors the KL kernel");
static_assert(BLVM_MAX_KL_LEVELS == kMaxLevels, "level cap");
static_assert(BLVM_MAX_SCALE_BUFFERS == kMaxScaleBuffers, "scale buffer cap");
static_assert(static_cast<int>(BLVM_FLAG_MASK_OUTPUT) == static_cast<int>(kFlagMaskOutput) &&
                  static_cast<int>(BLVM_FLAG_SKIP_PADDED) == static_cast<int>(kFlagSkipPadded), "flags");
using blvm_host::kMaxDevices;
using blvm_host::current_device;

namespace blvm_host {   // instantiated in blvm_dmol_f32.cu / _f16.cu / _bf16.cu
extern template int dmol_dispatch_tp<float>(const blvm::DmolArgs&, bool, int64_t, cudaStream_t);
extern template int dmol_dispatch_tp<__half>(const blvm::DmolArgs&, bool, int64_t, cudaStream_t);
extern template int dmol_dispatch_tp<__nv_bfloat16>(const blvm::DmolArgs&, bool, int64_t, cudaStream_t);
extern template int sample_dispatch_tp<float>(const blvm::SampleArgs&, int64_t, cudaStream_t);
extern template int sample_dispatch_tp<__half>(const blvm::SampleArgs&, int64_t, cudaStream_t);
extern template int sample_dispatch_tp<__nv_bfloat16>(const blvm::SampleArgs&, int64_t, cudaStream_t);
}  // namespace blvm_host

namespace {

using blvm_host::kTile;
using blvm_host::aligned;
using blvm_host::check_launch;
using blvm_host::fail;
using blvm_host::make_consts;

// 1 (default) = KL / finalize launches use programmatic dependent launch where the caller allows it; env BLVM_B200_PDL=0 disables
}  // namespace
namespace blvm_host {
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BLVM_B200_PDL");
    on = e ? (atoi(e) != 0) : 1;
  }
  return on != 0;
}

int sm_count() {
  static int n_dev[kMaxDevices] = {};
  const int dev = current_device();
  int& n = n_dev[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// utterances up to which ONE CTA finalizes the whole step (elbo_finalize_small_kernel); env BLVM_B200_FIN_SMALL_B overrides
int finalize_small_max_b() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BLVM_B200_FIN_SMALL_B");
    v = e ? atoi(e) : 64;
    if (v < 0) v = 0;
  }
  return v;
}

// 0 = tile kernel only, 1 = stream kernel where eligible (default), read once (A/B runs set it before the first call)
int g_stream_mode = -1;
int stream_mode() {
  if (g_stream_mode < 0) {
    const char* e = getenv("BLVM_B200_STREAM");
    g_stream_mode = e ? (atoi(e) != 0) : 1;
  }
  return g_stream_mode;
}
}  // namespace blvm_host
namespace {
using blvm_host::dmol_dispatch_tp;
using blvm_host::g_stream_mode;
using blvm_host::launch_ex;
using blvm_host::pdl_enabled;
using blvm_host::sample_dispatch_tp;
using blvm_host::sm_count;
using blvm_host::stream_mode;

// samples per partial sum: the register kernel's tile (128 * samples-per-thread, a function of K) or 128 (generic / DL)
int64_t dmol_tile_samples(int K, int D) {
  if (D == 1) {
    switch (K) {
#define BLVM_CASE(KK) \
  case KK:            \
    return static_cast<int64_t>(kTile) * DmolSpt<KK>::value;
      BLVM_FOR_EACH_K(BLVM_CASE)
#undef BLVM_CASE
      default: break;
    }
  }
  return kTile;
}

bool dmol_has_register_kernel(int K, int D) {
  if (D != 1) return false;
  switch (K) {
#define BLVM_CASE(KK) case KK:
    BLVM_FOR_EACH_K(BLVM_CASE)
#undef BLVM_CASE
    return true;
    default: return false;
  }
}

template <bool GRAD>
int dispatch_dmol(const DmolArgs& A, int raw_dtype, cudaStream_t st) {
  const int64_t tiles = A.B * A.chunks;
  if (tiles == 0) return BLVM_OK;
  if (tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles (%lld)", (long long)tiles);
  if (dmol_has_register_kernel(A.K, A.D)) {
    switch (raw_dtype) {
      case BLVM_DTYPE_F32: return dmol_dispatch_tp<float>(A, GRAD, tiles, st);
      case BLVM_DTYPE_F16: return dmol_dispatch_tp<__half>(A, GRAD, tiles, st);
      case BLVM_DTYPE_BF16: return dmol_dispatch_tp<__nv_bfloat16>(A, GRAD, tiles, st);
      default: return fail(BLVM_ERR_INVALID_ARGUMENT, "raw_dtype=%d", raw_dtype);
    }
  }
  if (raw_dtype != BLVM_DTYPE_F32)
    return fail(BLVM_ERR_UNSUPPORTED, "fp16/bf16 parameters need a register kernel (K=%d D=%d has none): upcast to fp32", A.K, A.D);
  dmol_generic_kernel<kTile, GRAD><<<static_cast<unsigned>(tiles), kTile, 0, st>>>(A);
  return check_launch("dmol_generic_kernel");
}

// ---- Gaussian mixture: same tile kernel, different component density; fp32 parameters, K in {1, 5, 10, 20} in
// registers, anything else through the generic kernel
#define BLVM_FOR_EACH_GMM_K(X) X(1) X(5) X(10) X(20)

template <int K, bool GRAD, int LIK>
int launch_gmm_tile(const DmolArgs& A, int64_t tiles, cudaStream_t st) {
  constexpr size_t smem = dmol_tile_smem_bytes<K, kTile, float>();
  auto kern = dmol_tile_kernel<K, kTile, GRAD, kUGeneral, float, LIK>;
  static bool configured[blvm_host::kMaxDevices] = {};
  const int dev = blvm_host::current_device();
  if (!configured[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured[dev] = true;
  }
  constexpr int G = DmolGroup<K, float>::value;
  DmolArgs A2 = A;
  A2.ctas_per_row = (A.chunks + G - 1) / G;
  kern<<<static_cast<unsigned>(A.B * A2.ctas_per_row), kTile, smem, st>>>(A2);
  return check_launch("dmol_tile_kernel<gmm>");
}

bool gmm_has_register_kernel(int K, int D) {
  if (D != 1) return false;
  switch (K) {
#define BLVM_CASE(KK) case KK:
    BLVM_FOR_EACH_GMM_K(BLVM_CASE)
#undef BLVM_CASE
    return true;
    default: return false;
  }
}

template <bool GRAD>
int dispatch_gmm(const DmolArgs& A, cudaStream_t st) {
  const int64_t tiles = A.B * A.chunks;
  if (tiles == 0) return BLVM_OK;
  if (tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles (%lld)", (long long)tiles);
  if (A.D == 1) {
    switch (A.K) {
#define BLVM_CASE(KK)                                                                          \
  case KK:                                                                                     \
    return A.lik == kLikGmmRaw ? launch_gmm_tile<KK, GRAD, kLikGmmRaw>(A, tiles, st)           \
                               : launch_gmm_tile<KK, GRAD, kLikGmmSd>(A, tiles, st);
      BLVM_FOR_EACH_GMM_K(BLVM_CASE)
#undef BLVM_CASE
      default: break;
    }
  }
  dmol_generic_kernel<kTile, GRAD><<<static_cast<unsigned>(tiles), kTile, 0, st>>>(A);
  return check_launch("dmol_generic_kernel<gmm>");
}

int validate_dmol(const float* y, const void* raw, int64_t B, int64_t T, int K, int D, int num_bins) {
  if (B < 0 || T < 0) return fail(BLVM_ERR_INVALID_ARGUMENT, "negative size B=%lld T=%lld", (long long)B, (long long)T);
  if (T > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "T=%lld: utterances are limited to 2^31 - 1 samples", (long long)T);
  if (K < 1 || D < 1) return fail(BLVM_ERR_INVALID_ARGUMENT, "K=%d D=%d must be >= 1", K, D);
  if (num_bins < 2) return fail(BLVM_ERR_INVALID_ARGUMENT, "num_bins=%d must be >= 2", num_bins);
  if (B * T > 0 && (!y || !raw)) return fail(BLVM_ERR_INVALID_ARGUMENT, "null y/raw");
  if (!aligned(y, 4) || !aligned(raw, 2)) return fail(BLVM_ERR_INVALID_ARGUMENT, "y must be 4-byte, raw element aligned");
  return BLVM_OK;
}

}  // namespace

namespace blvm_host {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return BLVM_OK;
}
const char* last_error() { return g_err; }
}  // namespace blvm_host

extern "C" {

int blvm_version(void) { return BLVM_B200_VERSION; }
const char* blvm_last_error_string(void) { return blvm_host::last_error(); }
int64_t blvm_dmol_chunks(int64_t T, int K, int D) {
  const int64_t ts = dmol_tile_samples(K, D);
  return (T + ts - 1) / ts;
}
int64_t blvm_dl_chunks(int64_t T) { return (T + kTile * kDlSpt - 1) / (kTile * kDlSpt); }
int blvm_set_stream_mode(int mode) {
  const int prev = stream_mode();
  g_stream_mode = mode < 0 ? -1 : (mode != 0);
  return prev;
}
int blvm_dmol_has_fast_path(int K, int D) { return dmol_has_register_kernel(K, D) ? 1 : 0; }
int64_t blvm_kl_chunks(int64_t row_elems) { return (row_elems + kKlChunk - 1) / kKlChunk; }

int blvm_dmol_fwd(const float* y, const void* raw, int raw_dtype, const int64_t* x_sl, int64_t B, int64_t T, int K, int D,
                  int num_bins, float log_epsilon, int flags, float* lp, double* partials, int* err_flag,
                  blvm_stream_t stream) {
  if (int rc = validate_dmol(y, raw, B, T, K, D, num_bins)) return rc;
  DmolArgs A{};
  A.y = y; A.raw = raw; A.x_sl = x_sl; A.gout = nullptr; A.gscale = 0.f; A.lp = lp; A.graw = nullptr;
  A.partials = partials; A.err_flag = err_flag; A.B = B; A.T = T; A.chunks = blvm_dmol_chunks(T, K, D); A.K = K; A.D = D;
  A.flags = flags; A.C = make_consts(num_bins, log_epsilon);
  return dispatch_dmol<false>(A, raw_dtype, static_cast<cudaStream_t>(stream));
}

int blvm_dmol_fwd_grad(const float* y, const void* raw, int raw_dtype, const int64_t* x_sl, const float* gout, float gscale,
                       const double* gscale_dev, int64_t B, int64_t T, int K, int D, int num_bins, float log_epsilon,
                       int flags, float* lp, void* graw, double* partials, int* err_flag, blvm_stream_t stream) {
  if (int rc = validate_dmol(y, raw, B, T, K, D, num_bins)) return rc;
  if (B * T > 0 && !graw) return fail(BLVM_ERR_INVALID_ARGUMENT, "null graw");
  if (!aligned(graw, 2)) return fail(BLVM_ERR_INVALID_ARGUMENT, "graw must be element aligned");
  DmolArgs A{};
  A.y = y; A.raw = raw; A.x_sl = x_sl; A.gout = gout; A.gscale = gscale; A.gscale_dev = gscale_dev; A.lp = lp; A.graw = graw;
  A.partials = partials; A.err_flag = err_flag; A.B = B; A.T = T; A.chunks = blvm_dmol_chunks(T, K, D); A.K = K; A.D = D;
  A.flags = flags; A.C = make_consts(num_bins, log_epsilon);
  return dispatch_dmol<true>(A, raw_dtype, static_cast<cudaStream_t>(stream));
}

int64_t blvm_gmm_chunks(int64_t T, int K, int D) {
  const int64_t ts = gmm_has_register_kernel(K, D) ? dmol_tile_samples(K, D) : kTile;
  return (T + ts - 1) / ts;
}

int blvm_gmm_fwd_grad(const float* y, const float* raw, const int64_t* x_sl, const float* gout, float gscale,
                      const double* gscale_dev, int64_t B, int64_t T, int K, int D, int from_raw, double softplus_beta,
                      double sd_add, double sd_floor, int flags, float* lp, float* graw, double* partials,
                      blvm_stream_t stream) {
  if (int rc = validate_dmol(y, raw, B, T, K, D, 2)) return rc;
  if (from_raw && !(softplus_beta > 0.0)) return fail(BLVM_ERR_INVALID_ARGUMENT, "softplus_beta=%g must be > 0", softplus_beta);
  DmolArgs A{};
  A.y = y; A.raw = raw; A.x_sl = x_sl; A.gout = gout; A.gscale = gscale; A.gscale_dev = gscale_dev; A.lp = lp; A.graw = graw;
  A.partials = partials; A.err_flag = nullptr; A.B = B; A.T = T; A.chunks = blvm_gmm_chunks(T, K, D); A.K = K; A.D = D;
  A.flags = flags; A.lik = from_raw ? kLikGmmRaw : kLikGmmSd;
  A.C = make_consts(256, -7.0f);
  A.C.sp_beta = static_cast<float>(softplus_beta); A.C.sp_inv_beta = static_cast<float>(1.0 / (from_raw ? softplus_beta : 1.0));
  A.C.sd_add = static_cast<float>(sd_add); A.C.sd_floor = static_cast<float>(sd_floor);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return graw ? dispatch_gmm<true>(A, st) : dispatch_gmm<false>(A, st);
}

int blvm_gaussian_ll(const float* y, const float* mu, const float* sd, const float* gout, int64_t n, double sd_floor,
                     float* lp, float* g_mu, float* g_sd, blvm_stream_t stream) {
  if (n < 0) return fail(BLVM_ERR_INVALID_ARGUMENT, "negative n");
  if (n > 0 && (!y || !mu || !sd)) return fail(BLVM_ERR_INVALID_ARGUMENT, "null input");
  if ((g_mu == nullptr) != (g_sd == nullptr)) return fail(BLVM_ERR_INVALID_ARGUMENT, "g_mu and g_sd must be given together");
  if (n == 0) return BLVM_OK;
  DmolConsts C = make_consts(256, -7.0f);
  C.sd_floor = static_cast<float>(sd_floor);
  const int64_t want = (n + 1023) / 1024;
  const unsigned blocks = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_mu) gaussian_ll_kernel<true><<<blocks, 256, 0, st>>>(y, mu, sd, gout, n, C, lp, g_mu, g_sd);
  else gaussian_ll_kernel<false><<<blocks, 256, 0, st>>>(y, mu, sd, gout, n, C, lp, nullptr, nullptr);
  return check_launch("gaussian_ll_kernel");
}

int blvm_dl_fwd_grad(const float* y, const float* raw, const int64_t* x_sl, const float* gout, float gscale, int64_t B,
                     int64_t T, int num_bins, float log_epsilon, int flags, float* lp, float* graw, double* partials,
                     int* err_flag, blvm_stream_t stream) {
  if (int rc = validate_dmol(y, raw, B, T, 1, 1, num_bins)) return rc;
  if (!aligned(raw, 8) || !aligned(graw, 8)) return fail(BLVM_ERR_INVALID_ARGUMENT, "raw/graw must be 8-byte aligned");
  DmolArgs A{};
  A.y = y; A.raw = raw; A.x_sl = x_sl; A.gout = gout; A.gscale = gscale; A.lp = lp; A.graw = graw;
  A.partials = partials; A.err_flag = err_flag; A.B = B; A.T = T; A.chunks = blvm_dl_chunks(T); A.K = 1; A.D = 1;
  A.flags = flags; A.C = make_consts(num_bins, log_epsilon);
  const int64_t tiles = A.B * A.chunks;
  if (tiles == 0) return BLVM_OK;
  if (tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (graw)
    dl_kernel<kTile, true><<<static_cast<unsigned>(tiles), kTile, 0, st>>>(A);
  else
    dl_kernel<kTile, false><<<static_cast<unsigned>(tiles), kTile, 0, st>>>(A);
  return check_launch("dl_kernel");
}

static bool kl_vec_ok(const KlArgs& A, bool grad) {
  bool vec = (A.row_elems % 4 == 0) && aligned(A.mu_q, 16) && aligned(A.sd_q, 16) && aligned(A.mu_p, 16) &&
             aligned(A.sd_p, 16) && aligned(A.kl, 16);
  if (grad) vec = vec && aligned(A.g_mu_q, 16) && aligned(A.g_sd_q, 16) && aligned(A.g_mu_p, 16) && aligned(A.g_sd_p, 16);
  return vec;
}

static int launch_kl(KlArgs& A, bool grad, cudaStream_t st, bool overlap_prev = false) {
  A.chunks = blvm_kl_chunks(A.row_elems);
  const int64_t tiles = A.B * A.chunks;
  if (tiles == 0) return BLVM_OK;
  if (tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles");
  A.vec = kl_vec_ok(A, grad) ? 1 : 0;
  const unsigned g = static_cast<unsigned>(tiles);
  const bool pdl = overlap_prev && pdl_enabled();
  const cudaError_t e = grad ? launch_ex(kl_kernel<true>, g, kKlTPB, 0, st, pdl, A) : launch_ex(kl_kernel<false>, g, kKlTPB, 0, st, pdl, A);
  if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "kl_kernel: %s", cudaGetErrorString(e));
  return check_launch("kl_kernel");
}

int blvm_kl_gaussian_fwd(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, int64_t n, float* kl,
                         blvm_stream_t stream) {
  if (n < 0) return fail(BLVM_ERR_INVALID_ARGUMENT, "negative n");
  if (n > 0 && (!mu_q || !sd_q || !mu_p || !sd_p || !kl)) return fail(BLVM_ERR_INVALID_ARGUMENT, "null pointer");
  KlArgs A{};
  A.mu_q = mu_q; A.sd_q = sd_q; A.mu_p = mu_p; A.sd_p = sd_p; A.kl = kl; A.B = 1; A.row_elems = n; A.Z = 1;
  return launch_kl(A, false, static_cast<cudaStream_t>(stream));
}

int blvm_kl_gaussian_bwd(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, const float* gout,
                         int64_t n, float* g_mu_q, float* g_sd_q, float* g_mu_p, float* g_sd_p, blvm_stream_t stream) {
  if (n < 0) return fail(BLVM_ERR_INVALID_ARGUMENT, "negative n");
  if (n > 0 && (!mu_q || !sd_q || !mu_p || !sd_p || !gout || !g_mu_q || !g_sd_q || !g_mu_p || !g_sd_p))
    return fail(BLVM_ERR_INVALID_ARGUMENT, "null pointer");
  KlArgs A{};
  A.mu_q = mu_q; A.sd_q = sd_q; A.mu_p = mu_p; A.sd_p = sd_p; A.gout = gout; A.gscale = 1.f;
  A.g_mu_q = g_mu_q; A.g_sd_q = g_sd_q; A.g_mu_p = g_mu_p; A.g_sd_p = g_sd_p; A.B = 1; A.row_elems = n; A.Z = 1;
  return launch_kl(A, true, static_cast<cudaStream_t>(stream));
}

int blvm_kl_elbo_fwd_grad(const float* mu_q, const float* sd_q, const float* mu_p, const float* sd_p, const int64_t* lens,
                          int64_t B, int64_t Tz, int64_t Z, double free_nats, float gscale, float* kl, float* g_mu_q,
                          float* g_sd_q, float* g_mu_p, float* g_sd_p, double* part_kl, double* part_klfn, int flags,
                          blvm_stream_t stream) {
  if (B < 0 || Tz < 0 || Z < 1) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad shape B=%lld Tz=%lld Z=%lld", (long long)B, (long long)Tz, (long long)Z);
  if (B * Tz > 0 && (!mu_q || !sd_q || !mu_p || !sd_p)) return fail(BLVM_ERR_INVALID_ARGUMENT, "null input");
  if (!part_kl || !part_klfn) return fail(BLVM_ERR_INVALID_ARGUMENT, "null partials");
  const bool grad = g_mu_q != nullptr;
  if (grad && (!g_sd_q || !g_mu_p || !g_sd_p)) return fail(BLVM_ERR_INVALID_ARGUMENT, "gradient outputs must be given together");
  KlArgs A{};
  A.mu_q = mu_q; A.sd_q = sd_q; A.mu_p = mu_p; A.sd_p = sd_p; A.lens = lens; A.gscale = gscale;
  A.fn_enabled = (free_nats != 0.0) ? 1 : 0;
  A.min_kl = static_cast<float>(free_nats / static_cast<double>(Z));  // python float / int, then torch.tensor(..., fp32)
  A.kl = kl; A.g_mu_q = g_mu_q; A.g_sd_q = g_sd_q; A.g_mu_p = g_mu_p; A.g_sd_p = g_sd_p;
  A.part_kl = part_kl; A.part_klfn = part_klfn; A.B = B; A.row_elems = Tz * Z; A.Z = Z;
  return launch_kl(A, grad, static_cast<cudaStream_t>(stream), (flags & BLVM_FLAG_OVERLAP_PREV) != 0);
}

int blvm_kl_reduce_fwd_grad(const float* kl, const int64_t* lens, int64_t B, int64_t Tz, int64_t Z, double free_nats,
                            float gscale, float* gkl, double* part_kl, double* part_klfn, int flags, blvm_stream_t stream) {
  if (B < 0 || Tz < 0 || Z < 1) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad shape");
  if (B * Tz > 0 && !kl) return fail(BLVM_ERR_INVALID_ARGUMENT, "null kl");
  if (!part_kl || !part_klfn) return fail(BLVM_ERR_INVALID_ARGUMENT, "null partials");
  KlArgs A{};
  A.kl_in = kl; A.lens = lens; A.gscale = gscale; A.fn_enabled = (free_nats != 0.0) ? 1 : 0;
  A.min_kl = static_cast<float>(free_nats / static_cast<double>(Z));
  A.gkl = gkl; A.part_kl = part_kl; A.part_klfn = part_klfn; A.B = B; A.row_elems = Tz * Z; A.Z = Z;
  A.chunks = blvm_kl_chunks(A.row_elems);
  const int64_t tiles = B * A.chunks;
  if (tiles == 0) return BLVM_OK;
  if (tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles");
  const bool pdl = (flags & BLVM_FLAG_OVERLAP_PREV) != 0 && pdl_enabled();
  const cudaError_t e = launch_ex(kl_kernel<false>, static_cast<unsigned>(tiles), kKlTPB, 0, static_cast<cudaStream_t>(stream), pdl, A);
  if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "kl_kernel (materialised KL): %s", cudaGetErrorString(e));
  return check_launch("kl_kernel (materialised KL)");
}

// Validate the level descriptors and lay out the concatenated tile ranges (shared by the level-array entry point and the
// single-launch step).
static int build_kl_multi(const blvm_kl_level_t* levels_host, int n_levels, int64_t B, float gscale, const double* gscale_dev, KlMultiArgs& M,
                          bool& any_grad, int64_t& total) {
  if (n_levels < 1 || n_levels > kMaxLevels) return fail(BLVM_ERR_INVALID_ARGUMENT, "n_levels=%d out of [1, %d]", n_levels, kMaxLevels);
  if (!levels_host || B < 0) return fail(BLVM_ERR_INVALID_ARGUMENT, "null levels / negative B");
  M.n_levels = n_levels;
  any_grad = false;
  total = 0;
  for (int l = 0; l < n_levels; ++l) {
    const blvm_kl_level_t& L = levels_host[l];
    if (L.Tz < 0 || L.Z < 1) return fail(BLVM_ERR_INVALID_ARGUMENT, "level %d: bad shape Tz=%lld Z=%lld", l, (long long)L.Tz, (long long)L.Z);
    if (!L.part_kl || !L.part_klfn) return fail(BLVM_ERR_INVALID_ARGUMENT, "level %d: null partials", l);
    KlArgs& A = M.level[l];
    A.lens = L.lens; A.gscale = gscale; A.gscale_dev = gscale_dev; A.fn_enabled = (L.free_nats != 0.0) ? 1 : 0;
    A.min_kl = static_cast<float>(L.free_nats / static_cast<double>(L.Z));
    A.part_kl = L.part_kl; A.part_klfn = L.part_klfn; A.B = B; A.row_elems = L.Tz * L.Z; A.Z = L.Z;
    A.chunks = blvm_kl_chunks(A.row_elems);
    bool grad = false;
    if (L.kl) {
      if (L.mu_q || L.sd_q || L.mu_p || L.sd_p) return fail(BLVM_ERR_INVALID_ARGUMENT, "level %d: give the four parameter tensors or kl, not both", l);
      A.kl_in = L.kl; A.gkl = L.g_kl;
    } else {
      if (B * L.Tz > 0 && (!L.mu_q || !L.sd_q || !L.mu_p || !L.sd_p)) return fail(BLVM_ERR_INVALID_ARGUMENT, "level %d: null input", l);
      grad = L.g_mu_q != nullptr;
      if (grad && (!L.g_sd_q || !L.g_mu_p || !L.g_sd_p)) return fail(BLVM_ERR_INVALID_ARGUMENT, "level %d: gradient outputs must be given together", l);
      A.mu_q = L.mu_q; A.sd_q = L.sd_q; A.mu_p = L.mu_p; A.sd_p = L.sd_p;
      A.g_mu_q = L.g_mu_q; A.g_sd_q = L.g_sd_q; A.g_mu_p = L.g_mu_p; A.g_sd_p = L.g_sd_p;
      A.z = L.z; A.g_z = L.z ? L.g_z : nullptr;
      A.vec = kl_vec_ok(A, grad) ? 1 : 0;
    }
    any_grad = any_grad || grad;
    M.tile_begin[l] = total;
    total += B * A.chunks;
  }
  M.tile_begin[n_levels] = total;
  if (total > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles");
  // a parameter level without gradient outputs inside a launch that computes gradients for another one would write through
  // null pointers: require consistency
  for (int l = 0; l < n_levels; ++l)
    if (!M.level[l].kl_in && any_grad && !M.level[l].g_mu_q) return fail(BLVM_ERR_INVALID_ARGUMENT, "level %d: gradient outputs missing", l);
  return BLVM_OK;
}

int blvm_kl_elbo_levels_fwd_grad(const blvm_kl_level_t* levels_host, int n_levels, int64_t B, float gscale, int flags,
                                 blvm_stream_t stream) {
  return blvm_kl_elbo_levels_fwd_grad_scaled(levels_host, n_levels, B, gscale, nullptr, flags, stream);
}

int blvm_kl_elbo_levels_fwd_grad_scaled(const blvm_kl_level_t* levels_host, int n_levels, int64_t B, float gscale, const double* gscale_dev,
                                        int flags, blvm_stream_t stream) {
  KlMultiArgs M{};
  bool any_grad = false;
  int64_t total = 0;
  if (int rc = build_kl_multi(levels_host, n_levels, B, gscale, gscale_dev, M, any_grad, total)) return rc;
  if (total == 0) return BLVM_OK;
  const bool pdl = (flags & BLVM_FLAG_OVERLAP_PREV) != 0 && pdl_enabled();
  const unsigned g = static_cast<unsigned>(total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e = any_grad ? launch_ex(kl_multi_kernel<true>, g, kKlTPB, 0, st, pdl, M) : launch_ex(kl_multi_kernel<false>, g, kKlTPB, 0, st, pdl, M);
  if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "kl_multi_kernel: %s", cudaGetErrorString(e));
  return check_launch("kl_multi_kernel");
}

static int finalize_impl(const double* logp_part, int64_t logp_chunks, const double* const* kl_part_host,
                         const double* const* klfn_part_host, const int64_t* kl_chunks_host, int n_levels,
                         const int64_t* x_sl, int64_t B, double beta, double denom, double* rows, double* scalars,
                         unsigned int* sync_counter, const ExchangeArgs& X, blvm_stream_t stream, int nansum_loss = 0) {
  if (n_levels < 0 || n_levels > kMaxLevels) return fail(BLVM_ERR_INVALID_ARGUMENT, "n_levels=%d out of [0, %d]", n_levels, kMaxLevels);
  if (B < 0 || !x_sl || !rows || !scalars || !sync_counter) return fail(BLVM_ERR_INVALID_ARGUMENT, "null x_sl/rows/scalars/sync_counter");
  FinalizeArgs A{};
  A.logp_part = logp_part; A.logp_chunks = logp_chunks; A.n_levels = n_levels; A.x_sl = x_sl; A.B = B; A.beta = beta;
  A.denom = denom; A.rows = rows; A.scalars = scalars; A.nansum_loss = nansum_loss;
  for (int l = 0; l < n_levels; ++l) {
    if (!kl_part_host[l] || !klfn_part_host[l]) return fail(BLVM_ERR_INVALID_ARGUMENT, "null KL partials at level %d", l);
    A.kl_part[l] = kl_part_host[l]; A.klfn_part[l] = klfn_part_host[l]; A.kl_chunks[l] = kl_chunks_host[l];
  }
  const unsigned blocks = static_cast<unsigned>(B > 0 ? (B + kFinWarps - 1) / kFinWarps : 1);
  // always a programmatic dependent: the kernel waits for its predecessor before its first read, so this is safe after
  // any kernel, and after a blvm kernel (which releases its dependents early) the CTAs are already resident when it ends
  if (B <= blvm_host::finalize_small_max_b()) {   // a handful of utterances: one CTA, no inter-CTA hand-off
    const cudaError_t e = launch_ex(elbo_finalize_small_kernel, 1u, kFinSmallTPB, 0, static_cast<cudaStream_t>(stream), pdl_enabled(), A, X);
    if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "elbo_finalize_small_kernel: %s", cudaGetErrorString(e));
    return check_launch("elbo_finalize_small_kernel");
  }
  const cudaError_t e = launch_ex(elbo_finalize_kernel, blocks, kFinTPB, 0, static_cast<cudaStream_t>(stream), pdl_enabled(), A, sync_counter, X);
  if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "elbo_finalize_kernel: %s", cudaGetErrorString(e));
  return check_launch("elbo_finalize_kernel");
}

int blvm_elbo_finalize(const double* logp_part, int64_t logp_chunks, const double* const* kl_part_host,
                       const double* const* klfn_part_host, const int64_t* kl_chunks_host, int n_levels,
                       const int64_t* x_sl, int64_t B, double beta, double denom, double* rows, double* scalars,
                       unsigned int* sync_counter, blvm_stream_t stream) {
  ExchangeArgs X{};
  return finalize_impl(logp_part, logp_chunks, kl_part_host, klfn_part_host, kl_chunks_host, n_levels, x_sl, B, beta, denom,
                       rows, scalars, sync_counter, X, stream);
}

int blvm_elbo_finalize_publish(const double* logp_part, int64_t logp_chunks, const double* const* kl_part_host,
                               const double* const* klfn_part_host, const int64_t* kl_chunks_host, int n_levels,
                               const int64_t* x_sl, int64_t B, double beta, double denom, double* rows, double* scalars,
                               unsigned int* sync_counter, void* const* peer_bases_host, int rank, int world,
                               unsigned long long* exchange_counters, double* prev_global_sums, int* err_flag,
                               blvm_stream_t stream) {
  if (world < 1 || world > kExMaxWorld || rank < 0 || rank >= world) return fail(BLVM_ERR_INVALID_ARGUMENT, "rank=%d world=%d (max %d)", rank, world, kExMaxWorld);
  if (!peer_bases_host || !exchange_counters) return fail(BLVM_ERR_INVALID_ARGUMENT, "null exchange buffers");
  ExchangeArgs X{};
  X.rank = rank; X.world = world; X.counters = exchange_counters; X.global_out = prev_global_sums; X.err = err_flag;
  for (int p = 0; p < world; ++p) {
    if (!peer_bases_host[p]) return fail(BLVM_ERR_INVALID_ARGUMENT, "null peer buffer %d", p);
    X.peer_base[p] = static_cast<double*>(peer_bases_host[p]);
  }
  return finalize_impl(logp_part, logp_chunks, kl_part_host, klfn_part_host, kl_chunks_host, n_levels, x_sl, B, beta, denom,
                       rows, scalars, sync_counter, X, stream);
}

int64_t blvm_exchange_buffer_bytes(void) { return static_cast<int64_t>(kExBufferBytes); }

// ---- the whole step from one call -----------------------------------------------------------------------------------
static int64_t step_logp_chunks(const blvm_elbo_step_t& S) {
  switch (S.likelihood) {
    case BLVM_LIK_DMOL: return blvm_dmol_chunks(S.T, S.K, S.D);
    case BLVM_LIK_DL: return blvm_dl_chunks(S.T);
    case BLVM_LIK_GMM: return blvm_gmm_chunks(S.T, S.K, S.D);
    default: return 0;
  }
}

static thread_local int g_last_step_launches = 0;
int blvm_last_step_launches(void) { return g_last_step_launches; }

int64_t blvm_elbo_step_workspace_doubles(const blvm_elbo_step_t* S) {
  if (!S || S->n_levels < 0 || S->n_levels > kMaxLevels || S->B < 0) return -1;
  int64_t n = 8 + (4 + static_cast<int64_t>(S->n_levels)) * S->B + S->B * step_logp_chunks(*S);
  for (int l = 0; l < S->n_levels; ++l) n += 2 * S->B * blvm_kl_chunks(S->levels[l].Tz * S->levels[l].Z);
  return n;
}

int blvm_elbo_step(const blvm_elbo_step_t* desc, blvm_stream_t stream) {
  if (!desc) return fail(BLVM_ERR_INVALID_ARGUMENT, "null descriptor");
  const blvm_elbo_step_t& S = *desc;
  const int L = S.n_levels;
  if (L < 0 || L > kMaxLevels) return fail(BLVM_ERR_INVALID_ARGUMENT, "n_levels=%d out of [0, %d]", L, kMaxLevels);
  if (S.B < 0 || !S.workspace || !S.x_sl || !S.sync_counter) return fail(BLVM_ERR_INVALID_ARGUMENT, "null workspace / x_sl / sync_counter or negative B");
  if (S.likelihood < BLVM_LIK_NONE || S.likelihood > BLVM_LIK_GMM) return fail(BLVM_ERR_INVALID_ARGUMENT, "likelihood=%d", S.likelihood);
  const bool has_lik = S.likelihood != BLVM_LIK_NONE;
  const bool lik_grad = has_lik && S.graw != nullptr;
  bool kl_grad = false;
  for (int l = 0; l < L; ++l) kl_grad = kl_grad || S.levels[l].g_mu_q != nullptr || S.levels[l].g_kl != nullptr;
  const double dn = S.denom;
  if ((lik_grad || kl_grad) && !(dn > 0.0)) return fail(BLVM_ERR_INVALID_ARGUMENT, "denom=%g must be > 0 when gradients are requested", dn);
  double* scalars = S.workspace;
  double* rows = S.workspace + 8;
  double* part = rows + (4 + static_cast<int64_t>(L)) * S.B;
  const int64_t logp_chunks = step_logp_chunks(S);
  double* logp_part = has_lik ? part : nullptr;
  part += S.B * logp_chunks;

  if (has_lik) {
    const int kflags = S.flags & (BLVM_FLAG_MASK_OUTPUT | BLVM_FLAG_SKIP_PADDED);
    const float gscale = lik_grad ? static_cast<float>(-1.0 / dn) : 0.f;
    int rc;
    if (S.likelihood == BLVM_LIK_DMOL) {
      rc = lik_grad ? blvm_dmol_fwd_grad(S.y, S.raw, S.raw_dtype, S.x_sl, nullptr, gscale, S.loss_scale, S.B, S.T, S.K, S.D, S.num_bins,
                                     S.log_epsilon, kflags, S.lp_twise, S.graw, logp_part, S.err_flag, stream)
                : blvm_dmol_fwd(S.y, S.raw, S.raw_dtype, S.x_sl, S.B, S.T, S.K, S.D, S.num_bins, S.log_epsilon, kflags, S.lp_twise,
                                logp_part, S.err_flag, stream);
    } else if (S.likelihood == BLVM_LIK_GMM) {
      if (S.raw_dtype != BLVM_DTYPE_F32) return fail(BLVM_ERR_UNSUPPORTED, "Gaussian-mixture parameters must be fp32");
      rc = blvm_gmm_fwd_grad(S.y, static_cast<const float*>(S.raw), S.x_sl, nullptr, gscale, nullptr, S.B, S.T, S.K, S.D, 1,
                             S.gmm_softplus_beta, S.gmm_sd_add, 0.0, kflags, S.lp_twise, static_cast<float*>(S.graw), logp_part, stream);
    } else {
      if (S.raw_dtype != BLVM_DTYPE_F32) return fail(BLVM_ERR_UNSUPPORTED, "discretized-logistic parameters must be fp32");
      rc = blvm_dl_fwd_grad(S.y, static_cast<const float*>(S.raw), S.x_sl, nullptr, gscale, S.B, S.T, S.num_bins, S.log_epsilon, kflags,
                            S.lp_twise, static_cast<float*>(S.graw), logp_part, S.err_flag, stream);
    }
    if (rc) return rc;
  }

  const double* kl_part[kMaxLevels];
  const double* klfn_part[kMaxLevels];
  int64_t kl_chunks[kMaxLevels];
  if (L > 0) {
    blvm_kl_level_t lv[kMaxLevels];
    for (int l = 0; l < L; ++l) {
      lv[l] = S.levels[l];
      kl_chunks[l] = blvm_kl_chunks(lv[l].Tz * lv[l].Z);
      lv[l].part_kl = part;
      lv[l].part_klfn = part + S.B * kl_chunks[l];
      kl_part[l] = lv[l].part_kl;
      klfn_part[l] = lv[l].part_klfn;
      part += 2 * S.B * kl_chunks[l];
    }
    // the launch just before is this step's likelihood kernel, which produces none of the KL inputs: overlap its tail
    // the loss scale (fp16 AMP) is applied to the KL gradients here as it is to the likelihood's: no rescale pass in the backward
    const int rc = blvm_kl_elbo_levels_fwd_grad_scaled(lv, L, S.B, kl_grad ? static_cast<float>(S.beta / dn) : 0.f, S.loss_scale,
                                                       has_lik ? BLVM_FLAG_OVERLAP_PREV : 0, stream);
    if (rc) return rc;
  }

  ExchangeArgs X{};
  if (S.world > 0) {
    if (S.world > kExMaxWorld || S.rank < 0 || S.rank >= S.world) return fail(BLVM_ERR_INVALID_ARGUMENT, "rank=%d world=%d (max %d)", S.rank, S.world, kExMaxWorld);
    if (!S.peer_bases_host || !S.exchange_counters) return fail(BLVM_ERR_INVALID_ARGUMENT, "null exchange buffers");
    X.rank = S.rank; X.world = S.world; X.counters = S.exchange_counters; X.global_out = S.prev_global_sums; X.err = S.exchange_err;
    for (int p = 0; p < S.world; ++p) {
      if (!S.peer_bases_host[p]) return fail(BLVM_ERR_INVALID_ARGUMENT, "null peer buffer %d", p);
      X.peer_base[p] = static_cast<double*>(S.peer_bases_host[p]);
    }
  }
  const int nansum = (S.flags & BLVM_FLAG_NANSUM_LOSS) ? 1 : 0;
  g_last_step_launches = (has_lik ? 1 : 0) + (L > 0 ? 1 : 0) + 1 + ((nansum && lik_grad) ? 1 : 0);
  if (int rc = finalize_impl(logp_part, logp_chunks, kl_part, klfn_part, kl_chunks, L, S.x_sl, S.B, S.beta, dn, rows, scalars,
                             S.sync_counter, X, stream, nansum))
    return rc;
  if (nansum && lik_grad) {
    const int P = S.likelihood == BLVM_LIK_DL ? 2 : S.K * (2 * S.D + 1);
    return blvm_row_gate_inplace(S.graw, S.raw_dtype, S.B, S.T * P, rows /* row 0 = log p per utterance */, stream);
  }
  return BLVM_OK;
}

int blvm_row_gate_inplace(void* buf, int dtype, int64_t B, int64_t row_elems, const double* row_values, blvm_stream_t stream) {
  if (B < 0 || row_elems < 0 || (B * row_elems > 0 && (!buf || !row_values))) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad arguments");
  if (B * row_elems == 0) return BLVM_OK;
  const int64_t chunks = (row_elems + kRowGateChunk - 1) / kRowGateChunk;
  if (B * chunks > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles");
  const unsigned g = static_cast<unsigned>(B * chunks), c = static_cast<unsigned>(chunks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case BLVM_DTYPE_F32: row_gate_kernel<float><<<g, 256, 0, st>>>(static_cast<float*>(buf), row_elems, c, row_values); break;
    case BLVM_DTYPE_F16: row_gate_kernel<__half><<<g, 256, 0, st>>>(static_cast<__half*>(buf), row_elems, c, row_values); break;
    case BLVM_DTYPE_BF16: row_gate_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(static_cast<__nv_bfloat16*>(buf), row_elems, c, row_values); break;
    default: return fail(BLVM_ERR_INVALID_ARGUMENT, "dtype=%d", dtype);
  }
  return check_launch("row_gate_kernel");
}

int blvm_exchange_consume(void* local_base, int world, unsigned long long* exchange_counters, int lag, double beta,
                          double* out_sums, int* err_flag, blvm_stream_t stream) {
  if (!local_base || !exchange_counters || !out_sums) return fail(BLVM_ERR_INVALID_ARGUMENT, "null pointer");
  if (world < 1 || world > kExMaxWorld || lag < 0) return fail(BLVM_ERR_INVALID_ARGUMENT, "world=%d lag=%d", world, lag);
  exchange_consume_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<double*>(local_base), world,
                                                                          exchange_counters, lag, beta, out_sums, err_flag);
  return check_launch("exchange_consume_kernel");
}

int blvm_quantize(const float* x, int64_t n, const float* boundaries, int64_t n_bins, int64_t* out, blvm_stream_t stream) {
  if (n < 0 || n_bins < 1) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad sizes");
  if (n > 0 && (!x || !boundaries || !out)) return fail(BLVM_ERR_INVALID_ARGUMENT, "null pointer");
  if (n == 0) return BLVM_OK;
  const int64_t blocks = (n + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many elements");
  quantize_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, boundaries, n_bins, out);
  return check_launch("quantize_kernel");
}

int blvm_dmol_sample_mode(const void* raw, int raw_dtype, int64_t N, int K, int D, float log_epsilon, uint64_t seed,
                          uint64_t offset, float* sample, float* mode, int32_t* mode_index, blvm_stream_t stream) {
  if (N < 0 || K < 1 || D < 1) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad shape N=%lld K=%d D=%d", (long long)N, K, D);
  if (N > 0 && !raw) return fail(BLVM_ERR_INVALID_ARGUMENT, "null raw");
  if (N == 0) return BLVM_OK;
  SampleArgs A{};
  A.raw = raw; A.N = N; A.K = K; A.D = D; A.log_eps = log_epsilon; A.seed = seed; A.offset = offset;
  A.sample = sample; A.mode = mode; A.mode_index = mode_index;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // register-kernel shapes whose slabs are 16-byte aligned: TMA-staged tiles; both kernels draw identical samples
  const int64_t esz = raw_dtype == BLVM_DTYPE_F32 ? 4 : 2;
  if (D == 1 && dmol_has_register_kernel(K, D) && aligned(raw, 16) && (N * 3 * K * esz) % 16 == 0) {
    const int64_t tile = dmol_tile_samples(K, D);
    const int64_t tiles = (N + tile - 1) / tile;
    if (tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many samples");
    switch (raw_dtype) {
      case BLVM_DTYPE_F32: return sample_dispatch_tp<float>(A, tiles, st);
      case BLVM_DTYPE_F16: return sample_dispatch_tp<__half>(A, tiles, st);
      case BLVM_DTYPE_BF16: return sample_dispatch_tp<__nv_bfloat16>(A, tiles, st);
      default: return fail(BLVM_ERR_INVALID_ARGUMENT, "raw_dtype=%d", raw_dtype);
    }
  }
  const int64_t blocks = (N + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many samples");
  const unsigned g = static_cast<unsigned>(blocks);
  switch (raw_dtype) {
    case BLVM_DTYPE_F32: dmol_sample_mode_kernel<float><<<g, 256, 0, st>>>(A); break;
    case BLVM_DTYPE_F16: dmol_sample_mode_kernel<__half><<<g, 256, 0, st>>>(A); break;
    case BLVM_DTYPE_BF16: dmol_sample_mode_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(A); break;
    default: return fail(BLVM_ERR_INVALID_ARGUMENT, "raw_dtype=%d", raw_dtype);
  }
  return check_launch("dmol_sample_mode_kernel");
}

int blvm_scale_inplace(float* buf, int64_t n, const double* scale, blvm_stream_t stream) {
  if (n < 0 || (n > 0 && (!buf || !scale))) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad arguments");
  if (n == 0) return BLVM_OK;
  const int64_t want = (n + 1023) / 1024;
  const unsigned blocks = static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8);
  scale_inplace_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(buf, n, scale);
  return check_launch("scale_inplace_kernel");
}

int blvm_scale_inplace_multi(void* const* bufs_host, const int64_t* ns_host, const int* dtypes_host, int count,
                             const double* scale, blvm_stream_t stream) {
  if (count < 0 || count > kMaxScaleBuffers) return fail(BLVM_ERR_INVALID_ARGUMENT, "count=%d out of [0, %d]", count, kMaxScaleBuffers);
  if (count == 0) return BLVM_OK;
  if (!scale) return fail(BLVM_ERR_INVALID_ARGUMENT, "null scale");
  ScaleMultiArgs A{};
  A.count = count; A.scale = scale;
  int64_t nmax = 0;
  for (int i = 0; i < count; ++i) {
    if (ns_host[i] < 0 || (ns_host[i] > 0 && !bufs_host[i])) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad buffer %d", i);
    A.buf[i] = bufs_host[i]; A.n[i] = ns_host[i]; A.dtype[i] = dtypes_host ? dtypes_host[i] : 0;
    if (A.dtype[i] < 0 || A.dtype[i] > 2) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad dtype for buffer %d", i);
    nmax = ns_host[i] > nmax ? ns_host[i] : nmax;
  }
  const int64_t want = (nmax + 1023) / 1024;   // a CTA iteration covers 256 threads x 4 elements
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  const unsigned grid = static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
  scale_inplace_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  return check_launch("scale_inplace_multi_kernel");
}

}  // extern "C"
