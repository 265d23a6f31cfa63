// C ABI of the fused likelihood head (linear_dmol_kernel.cuh): nn.Linear(x_dim -> 3K) + DMoL value / gradient + the Linear's
// backward on the tcgen05 tensor cores.  Validation + launch only: no allocation, no synchronisation.
#include "../../include/blvm_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "host_common.h"
#include "linear_dmol_kernel.cuh"

using namespace blvm;
using blvm_host::aligned;
using blvm_host::check_launch;
using blvm_host::current_device;
using blvm_host::fail;
using blvm_host::kMaxDevices;

namespace blvm_host {
int sm_count();   // blvm_b200.cu
}

namespace {

constexpr int kLinearK = 10;   // num_mix of every audio model of the reference (vrnn.py:467, srnn.py:436, stcn.py:199, ...)

int padded_dim(int64_t Din) {   // smallest instantiated DP with DP >= Din + 1 (the ones column), 0 = unsupported
  if (Din < 1 || Din % 2 != 0) return 0;
  if (Din + 1 <= 32) return 32;
  if (Din + 1 <= 80) return 80;
  return 0;
}

template <int DP, bool GRAD, int UMODE, typename TP, int DIN = 0>
int launch_linear(const LinearDmolArgs& A, cudaStream_t st, unsigned* grid_out) {
  auto kern = linear_dmol_kernel<kLinearK, DP, GRAD, UMODE, TP, DIN>;
  constexpr int smem = LinearSmem<DP>::bytes;
  constexpr int tmem_cols = LinearSmem<DP>::tmem_cols;
  static int resident_dev[kMaxDevices] = {};
  int& resident = resident_dev[current_device()];
  if (resident == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // as many CTAs as shared memory can hold
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem);
    if (e != cudaSuccess || occ < 1) return fail(BLVM_ERR_CUDA, "occupancy query (smem=%d): %s", smem, cudaGetErrorString(e));
    // The occupancy API answers 1 for any kernel that allocates tensor memory; the real limits are the tensor-memory columns (a
    // CTA keeps them for its whole persistent life), shared memory and the 64-register budget (8 CTAs of 128 threads)
    const int by_tmem = 512 / tmem_cols, by_smem = (227 * 1024) / (smem + 1024);
    resident = by_tmem < by_smem ? by_tmem : by_smem;
    if (resident > 8) resident = 8;   // the register budget: 64 registers x 128 threads (__launch_bounds__(128, 8) for DP = 32)
    if (resident < 1) resident = 1;
    if (const char* e = getenv("BLVM_B200_LINEAR_CTAS_PER_SM")) {   // A/B knob
      const int v = atoi(e);
      if (v > 0) resident = v;
    }
    if (getenv("BLVM_B200_DEBUG")) fprintf(stderr, "[blvm] linear_dmol_kernel<DP=%d>: occupancy %d CTAs/SM, tensor memory allows %d, using %d\n", DP, occ, by_tmem, resident);
  }
  const int64_t slots = static_cast<int64_t>(blvm_host::sm_count()) * resident;
  const unsigned grid = static_cast<unsigned>(A.tiles < slots ? A.tiles : slots);
  if (grid_out) *grid_out = grid;
  if (A.tiles == 0) return BLVM_OK;
  kern<<<grid, 128, smem, st>>>(A);
  return check_launch("linear_dmol_kernel");
}

template <int DP, typename TP>
int dispatch_mode(const LinearDmolArgs& A, bool grad, cudaStream_t st, unsigned* grid_out) {
  const bool tiny = blvm_host::u_is_tiny(A.C);
  if constexpr (DP == 32) {   // the reference's head: x_dim = 3 * num_mix = 30, 16-bit audio, training step
    if (A.Din == 30 && tiny && grad) return launch_linear<DP, true, kUTiny, TP, 30>(A, st, grid_out);
  }
  if (grad) return tiny ? launch_linear<DP, true, kUTiny, TP>(A, st, grid_out) : launch_linear<DP, true, kUGeneral, TP>(A, st, grid_out);
  return tiny ? launch_linear<DP, false, kUTiny, TP>(A, st, grid_out) : launch_linear<DP, false, kUGeneral, TP>(A, st, grid_out);
}

template <typename TP>
int dispatch_dp(const LinearDmolArgs& A, int DP, bool grad, cudaStream_t st, unsigned* grid_out) {
  switch (DP) {
    case 32: return dispatch_mode<32, TP>(A, grad, st, grid_out);
    case 80: return dispatch_mode<80, TP>(A, grad, st, grid_out);
    default: return fail(BLVM_ERR_UNSUPPORTED, "x_dim=%d has no fused-head kernel", A.Din);
  }
}

}  // namespace

extern "C" {

int blvm_linear_dmol_padded_dim(int K, int64_t Din) { return K == kLinearK ? padded_dim(Din) : 0; }

int64_t blvm_linear_dmol_max_ctas(void) { return static_cast<int64_t>(blvm_host::sm_count()) * 8; }

int blvm_linear_dmol_fwd_grad(const float* y, const void* x, const void* W, const float* bias, int dtype, const int64_t* x_sl,
                              float gscale, const double* gscale_dev, int64_t B, int64_t T, int64_t Din, int K, int num_bins,
                              float log_epsilon, int flags, float* lp, void* dx, float* dw_partial, int64_t dw_partial_ctas,
                              double* partials, int* err_flag, float* raw_debug, int64_t* ctas_used_host, blvm_stream_t stream) {
  if (B < 0 || T < 0 || T > 0x7fffffffLL) return fail(BLVM_ERR_INVALID_ARGUMENT, "bad shape B=%lld T=%lld", (long long)B, (long long)T);
  if (K != kLinearK) return fail(BLVM_ERR_UNSUPPORTED, "the fused head is instantiated for num_mix=%d (got %d)", kLinearK, K);
  const int DP = padded_dim(Din);
  if (DP == 0) return fail(BLVM_ERR_UNSUPPORTED, "x_dim=%lld: the fused head needs an even x_dim <= 79", (long long)Din);
  if (dtype != BLVM_DTYPE_F16 && dtype != BLVM_DTYPE_BF16) return fail(BLVM_ERR_UNSUPPORTED, "the fused head takes fp16 / bf16 activations and weights (AMP)");
  if (num_bins < 2) return fail(BLVM_ERR_INVALID_ARGUMENT, "num_bins=%d", num_bins);
  if (B * T > 0 && (!y || !x || !W)) return fail(BLVM_ERR_INVALID_ARGUMENT, "null y / x / W");
  if (!aligned(x, 4) || !aligned(W, 2) || !aligned(dx, 4)) return fail(BLVM_ERR_INVALID_ARGUMENT, "x / dx must be 4-byte, W 2-byte aligned");
  if ((dx == nullptr) != (dw_partial == nullptr)) return fail(BLVM_ERR_INVALID_ARGUMENT, "dx and dw_partial must be given together");
  LinearDmolArgs A{};
  A.y = y; A.x = x; A.W = W; A.bias = bias; A.x_sl = x_sl; A.gscale = gscale; A.gscale_dev = gscale_dev; A.lp = lp; A.dx = dx;
  A.dw_partial = dw_partial; A.partials = partials; A.err_flag = err_flag; A.raw_debug = raw_debug; A.B = B; A.T = T;
  A.chunks = (T + 127) / 128; A.tiles = B * A.chunks; A.Din = static_cast<int>(Din); A.flags = flags;
  A.C = blvm_host::make_consts(num_bins, log_epsilon);
  if (A.tiles > 0x7fffffffLL) return fail(BLVM_ERR_UNSUPPORTED, "too many tiles");
  if (dw_partial && dw_partial_ctas < blvm_linear_dmol_max_ctas()) return fail(BLVM_ERR_INVALID_ARGUMENT, "dw_partial needs room for %lld CTAs", (long long)blvm_linear_dmol_max_ctas());
  unsigned grid = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool grad = dx != nullptr;
  const int rc = dtype == BLVM_DTYPE_BF16 ? dispatch_dp<__nv_bfloat16>(A, DP, grad, st, &grid) : dispatch_dp<__half>(A, DP, grad, st, &grid);
  if (ctas_used_host) *ctas_used_host = grid;
  return rc;
}

int blvm_linear_dmol_reduce_dw(const float* dw_partial, int64_t ctas, int64_t Din, int K, float* dW, float* db, blvm_stream_t stream) {
  const int DP = K == kLinearK ? padded_dim(Din) : 0;
  if (DP == 0) return fail(BLVM_ERR_UNSUPPORTED, "unsupported K / x_dim");
  if (ctas < 0 || !dw_partial || !dW) return fail(BLVM_ERR_INVALID_ARGUMENT, "null pointer / negative count");
  const int P = 3 * K;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (DP == 32) linear_dmol_reduce_kernel<32><<<P, 1024, 0, st>>>(dw_partial, ctas, static_cast<int>(Din), P, dW, db);
  else linear_dmol_reduce_kernel<80><<<P, 1024, 0, st>>>(dw_partial, ctas, static_cast<int>(Din), P, dW, db);
  return check_launch("linear_dmol_reduce_kernel");
}

}  // extern "C"
