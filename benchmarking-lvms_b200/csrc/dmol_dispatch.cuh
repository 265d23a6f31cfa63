// Host-side launchers of the DMoL register kernels, templated on the parameter element type.  Included by one translation
// unit per element type (blvm_dmol_f32.cu / _f16.cu / _bf16.cu), each of which explicitly instantiates
// dmol_dispatch_tp<TP> and sample_dispatch_tp<TP>: the ~250 kernel instantiations compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "host_common.h"

#include "dmol_kernels.cuh"
#include "dmol_stream_kernel.cuh"
#include "sample_kernels.cuh"

namespace blvm_host {

using namespace blvm;

constexpr int kTile = BLVM_DMOL_TILE;

bool pdl_enabled();      // defined in blvm_b200.cu
int sm_count();
int stream_mode();

template <int K, bool GRAD, int UMODE, typename TP>
int launch_tile_mode(const DmolArgs& A, int64_t tiles, cudaStream_t st) {
  constexpr size_t smem = dmol_tile_smem_bytes<K, kTile, TP>();
  auto kern = dmol_tile_kernel<K, kTile, GRAD, UMODE, TP>;
  static bool configured[kMaxDevices] = {};  // per instantiation and device; benign race (idempotent attribute)
  const int dev = current_device();
  if (!configured[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured[dev] = true;
  }
  constexpr int G = DmolGroup<K, TP>::value;           // chunks per CTA (dmol_kernels.cuh)
  DmolArgs A2 = A;
  A2.ctas_per_row = (A.chunks + G - 1) / G;
  (void)tiles;
  kern<<<static_cast<unsigned>(A.B * A2.ctas_per_row), kTile, smem, st>>>(A2);
  return check_launch("dmol_tile_kernel");
}

// ---- persistent pipelined variant (dmol_stream_kernel.cuh) ------------------------------------------------------------
// Ring depth and CTA size.  Defaults: 2 stages / lookahead 1 / 128 threads (round-1 A/B, profiles/r1_ab_stream_kernel.log).  fp32 with
// K = 3..5 (32 KB stages, 3 CTAs per SM by shared memory at 2 stages): 256 threads and 3 stages -- twice the warps per CTA over two
// resident CTAs, the store of tile i-1 no longer in front of the load of tile i+1 -- measured K = 5 95.0 -> 88.7 us, K = 4 75.4 -> 71.5,
// K = 3 60.2 -> 58.3 (T = 64000, K = 5: 378 -> 356 us); 16-bit parameters and K <= 2 are faster with the default (bf16 K = 5 59.7 vs 63.9 us).
// -DBLVM_STREAM_STAGES / _LOOKAHEAD / _TPB override every instantiation (A/B builds).
#if defined(BLVM_STREAM_STAGES) || defined(BLVM_STREAM_TPB)
#ifndef BLVM_STREAM_STAGES
#define BLVM_STREAM_STAGES 2
#endif
#ifndef BLVM_STREAM_LOOKAHEAD
#define BLVM_STREAM_LOOKAHEAD 1
#endif
#ifndef BLVM_STREAM_TPB
#define BLVM_STREAM_TPB 128
#endif
template <int K, typename TP>
struct StreamCfg {
  static constexpr int tpb = BLVM_STREAM_TPB, stages = BLVM_STREAM_STAGES, lookahead = BLVM_STREAM_LOOKAHEAD;
};
#else
template <int K, typename TP>
struct StreamCfg {
  static constexpr bool wide = sizeof(TP) == 4 && K >= 3;
  static constexpr int tpb = wide ? 256 : 128, stages = wide ? 3 : 2, lookahead = 1;
};
#endif
#ifndef BLVM_STREAM_MAX_K
#define BLVM_STREAM_MAX_K 5      // K above this keeps the one-tile-per-CTA kernel (already at the HBM roofline)
#endif

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute (ptx_sm100.cuh: pdl_*).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}


template <typename TP>
bool stream_eligible(const DmolArgs& A, int K) {
  const int64_t row_bytes = A.T * 3 * K * static_cast<int64_t>(sizeof(TP));
  return A.T % 4 == 0 && row_bytes % 16 == 0 && aligned(A.raw, 16) && aligned(A.y, 16) && (!A.graw || aligned(A.graw, 16));
}

template <int K, bool GRAD, int UMODE, typename TP, int LIK = kLikDmol>
int launch_stream(const DmolArgs& A, int64_t tiles, cudaStream_t st) {
  using Cfg = StreamCfg<K, TP>;
  constexpr int TPB = (128 * DmolSpt<K>::value) % Cfg::tpb == 0 ? Cfg::tpb : 128;
  constexpr int S = Cfg::stages, LA = Cfg::lookahead;
  constexpr size_t smem = StreamLayout<K, TPB, TP>::bytes(S);
  auto kern = dmol_stream_kernel<K, TPB, S, LA, GRAD, UMODE, TP, LIK>;
  static int resident_dev[kMaxDevices] = {};  // CTAs per SM; per instantiation and device, benign race
  int& resident = resident_dev[current_device()];
  if (resident == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, smem);
    if (e != cudaSuccess || occ < 1) return fail(BLVM_ERR_CUDA, "occupancy query (smem=%zu): %s", smem, cudaGetErrorString(e));
    resident = occ;
  }
  const int64_t slots = static_cast<int64_t>(sm_count()) * resident;
  const unsigned grid = static_cast<unsigned>(tiles < slots ? tiles : slots);
  kern<<<grid, TPB, smem, st>>>(A, tiles);
  return check_launch("dmol_stream_kernel");
}

// u = h / s <= h * exp(-log_epsilon) for every element: if that bound is tiny (16-bit bins with the -7 clamp: 0.0167)
// the kernel specialisation without the large-u code is exact to O(u^4) ~ 1e-7 (blvm_math.cuh).
template <int K, bool GRAD, typename TP>
int launch_tile_dtype(const DmolArgs& A, int64_t tiles, cudaStream_t st) {
  if constexpr (K <= BLVM_STREAM_MAX_K) {
    if (stream_mode() && stream_eligible<TP>(A, K)) {
      if (blvm_host::u_is_tiny(A.C)) return launch_stream<K, GRAD, kUTiny, TP>(A, tiles, st);
      return launch_stream<K, GRAD, kUGeneral, TP>(A, tiles, st);
    }
  }
  if (blvm_host::u_is_tiny(A.C)) return launch_tile_mode<K, GRAD, kUTiny, TP>(A, tiles, st);
  return launch_tile_mode<K, GRAD, kUGeneral, TP>(A, tiles, st);
}

#define BLVM_FOR_EACH_K(X) X(1) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(12) X(16) X(20) X(30)

template <int K, typename TP>
int launch_sample_tile(const SampleArgs& A, int64_t tiles, cudaStream_t st) {
  constexpr size_t smem = ((size_t(kTile) * DmolSpt<K>::value * 3 * K * sizeof(TP) + 15) / 16) * 16 + 16;
  auto kern = dmol_sample_mode_tile_kernel<K, TP>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(BLVM_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured[dev] = true;
  }
  kern<<<static_cast<unsigned>(tiles), kTile, smem, st>>>(A);
  return check_launch("dmol_sample_mode_tile_kernel");
}


// ---- per element type: everything above behind two plain functions ---------------------------------------------------
template <typename TP>
int dmol_dispatch_tp(const DmolArgs& A, bool grad, int64_t tiles, cudaStream_t st) {
  switch (A.K) {
#define BLVM_CASE(KK) \
  case KK:            \
    return grad ? launch_tile_dtype<KK, true, TP>(A, tiles, st) : launch_tile_dtype<KK, false, TP>(A, tiles, st);
    BLVM_FOR_EACH_K(BLVM_CASE)
#undef BLVM_CASE
    default: return fail(BLVM_ERR_UNSUPPORTED, "no register kernel for K=%d", A.K);
  }
}

template <typename TP>
int sample_dispatch_tp(const SampleArgs& A, int64_t tiles, cudaStream_t st) {
  switch (A.K) {
#define BLVM_CASE(KK) \
  case KK:            \
    return launch_sample_tile<KK, TP>(A, tiles, st);
    BLVM_FOR_EACH_K(BLVM_CASE)
#undef BLVM_CASE
    default: return fail(BLVM_ERR_UNSUPPORTED, "no register kernel for K=%d", A.K);
  }
}

}  // namespace blvm_host
