// linear_dmol_kernel: the likelihood HEAD in one kernel for sm_100a -- nn.Linear(x_dim -> 3K) on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), the DMoL value + gradient in registers, and the Linear's backward (dx, dW, db) on the
// tensor cores again, without the (B, T, 3K) parameter tensor or its gradient ever touching HBM.
//
// Replaces, for 16-bit (AMP) activations, the chain
//   DiscretizedLogisticMixtureDense.forward  blvm/modules/distributions.py:381-387   raw = x W^T + b, split, clamp
//   discretized_logistic_mixture_ll          blvm/utils/log_likelihoods.py:170-231   + mask / row sums (vrnn.py:266-269)
//   autograd backward of both                                                         d raw, d x = d raw W, d W = d raw^T x, d b
// HBM traffic per sample (bf16, x_dim = 3K = 30): unfused 548 B (x 60 | raw 60 w + 60 r | graw 60 w + 2 x 60 r | x 60 r again |
// dx 60 | y, lp 8) -> fused 128 B (x 60, y 4, dx 60, lp 4).
//
// One persistent CTA of 128 threads walks tiles of 128 consecutive samples of one utterance; thread t owns sample t.
//   1. the tile's x slab lands in a linear staging buffer by one 1-D TMA bulk copy (UBLKCP); each thread re-lays its row out
//      into the no-swizzle UMMA "core matrix" layout (8 rows x 16 bytes), with a ones column appended (bias / db for free);
//   2. thread 0 issues  RAW[128 x 32] = X[128 x Dp] . W^T  (tcgen05.mma kind::f16, M = 128, N = 32), commit -> mbarrier;
//   3. every thread pulls ITS row of RAW out of TMEM (tcgen05.ld 32x32b.x32: lane = sample, 32 columns = the 3K parameters)
//      and runs the same dmol_sample<> as the tile kernel: value, masked fp64 row partial, gradient row G[t, 0:32];
//   4. G goes to shared memory as bf16/fp16 in core-matrix layout and thread 0 issues
//        DX[128 x Dp]  = G[128 x 32] . W          (A K-major, B MN-major: the SAME W buffer read transposed)
//        DW[64 x Dp]  += G^T[64 x 128] . X        (A, B MN-major: the SAME G and X buffers; accumulates over ALL tiles of the CTA)
//      A no-swizzle 8x8 core matrix stored [i][j] is K-major for (MN = i, K = j) and MN-major for (MN = j, K = i): no buffer is
//      ever transposed, only the descriptors' leading / stride offsets swap.
//   5. DX rows come back through tcgen05.ld, are rounded to the activation dtype and leave by one bulk store; after its last
//      tile the CTA writes its DW / db partial (a second tiny kernel adds the partials in CTA order: deterministic).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "blvm_math.cuh"
#include "dmol_kernels.cuh"
#include "ptx_sm100.cuh"

namespace blvm {

struct LinearDmolArgs {
  const float* y;            // (B*T)
  const void* x;             // (B*T, Din) bf16 / fp16
  const void* W;             // (P, Din)   same dtype (the autocast copy of params.weight), P = 3K
  const float* bias;         // (P) fp32, nullable
  const int64_t* x_sl;       // (B), nullable
  float gscale;
  const double* gscale_dev;  // nullable
  float* lp;                 // (B*T) nullable
  void* dx;                  // (B*T, Din) nullable (forward only)
  float* dw_partial;         // (gridDim.x, 32, DP) fp32, nullable with dx
  double* partials;          // (B, chunks) nullable
  int* err_flag;
  float* raw_debug;          // (B*T, 32) nullable: the tensor-core result (tests)
  int64_t B, T, chunks, tiles;
  int Din, flags;
  DmolConsts C;
};

namespace tc {   // tcgen05 / TMEM wrappers (PTX ISA 8.6, sm_100a)

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void ld_row32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start / leading / stride byte offsets in 16-byte units,
// version 1 = Blackwell, layout type 0).  K-major operand: `sbo` = distance between 8-row groups along M/N, `lbo` = distance between the
// two 8-element core matrices along K.  MN-major operand: `sbo` = distance between 8-element groups along M/N, `lbo` = distance between
// 8-element groups along K.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (sm_100)
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, 16-bit A/B, majors, N >> 3, M >> 4.
__host__ __device__ constexpr uint32_t instr_desc(int fmt16, int a_mn_major, int b_mn_major, int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(fmt16) << 7) | (static_cast<uint32_t>(fmt16) << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace tc

template <typename TP>
struct Fmt16;
template <>
struct Fmt16<__half> {
  static constexpr int value = 0;
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint16_t one() { return 0x3c00; }
};
template <>
struct Fmt16<__nv_bfloat16> {
  static constexpr int value = 1;
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint16_t one() { return 0x3f80; }
};

// Shared-memory plan (bytes), DP = padded x_dim (multiple of 16, >= Din + 1 for the ones column)
template <int DP>
struct LinearSmem {
  static constexpr int kStage = 128 * DP * 2;          // linear x slab in (Din <= DP - 1 columns used); a second one for the dx slab out
  static constexpr int kX = 128 * DP * 2;              // X  [128 s][DP d]   core-matrix layout
  static constexpr int kW = 32 * DP * 2;               // W  [32 p][DP d]
  static constexpr int kG = 128 * 32 * 2;              // G  [128 s][32 p]
  // G sits in FRONT of X: the M = 64 operand of the DW product addresses 64 columns of G (32 exist), i.e. reads up to 512 bytes past
  // the end of G -- these must be readable shared memory (the first rows of X; they only feed accumulator rows nobody reads)
  // When the dx slab fits (DP = 32: 7.5 KB <= 8 KB) it is staged in G's buffer, which is idle between this tile's DW product and the
  // next tile's gradient rows: 26 KB per CTA, 8 CTAs per SM.
  static constexpr bool kOutInG = kStage <= kG;
  static constexpr int oStage = 0, oOut = kOutInG ? oStage + kStage : oStage + kStage, oG = kOutInG ? oOut : oOut + kStage, oX = oG + kG,
                       oW = oX + kX, oBar = oW + kW, oMisc = oBar + 32;
  static constexpr int bytes = oMisc + 64;
  // tensor memory: RAW (32 columns) is dead once every thread has read its row, so DX (DP columns) reuses its columns; DW: DP more
  static constexpr int tmem_cols = 2 * DP <= 64 ? 64 : (2 * DP <= 128 ? 128 : (2 * DP <= 256 ? 256 : 512));
};

// byte offset of element (row i, col j) in a core-matrix buffer with `cols` 16-bit columns: 8 x 8 blocks of 128 bytes, row-blocks
// outermost; inside a block row-major (16 bytes per row)
__device__ __forceinline__ uint32_t cm_off(int i, int j, int cols) {
  return static_cast<uint32_t>(((i >> 3) * (cols >> 3) + (j >> 3)) * 128 + (i & 7) * 16 + (j & 7) * 2);
}

#ifndef BLVM_LINEAR_BACKOFF
#define BLVM_LINEAR_BACKOFF 0
#endif
#if BLVM_LINEAR_BACKOFF
#define BLVM_LINEAR_WAIT(bar, ph) ptx::mbar_wait_backoff(bar, ph)
#else
#define BLVM_LINEAR_WAIT(bar, ph) ptx::mbar_wait(bar, ph)
#endif
#ifndef BLVM_LINEAR_HEAD_LIN
#define BLVM_LINEAR_HEAD_LIN 0   // linear-domain sample evaluation in the fused head: 151 -> 143 us when no sample needs the log-domain
                                 // fallback, 164 us when most do (random weights); off: the head keeps one body
#endif
#ifndef BLVM_LINEAR_MINB
#define BLVM_LINEAR_MINB 8     // CTAs per SM the register allocation is capped for (DP = 32): 8 -> 64 registers.  Measured (B = 256 x
                               // 16000, x_dim 30, bf16, grid = cap x 148): 5 -> 183.6 us (96 registers), 7 -> 169.6, 8 -> 166.9
#endif
// DIN: x_dim known at compile time (30 = 3 * num_mix: every VRNN / SRNN / STCN / LSTM head of the reference, vrnn.py:464-469) so that
// the row re-layout and the dx packing unroll without per-word predicates; 0 = read it from the arguments.
template <int K, int DP, bool GRAD, int UMODE, typename TP, int DIN = 0>
__global__ void __launch_bounds__(128, LinearSmem<DP>::kOutInG ? BLVM_LINEAR_MINB : 2) linear_dmol_kernel(const __grid_constant__ LinearDmolArgs A) {
  static_assert(3 * K <= 32, "the parameter row must fit the N = 32 accumulator tile");
  static_assert(DP % 16 == 0 && DP >= 32 && DP <= 240, "padded x_dim");
  constexpr int P = 3 * K;
  using S = LinearSmem<DP>;
  constexpr uint32_t kTmemCols = S::tmem_cols;
  constexpr uint32_t XS = DP * 16;        // byte stride between 8-row blocks of the X / W buffers
  constexpr uint32_t GS = 32 * 16;        // ... of the G buffer
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* stage = smem + S::oStage;
  unsigned char* stage_out = smem + S::oOut;
  unsigned char* sX = smem + S::oX;
  unsigned char* sW = smem + S::oW;
  unsigned char* sG = smem + S::oG;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + S::oBar);       // TMA slab landed
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + S::oBar + 8);    // RAW resp. DX complete
  uint64_t* bar_dw = reinterpret_cast<uint64_t*>(smem + S::oBar + 16);    // the DW product of a tile complete (X and G reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::oMisc);
  double* scratch = reinterpret_cast<double*>(smem + S::oMisc + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Din = DIN ? DIN : A.Din;
  const TP* Wg = static_cast<const TP*>(A.W);

  // ---- one-time setup: TMEM, barriers, W (+ bias column) and the constant parts of X / G in core-matrix layout ----------
  if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    ptx::mbar_init(bar_load, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::mbar_init(bar_dw, 1);
    ptx::fence_mbar_init();
  }
  for (int i = tid; i < S::kW / 4; i += 128) reinterpret_cast<uint32_t*>(sW)[i] = 0u;
  for (int i = tid; i < S::kG / 4; i += 128) reinterpret_cast<uint32_t*>(sG)[i] = 0u;
  for (int i = tid; i < S::kX / 4; i += 128) reinterpret_cast<uint32_t*>(sX)[i] = 0u;
  __syncthreads();
  for (int i = tid; i < P * Din; i += 128) {
    const int p = i / Din, d = i - p * Din;
    *reinterpret_cast<TP*>(sW + cm_off(p, d, DP)) = Wg[i];
  }
  if (tid < P) {   // bias rides in column Din (the ones column of X)
    const float bv = A.bias ? A.bias[tid] : 0.f;
    const uint32_t w = Fmt16<TP>::pack(bv, 0.f);
    *reinterpret_cast<uint16_t*>(sW + cm_off(tid, Din, DP)) = static_cast<uint16_t>(w & 0xffffu);
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_raw = tmem, t_dx = tmem, t_dw = tmem + DP;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t x_row_off = cm_off(tid, 0, DP), g_row_off = cm_off(tid, 0, 32);   // this thread's row in the blocked X / G buffers

  constexpr uint32_t idesc_fwd = tc::instr_desc(Fmt16<TP>::value, 0, 0, 128, 32);
  constexpr uint32_t idesc_dx = tc::instr_desc(Fmt16<TP>::value, 0, 1, 128, DP);
  constexpr uint32_t idesc_dw = tc::instr_desc(Fmt16<TP>::value, 1, 1, 64, DP);
  const uint32_t aX = ptx::smem_u32(sX), aW = ptx::smem_u32(sW), aG = ptx::smem_u32(sG);

  float gs = A.gscale;
  if (GRAD && A.gscale_dev) gs *= static_cast<float>(*A.gscale_dev);
  const uint64_t pol = ptx::policy_evict_first();
  uint32_t ph_load = 0, ph_mma = 0, ph_dw = 0;
  bool dw_started = false, store_pending = false;
  const int row_bytes = Din * 2;

  // The x slab of a tile: contiguous n * Din 16-bit values.  `issue_load` starts its bulk copy into the (single) input stage; the
  // stage is free again as soon as every thread has re-laid its row out, so the NEXT tile's slab is requested right then and
  // lands while this tile is being evaluated.
  const unsigned chunks32 = static_cast<unsigned>(A.chunks);
  auto tile_geom = [&](int64_t tile, unsigned& b, int& c, int& n, int64_t& s0) {   // 32-bit: the host rejects more than 2^31 - 1 tiles
    b = static_cast<unsigned>(tile) / chunks32;
    c = static_cast<int>(static_cast<unsigned>(tile) - b * chunks32);
    n = min(128, static_cast<int>(A.T) - c * 128);
    s0 = static_cast<int64_t>(b) * A.T + c * 128;
  };
  auto issue_load = [&](int64_t tile) -> bool {   // thread 0 only; returns whether the slab travels by TMA
    unsigned b; int c, n; int64_t s0;
    tile_geom(tile, b, c, n, s0);
    const unsigned char* gsrc = static_cast<const unsigned char*>(A.x) + s0 * row_bytes;
    const uint32_t bytes = static_cast<uint32_t>(n) * row_bytes;
    if (((reinterpret_cast<uintptr_t>(gsrc) | bytes) & 15u) != 0) return false;
    ptx::mbar_arrive_expect_tx(bar_load, bytes);
    ptx::bulk_g2s(stage, gsrc, bytes, bar_load, pol);
    return true;
  };
  if (tid == 0 && static_cast<int64_t>(blockIdx.x) < A.tiles) issue_load(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
    unsigned b; int c, n; int64_t s0;
    tile_geom(tile, b, c, n, s0);
    const int t0 = c * 128;
    int64_t len64 = A.x_sl ? A.x_sl[b] : A.T;
    const int len = static_cast<int>(len64 < 0 ? 0 : (len64 > A.T ? A.T : len64));
    const int nvalid = max(0, min(n, len - t0));
    const unsigned char* gsrc = static_cast<const unsigned char*>(A.x) + s0 * row_bytes;
    const uint32_t bytes = static_cast<uint32_t>(n) * row_bytes;
    const bool bulk_in = ((reinterpret_cast<uintptr_t>(gsrc) | bytes) & 15u) == 0;

    // ---- 1. x slab -> staging (TMA, requested one tile ahead) -> core-matrix layout ---------------------------------------
    float yv = 0.f;
    if (tid < n) yv = ptx::ldg_stream(A.y + s0 + tid);
    if (bulk_in) {
      ptx::mbar_wait(bar_load, ph_load);
      ph_load ^= 1u;
    } else {
      for (uint32_t i = tid; i < bytes / 2; i += 128) reinterpret_cast<uint16_t*>(stage)[i] = reinterpret_cast<const uint16_t*>(gsrc)[i];
      __syncthreads();
    }
    if (GRAD && dw_started && !S::kOutInG) {   // the previous tile's DW product still reads X and G: it has had a whole epilogue to finish
      ptx::mbar_wait(bar_dw, ph_dw);
      ph_dw ^= 1u;
    }
    {
      // this thread's row as 32-bit words (Din is even: rows are 4-byte aligned), 16-byte chunks of 8 values into the blocked layout;
      // the word after the row carries the ones column (bias / db), the rest of the padding is zero
      const uint32_t* row = reinterpret_cast<const uint32_t*>(stage) + tid * (Din >> 1);
      const int nw = (tid < n) ? (Din >> 1) : -1;
      const uint32_t one = static_cast<uint32_t>(Fmt16<TP>::one());
#pragma unroll
      for (int cb = 0; cb < DP / 8; ++cb) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = 4 * cb + q;
          w[q] = (j < nw) ? row[j] : ((j == (Din >> 1)) ? one : 0u);
        }
        *reinterpret_cast<uint4*>(sX + x_row_off + cb * 128) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    ptx::fence_proxy_async_smem();     // generic-proxy writes of X (and, the first time, W / G) -> visible to the tensor core
    tc::fence_before();
    __syncthreads();

    // ---- 2. RAW = X . W^T ----------------------------------------------------------------------------------------------
    if (tid == 0) {
      if (tile + gridDim.x < A.tiles) issue_load(tile + gridDim.x);   // the input stage is free: request the next slab now
      tc::fence_after();
#pragma unroll
      for (int k = 0; k < DP / 16; ++k)
        tc::mma_f16(t_raw, tc::smem_desc(aX + k * 256, 128, XS), tc::smem_desc(aW + k * 256, 128, XS), idesc_fwd, k > 0);
      tc::commit(bar_mma);
    }
    BLVM_LINEAR_WAIT(bar_mma, ph_mma);
    ph_mma ^= 1u;
    tc::fence_after();

    // ---- 3. this thread's sample ---------------------------------------------------------------------------------------
    float r[32];
    tc::ld_row32(t_raw + lane_base, r);
    if (A.raw_debug && tid < n) {
#pragma unroll
      for (int q = 0; q < 32; ++q) A.raw_debug[(s0 + tid) * 32 + q] = r[q];
    }
    float L = 0.f;
    {
      float g = 0.f;
      if (tid < n) {
        if (!(yv <= 1.0f && yv >= -1.0f) && A.err_flag) atomicOr(A.err_flag, 1);
        if (GRAD) g = (tid < nvalid) ? gs : 0.f;
      }
      float rr[P];
#pragma unroll
      for (int q = 0; q < P; ++q) rr[q] = r[q];
      if constexpr (DmolEvalTraits<K, GRAD, UMODE, kLikDmol>::kLin && BLVM_LINEAR_HEAD_LIN) {
        // linear-domain evaluation; a sample it cannot represent (blvm_math.cuh) is re-read from tensor memory -- a warp-collective
        // load, so the whole warp re-reads when any lane needs it -- and evaluated in the log domain
        const bool ok = dmol_sample_lin<K, GRAD>(yv, rr, g, A.C, L);
        if (__any_sync(0xffffffffu, !ok)) {
          float r2[32];
          tc::ld_row32(t_raw + lane_base, r2);
          if (!ok) {
#pragma unroll
            for (int q = 0; q < P; ++q) rr[q] = r2[q];
            L = dmol_sample<K, GRAD, UMODE>(yv, rr, g, A.C);
          }
        }
      } else {
        L = dmol_sample<K, GRAD, UMODE>(yv, rr, g, A.C);
      }
      if (GRAD) {   // rows beyond the tile / the utterance carry g = 0: their gradient rows are exact zeros already
#pragma unroll
        for (int q = 0; q < P; ++q) r[q] = rr[q];
      }
    }
    const float Lm = (tid < nvalid) ? L : L * 0.0f;
    if (tid < n && A.lp) A.lp[s0 + tid] = (A.flags & kFlagMaskOutput) ? Lm : L;
    double acc = (tid < n) ? static_cast<double>(Lm) : 0.0;

    if (GRAD) {
      // ---- 4. G -> shared memory (core-matrix layout, 4 x 16 bytes per row), then DX and DW on the tensor cores ----------
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        uint4 v;
        v.x = Fmt16<TP>::pack(r[8 * cb + 0], r[8 * cb + 1]);
        v.y = Fmt16<TP>::pack(r[8 * cb + 2], r[8 * cb + 3]);
        v.z = Fmt16<TP>::pack(r[8 * cb + 4], (8 * cb + 5 < P) ? r[8 * cb + 5] : 0.f);
        v.w = Fmt16<TP>::pack((8 * cb + 6 < P) ? r[8 * cb + 6] : 0.f, (8 * cb + 7 < P) ? r[8 * cb + 7] : 0.f);
        *reinterpret_cast<uint4*>(sG + g_row_off + cb * 128) = v;
      }
      // the masked fp64 tile sum rides on the same barrier: warp sums now, thread 0 adds them after it (fixed order)
      if (A.partials) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) scratch[warp] = acc;
      }
      if (tid == 0 && store_pending) ptx::bulk_wait_read0();   // the previous tile's dx slab has long left the output stage
      ptx::fence_proxy_async_smem();
      tc::fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after();
#pragma unroll
        for (int k = 0; k < 2; ++k)      // K = p: 32 = 2 x 16
          tc::mma_f16(t_dx, tc::smem_desc(aG + k * 256, 128, GS), tc::smem_desc(aW + k * 2 * XS, XS, 128), idesc_dx, k > 0);
        tc::commit(bar_mma);             // DX is what the epilogue waits for; DW gets its own barrier (waited for by the next tile)
        if (A.partials) A.partials[static_cast<int64_t>(b) * A.chunks + c] = ((scratch[0] + scratch[1]) + scratch[2]) + scratch[3];
#pragma unroll
        // DW: M = 64 rows of G^T are addressed but G has 32 columns: rows 32-63 of the accumulator read the neighbouring row block
        // (finite garbage) and are never read back
#pragma unroll
        for (int k = 0; k < 8; ++k)      // K = samples: 128 = 8 x 16; accumulates over all tiles of this CTA
          tc::mma_f16(t_dw, tc::smem_desc(aG + k * 2 * GS, GS, 128), tc::smem_desc(aX + k * 2 * XS, XS, 128), idesc_dw, dw_started || k > 0);
        tc::commit(bar_dw);
      }
      dw_started = true;
      BLVM_LINEAR_WAIT(bar_mma, ph_mma);
      ph_mma ^= 1u;
      if constexpr (S::kOutInG) {        // the dx slab is staged in G's buffer: the DW product (8 small MMAs behind DX) must have read it
        ptx::mbar_wait(bar_dw, ph_dw);
        ph_dw ^= 1u;
      }
      tc::fence_after();

      // ---- 5. DX rows: TMEM -> registers -> activation dtype -> staging -> bulk store -----------------------------------
      unsigned char* gdst = static_cast<unsigned char*>(A.dx) + s0 * row_bytes;
      const bool bulk_out = ((reinterpret_cast<uintptr_t>(gdst) | bytes) & 15u) == 0;
      uint32_t* orow = reinterpret_cast<uint32_t*>(stage_out) + tid * (Din >> 1);
      const int nw_out = Din >> 1;
#pragma unroll
      for (int cb = 0; cb < DP / 32; ++cb) {
        float v[32];
        tc::ld_row32(t_dx + lane_base + cb * 32, v);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (cb * 16 + q < nw_out) orow[cb * 16 + q] = Fmt16<TP>::pack(v[2 * q], v[2 * q + 1]);
      }
      if constexpr (DP % 32 != 0) {
        float v[32];
        tc::ld_row32(t_dx + lane_base + (DP / 32) * 32 - 16, v);   // last 16 columns (overlapping read keeps the x32 shape)
#pragma unroll
        for (int q = 8; q < 16; ++q)
          if ((DP / 32) * 16 - 8 + q < nw_out) orow[(DP / 32) * 16 - 8 + q] = Fmt16<TP>::pack(v[2 * q], v[2 * q + 1]);
      }
      if (bulk_out) {
        ptx::fence_proxy_async_smem();
        tc::fence_before();
        __syncthreads();
        if (tid == 0) {
          ptx::bulk_s2g(gdst, stage_out, bytes, pol);
          ptx::bulk_commit();            // its shared-memory reads are waited for just before the output stage is written again
        }
        store_pending = true;
      } else {
        tc::fence_before();
        __syncthreads();
        for (uint32_t i = tid; i < bytes / 2; i += 128) reinterpret_cast<uint16_t*>(gdst)[i] = reinterpret_cast<const uint16_t*>(stage_out)[i];
      }
    }
    if (!GRAD && A.partials) {
      const double s = block_sum_f64<4>(acc, scratch);
      if (tid == 0) A.partials[static_cast<int64_t>(b) * A.chunks + c] = s;
    }
    if (!GRAD) {
      tc::fence_before();
      __syncthreads();                   // TMEM reads of this tile are complete before the next RAW product overwrites them
    }
    // (GRAD: the barrier before the bulk store already separates this tile's TMEM / stage accesses from the next tile's writes)
  }
  if (GRAD && dw_started && !S::kOutInG) {   // the last DW product
    ptx::mbar_wait(bar_dw, ph_dw);
    ph_dw ^= 1u;
  }
  if (GRAD && tid == 0 && store_pending) ptx::bulk_wait_read0();

  // ---- DW / db partial of this CTA: rows p (M = 64 accumulator: row m lives in TMEM lane (m % 16) + 32 (m / 16)) -----------
  if (GRAD && A.dw_partial) {
    tc::fence_after();
    float* out = A.dw_partial + static_cast<int64_t>(blockIdx.x) * 32 * DP;
    const int p = 16 * warp + lane;      // valid for lane < 16, warp < 2
#pragma unroll
    for (int cb = 0; cb < DP / 32; ++cb) {
      float v[32];
      tc::ld_row32(t_dw + lane_base + cb * 32, v);
      if (lane < 16 && warp < 2) {
#pragma unroll
        for (int q = 0; q < 32; ++q) out[p * DP + cb * 32 + q] = dw_started ? v[q] : 0.f;
      }
    }
    if constexpr (DP % 32 != 0) {
      float v[32];
      tc::ld_row32(t_dw + lane_base + (DP / 32) * 32 - 16, v);
      if (lane < 16 && warp < 2) {
#pragma unroll
        for (int q = 16; q < 32; ++q) out[p * DP + (DP / 32) * 32 - 16 + q] = dw_started ? v[q] : 0.f;
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

// dW (P, Din), db (P) = sum over CTAs of the partials.  One block per output row p: thread (g, d) adds the partials of CTAs g, g+G, ...
// for column d (independent coalesced loads), then the G group sums are added in group order -- a fixed order, so the result is
// deterministic for a given CTA count.  (A single thread per element walking all ~1200 CTAs took 50 us: a chain of dependent L2 loads.)
template <int DP>
__global__ void __launch_bounds__(1024) linear_dmol_reduce_kernel(const float* __restrict__ part, int64_t ctas, int Din, int P,
                                                                   float* __restrict__ dW, float* __restrict__ db) {
  constexpr int DL = DP <= 32 ? 32 : 128;   // threads along d
  constexpr int G = 1024 / DL;              // CTA groups
  __shared__ float acc[G][DL];
  const int p = blockIdx.x, d = threadIdx.x % DL, g = threadIdx.x / DL;
  float s = 0.f;
  if (d <= Din) {
    const float* src = part + (static_cast<int64_t>(g) * 32 + p) * DP + d;
    constexpr int64_t kStep = static_cast<int64_t>(G) * 32 * DP;
    int64_t c = g;
#pragma unroll 1
    for (; c + 3 * G < ctas; c += 4 * G, src += 4 * kStep) {
      const float a0 = src[0], a1 = src[kStep], a2 = src[2 * kStep], a3 = src[3 * kStep];
      s += a0; s += a1; s += a2; s += a3;
    }
    for (; c < ctas; c += G, src += kStep) s += src[0];
  }
  acc[g][d] = s;
  __syncthreads();
  if (g == 0 && d <= Din) {
    float t = acc[0][d];
#pragma unroll
    for (int j = 1; j < G; ++j) t += acc[j][d];
    if (d < Din) dW[p * Din + d] = t;
    else if (db) db[p] = t;
  }
  (void)P;
}

}  // namespace blvm
