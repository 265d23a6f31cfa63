// dmol_stream_kernel: persistent, software-pipelined variant of dmol_tile_kernel for sm_100a.
//
// The tile kernel runs one tile per CTA: load -> compute -> store are serialised inside the CTA and only overlap across
// the CTAs resident on an SM.  That is enough while a tile carries a lot of arithmetic (K >= 8), but for small K the
// per-tile issue time and the memory time are of the same size, the phases no longer hide each other, and a short
// launch loses a further ~0.3 wave to quantisation (4096 CTAs over 148 x 16 slots).  Here the grid is
// (#SMs x resident CTAs) and every CTA walks tiles blockIdx.x, +gridDim.x, ... through a ring of STAGES shared-memory
// stages:
//
//   thread 0 (producer side)                        all threads (consumer side)
//   ------------------------------------------      -------------------------------------------------------------
//   wait until the bulk STORE that last used   \
//   the stage has read it (wait_group.read)     |   mbarrier wait on full[stage]  (TMA transaction bytes)
//   arm full[stage'] with expect_tx             |   rows -> registers, value + gradient (blvm_math.cuh), gradient
//   cp.async.bulk  params slab + y slab        /    rows back IN PLACE, fence.proxy.async, ONE __syncthreads
//   (LOOKAHEAD tiles ahead)                         thread 0: cp.async.bulk shared -> global of the gradient slab
//
// so the loads of tile i+LOOKAHEAD and the store of tile i-1 are in flight while tile i is being evaluated.  y travels
// through the same TMA pipeline (its global-load latency would otherwise be exposed once per tile).  The masked fp64
// tile sum uses a double-buffered scratch so that the loop has a single CTA barrier per tile.
//
// Eligibility (checked on the host): every tile slab 16-byte aligned, i.e. T % 4 == 0, (T * 3K * sizeof(TP)) % 16 == 0 and
// 16-byte aligned base pointers; other shapes take the tile kernel (which has the element-wise fallback).  Same tiling,
// same partials layout; values and gradients bit-identical to the tile kernel (the fp64 tile sums too when the CTA has 128 threads).
#pragma once
#include "dmol_kernels.cuh"

namespace blvm {

template <int N>
__device__ __forceinline__ void bulk_wait_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int K, int TPB, typename TP>
struct StreamLayout {
  static constexpr int P = 3 * K;
  static constexpr int TILE = 128 * DmolSpt<K>::value;            // same tiling as the tile kernel (partials layout)
  static constexpr int SPT = TILE / TPB;
  static constexpr size_t kSlab = ((size_t(TILE) * P * sizeof(TP) + 127) / 128) * 128;
  static constexpr size_t kY = size_t(TILE) * sizeof(float);
  static constexpr size_t kStage = kSlab + kY;
  static constexpr size_t bytes(int stages) { return kStage * stages + 8 * stages + 2 * (TPB / 32) * sizeof(double) + 16; }
  static_assert(TILE % TPB == 0, "tile must be a multiple of the CTA size");
};

struct StreamTile {
  int64_t s0;     // first flat sample
  int n, nvalid;  // samples in the tile, valid (unmasked) prefix
  bool skip;
};

// Position of a tile as (utterance b, chunk c), advanced by the grid stride without a division: the persistent loop used to pay a
// 64-bit division per thread and tile (~100 instructions, 15 % of a K = 1 tile).  32-bit: the host rejects T >= 2^31 and more than
// 2^31 - 1 tiles.
struct TileCursor {
  unsigned b, c;
};
__device__ __forceinline__ TileCursor cursor_at(unsigned tile, unsigned chunks) {
  TileCursor k;
  k.b = tile / chunks;
  k.c = tile - k.b * chunks;
  return k;
}
__device__ __forceinline__ TileCursor cursor_next(TileCursor k, unsigned step_b, unsigned step_c, unsigned chunks) {
  k.b += step_b;
  k.c += step_c;
  if (k.c >= chunks) {
    k.c -= chunks;
    ++k.b;
  }
  return k;
}
// the load and the clamp are separate so that a prefetched length is not touched (= waited for) before it is needed
__device__ __forceinline__ int64_t row_len_raw(const DmolArgs& A, unsigned b) { return A.x_sl ? A.x_sl[b] : A.T; }
__device__ __forceinline__ int row_len_clamp(const DmolArgs& A, int64_t len) {
  return static_cast<int>(len < 0 ? 0 : (len > A.T ? A.T : len));
}
__device__ __forceinline__ int row_len(const DmolArgs& A, unsigned b) { return row_len_clamp(A, row_len_raw(A, b)); }

template <int TILE>
__device__ __forceinline__ StreamTile stream_tile(const DmolArgs& A, TileCursor k, int len) {
  StreamTile t;
  const int t0 = static_cast<int>(k.c) * TILE;
  t.n = min(TILE, static_cast<int>(A.T) - t0);
  t.s0 = static_cast<int64_t>(k.b) * A.T + t0;
  t.nvalid = max(0, min(t.n, len - t0));
  t.skip = (A.flags & kFlagSkipPadded) && t.nvalid == 0;
  return t;
}

template <int K, int TPB, int STAGES, int LOOKAHEAD, bool GRAD, int UMODE, typename TP, int LIK = kLikDmol>
__global__ void __launch_bounds__(TPB) dmol_stream_kernel(const DmolArgs A, const int64_t n_tiles) {
  using L = StreamLayout<K, TPB, TP>;
  constexpr int P = L::P, TILE = L::TILE, SPT = L::SPT, NW = TPB / 32;
  static_assert(LOOKAHEAD >= 1 && LOOKAHEAD < STAGES, "the stage being evaluated must not be a load target");
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kStage * STAGES);
  double* scratch = reinterpret_cast<double*>(smem + L::kStage * STAGES + 8 * STAGES + (8 * STAGES % 16 ? 8 : 0));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  ptx::pdl_launch_dependents();   // persistent grid: every CTA is resident, dependents may queue behind it right away
  if (first >= n_tiles) return;
  const int64_t count = (n_tiles - first + stride - 1) / stride;   // tiles of this CTA
  const uint64_t pol = ptx::policy_evict_first();
  float gs = A.gscale;
  if (GRAD && A.gscale_dev) gs *= static_cast<float>(*A.gscale_dev);

  // producer: arm the stage's barrier and start the two bulk loads of tile `it` (a fully padded tile under
  // kFlagSkipPadded is not read: the barrier still completes its phase so that the parity bookkeeping stays uniform)
  const unsigned chunks32 = static_cast<unsigned>(A.chunks);
  const unsigned step_b = static_cast<unsigned>(stride) / chunks32, step_c = static_cast<unsigned>(stride) - step_b * chunks32;
  int st_issue = 0;   // thread 0 only: ring position of the next load
  TileCursor pcur = cursor_at(static_cast<unsigned>(first), chunks32);   // thread 0 only: the next tile to load
  auto issue = [&]() {
    // the row length only matters to the producer when fully padded tiles are skipped (otherwise every tile is read)
    const StreamTile t = stream_tile<TILE>(A, pcur, (A.flags & kFlagSkipPadded) ? row_len(A, pcur.b) : static_cast<int>(A.T));
    pcur = cursor_next(pcur, step_b, step_c, chunks32);
    const int st = st_issue;
    st_issue = (st_issue + 1 == STAGES) ? 0 : st_issue + 1;
    unsigned char* base = smem + L::kStage * st;
    if (t.skip) {
      ptx::mbar_arrive_expect_tx(full + st, 0);
      return;
    }
    const uint32_t pbytes = static_cast<uint32_t>(t.n) * P * sizeof(TP), ybytes = static_cast<uint32_t>(t.n) * sizeof(float);
    ptx::mbar_arrive_expect_tx(full + st, pbytes + ybytes);
    ptx::bulk_g2s(base, static_cast<const TP*>(A.raw) + t.s0 * P, pbytes, full + st, pol);
    ptx::bulk_g2s(base + L::kSlab, A.y + t.s0, ybytes, full + st, pol);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(full + s, 1);
    ptx::fence_mbar_init();
    for (int64_t it = 0; it < LOOKAHEAD && it < count; ++it) issue();
  }
  __syncthreads();

  int st = 0;
  uint32_t parity = 0;
  TileCursor cur = cursor_at(static_cast<unsigned>(first), chunks32);
  int64_t len_next = row_len_raw(A, cur.b);
  for (int64_t it = 0; it < count; ++it) {
    // the row length of the NEXT tile is requested now and used one iteration later: its load latency used to sit in front of the
    // interior / tail decision of every tile
    const int len = row_len_clamp(A, len_next);
    const TileCursor nxt = cursor_next(cur, step_b, step_c, chunks32);
    if (it + 1 < count) len_next = row_len_raw(A, nxt.b);
    const StreamTile t = stream_tile<TILE>(A, cur, len);
    const int64_t tile_id = static_cast<int64_t>(cur.b) * A.chunks + cur.c;
    cur = nxt;
    TP* tile = reinterpret_cast<TP*>(smem + L::kStage * st);
    const float* ytile = reinterpret_cast<const float*>(smem + L::kStage * st + L::kSlab);

    if (tid == 0 && it + LOOKAHEAD < count) {
      // the target stage was last used by tile it+LOOKAHEAD-STAGES; STAGES-LOOKAHEAD-1 younger stores may still be reading
      if (GRAD) bulk_wait_read_n<STAGES - LOOKAHEAD - 1>();
      issue();
    }
    ptx::mbar_wait(full + st, parity);

    double acc = 0.0;
    // Interior tiles (full, fully valid, no per-sample upstream gradient) are the overwhelming majority: they run without
    // the per-sample bound / mask / skip tests, and the range check of y is folded into one flag per thread.
    const bool interior = !t.skip && t.n == TILE && t.nvalid == TILE && A.gout == nullptr;
    if (interior) {
      bool bad = false;
      if constexpr (K == 1 && UMODE == kUTiny && LIK == kLikDmol && SPT % 2 == 0) {
        // one component: two SAMPLES per packed instruction (bit-identical to the per-sample evaluation)
#pragma unroll
        for (int j = 0; j < SPT; j += 2) {
          const int ia = j * TPB + tid, ib = ia + TPB;
          const float ya = ytile[ia], yb = ytile[ib];
          bad |= !(ya <= 1.0f && ya >= -1.0f) | !(yb <= 1.0f && yb >= -1.0f);
          float ra[P], rb[P], La, Lb;
          RowIO<TP, P>::load(tile + ia * P, ra);
          RowIO<TP, P>::load(tile + ib * P, rb);
          dmol_k1_two_samples<GRAD>(ya, yb, ra, rb, gs, gs, A.C, La, Lb);
          if (GRAD) {
            RowIO<TP, P>::store(tile + ia * P, ra);
            RowIO<TP, P>::store(tile + ib * P, rb);
          }
          if (A.lp) {
            A.lp[t.s0 + ia] = La;
            A.lp[t.s0 + ib] = Lb;
          }
          acc += static_cast<double>(La);
          acc += static_cast<double>(Lb);
        }
      } else {
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          const int i = j * TPB + tid;
          const float yv = ytile[i];
          bad |= !(yv <= 1.0f && yv >= -1.0f);
          float r[P];
          TP* row = tile + i * P;
          RowIO<TP, P>::load(row, r);
          const float Lv = dmol_eval<K, GRAD, UMODE, LIK, kLinTP<TP, K>>(yv, r, gs, A.C, [&](float (&rr)[P]) { RowIO<TP, P>::load(row, rr); });
          if (GRAD) RowIO<TP, P>::store(row, r);
          if (A.lp) A.lp[t.s0 + i] = Lv;
          acc += static_cast<double>(Lv);
        }
      }
      if (LIK == kLikDmol && bad && A.err_flag) atomicOr(A.err_flag, 1);
    } else {
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const int i = j * TPB + tid;
        if (i < t.n) {
          float Lv = 0.f;
          float r[P];
          TP* row = tile + i * P;
          if (!t.skip) {
            const float yv = ytile[i];
            if (LIK == kLikDmol && !(yv <= 1.0f && yv >= -1.0f) && A.err_flag) atomicOr(A.err_flag, 1);
            float g = 0.f;
            if (GRAD) {
              g = (i < t.nvalid) ? gs : 0.f;
              if (A.gout) g *= ptx::ldg_stream(A.gout + t.s0 + i);
            }
            RowIO<TP, P>::load(row, r);
            Lv = dmol_eval<K, GRAD, UMODE, LIK, kLinTP<TP, K>>(yv, r, g, A.C, [&](float (&rr)[P]) { RowIO<TP, P>::load(row, rr); });
          } else {
#pragma unroll
            for (int q = 0; q < P; ++q) r[q] = 0.f;
          }
          if (GRAD) RowIO<TP, P>::store(row, r);
          const float Lm = (i < t.nvalid) ? Lv : Lv * 0.0f;   // log_prob * mask (NaN/inf propagate like `* 0`), vrnn.py:268
          if (A.lp) A.lp[t.s0 + i] = (A.flags & kFlagMaskOutput) ? Lm : Lv;
          acc += static_cast<double>(Lm);
        }
      }
    }

    double* sc = scratch + (it & 1) * NW;
    if (A.partials) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) sc[warp] = acc;
    }
    if (GRAD) ptx::fence_proxy_async_smem();   // this thread's gradient rows -> visible to the TMA unit
    __syncthreads();                           // the only CTA barrier per tile: rows written, warp sums posted, stage consumed
    if (tid == 0) {
      if (GRAD) {
        ptx::bulk_s2g(static_cast<TP*>(A.graw) + t.s0 * P, tile, static_cast<uint32_t>(t.n) * P * sizeof(TP), pol);
        ptx::bulk_commit();
      }
      if (A.partials) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += sc[w];   // same order as block_sum_f64: with TPB = 128 bit-identical to the tile kernel
        A.partials[tile_id] = s;
      }
    }
    if (++st == STAGES) {
      st = 0;
      parity ^= 1u;
    }
  }
  if (GRAD && tid == 0) bulk_wait_read_n<0>();   // shared memory must outlive the last stores' reads
}

}  // namespace blvm
